"""Builders shared by the DDPG tests: the same agent twice, once on the CPU oracle, once on the GPU drop-in."""
import numpy as np

from curious_b200 import synth


def ddpg_kwargs(n_modules=4, structure='curious', task_replay='replay_task_cp_buffer', normalize_obs=False,
                batch_size=256, hidden=256, layers=3, T=50, relative_goals=False, dimo=None):
    dims = synth.arm_dims(n_modules, dimo)
    ag_ids, g_ids = synth.arm_task_ids(n_modules)
    if structure == 'flat':
        dims = {k: v for k, v in dims.items() if k != 'task_descr'}
    gamma = 1. - 1. / T
    kw = dict(input_dims=dims, hidden=hidden, layers=layers, polyak=0.95, batch_size=batch_size, Q_lr=0.001,
              pi_lr=0.001, norm_eps=0.01, norm_clip=5, max_u=1., action_l2=1.0, clip_obs=200., T=T,
              rollout_batch_size=2, relative_goals=relative_goals, clip_pos_returns=True,
              clip_return=1. / (1. - gamma), normalize_obs=normalize_obs, gamma=gamma, structure=structure,
              tasks_ag_id=ag_ids, tasks_g_id=g_ids, task_replay=task_replay, eps_task=0.4)
    return kw, dims, ag_ids, g_ids


def goal_subtract(a, b):
    return a - b


def make_oracle_agent(kw, dims, ag_ids, g_ids, buffer_episodes=40, seed=0):
    from oracle import ddpg_oracle, her_oracle, replay_oracle
    from oracle.reward_oracle import ModuleDistanceReward
    T = kw['T']
    flat = kw['structure'] == 'flat'
    reward = ModuleDistanceReward(ag_ids, g_ids)
    if flat:
        sampler = her_oracle.make_sample_her_transitions('her', 4, reward, kw['task_replay'], tasks_ag_id=ag_ids,
                                                         tasks_g_id=g_ids)
    else:
        sampler = her_oracle.make_sample_multi_task_her_transitions('her', 4, kw['task_replay'], reward,
                                                                    tasks_ag_id=ag_ids, tasks_g_id=g_ids)
    shapes = synth.buffer_shapes(dims, T)
    if flat:
        shapes = {k: v for k, v in shapes.items() if k not in ('task_descr', 'change')}
    if 'buffer' in kw['task_replay']:
        buffers = [replay_oracle.ReplayBufferOracle(shapes, buffer_episodes * T, T, sampler)
                   for _ in range(len(g_ids) + 1)]
    else:
        buffers = replay_oracle.ReplayBufferOracle(shapes, buffer_episodes * T, T, sampler)
    agent = ddpg_oracle.DDPGOracle(sample_transitions=sampler, buffers=buffers,
                                   weights_rng=np.random.RandomState(seed), **kw)
    return agent


def make_gpu_agent(kw, dims, ag_ids, g_ids, buffer_episodes=40, seed=0, her_rng='numpy', **extra):
    from curious_b200 import her
    from curious_b200.ddpg import DDPG
    from curious_b200.replay_buffer import ReplayBuffer
    from curious_b200.reward import ModuleDistanceReward
    T = kw['T']
    flat = kw['structure'] == 'flat'
    reward = ModuleDistanceReward(ag_ids, g_ids)
    if flat:
        sampler = her.make_sample_her_transitions('her', 4, reward, kw['task_replay'], tasks_ag_id=ag_ids,
                                                  tasks_g_id=g_ids)
        net = 'baselines.her.actor_critic:ActorCritic'
    else:
        sampler = her.make_sample_multi_task_her_transitions('her', 4, kw['task_replay'], reward,
                                                             tasks_ag_id=ag_ids, tasks_g_id=g_ids)
        net = 'baselines.her.actor_critic:MultiTaskActorCritic'
    shapes = synth.buffer_shapes(dims, T)
    if flat:
        shapes = {k: v for k, v in shapes.items() if k not in ('task_descr', 'change')}
    if 'buffer' in kw['task_replay']:
        buffers = [ReplayBuffer(shapes, buffer_episodes * T, T, sampler) for _ in range(len(g_ids) + 1)]
    else:
        buffers = ReplayBuffer(shapes, buffer_episodes * T, T, sampler)
    agent = DDPG(network_class=net, scope='ddpg', subtract_goals=goal_subtract, sample_transitions=sampler,
                 buffers=buffers, seed=seed, her_rng=her_rng, **kw, **extra)
    return agent


def episode_stream(dims, T, n_calls, rollout_batch_size=2, seed=123, flat=False):
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n_calls):
        ep = synth.make_episodes(rng, rollout_batch_size, T, dims, change_dtype=bool)
        if flat:
            ep = {k: v for k, v in ep.items() if k not in ('task_descr', 'change')}
        out.append(ep)
    return out


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
