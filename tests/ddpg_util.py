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


# ------------------------------------------------------------------------------------------------
# The UNMODIFIED reference agent, run live over oracle/tf1_shim.py (build container only: needs /root/reference)
# ------------------------------------------------------------------------------------------------
REFERENCE_ROOT = '/root/reference'


class _NumPy1(object):
    """NumPy as the reference saw it (1.x), at the module boundary only: `np.int` exists (ddpg.py:282-318), and np.sqrt of a
    Python float hands back a Python float, which reproduces value-based casting in `(-a) * self.m` (mpi_adam.py:31-34:
    the step stays float32; NumPy >= 2 would promote it to float64)."""
    int = int

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def sqrt(x):
        return float(np.sqrt(x)) if isinstance(x, float) else np.sqrt(x)


def reference_modules():
    """(tf shim, baselines.her.ddpg, baselines.her.her, baselines.her.replay_buffer) with the reference source as it lies."""
    from oracle import tf1_shim
    tf = tf1_shim.install(REFERENCE_ROOT)
    import baselines.common.mpi_adam as ref_adam
    import baselines.her.ddpg as ref_ddpg
    import baselines.her.her as ref_her
    import baselines.her.replay_buffer as ref_rb
    from baselines import logger
    logger.set_level(logger.DISABLED)
    ref_ddpg.np = _NumPy1()
    ref_adam.np = _NumPy1()
    return tf, ref_ddpg, ref_her, ref_rb


def make_reference_agent(kw, dims, ag_ids, g_ids, buffer_episodes=40, init_seed=0):
    """baselines.her.ddpg.DDPG itself (graph built by the reference's own _create_network / actor_critic.py / util.nn* /
    normalizer.py / mpi_adam.py over the TF1 stand-in), its own ReplayBuffer and HER sampler; the reward closure is the
    oracle's (gym_flowers is absent)."""
    from oracle.reward_oracle import ModuleDistanceReward
    tf, ref_ddpg, ref_her, ref_rb = reference_modules()
    tf.reset_default_graph()
    tf.set_random_seed(init_seed)
    T = kw['T']
    flat = kw['structure'] == 'flat'
    reward = ModuleDistanceReward(ag_ids, g_ids)
    if flat:
        sampler = ref_her.make_sample_her_transitions('her', 4, reward, kw['task_replay'], tasks_ag_id=ag_ids,
                                                      tasks_g_id=g_ids)
        net = 'baselines.her.actor_critic:ActorCritic'
    else:
        sampler = ref_her.make_sample_multi_task_her_transitions('her', 4, kw['task_replay'], reward,
                                                                 tasks_ag_id=ag_ids, tasks_g_id=g_ids)
        net = 'baselines.her.actor_critic:MultiTaskActorCritic'
    shapes = synth.buffer_shapes(dims, T)
    if flat:
        shapes = {k: v for k, v in shapes.items() if k not in ('task_descr', 'change')}
    if 'buffer' in kw['task_replay']:
        buffers = [ref_rb.ReplayBuffer(shapes, buffer_episodes * T, T, sampler) for _ in range(len(g_ids) + 1)]
    else:
        buffers = ref_rb.ReplayBuffer(shapes, buffer_episodes * T, T, sampler)
    return ref_ddpg.DDPG(network_class=net, scope='ddpg', subtract_goals=goal_subtract, sample_transitions=sampler,
                         buffers=buffers, **kw)


def reference_flat(agent, which, target=False):
    """Flat parameter vector of the reference agent in its own GetFlat order (common/tf_util.py:239-246)."""
    vs = agent._vars(('target/' if target else 'main/') + which)
    return np.concatenate([v.value.numpy().astype(np.float32).ravel() for v in vs])


def load_reference_flat(agent, which, flat, target=False):
    """Write a flat float32 vector into the reference agent's variables (GetFlat order)."""
    k = 0
    for v in agent._vars(('target/' if target else 'main/') + which):
        n = int(np.prod(v.value.shape)) if v.value.ndim else 1
        v.load(np.asarray(flat[k:k + n], np.float32).reshape(tuple(v.value.shape)))
        k += n
    assert k == len(flat)


# ------------------------------------------------------------------------------------------------
# Seeded inputs of the reference-graph fixtures (tests/golden/ddpg/*.npz hold only the reference's OUTPUTS; every input
# is rebuilt from the seed by these recipes - NumPy's legacy RandomState stream is frozen).
# ------------------------------------------------------------------------------------------------
def seeded_flat(seed, n, scale):
    return (np.random.RandomState(seed).uniform(-1.0, 1.0, n) * scale).astype(np.float32)


def seeded_net_flats(seed, sizes, hidden):
    """{('Q'|'pi', target): flat}: four DIFFERENT parameter vectors (a main/target mix-up must show), biases included,
    scaled like a Xavier layer of this width so that activations stay O(1)."""
    scale = float(np.sqrt(6.0 / (2 * hidden)))
    return {(w, t): seeded_flat(seed + 10 * i + (5 if t else 0), sizes[w], scale)
            for i, w in enumerate(('Q', 'pi')) for t in (False, True)}


def seeded_stats(seed, size):
    """Normaliser state [sum, sumsq, count, mean, std] (normalizer.py:33-47 variable order) with a non-trivial mean/std."""
    rng = np.random.RandomState(seed)
    mean = rng.normal(0.0, 0.5, size).astype(np.float32)
    std = rng.uniform(0.5, 2.0, size).astype(np.float32)
    count = np.array([1000.0], np.float32)
    return [mean * count, (std * std + mean * mean) * count, count, mean, std]


def seeded_batch(seed, stage_keys, dims, n):
    """One staged batch in stage order (ddpg.py:76-84): o, o_2 ~ 2 N(0,1), goals ~ 0.3 U(-1,1), u ~ U(-1,1), one-hot
    task_descr, r in {-1, 0}."""
    rng = np.random.RandomState(seed)
    out = []
    td = None
    for k in stage_keys:
        base = k[:-2] if k.endswith('_2') else k
        if base == 'o':
            a = 2.0 * rng.standard_normal((n, dims['o']))
        elif base in ('g', 'ag'):
            a = 0.3 * rng.uniform(-1, 1, (n, dims[base]))
        elif base == 'u':
            a = rng.uniform(-1, 1, (n, dims['u']))
        elif base == 'task_descr':
            if td is None:
                td = np.eye(dims['task_descr'])[rng.randint(0, dims['task_descr'], n)]
            a = td
        elif base == 'r':
            a = -(rng.uniform(0, 1, (n, 1)) < 0.7).astype(np.float64)
        else:
            a = rng.standard_normal((n, dims[base]))
        out.append(np.asarray(a, np.float32))
    return out
