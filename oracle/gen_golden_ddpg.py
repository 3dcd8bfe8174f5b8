"""Generate tests/golden/ddpg/*.npz: outputs of the UNMODIFIED reference learner (test infrastructure).

Run in the build container only (needs /root/reference):      python oracle/gen_golden_ddpg.py

baselines/her/ddpg.py, actor_critic.py, util.py (nn, nn_modular_her, flatten_grads), normalizer.py, common/tf_util.py
(GetFlat / SetFromFlat) and common/mpi_adam.py are imported as they lie and build their TF1 graph over oracle/tf1_shim.py
(TensorFlow 1.x is not installable here; the stand-in supplies the TF primitives, the reference supplies the graph).
Per case the script instantiates baselines.her.ddpg.DDPG, loads seeded parameters / normaliser statistics (recipes in
tests/ddpg_util.py), stages seeded batches with the reference's own stage_batch, and records what the reference computes:

    Q_loss[k], pi_loss[k], Q_pi[k]      float64 values of Q_loss_tf, pi_loss_tf, main.Q_pi_tf at update k (ddpg.py:412-449)
    Q_grad0, pi_grad0                   the flattened gradients of update 0 (ddpg.py:443-449), float32 storage
    main_* / target_* after             the parameters after all updates (MpiAdam.update, update_target_net every 2nd update),
                                        full vectors for the small nets, every 8th element for hidden = 256
    act_u*, act_q*                      get_actions(compute_Q=True) of the main and the target network (ddpg.py:129-161)
    weights_pkl                         (small nets) the bytes of the file the reference's save_weights wrote at the end
    relu_margin0                        min |hidden pre-activation| of update 0 - computed by oracle/ddpg_oracle.py on the same
                                        inputs (the reference does not expose it); lets a test tell a ReLU-kink flip from an error
    meta                                JSON: the DDPG kwargs, seeds, number of updates

`agent_*` files are whole-agent trajectories (run_agent_case): the reference's store_episode (module routing, the normaliser
update through its HER sampler) and train() WITH its own sample_batch, ReplayBuffer and HER sampler under recorded
np.random seeds: normaliser statistics, buffer fill levels, LP proportions, Q_loss / Q_pi per update, parameters after, and the
np.random state the run ends on.

Nothing from the reference is copied into the repo: only its outputs.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'tests', 'golden', 'ddpg')

CASES = [
    dict(name='arm4_h64', n_modules=4, structure='curious', hidden=64, layers=3, batch=64, normalize_obs=True, updates=4),
    dict(name='arm4_h64_raw_rel', n_modules=4, structure='curious', hidden=64, layers=3, batch=64, normalize_obs=False,
         relative_goals=True, updates=3),
    dict(name='arm8_h64_l2', n_modules=8, structure='curious', hidden=64, layers=2, batch=48, normalize_obs=True, updates=3),
    dict(name='arm4_h64_l4_posret', n_modules=4, structure='curious', hidden=64, layers=4, batch=64, normalize_obs=True,
         clip_pos_returns=False, updates=3),
    dict(name='experts_h64', n_modules=4, structure='task_experts', hidden=64, layers=3, batch=64, normalize_obs=True, updates=3),
    dict(name='flat_h64', n_modules=4, structure='flat', hidden=64, layers=3, batch=64, normalize_obs=True, updates=4),
    dict(name='arm4_h256_b256', n_modules=4, structure='curious', hidden=256, layers=3, batch=256, normalize_obs=True, updates=3),
    dict(name='flat_h256_b256', n_modules=4, structure='flat', hidden=256, layers=3, batch=256, normalize_obs=True, updates=2),
    dict(name='arm4_h256_b1024', n_modules=4, structure='curious', hidden=256, layers=3, batch=1024, normalize_obs=True,
         updates=2),
    dict(name='arm8_h256_b256', n_modules=8, structure='curious', hidden=256, layers=3, batch=256, normalize_obs=False, updates=2),
]


def case_kwargs(case):
    from tests.ddpg_util import ddpg_kwargs
    flat = case['structure'] == 'flat'
    kw, dims, ag_ids, g_ids = ddpg_kwargs(case['n_modules'], structure=case['structure'],
                                          task_replay='' if flat else 'replay_task_cp_buffer',
                                          normalize_obs=case['normalize_obs'], batch_size=case['batch'],
                                          hidden=case['hidden'], layers=case['layers'],
                                          relative_goals=case.get('relative_goals', False))
    kw['clip_pos_returns'] = case.get('clip_pos_returns', True)
    if case['structure'] == 'task_experts':
        kw['t_id'] = 1
    return kw, dims, ag_ids, g_ids


def action_inputs(seed, dims, n):
    rng = np.random.RandomState(seed)
    o = (2.0 * rng.standard_normal((n, dims['o']))).astype(np.float32)
    ag = (0.3 * rng.uniform(-1, 1, (n, dims['ag']))).astype(np.float32)
    g = (0.3 * rng.uniform(-1, 1, (n, dims['g']))).astype(np.float32)
    td = None
    if 'task_descr' in dims:
        td = np.eye(dims['task_descr'], dtype=np.float32)[rng.randint(0, dims['task_descr'], n)]
    return o, ag, g, td


def run_case(case, seed):
    from tests.ddpg_util import (load_reference_flat, make_oracle_agent, make_reference_agent, reference_flat, seeded_batch,
                                 seeded_net_flats, seeded_stats)
    from oracle import ddpg_oracle as D
    kw, dims, ag_ids, g_ids = case_kwargs(case)
    ref = make_reference_agent(kw, dims, ag_ids, g_ids, buffer_episodes=2)
    sizes = {w: reference_flat(ref, w).size for w in ('Q', 'pi')}
    flats = seeded_net_flats(seed, sizes, case['hidden'])
    for (w, t), f in flats.items():
        load_reference_flat(ref, w, f, t)
    for stats, sd, size in ((ref.o_stats, seed + 100, dims['o']), (ref.g_stats, seed + 101, dims['g'])):
        for v, a in zip((stats.sum_tf, stats.sumsq_tf, stats.count_tf, stats.mean, stats.std), seeded_stats(sd, size)):
            v.load(a)
    stage_keys = list(ref.stage_shapes.keys())
    rec = {}
    K = case['updates']
    ql, pl, qpi = [], [], []
    for k in range(K):
        batch = seeded_batch(seed + 1000 + k, stage_keys, dims, case['batch'])
        ref.stage_batch(batch)
        v = ref.sess.run64([ref.Q_loss_tf, ref.pi_loss_tf, ref.main.Q_pi_tf, ref.Q_grad_tf, ref.pi_grad_tf])
        ql.append(v[0]); pl.append(v[1]); qpi.append(v[2])
        if k == 0:
            rec['Q_grad0'], rec['pi_grad0'] = v[3].astype(np.float32), v[4].astype(np.float32)
            ora = make_oracle_agent(kw, dims, ag_ids, g_ids, buffer_episodes=2)
            ora.main_Q, ora.main_pi = D.unflatten(flats[('Q', False)], ora.ac.Q_shapes), D.unflatten(flats[('pi', False)], ora.ac.pi_shapes)
            ora.target_Q, ora.target_pi = D.unflatten(flats[('Q', True)], ora.ac.Q_shapes), D.unflatten(flats[('pi', True)], ora.ac.pi_shapes)
            for stats, sd, size in ((ora.o_stats, seed + 100, dims['o']), (ora.g_stats, seed + 101, dims['g'])):
                stats.sum, stats.sumsq, stats.count, stats.mean, stats.std = seeded_stats(sd, size)
            rec['relu_margin0'] = np.float64(ora.grads(batch)['relu_margin'])
        ref.train(stage=False)                      # the reference's own _grads + MpiAdam.update (ddpg.py:367-373)
        if k % 2 == 1:
            ref.update_target_net()
    rec['Q_loss'], rec['pi_loss'], rec['Q_pi'] = np.array(ql), np.array(pl), np.stack(qpi)
    stride = 8 if case['hidden'] >= 256 else 1
    for w in ('Q', 'pi'):
        rec['main_%s_after' % w] = reference_flat(ref, w)[::stride]
        rec['target_%s_after' % w] = reference_flat(ref, w, True)[::stride]
    o, ag, g, td = action_inputs(seed + 2000, dims, 7)
    for tgt in (False, True):
        u, q = ref.get_actions(o, ag, g, task_descr=td, use_target_net=tgt, compute_Q=True)
        rec['act_u_%s' % ('target' if tgt else 'main')] = np.asarray(u, np.float64)
        rec['act_q_%s' % ('target' if tgt else 'main')] = np.asarray(q, np.float64)
    if case['hidden'] < 256:
        import tempfile
        with tempfile.TemporaryDirectory() as d:         # the reference's own on-disk format (ddpg.py:481-497)
            ref.save_weights(os.path.join(d, 'policy'))
            rec['weights_pkl'] = np.frombuffer(open(os.path.join(d, 'policy_weights.pkl'), 'rb').read(), np.uint8)
    meta = dict(case=case, seed=seed, stride=stride, stage_keys=stage_keys,
                variables_Q=[v.name for v in ref._vars('main/Q')], variables_pi=[v.name for v in ref._vars('main/pi')],
                stats_variables=[v.name for v in ref._global_vars('o_stats')])
    rec['meta'] = json.dumps(meta)
    return rec


# ---------------------------------------------------------------------------------------------------------------
# whole-agent trajectories: store_episode (routing, normaliser update through the HER sampler) -> train() with sampling
# ---------------------------------------------------------------------------------------------------------------
AGENT_CASES = [
    dict(name='agent_arm4_h64', n_modules=4, structure='curious', task_replay='replay_task_cp_buffer', hidden=64, batch=64,
         normalize_obs=True, stores=6, updates=5),
    dict(name='agent_arm8_h64_random_buffer', n_modules=8, structure='curious', task_replay='replay_task_random_buffer',
         hidden=64, batch=64, normalize_obs=True, stores=8, updates=4),
    dict(name='agent_arm4_h64_cp_transition', n_modules=4, structure='curious', task_replay='replay_cp_task_transition',
         hidden=64, batch=64, normalize_obs=True, relative_goals=True, stores=5, updates=4),
    dict(name='agent_flat_h64', n_modules=4, structure='flat', task_replay='', hidden=64, batch=64, normalize_obs=True,
         stores=5, updates=4),
    dict(name='agent_arm4_h256_b256', n_modules=4, structure='curious', task_replay='replay_task_cp_buffer', hidden=256,
         batch=256, normalize_obs=True, stores=8, updates=4),
]


def agent_kwargs(case):
    from tests.ddpg_util import ddpg_kwargs
    return ddpg_kwargs(case['n_modules'], structure=case['structure'], task_replay=case['task_replay'],
                       normalize_obs=case['normalize_obs'], batch_size=case['batch'], hidden=case['hidden'],
                       relative_goals=case.get('relative_goals', False))


def run_agent_case(case, seed):
    """The reference agent end to end.  Inputs by recipe: parameters seeded_net_flats(seed), episodes
    episode_stream(seed=seed + 1), np.random.seed(seed + 2) before the stores, np.random.seed(seed + 10 + k) before train k."""
    from tests.ddpg_util import (episode_stream, load_reference_flat, make_reference_agent, reference_flat, seeded_net_flats)
    kw, dims, ag_ids, g_ids = agent_kwargs(case)
    ref = make_reference_agent(kw, dims, ag_ids, g_ids, buffer_episodes=40)
    sizes = {w: reference_flat(ref, w).size for w in ('Q', 'pi')}
    for (w, t), f in seeded_net_flats(seed, sizes, case['hidden']).items():
        load_reference_flat(ref, w, f, t)
    cp = np.linspace(0.02, 0.3, case['n_modules'])
    flat = case['structure'] == 'flat'
    state = np.random.get_state()
    rec = {}
    try:
        np.random.seed(seed + 2)
        n = 0
        for ep in episode_stream(dims, kw['T'], case['stores'], seed=seed + 1, flat=flat):
            n += 2
            ref.store_episode({k: v.copy() for k, v in ep.items()}, cp, n)
        for tag, st in (('o', ref.o_stats), ('g', ref.g_stats)):
            rec['stats_%s_mean' % tag] = st.mean.value.numpy().astype(np.float64)
            rec['stats_%s_std' % tag] = st.std.value.numpy().astype(np.float64)
            rec['stats_%s_count' % tag] = st.count_tf.value.numpy().astype(np.float64)
        bufs = ref.buffer if isinstance(ref.buffer, list) else [ref.buffer]
        rec['buffer_sizes'] = np.array([b.current_size for b in bufs])
        ql, qpi, props = [], [], []
        for k in range(case['updates']):
            np.random.seed(seed + 10 + k)
            ref.stage_batch()                               # sample_batch + stage (ddpg.py:362-366)
            v = ref.sess.run64([ref.Q_loss_tf, ref.main.Q_pi_tf])
            ql.append(v[0]); qpi.append(v[1])
            if hasattr(ref, 'proportions'):
                props.append(np.asarray(ref.proportions, np.int64))
            ref.train(stage=False)
            if k % 2 == 1:
                ref.update_target_net()
        rec['Q_loss'], rec['Q_pi'] = np.array(ql), np.stack(qpi)
        if props:
            rec['proportions'] = np.stack(props)
        rec['rng_after'] = np.random.get_state()[1].copy()
    finally:
        np.random.set_state(state)
    stride = 8 if case['hidden'] >= 256 else 1
    for w in ('Q', 'pi'):
        rec['main_%s_after' % w] = reference_flat(ref, w)[::stride]
        rec['target_%s_after' % w] = reference_flat(ref, w, True)[::stride]
    rec['meta'] = json.dumps(dict(case=case, seed=seed, stride=stride, cp=list(cp)))
    return rec


def main():
    os.makedirs(OUT, exist_ok=True)
    for i, case in enumerate(AGENT_CASES):
        rec = run_agent_case(case, seed=9000 + 41 * i)
        path = os.path.join(OUT, case['name'] + '.npz')
        np.savez_compressed(path, **rec)
        print('%-30s Q_loss %s  %.0f KB' % (case['name'], np.round(rec['Q_loss'], 6), os.path.getsize(path) / 1024))
    if '--agents-only' in sys.argv:
        return
    for i, case in enumerate(CASES):
        rec = run_case(case, seed=7000 + 37 * i)
        path = os.path.join(OUT, case['name'] + '.npz')
        np.savez_compressed(path, **rec)
        print('%-22s Q_loss %s  %.0f KB' % (case['name'], np.round(rec['Q_loss'], 6), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
