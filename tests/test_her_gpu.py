"""GPU parity of the fused HER kernel (through the C ABI) against the golden vectors recorded from the
unmodified reference and against the NumPy oracle.  Bit-exact (integer / float32-representable data)."""
import numpy as np
import pytest

from tests.golden_util import future_p, load_case, per_row_choices, sampler_cases

pytestmark = pytest.mark.gpu


def _mk_sampler(meta, rng='numpy'):
    from curious_b200 import her
    from curious_b200.reward import ModuleDistanceReward
    reward = ModuleDistanceReward(meta['tasks_ag_id'], meta['tasks_g_id'], meta['threshold'])
    if meta['flat']:
        s = her.make_sample_her_transitions(meta['goal_replay'], meta['her_replay_k'], reward, meta['task_replay'],
                                            tasks_ag_id=meta['tasks_ag_id'], tasks_g_id=meta['tasks_g_id'])
    else:
        s = her.make_sample_multi_task_her_transitions(meta['goal_replay'], meta['her_replay_k'],
                                                       meta['task_replay'], reward,
                                                       tasks_ag_id=meta['tasks_ag_id'], tasks_g_id=meta['tasks_g_id'])
    s.rng = rng
    return s


@pytest.mark.parametrize('name', sampler_cases())
def test_dropin_matches_reference_golden(name):
    """Seed np.random like the reference run and call the drop-in the way the reference is called:
    the kernel must return the reference's transitions bit for bit."""
    from curious_b200.replay_buffer import ReplayBuffer
    meta, eps, stream, ref = load_case(name)
    sampler = _mk_sampler(meta)
    T = meta['T']
    kw = {} if meta['flat'] else dict(task_to_replay=meta['task_to_replay'], cp_proba=meta['cp_proba'])
    np.random.seed(meta['seed'])
    if meta['via_buffer']:
        shapes = {k: v.shape[1:] for k, v in eps.items()}
        buf = ReplayBuffer(shapes, (meta['E'] + 3) * T, T, sampler)
        buf.store_episode({k: v.copy() for k, v in eps.items()})
        out = buf.sample(meta['B'], **kw)
    else:
        batch = {k: v.copy() for k, v in eps.items()}
        batch['o_2'] = batch['o'][:, 1:, :]
        batch['ag_2'] = batch['ag'][:, 1:, :]
        out = sampler(batch, meta['B'], **kw)
    assert set(ref.keys()) <= set(out.keys())
    for k in ref:
        assert out[k].shape == ref[k].shape, k
        assert np.array_equal(out[k], ref[k].astype(np.float64)), k


def _arm_setup(n_modules, E, T, seed=0):
    from curious_b200 import synth
    dims = synth.arm_dims(n_modules)
    ag_ids, g_ids = synth.arm_task_ids(n_modules)
    eps = synth.make_episodes(np.random.RandomState(seed), E, T, dims)
    return dims, ag_ids, g_ids, eps


@pytest.mark.parametrize('task_replay,ttr', [('replay_task_cp_buffer', 2), ('replay_task_cp_buffer', None),
                                             ('replay_random_task_transition', None),
                                             ('replay_cp_task_transition', None),
                                             ('replay_current_task_transition', None)])
@pytest.mark.parametrize('n_modules', [4, 8])
def test_kernel_vs_oracle_large(task_replay, ttr, n_modules):
    """Arm4/Arm8-shaped buffers, 8192 rows per call, same np.random seed on both sides."""
    from curious_b200 import her
    from curious_b200.replay_buffer import ReplayBuffer
    from curious_b200.reward import ModuleDistanceReward
    from oracle import her_oracle, replay_oracle
    from oracle.reward_oracle import ModuleDistanceReward as OracleReward
    T, E, B = 50, 300, 8192
    dims, ag_ids, g_ids, eps = _arm_setup(n_modules, E, T)
    cp = np.linspace(1, 2, n_modules)
    cp = cp / cp.sum()
    cp_proba = cp if task_replay == 'replay_cp_task_transition' else None
    shapes = {k: v.shape[1:] for k, v in eps.items()}

    o_s = her_oracle.make_sample_multi_task_her_transitions('her', 4, task_replay, OracleReward(ag_ids, g_ids),
                                                            tasks_ag_id=ag_ids, tasks_g_id=g_ids)
    obuf = replay_oracle.ReplayBufferOracle(shapes, E * T, T, o_s)
    obuf.store_episode({k: v.copy() for k, v in eps.items()})
    np.random.seed(77)
    ref = obuf.sample(B, task_to_replay=ttr, cp_proba=cp_proba)

    g_s = her.make_sample_multi_task_her_transitions('her', 4, task_replay, ModuleDistanceReward(ag_ids, g_ids),
                                                     tasks_ag_id=ag_ids, tasks_g_id=g_ids)
    gbuf = ReplayBuffer(shapes, E * T, T, g_s)
    gbuf.store_episode({k: v.copy() for k, v in eps.items()})
    np.random.seed(77)
    out = gbuf.sample(B, task_to_replay=ttr, cp_proba=cp_proba)
    for k in ref:
        assert np.array_equal(out[k], np.asarray(ref[k], np.float64)), k
    # both reward branches must be exercised for the comparison to mean something
    assert 0.05 < (-ref['r']).mean() < 0.95


@pytest.mark.parametrize('task_replay,ttr,flat', [('replay_task_cp_buffer', 1, False), ('replay_task_cp_buffer', None, False),
                                                  ('replay_random_task_transition', None, False),
                                                  ('replay_current_task_transition', None, False), ('', None, True)])
def test_reward_table_kinds_vs_oracle(task_replay, ttr, flat):
    """Every rule of the reward table (distance with per-module thresholds, pair = offset between two achieved-goal
    slices, info passthrough) inside the fused kernel vs the oracle's restatement, all sampler modes; bit-exact.
    (Rules are restatements: parity with gym_flowers' compute_reward is unpinned)"""
    from curious_b200 import her
    from curious_b200.replay_buffer import ReplayBuffer
    from curious_b200.reward import ModuleRewardTable
    from oracle import her_oracle, replay_oracle
    from oracle.reward_oracle import ModuleRewardTable as OracleTable
    T, E, B, N = 50, 200, 8192, 4
    dims, ag_ids, g_ids, eps = _arm_setup(N, E, T, seed=21)
    spec = dict(threshold=[0.05, 0.25, 0.05, 0.12], kinds=['distance', 'pair', 'info', 'distance'],
                ref_ag_id=[None, [0, 1, 2], None, None], info_keys=[None, None, 'is_success', None], flat_threshold=0.35)
    shapes = {k: v.shape[1:] for k, v in eps.items()}
    if flat:
        o_s = her_oracle.make_sample_her_transitions('her', 4, OracleTable(ag_ids, g_ids, **spec), '', tasks_ag_id=ag_ids,
                                                     tasks_g_id=g_ids)
        g_s = her.make_sample_her_transitions('her', 4, ModuleRewardTable(ag_ids, g_ids, **spec), '', tasks_ag_id=ag_ids,
                                              tasks_g_id=g_ids)
        kw = {}
    else:
        o_s = her_oracle.make_sample_multi_task_her_transitions('her', 4, task_replay, OracleTable(ag_ids, g_ids, **spec),
                                                                tasks_ag_id=ag_ids, tasks_g_id=g_ids)
        g_s = her.make_sample_multi_task_her_transitions('her', 4, task_replay, ModuleRewardTable(ag_ids, g_ids, **spec),
                                                         tasks_ag_id=ag_ids, tasks_g_id=g_ids)
        kw = dict(task_to_replay=ttr)
    obuf = replay_oracle.ReplayBufferOracle(shapes, E * T, T, o_s)
    obuf.store_episode({k: v.copy() for k, v in eps.items()})
    np.random.seed(5)
    ref = obuf.sample(B, **kw)
    gbuf = ReplayBuffer(shapes, E * T, T, g_s)
    gbuf.store_episode({k: v.copy() for k, v in eps.items()})
    np.random.seed(5)
    out = gbuf.sample(B, **kw)
    for k in ref:
        assert np.array_equal(out[k], np.asarray(ref[k], np.float64)), k
    if not flat:
        mod = np.argmax(ref['task_descr'], axis=1)
        for m in (range(N) if ttr is None else [ttr]):
            rows = mod == m
            assert rows.sum() > 50 and 0.0 < (-ref['r'][rows]).mean() < 1.0, (m, rows.sum(), (-ref['r'][rows]).mean())
        rows = mod == 2                                              # info rule: r = info_is_success - 1
        assert np.array_equal(ref['r'][rows, 0], ref['info_is_success'][rows, 0] - 1.0)


def test_store_roundtrip_and_overwrite_policy():
    """store_episode -> packed rows -> .buffers gives back what was stored; slots follow the reference's
    sequential-then-random policy with the same np.random seed."""
    from curious_b200.replay_buffer import ReplayBuffer
    from oracle import replay_oracle
    T = 50
    dims, ag_ids, g_ids, eps = _arm_setup(4, 9, T, seed=3)
    shapes = {k: v.shape[1:] for k, v in eps.items()}
    gbuf = ReplayBuffer(shapes, 5 * T, T, None)
    obuf = replay_oracle.ReplayBufferOracle(shapes, 5 * T, T, None)
    np.random.seed(5)
    for e in range(9):
        gbuf.store_episode({k: v[e:e + 1] for k, v in eps.items()})
    np.random.seed(5)
    for e in range(9):
        obuf.store_episode({k: v[e:e + 1] for k, v in eps.items()})
    assert gbuf.current_size == obuf.current_size == 5
    assert gbuf.get_transitions_stored() == obuf.get_transitions_stored()
    host = gbuf.buffers
    for k in shapes:
        assert np.array_equal(host[k], obuf.buffers[k]), k


@pytest.mark.parametrize('mode', ['buffer', 'random', 'cp'])
def test_philox_draws_and_relabel(mode):
    """Philox mode: the kernel's own draws equal the NumPy restatement of the documented mapping, and the
    transitions equal the oracle fed with those draws."""
    from curious_b200 import her
    from curious_b200.replay_buffer import ReplayBuffer
    from curious_b200.reward import ModuleDistanceReward
    from oracle import her_oracle, philox_oracle, replay_oracle
    from oracle.reward_oracle import ModuleDistanceReward as OracleReward
    task_replay = {'buffer': 'replay_task_cp_buffer', 'random': 'replay_random_task_transition',
                   'cp': 'replay_cp_task_transition'}[mode]
    T, E, B, N = 50, 123, 5000, 4
    dims, ag_ids, g_ids, eps = _arm_setup(N, E, T, seed=9)
    shapes = {k: v.shape[1:] for k, v in eps.items()}
    cp = np.array([0.1, 0.4, 0.3, 0.2])
    g_s = her.make_sample_multi_task_her_transitions('her', 4, task_replay, ModuleDistanceReward(ag_ids, g_ids),
                                                     tasks_ag_id=ag_ids, tasks_g_id=g_ids)
    g_s.rng, g_s.seed, g_s.calls = 'philox', 0x1234567890ABCDEF, 41
    gbuf = ReplayBuffer(shapes, E * T, T, g_s)
    gbuf.store_episode({k: v.copy() for k, v in eps.items()})
    ttr = 1 if mode == 'buffer' else None
    res = g_s.sample_device([(gbuf.device_view(), B, ttr)], B, cp_proba=cp if mode == 'cp' else None, want_idx=True)
    idx = res['idx'].cpu().numpy()
    cdf = cp.cumsum() / cp.cumsum()[-1]
    d = philox_oracle.draws(np.arange(B), np.full(B, E), T, 0x1234567890ABCDEF, 41,
                            mode=mode if mode != 'buffer' else None, n_tasks=N, cdf=cdf)
    assert np.array_equal(idx[:, 0], d['ep'])
    assert np.array_equal(idx[:, 1], d['t'])
    her_rows = d['u_her'] < 0.8
    ft = np.where(her_rows, d['t'] + 1 + (d['u_off'] * (T - d['t'])).astype(int), -1)
    assert np.array_equal(idx[:, 2], ft)
    assert 0.77 < her_rows.mean() < 0.83
    # oracle on the same draws
    o_s = her_oracle.make_sample_multi_task_her_transitions('her', 4, task_replay, OracleReward(ag_ids, g_ids),
                                                            tasks_ag_id=ag_ids, tasks_g_id=g_ids)
    obuf = replay_oracle.ReplayBufferOracle(shapes, E * T, T, o_s)
    obuf.store_episode({k: v.copy() for k, v in eps.items()})
    choice = np.where(her_rows, d['choice'], -1)
    ref = obuf.sample(B, task_to_replay=ttr, cp_proba=cp if mode == 'cp' else None,
                      stream=her_oracle.HerStream(d['ep'], d['t'], d['u_her'], d['u_off'], choice))
    out = g_s.to_host_dict(res, gbuf.device_view())
    for k in ref:
        assert np.array_equal(out[k], np.asarray(ref[k], np.float64)), k


def test_philox_device_known_answers():
    """Random123 known-answer vectors through the device function the kernels draw with (cur_philox4x32_10), plus random
    (counter, key) pairs against the oracle restatement."""
    import ctypes as C
    import torch
    from curious_b200 import _lib
    from oracle import philox_oracle
    from tests.test_oracle_golden import PHILOX_KAT
    rng = np.random.RandomState(0)
    extra = rng.randint(0, 2 ** 32, size=(4096, 6), dtype=np.uint64).astype(np.uint32)
    inp = np.concatenate([np.array([list(c) + list(k) for c, k, _ in PHILOX_KAT], np.uint32), extra])
    d_in = torch.from_numpy(inp.view(np.int32)).cuda()
    d_out = torch.zeros((inp.shape[0], 4), dtype=torch.int32, device='cuda')
    _lib.check(_lib.load().cur_philox4x32_10(_lib.stream_ptr(), d_in.data_ptr(), inp.shape[0], d_out.data_ptr()),
               'cur_philox4x32_10')
    out = d_out.cpu().numpy().view(np.uint32)
    for i, (_, _, want) in enumerate(PHILOX_KAT):
        assert tuple(int(x) for x in out[i]) == want
    for j in range(0, extra.shape[0], 16):                 # the oracle takes one key per call
        c = extra[j]
        got = philox_oracle.philox4x32_10(*[np.array([int(x)]) for x in c[:4]], int(c[4]), int(c[5]))
        assert tuple(int(x[0]) for x in got) == tuple(int(x) for x in out[len(PHILOX_KAT) + j])


def test_full_size_properties():
    """BASELINE config 2 shape: 1e6-transition Arm4 buffer (20 000 episodes x T=50), 2^20 rows per launch.
    Size-independent properties instead of a CPU comparison."""
    import torch
    from curious_b200 import her, synth
    from curious_b200.replay_buffer import ReplayBuffer
    from curious_b200.reward import ModuleDistanceReward
    T, N = 50, 4
    dims = synth.arm_dims(N)
    ag_ids, g_ids = synth.arm_task_ids(N)
    shapes = synth.buffer_shapes(dims, T)
    s = her.make_sample_multi_task_her_transitions('her', 4, 'replay_task_cp_buffer',
                                                   ModuleDistanceReward(ag_ids, g_ids), tasks_ag_id=ag_ids,
                                                   tasks_g_id=g_ids)
    s.rng = 'philox'
    buf = ReplayBuffer(shapes, 1000000, T, s)
    rng = np.random.RandomState(0)
    chunk = 2000
    for _ in range(buf.size // chunk):
        buf.store_episode(synth.make_episodes(rng, chunk, T, dims))
    assert buf.full
    B = 1 << 20
    s.calls = 7
    a = s.sample_device([(buf.device_view(), B, 2)], B, want_idx=True)
    s.calls = 7
    b = s.sample_device([(buf.device_view(), B, 2)], B, want_idx=True)
    torch.cuda.synchronize()
    for k in ('o', 'g', 'td', 'r', 'o_2', 'idx'):
        assert torch.equal(a[k], b[k]), 'same counter must give the same sample: ' + k
    idx = a['idx'].cpu().numpy()
    ep, t, ft, mod = idx.T
    assert ep.min() >= 0 and ep.max() < buf.size and t.min() >= 0 and t.max() < T
    her_rows = ft >= 0
    assert abs(her_rows.mean() - 0.8) < 0.01
    assert (ft[her_rows] >= t[her_rows] + 1).all() and (ft[her_rows] <= T).all()
    assert (mod[her_rows] == 2).all() and (mod[~her_rows] == -1).all()
    td = a['td'].cpu().numpy()
    g = a['g'].cpu().numpy()
    r = a['r'].cpu().numpy()
    assert np.array_equal(td.sum(1), np.ones(B, np.float32))
    assert (td[her_rows, 2] == 1).all()
    off = np.ones(dims['g'], bool)
    off[g_ids[2]] = False
    assert (g[her_rows][:, off] == 0).all()                      # whole goal cleared outside the module slice
    assert set(np.unique(r)) <= {-1.0, 0.0}
    assert 0.02 < (r == 0).mean() < 0.98
    # future_t == T rows must reproduce the final achieved goal exactly -> recomputed reward is 0 there only if
    # the row's ag_2 equals it; check the gather itself on a sample of rows instead
    host = buf.storage.view(buf.size, T, buf.layout.trans_stride)
    sel = np.where(her_rows)[0][:4096]
    L = buf.layout
    a0 = (dims['o'] + 3) // 4 * 4 + L.off_ag                  # ag(t+1) block of a transition row; ag(ft) sits in transition ft - 1
    fut = host[torch.from_numpy(ep[sel]).long().cuda(), torch.from_numpy(ft[sel] - 1).long().cuda()][:, a0:a0 + dims['ag']]
    assert torch.equal(a['g'][torch.from_numpy(sel).cuda()][:, g_ids[2]], fut[:, ag_ids[2]])


def test_error_behaviour_follows_the_reference():
    """The reference signals misuse with Python asserts / KeyErrors (SURVEY 8b "Errors"); the drop-in raises the same
    exception types at the same call sites, and both sides agree case by case."""
    from curious_b200.replay_buffer import ReplayBuffer
    from oracle import replay_oracle
    T = 50
    dims, ag_ids, g_ids, eps = _arm_setup(4, 7, T, seed=4)
    shapes = {k: v.shape[1:] for k, v in eps.items()}
    meta = dict(tasks_ag_id=ag_ids, tasks_g_id=g_ids, threshold=0.05, flat=False, goal_replay='her', her_replay_k=4,
                task_replay='replay_task_cp_buffer')
    for make in (lambda: ReplayBuffer(shapes, 5 * T, T, _mk_sampler(meta, rng='philox')),
                 lambda: replay_oracle.ReplayBufferOracle(shapes, 5 * T, T, None)):
        buf = make()
        with pytest.raises(AssertionError):                       # replay_buffer.py:43: sampling an empty buffer
            buf.sample(8)
        with pytest.raises(AssertionError):                       # replay_buffer.py:62: ragged episode batch
            buf.store_episode({k: (v[:2] if k != 'u' else v[:1]) for k, v in eps.items()})
        with pytest.raises(AssertionError):                       # replay_buffer.py:92: more episodes than slots
            buf.store_episode({k: v[:6] for k, v in eps.items()})
        with pytest.raises(KeyError):                             # replay_buffer.py:70: a key of buffer_shapes is missing
            buf.store_episode({k: v[:1] for k, v in eps.items() if k != 'g'})
        assert buf.get_current_size() == 0 or buf.get_current_size() == T   # (the reference reserves the slot first)
        buf.clear_buffer()
        assert buf.get_current_size() == 0 and not buf.full
    # sampler: a segment table that does not add up to the batch (ddpg.py:323) and an empty segment (replay_buffer.py:43)
    gbuf = ReplayBuffer(shapes, 5 * T, T, _mk_sampler(meta, rng='philox'))
    gbuf.store_episode({k: v[:2] for k, v in eps.items()})
    s = gbuf.sample_transitions
    with pytest.raises(AssertionError):
        s.sample_device([(gbuf.device_view(), 5, 0)], 8)
    empty = ReplayBuffer(shapes, 5 * T, T, s)
    with pytest.raises(AssertionError):
        s.sample_device([(empty.device_view(), 8, 0)], 8)
    out = gbuf.sample(8, task_to_replay=1)                        # and the well-formed call still works afterwards
    assert out['r'].shape == (8, 1) and set(shapes) <= set(out)
