"""The oracle restatement vs golden vectors recorded from the UNMODIFIED reference
(baselines/her/her.py, replay_buffer.py) - bit exact.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import her_oracle, replay_oracle
from oracle.reward_oracle import ModuleDistanceReward
from tests.golden_util import GOLDEN, load_case, per_row_choices, sampler_cases


def build_oracle_sampler(meta):
    reward = ModuleDistanceReward(meta['tasks_ag_id'], meta['tasks_g_id'], meta['threshold'])
    if meta['flat']:
        s = her_oracle.make_sample_her_transitions(meta['goal_replay'], meta['her_replay_k'], reward,
                                                   meta['task_replay'], tasks_ag_id=meta['tasks_ag_id'],
                                                   tasks_g_id=meta['tasks_g_id'])
    else:
        s = her_oracle.make_sample_multi_task_her_transitions(
            meta['goal_replay'], meta['her_replay_k'], meta['task_replay'], reward,
            tasks_ag_id=meta['tasks_ag_id'], tasks_g_id=meta['tasks_g_id'])
    return s, reward


def run_oracle(meta, eps, stream=None):
    sampler, reward = build_oracle_sampler(meta)
    T = meta['T']
    kw = {}
    if not meta['flat']:
        kw = dict(task_to_replay=meta['task_to_replay'], cp_proba=meta['cp_proba'])
    if stream is not None:
        kw['stream'] = stream
    if meta['via_buffer']:
        shapes = {k: v.shape[1:] for k, v in eps.items()}
        buf = replay_oracle.ReplayBufferOracle(shapes, (meta['E'] + 3) * T, T, sampler)
        buf.store_episode({k: v.copy() for k, v in eps.items()})
        if stream is None:
            np.random.seed(meta['seed'])
            # the reference seeded before store_episode; storing E episodes in order draws nothing
        out = buf.sample(meta['B'], **kw)
    else:
        batch = {k: v.copy() for k, v in eps.items()}
        batch['o_2'] = batch['o'][:, 1:, :]
        batch['ag_2'] = batch['ag'][:, 1:, :]
        if stream is None:
            np.random.seed(meta['seed'])
        out = sampler(batch, meta['B'], **kw)
    return out, sampler, reward


@pytest.mark.parametrize('name', sampler_cases())
def test_oracle_matches_reference_with_own_draws(name):
    """Seeded like the reference run: same RNG consumption order, same outputs."""
    meta, eps, stream, ref = load_case(name)
    out, sampler, reward = run_oracle(meta, eps)
    s = sampler.last_stream
    assert np.array_equal(s.ep, stream['s_ep'])
    assert np.array_equal(s.t, stream['s_t'])
    assert np.array_equal(s.u_her, stream['s_uher'])
    assert np.array_equal(s.u_off, stream['s_uoff'])
    assert set(out.keys()) == set(ref.keys())
    for k in ref:
        assert out[k].shape == ref[k].shape, k
        assert out[k].dtype == ref[k].dtype, k
        assert np.array_equal(out[k], ref[k]), k
    assert reward.n_calls == meta['reward_calls']
    assert (reward.last_kwargs['task_descr'] is None) == meta['reward_td_none']


@pytest.mark.parametrize('name', sampler_cases())
def test_oracle_matches_reference_with_injected_stream(name):
    meta, eps, stream, ref = load_case(name)
    inj = her_oracle.HerStream(stream['s_ep'], stream['s_t'], stream['s_uher'], stream['s_uoff'],
                               per_row_choices(meta, stream))
    state = np.random.get_state()[1].copy()
    out, _, _ = run_oracle(meta, eps, stream=inj)
    assert np.array_equal(np.random.get_state()[1], state), "injected mode must not touch np.random"
    for k in ref:
        assert np.array_equal(out[k], ref[k]), k


@pytest.mark.parametrize('name', ['storage_single', 'storage_batched'])
def test_storage_index_policy(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    meta = json.loads(str(z['meta']))
    T = meta['T']
    buf = replay_oracle.ReplayBufferOracle({'o': (T + 1, 2), 'u': (T, 1)}, meta['size_ep'] * T, T, None)
    np.random.seed(meta['seed'])
    for k, inc in enumerate(meta['increments']):
        ep = {'o': np.full((inc, T + 1, 2), float(k)), 'u': np.full((inc, T, 1), float(k))}
        idx = buf.store_episode(ep)
        assert np.array_equal(np.atleast_1d(idx), z['idx_%d' % k]), k
        assert buf.get_current_episode_size() == z['sizes'][k]
        assert buf.get_transitions_stored() == z['stored'][k]
    assert np.array_equal(buf.buffers['u'][:buf.current_size, 0, 0], z['final_u'])


def test_invariants_on_golden():
    """Hand-derived invariants: future_t in [t+1, T], td one-hot, g zero off-module."""
    for name in sampler_cases():
        meta, eps, stream, ref = load_case(name)
        T = meta['T']
        off = (stream['s_uoff'] * (T - stream['s_t'])).astype(int)
        ft = stream['s_t'] + 1 + off
        assert (ft >= stream['s_t'] + 1).all() and (ft <= T).all()
        assert ref['r'].shape == (meta['B'], 1)
        assert set(np.unique(ref['r'])) <= {-1.0, 0.0}
        if not meta['flat']:
            assert np.array_equal(ref['task_descr'].sum(1), np.ones(meta['B']))


@pytest.mark.skipif(not os.path.isdir('/root/reference/baselines/her'), reason='needs the reference checkout (build container)')
def test_committed_fixtures_are_what_the_unmodified_reference_produces(tmp_path):
    """Re-runs oracle/gen_golden.py (which imports her.py / replay_buffer.py from /root/reference unmodified) into a scratch
    directory and compares every array of every fixture with the committed tests/golden/*.npz: the fixtures the GPU box
    tests against are the reference's outputs, not something edited by hand."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('gen_golden', os.path.join(os.path.dirname(GOLDEN), '..', 'oracle',
                                                                             'gen_golden.py'))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    state = np.random.get_state()
    try:
        gen.main(str(tmp_path))
    finally:
        np.random.set_state(state)
    manifest = json.load(open(os.path.join(GOLDEN, 'MANIFEST.json')))
    assert len(manifest['cases']) >= 15
    for name in manifest['cases']:
        new, old = np.load(str(tmp_path / (name + '.npz')), allow_pickle=True), np.load(os.path.join(GOLDEN, name + '.npz'),
                                                                                       allow_pickle=True)
        assert sorted(new.files) == sorted(old.files), name
        for key in old.files:
            a, b = new[key], old[key]
            assert a.dtype == b.dtype and a.shape == b.shape, (name, key)
            assert np.array_equal(a, b) or (a.dtype.kind == 'f' and np.array_equal(a, b, equal_nan=True)), (name, key)


@pytest.mark.skipif(not os.path.isdir('/root/reference/baselines/her'), reason='needs the reference checkout (build container)')
def test_queue_and_logger_fixtures_regenerate_identically(tmp_path):
    """Same for tests/golden/competence_queue.npz (reference queues.py) and progress_golden.csv (reference logger.py)."""
    import importlib.util
    oracle_dir = os.path.join(os.path.dirname(GOLDEN), '..', 'oracle')
    for script in ('gen_golden_queue.py', 'gen_golden_progress.py'):
        spec = importlib.util.spec_from_file_location(script[:-3], os.path.join(oracle_dir, script))
        gen = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(gen)
        gen.main(str(tmp_path))
    new, old = np.load(str(tmp_path / 'competence_queue.npz')), np.load(os.path.join(GOLDEN, 'competence_queue.npz'))
    assert sorted(new.files) == sorted(old.files)
    for key in old.files:
        assert np.array_equal(new[key], old[key]), key
    for name in ('progress_golden.csv', 'progress_rows.json'):
        assert open(str(tmp_path / name)).read() == open(os.path.join(GOLDEN, name)).read(), name


@pytest.mark.skipif(not os.path.isdir('/root/reference/baselines/her'), reason='needs the reference checkout (build container)')
def test_oracle_equals_the_live_reference_on_random_scenarios(tmp_path):
    """Beyond the 15 committed fixtures: 60 randomly drawn scenarios (sizes, modules, every task_replay mode, with and
    without HER, flat, longer ag slices, buffer / direct calls) run through the UNMODIFIED reference here and through the
    oracle with the same seed - draws, outputs, dtypes and reward-call contract must agree bit for bit."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('gen_golden', os.path.join(os.path.dirname(GOLDEN), '..', 'oracle',
                                                                             'gen_golden.py'))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    ref_her, ref_rb = gen.import_reference()
    rng = np.random.RandomState(20241017)
    state = np.random.get_state()
    modes = ['replay_task_cp_buffer', 'replay_task_random_buffer', 'replay_random_task_transition',
             'replay_cp_task_transition', 'replay_current_task_transition', 'hand_designed']
    try:
        for i in range(60):
            flat = bool(rng.rand() < 0.2)
            n_modules = int(rng.randint(1, 7))
            mode = '' if flat else modes[int(rng.randint(len(modes)))]
            kw = dict(n_modules=n_modules, dimo=int(rng.randint(1, 12)), E=int(rng.randint(1, 9)), T=int(rng.randint(1, 14)),
                      B=int(rng.randint(1, 130)), seed=int(rng.randint(1 << 30)), data_seed=int(rng.randint(1 << 30)),
                      goal_replay='her' if rng.rand() < 0.8 else 'none', task_replay=mode, flat=flat,
                      longer_ag=bool(rng.rand() < 0.3) and not flat, via_buffer=bool(rng.rand() < 0.7))
            if not flat:
                if ('buffer' in mode or mode == 'hand_designed') and rng.rand() < 0.7:
                    kw['task_to_replay'] = int(rng.randint(n_modules))
                if mode == 'replay_cp_task_transition':
                    p = rng.rand(n_modules) + 0.05
                    kw['cp_proba'] = list(p / p.sum())
            name = 'rand_%02d' % i
            gen.run_sampler_case(name, ref_her, ref_rb, outdir=str(tmp_path), **kw)
            meta, eps, stream, ref = load_case(name, str(tmp_path))
            out, sampler, reward = run_oracle(meta, eps)
            s = sampler.last_stream
            assert np.array_equal(s.ep, stream['s_ep']) and np.array_equal(s.t, stream['s_t']), (i, kw)
            assert np.array_equal(s.u_her, stream['s_uher']) and np.array_equal(s.u_off, stream['s_uoff']), (i, kw)
            assert set(out.keys()) == set(ref.keys()), (i, kw)
            for k in ref:
                assert out[k].shape == ref[k].shape and out[k].dtype == ref[k].dtype, (i, k, kw)
                assert np.array_equal(out[k], ref[k]), (i, k, kw)
            assert reward.n_calls == meta['reward_calls'], (i, kw)
    finally:
        np.random.set_state(state)


@pytest.mark.skipif(not os.path.isdir('/root/reference/baselines/her'), reason='needs the reference checkout (build container)')
def test_storage_oracle_equals_the_live_reference_on_random_sequences():
    """ReplayBuffer slot policy and counters (replay_buffer.py:57-109) on 40 random capacity / batch-size sequences,
    reference and oracle seeded alike: same slots, same sizes, same stored data, same failure on an oversize batch."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('gen_golden', os.path.join(os.path.dirname(GOLDEN), '..', 'oracle',
                                                                             'gen_golden.py'))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    _, ref_rb = gen.import_reference()
    rng = np.random.RandomState(99)
    state = np.random.get_state()
    try:
        for case in range(40):
            T, size_ep = int(rng.randint(1, 6)), int(rng.randint(1, 12))
            shapes = {'o': (T + 1, 2), 'u': (T, 1)}
            bufs = [ref_rb.ReplayBuffer(shapes, size_ep * T, T, None), replay_oracle.ReplayBufferOracle(shapes, size_ep * T, T, None)]
            seed = int(rng.randint(1 << 30))
            incs = [int(rng.randint(1, size_ep + 2)) for _ in range(int(rng.randint(3, 25)))]
            results = []
            for buf in bufs:
                np.random.seed(seed)
                log = []
                for k, inc in enumerate(incs):
                    ep = {'o': np.full((inc, T + 1, 2), float(k)), 'u': np.full((inc, T, 1), float(k))}
                    try:
                        buf.store_episode(ep)
                        log.append(('ok', buf.get_current_episode_size(), buf.get_current_size(), buf.get_transitions_stored(),
                                    bool(buf.full)))
                    except AssertionError:
                        log.append(('too large',))
                n = buf.get_current_episode_size()
                results.append((log, buf.buffers['u'][:n, 0, 0].copy(), buf.buffers['o'][:n, -1, 1].copy(),
                                np.random.get_state()[1].copy()))
            (la, ua, oa, sa), (lb, ub, ob, sb) = results
            assert la == lb, (case, incs)
            assert np.array_equal(ua, ub) and np.array_equal(oa, ob), (case, incs)
            assert np.array_equal(sa, sb), 'np.random consumed differently'
            assert any(inc > size_ep for inc in incs) == any(x == ('too large',) for x in la)
    finally:
        np.random.set_state(state)


# Random123 (D. E. Shaw Research) known-answer vectors for philox4x32, 10 rounds: (counter, key) -> output
PHILOX_KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_oracle_known_answers():
    """The counter-based generator of the Philox draw mode is Philox4x32-10 as published: the oracle restatement
    reproduces the three Random123 known-answer vectors (the kernel's device function is checked against the same
    vectors in tests/test_her_gpu.py::test_philox_device_known_answers)."""
    from oracle import philox_oracle
    for ctr, key, want in PHILOX_KAT:
        got = philox_oracle.philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(x[0]) for x in got) == want
