"""The NumPy-only methods of the UNMODIFIED reference executed live against the oracle restatement and the product's host
arithmetic: DDPG.sample_batch (ddpg.py:251-360), DDPG.store_episode (:163-223), DDPG.get_actions post-processing (:129-161),
DDPG._preprocess_og (:118-127) and the NumPy half of Normalizer (normalizer.py:64-70,84-118).

baselines/her/ddpg.py and normalizer.py cannot be imported (TensorFlow 1.x), but these methods are plain NumPy: their
source text is cut out of the files with `ast`, compiled as it stands and run on a stand-in `self` whose buffers, TF
session and MPI communicator are recording stubs.  The only shim is `np.int` (removed from NumPy >= 1.24, used at
ddpg.py:282,286,314,318) -> `int`.  Build container only (needs /root/reference); CPU."""
import ast
import os
import textwrap
from collections import OrderedDict

import numpy as np
import pytest

from curious_b200 import apportion
from oracle.ddpg_oracle import DDPGOracle

REF = '/root/reference/baselines/her/ddpg.py'
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason='needs the reference checkout (build container)')


class _NumpyWithInt(object):
    int = int

    def __getattr__(self, name):
        return getattr(np, name)


class _FakeMPI(object):
    """ddpg.py:185 all-reduces a counter whose result is never read."""
    SUM = 'sum'

    class COMM_WORLD(object):
        @staticmethod
        def Allreduce(src, dst, op=None):
            dst[...] = src


def _reference_namespace():
    """sample_batch / _preprocess_og / store_episode / get_actions / _random_action of the reference class and
    transitions_in_episode_batch of her/util.py, compiled from their unmodified source text."""
    ns = {'np': _NumpyWithInt(), 'MPI': _FakeMPI}
    src = open(REF).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == 'DDPG'][0]
    wanted = ('sample_batch', '_preprocess_og', 'store_episode', 'get_actions', '_random_action')
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in wanted:
            exec(compile(textwrap.dedent(ast.get_source_segment(src, fn)), REF, 'exec'), ns)
    util = os.path.join(os.path.dirname(REF), 'util.py')
    usrc = open(util).read()
    for fn in ast.parse(usrc).body:
        if isinstance(fn, ast.FunctionDef) and fn.name == 'transitions_in_episode_batch':
            exec(compile(ast.get_source_segment(usrc, fn), util, 'exec'), ns)
    assert all(k in ns for k in wanted + ('transitions_in_episode_batch',))
    return ns


def _reference_methods():
    ns = _reference_namespace()
    return ns['sample_batch'], ns['_preprocess_og']


class _Buffer(object):
    """Records sample() calls and returns transitions whose values say where they came from."""

    def __init__(self, index, episodes, dims, log):
        self.index, self.current_size, self.dims, self.log = index, episodes, dims, log

    def sample(self, n, task_to_replay=None, cp_proba=None):
        n = int(n)
        self.log.append((self.index, n, task_to_replay, None if cp_proba is None else np.array(cp_proba).tolist()))
        base = 1000.0 * self.index + np.arange(n, dtype=np.float64).reshape(-1, 1)
        out = {}
        for k, d in self.dims.items():
            out[k] = base + 0.001 * np.arange(d) - (500.0 if k in ('o', 'g') else 0.0)      # some beyond +-clip_obs
        out['o_2'], out['ag_2'], out['r'] = out['o'] + 0.5, out['ag'] + 0.25, -np.ones((n, 1))
        return out


class _Self(object):
    pass


def _make(structure, task_replay, rng, log, relative_goals):
    nb = int(rng.randint(1, 7))
    dims = OrderedDict(o=int(rng.randint(1, 6)), g=3 * nb, ag=3 * nb, u=2)
    if structure != 'flat':
        dims['task_descr'] = nb
    s = _Self()
    s.structure, s.task_replay, s.nb_tasks, s.T = structure, task_replay, nb, int(rng.randint(1, 60))
    s.batch_size = int(rng.randint(1, 300))
    s.eps_task, s.t_id = float(rng.choice([0.0, 0.4, 1.0])), int(rng.randint(nb))
    s.cp = rng.rand(nb) * (rng.rand(nb) < 0.7) if rng.rand() < 0.8 else np.zeros(nb)
    s.clip_obs, s.relative_goals, s.dimg, s.dimag = 200.0, relative_goals, dims['g'], dims['ag']
    s.subtract_goals = lambda a, b: a - b
    keys = sorted(k for k in dims) + ['o_2', 'g_2', 'r']                   # ddpg.py:73-83
    s.stage_shapes = OrderedDict((k, None) for k in keys)
    s.stage_keys = keys
    s.modular = structure != 'flat'
    if structure != 'flat' and ('buffer' in task_replay):
        sizes = [0] + [int(rng.randint(0, 4)) * int(rng.rand() < 0.7) for _ in range(nb)]
        if structure == 'task_experts' and rng.rand() < 0.5:
            sizes[0] = int(rng.randint(0, 3))                               # never filled in practice; the code allows it
        if sum(sizes[1:]) == 0:
            sizes[1 + int(rng.randint(nb))] = 2
        s.buffer = [_Buffer(i, n, dims, log) for i, n in enumerate(sizes)]
    else:
        s.buffer = _Buffer(0, 3, dims, log)
    return s


CASES = [('curious', 'replay_task_cp_buffer'), ('curious', 'replay_task_random_buffer'),
         ('curious', 'replay_cp_task_transition'), ('curious', 'replay_random_task_transition'),
         ('curious', 'replay_current_task_transition'), ('task_experts', 'replay_current_task_buffer'), ('flat', '')]


@pytest.mark.parametrize('structure,task_replay', CASES)
def test_oracle_sample_batch_equals_reference_method(structure, task_replay):
    ref_sample_batch, ref_preprocess = _reference_methods()
    import zlib
    rng = np.random.RandomState(zlib.crc32((structure + task_replay).encode()))
    state = np.random.get_state()
    try:
        for case in range(50):
            seed = int(rng.randint(1 << 30))
            snapshot = rng.get_state()
            runs = []
            for which in ('reference', 'oracle'):
                rng.set_state(snapshot)
                log = []
                s = _make(structure, task_replay, rng, log, relative_goals=bool(case % 3 == 0))
                np.random.seed(seed)
                if which == 'reference':
                    s._preprocess_og = lambda o, ag, g, s=s: ref_preprocess(s, o, ag, g)
                    batch = ref_sample_batch(s)
                else:
                    batch = DDPGOracle.sample_batch(s)
                prop = None if not hasattr(s, 'proportions') else np.asarray(s.proportions).tolist()
                runs.append((batch, log, prop, np.random.get_state()[1].copy(), s))
            (rb, rlog, rprop, rstate, rs), (ob, olog, oprop, ostate, _) = runs
            assert rlog == olog, (case, rlog, olog)
            assert rprop == oprop, case
            assert np.array_equal(rstate, ostate), 'np.random consumed differently'
            assert len(rb) == len(ob) == len(rs.stage_keys)
            for key, a, b in zip(rs.stage_keys, rb, ob):
                assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b), (case, key)
            # the product's host arithmetic (curious_b200/apportion.py) against the same reference run
            if rprop is not None:
                sizes = [b.current_size for b in rs.buffer]
                if structure == 'curious':
                    mine = apportion.proportions_curious(sizes, rs.T, rs.batch_size, task_replay, rs.cp, rs.eps_task)
                else:
                    mine = apportion.proportions_task_expert(sizes, rs.T, rs.batch_size, rs.t_id)
                assert np.asarray(mine).tolist() == rprop, (case, sizes)
            elif task_replay == 'replay_cp_task_transition':
                assert np.array(apportion.cp_probabilities(rs.cp, rs.eps_task)).tolist() == rlog[0][3]
    finally:
        np.random.set_state(state)


class _RecordingBuffer(object):
    def __init__(self, index, log):
        self.index, self.log = index, log

    def store_episode(self, ep):
        self.log.append((self.index, {k: np.array(v, copy=True) for k, v in ep.items()}))


class _RecordingStats(object):
    def __init__(self, name, log):
        self.name, self.log = name, log

    def update(self, v):
        self.log.append((self.name, 'update', np.array(v, copy=True)))

    def recompute_stats(self):
        self.log.append((self.name, 'recompute', None))


@pytest.mark.parametrize('structure,task_replay', [('curious', 'replay_task_cp_buffer'), ('curious', 'replay_cp_task_transition'),
                                                   ('task_experts', 'replay_current_task_buffer'), ('flat', '')])
@pytest.mark.parametrize('nb_tasks', [3, 8])
def test_oracle_store_episode_equals_reference_method(structure, task_replay, nb_tasks):
    """DDPG.store_episode (ddpg.py:163-223) live: per-module routing by `change`, the j < 5 cap for 5+ modules, the
    duplication into every active module's buffer, and the normaliser batch drawn by the (reference) HER sampler."""
    import importlib.util
    from curious_b200 import synth
    from oracle.reward_oracle import ModuleDistanceReward
    ns = _reference_namespace()
    spec = importlib.util.spec_from_file_location('gen_golden', os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), 'oracle', 'gen_golden.py'))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    ref_her, _ = gen.import_reference()
    dims = synth.arm_dims(nb_tasks, 7)
    ag_ids, g_ids = synth.arm_task_ids(nb_tasks)
    if structure == 'flat':
        dims = {k: v for k, v in dims.items() if k != 'task_descr'}
    rng = np.random.RandomState(17 + nb_tasks)
    state = np.random.get_state()
    try:
        for case in range(12):
            eps = synth.make_episodes(rng, 2, 6, dims, still_prob=0.5, change_dtype=bool)
            if structure == 'flat':
                eps = {k: v for k, v in eps.items() if k not in ('task_descr', 'change')}
            seed = int(rng.randint(1 << 30))
            runs = []
            for which in ('reference', 'oracle'):
                log, slog = [], []
                s = _Self()
                s.structure, s.task_replay, s.nb_tasks = structure, task_replay, nb_tasks
                s.tasks_ag_id, s.tasks_g_id, s.modular = ag_ids, g_ids, structure != 'flat'
                s.clip_obs, s.relative_goals, s.dimg, s.dimag = 200.0, False, dims['g'], dims['ag']
                s.subtract_goals = lambda a, b: a - b
                reward = ModuleDistanceReward(ag_ids, g_ids, 0.05)
                if structure == 'flat':
                    s.sample_transitions = ref_her.make_sample_her_transitions('her', 4, reward, '', tasks_ag_id=ag_ids,
                                                                               tasks_g_id=g_ids)
                else:
                    s.sample_transitions = ref_her.make_sample_multi_task_her_transitions(
                        'her', 4, task_replay, reward, tasks_ag_id=ag_ids, tasks_g_id=g_ids)
                multi = 'buffer' in task_replay
                s.buffer = [_RecordingBuffer(i, log) for i in range(nb_tasks + 1)] if multi else _RecordingBuffer(0, log)
                s.o_stats, s.g_stats = _RecordingStats('o', slog), _RecordingStats('g', slog)
                batch = {k: v.copy() for k, v in eps.items()}
                np.random.seed(seed)
                if which == 'reference':
                    s._preprocess_og = lambda o, ag, g, s=s: ns['_preprocess_og'](s, o, ag, g)
                    ns['store_episode'](s, batch, np.arange(nb_tasks) / 10.0, 40 + case)
                else:
                    DDPGOracle.store_episode(s, batch, np.arange(nb_tasks) / 10.0, 40 + case)
                runs.append((log, slog, s.n_episodes, np.random.get_state()[1].copy(), sorted(batch.keys())))
            (rl, rs, rn, rstate, rkeys), (ol, os_, on, ostate, okeys) = runs
            assert [i for i, _ in rl] == [i for i, _ in ol], case                     # same buffers, same order
            for (_, a), (_, b) in zip(rl, ol):
                assert a.keys() == b.keys()
                for k in a:
                    assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), (case, k)
            if 'buffer' in task_replay:                                             # the product's routing rule as well
                want = [j + 1 for e in range(2) for j in apportion.active_modules(eps['change'][e, -1], ag_ids, g_ids)]
                assert [i for i, _ in rl] == want, case
            if nb_tasks >= 5 and 'buffer' in task_replay:
                assert all(i <= 5 for i, _ in rl)                                   # ddpg.py:183: modules 5.. are never stored
            assert [(n, what) for n, what, _ in rs] == [(n, what) for n, what, _ in os_]
            for (_, _, a), (_, _, b) in zip(rs, os_):
                assert (a is None and b is None) or (a.dtype == b.dtype and np.array_equal(a, b)), case
            assert rn == on and rkeys == okeys and np.array_equal(rstate, ostate)
    finally:
        np.random.set_state(state)


def test_oracle_get_actions_postprocessing_equals_reference_method():
    """DDPG.get_actions (ddpg.py:129-161) live with the TF session stubbed out by the oracle's own network output: the feed
    (preprocessed o, g, zero u, task_descr), Gaussian noise added in place to the float32 action, clipping, eps-greedy
    replacement, the single-row squeeze and the host RNG consumption."""
    from tests.ddpg_util import ddpg_kwargs, make_oracle_agent
    ns = _reference_namespace()
    kw, dims, ag_ids, g_ids = ddpg_kwargs(3, hidden=16, layers=2, dimo=9)
    ora = make_oracle_agent(kw, dims, ag_ids, g_ids, buffer_episodes=2)
    rng = np.random.RandomState(3)
    state = np.random.get_state()

    class Net(object):
        pi_tf, Q_pi_tf, o_tf, g_tf, u_tf, td_tf = 'pi', 'Q_pi', 'o', 'g', 'u', 'td'

    class Session(object):
        def run(self, vals, feed_dict):
            self.feed = feed_dict
            return [self.outputs[v].copy() for v in vals]
    try:
        for n in (1, 2, 17):
            for use_target in (False, True):
                for compute_Q in (False, True):
                    o = (rng.standard_normal((n, dims['o'])) * 150).astype(np.float32)
                    ag = rng.uniform(-1, 1, (n, dims['ag'])).astype(np.float32)
                    g = rng.uniform(-1, 1, (n, dims['g'])).astype(np.float32)
                    td = np.eye(3, dtype=np.float32)[rng.randint(0, 3, n)]
                    clean = ora.get_actions(o, ag, g, task_descr=td, use_target_net=use_target, compute_Q=True)
                    pi, q = np.asarray(clean[0], np.float32).reshape(n, -1), np.asarray(clean[1])
                    s = _Self()
                    s.main, s.target, s.sess = Net(), Net(), Session()
                    s.sess.outputs = {'pi': pi, 'Q_pi': q}
                    s.structure, s.dimo, s.dimg, s.dimu, s.dimtd, s.dimag = 'curious', dims['o'], dims['g'], dims['u'], 3, dims['ag']
                    s.max_u, s.clip_obs, s.relative_goals = 1.0, 200.0, False
                    s._preprocess_og = lambda o, ag, g, s=s: ns['_preprocess_og'](s, o, ag, g)
                    s._random_action = lambda k, s=s: ns['_random_action'](s, k)
                    seed = int(rng.randint(1 << 30))
                    np.random.seed(seed)
                    ref = ns['get_actions'](s, o, ag, g, task_descr=td, noise_eps=0.2, random_eps=0.3,
                                            use_target_net=use_target, compute_Q=compute_Q)
                    ref_state = np.random.get_state()[1].copy()
                    np.random.seed(seed)
                    mine = ora.get_actions(o, ag, g, task_descr=td, noise_eps=0.2, random_eps=0.3,
                                           use_target_net=use_target, compute_Q=compute_Q)
                    assert np.array_equal(np.random.get_state()[1], ref_state)
                    ref_u, mine_u = (ref[0], mine[0]) if compute_Q else (ref, mine)
                    assert ref_u.shape == mine_u.shape == ((dims['u'],) if n == 1 else (n, dims['u']))
                    assert ref_u.dtype == mine_u.dtype and np.array_equal(ref_u, mine_u)
                    if compute_Q:
                        assert np.array_equal(ref[1], mine[1])
                    feed = s.sess.feed
                    assert np.array_equal(feed['o'], np.clip(o, -200, 200)) and np.array_equal(feed['g'], g)
                    assert feed['u'].shape == (n, dims['u']) and not feed['u'].any() and np.array_equal(feed['td'], td)
    finally:
        np.random.set_state(state)


def test_oracle_normalizer_accumulation_equals_reference_methods():
    """Normalizer.update / _mpi_average / synchronize / recompute_stats (normalizer.py:64-70,84-118) live - the NumPy half
    of the class (float32 accumulators fed with float64 or float32 batches, snapshot + reset, mean over ranks of sum, sumsq,
    count in that order) with a 3-rank world emulated by the fake communicator; the TF half (running sums, mean, std) is
    restated in the oracle and checked bit for bit against the CUDA kernels elsewhere."""
    import threading
    from oracle.ddpg_oracle import NormalizerOracle
    nsrc = open(os.path.join(os.path.dirname(REF), 'normalizer.py')).read()
    cls = [n for n in ast.parse(nsrc).body if isinstance(n, ast.ClassDef) and n.name == 'Normalizer'][0]
    rng = np.random.RandomState(8)
    contributions = []

    class World3(object):
        SUM = 'sum'

        class COMM_WORLD(object):
            @staticmethod
            def Allreduce(src, dst, op=None):
                other = contributions.pop(0)
                dst[...] = src + other

            @staticmethod
            def Get_size():
                return 3
    ns = {'np': np, 'MPI': World3}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ('update', '_mpi_average', 'synchronize', 'recompute_stats'):
            exec(compile(textwrap.dedent(ast.get_source_segment(nsrc, fn)), REF, 'exec'), ns)
    size = 5
    ref = _Self()
    ref.size, ref.lock = size, threading.Lock()
    ref.local_sum, ref.local_sumsq, ref.local_count = np.zeros(size, np.float32), np.zeros(size, np.float32), np.zeros(1, np.float32)
    ref.count_pl, ref.sum_pl, ref.sumsq_pl, ref.update_op, ref.recompute_op = 'count', 'sum', 'sumsq', 'update', 'recompute'
    feeds = []

    class Session(object):
        def run(self, op, feed_dict=None):
            if feed_dict is not None:
                feeds.append({k: np.array(v, copy=True) for k, v in feed_dict.items()})
    ref.sess = Session()
    ref._mpi_average = lambda x: ns['_mpi_average'](ref, x)
    ref.synchronize = lambda **kw: ns['synchronize'](ref, **kw)
    synced = []
    others = []

    def mean_over_ranks(x):
        out = (x + others.pop(0)) / 3
        synced.append(np.array(out, copy=True))
        return out
    ora = NormalizerOracle(size, mean_over_ranks=mean_over_ranks)
    for round_ in range(6):
        for _ in range(int(rng.randint(1, 4))):
            v = rng.standard_normal((int(rng.randint(1, 40)), size)) * 3
            v = v.astype(np.float32) if rng.rand() < 0.5 else v
            ns['update'](ref, v.copy())
            ora.update(v.copy())
        assert np.array_equal(ref.local_sum, ora.local_sum) and np.array_equal(ref.local_sumsq, ora.local_sumsq)
        assert np.array_equal(ref.local_count, ora.local_count) and ref.local_sum.dtype == ora.local_sum.dtype == np.float32
        peers = [rng.standard_normal(size).astype(np.float32), np.abs(rng.standard_normal(size)).astype(np.float32),
                 np.array([float(rng.randint(1, 50))], np.float32)]
        contributions[:] = [p.copy() for p in peers]
        others[:] = [p.copy() for p in peers]
        ns['recompute_stats'](ref)
        ora.recompute_stats()
        feed = feeds[-1]
        assert np.array_equal(feed['sum'], synced[-3]) and np.array_equal(feed['sumsq'], synced[-2]), round_
        assert np.array_equal(feed['count'], synced[-1]) and feed['sum'].dtype == np.float32
        assert not ref.local_sum.any() and not ora.local_sum.any() and ref.local_count[0] == ora.local_count[0] == 0


def test_util_glue_equals_reference_functions():
    """her/util.py store_args / convert_episode_to_batch_major / transitions_in_episode_batch (plain Python, cut out of the
    unmodified file) against curious_b200/util.py."""
    import functools
    import inspect
    from curious_b200 import util as mine
    upath = os.path.join(os.path.dirname(REF), 'util.py')
    usrc = open(upath).read()
    ns = {'np': np, 'inspect': inspect, 'functools': functools}
    for fn in ast.parse(usrc).body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ('store_args', 'convert_episode_to_batch_major',
                                                           'transitions_in_episode_batch'):
            exec(compile(ast.get_source_segment(usrc, fn), upath, 'exec'), ns)

    def make(decorator):
        class K(object):
            @decorator
            def __init__(self, a, b, c=3, *, d=4, **kwargs):
                self.seen = dict(kwargs)
        return K
    for args, kwargs in (((1, 2), {}), ((1,), {'b': 5, 'd': 6}), ((1, 2, 9), {'extra': 'x', 'structure': 'curious'})):
        r, m = make(ns['store_args'])(*args, **kwargs), make(mine.store_args)(*args, **kwargs)
        assert r.__dict__ == m.__dict__
    rng = np.random.RandomState(0)
    episode = {'o': [rng.rand(2, 3) for _ in range(5)], 'u': [rng.rand(2, 1) for _ in range(4)]}
    a, b = ns['convert_episode_to_batch_major'](episode), mine.convert_episode_to_batch_major(episode)
    assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) and a[k].shape == b[k].shape for k in a)
    assert a['o'].shape == (2, 5, 3)
    assert ns['transitions_in_episode_batch'](a) == mine.transitions_in_episode_batch(b) == 8


def _reference_lp_block():
    """The learning-progress block at the end of RolloutWorker.generate_rollouts (rollout.py:316-404), cut out of the
    unmodified file and wrapped into a function of (self, successful, achieved_goals)."""
    rpath = os.path.join(os.path.dirname(REF), 'rollout.py')
    rsrc = open(rpath).read()
    cls = [n for n in ast.parse(rsrc).body if isinstance(n, ast.ClassDef) and n.name == 'RolloutWorker'][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == 'generate_rollouts'][0]
    blocks = [n for n in fn.body if isinstance(n, ast.If) and 'task_experts' in ast.get_source_segment(rsrc, n.test)
              and 'competence_computers' in ast.get_source_segment(rsrc, n)]
    assert len(blocks) == 1
    lines = rsrc.splitlines()[blocks[0].lineno - 1:blocks[0].end_lineno]
    body = textwrap.indent(textwrap.dedent('\n'.join(lines)), '    ')
    ns = {}
    exec(compile('def lp_block(self, successful, achieved_goals, MPI, np):\n' + body + '\n', rpath, 'exec'), ns)
    return ns['lp_block']


@pytest.mark.parametrize('structure,task_selection,eval_', [('curious', 'active_competence_progress', False),
                                                            ('curious', 'random', False),
                                                            ('curious', 'active_competence_progress', True),
                                                            ('task_experts', 'active_competence_progress', False)])
def test_competence_tracker_equals_reference_lp_block(structure, task_selection, eval_):
    """curious_b200.queues.CompetenceTracker.update against the reference's own LP block run live with the reference's own
    CompetenceQueue objects: C, CP and the task-selection probabilities p after every rollout."""
    import importlib.util
    from curious_b200.queues import CompetenceTracker
    spec = importlib.util.spec_from_file_location('gen_golden_queue', os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), 'oracle', 'gen_golden_queue.py'))
    gq = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gq)
    ref_queue = gq.import_reference_queue().CompetenceQueue
    lp_block = _reference_lp_block()

    class OneRank(object):
        class COMM_WORLD(object):
            gather = staticmethod(lambda x, root=0: [x])
            bcast = staticmethod(lambda x, root=0: x)
    nb, B, window = 4, 3, 6
    rng = np.random.RandomState(11)

    class Env(object):
        def __init__(self):
            self.unwrapped = self
            self.task, self.goal = 0, np.zeros(3 * nb)
    s = _Self()
    s.structure, s.task_selection, s.eval, s.goal_selection = structure, task_selection, eval_, 'random'
    s.rollout_batch_size, s.rank, s.nb_tasks, s.unique_task = B, 0, nb, 2
    s.envs = [Env() for _ in range(B)]
    s.tasks_g_id = [[3 * j, 3 * j + 1, 3 * j + 2] for j in range(nb)]
    s.competence_computers = [ref_queue(window=window) for _ in range(nb)]
    s.get_C = lambda: [q.C for q in s.competence_computers]
    s.get_CP = lambda: [q.CP for q in s.competence_computers]
    s.task_history, s.goal_history, s.tasks, s.goals = [], [], [0] * B, [None] * B
    s.p, s.CP, s.C = np.ones(nb) / nb, np.zeros(nb), np.zeros(nb)
    mine = CompetenceTracker(nb, queue_length=window, task_selection=task_selection, structure=structure,
                             unique_task=2, eval=eval_, comm=False)
    skill = np.array([0.9, 0.5, 0.1, 0.0])
    seen_cp, seen_skewed = 0.0, False
    for step in range(80):
        s.exploit = bool(rng.rand() < 0.6) or eval_
        for e in s.envs:
            e.task = int(rng.randint(nb)) if structure == 'curious' else 2
        skill = np.clip(skill + rng.uniform(-0.1, 0.15, nb), 0, 1)                  # competence drifts: CP becomes non-zero
        successful = (rng.rand(B) < skill[[e.task for e in s.envs]]).astype(np.float64)
        lp_block(s, successful, [np.zeros((B, 3 * nb))], OneRank, np)
        tasks, succ = ([e.task for e in s.envs], successful.tolist()) if s.exploit else ([], [])
        CP, p = mine.update(tasks, succ)
        assert np.array_equal(np.asarray(mine.C, np.float64), np.asarray(s.C, np.float64)), step
        if not eval_:
            assert np.array_equal(np.asarray(CP, np.float64), np.asarray(s.CP, np.float64)), step
            assert np.array_equal(np.asarray(p, np.float64), np.asarray(s.p, np.float64)), step
            assert abs(np.sum(p) - 1.0) < 1e-12
            seen_cp = max(seen_cp, float(np.max(s.CP)))
            seen_skewed = seen_skewed or not np.allclose(s.p, 1.0 / nb)
    if structure == 'curious' and task_selection == 'active_competence_progress' and not eval_:
        assert seen_cp > 0 and seen_skewed                                          # the interesting branch was exercised


def test_expert_selection_equals_reference_block():
    """The expert-selection statement of the reference's train loop (experiment/train.py:79-104) cut out and run live
    against curious_b200.train.train: same experts drawn from the same np.random stream, epoch after epoch.  (The
    reference computes a CP-weighted `proba` but draws from `p`, which stays uniform - mirrored as it behaves.)"""
    import warnings
    from tests.test_train_loop_cpu import _workers
    from curious_b200.train import train
    tpath = os.path.join(os.path.dirname(REF), 'experiment', 'train.py')
    tsrc = open(tpath).read()
    fn = [n for n in ast.parse(tsrc).body if isinstance(n, ast.FunctionDef) and n.name == 'train'][0]
    branch = [n for n in fn.body if isinstance(n, ast.If) and 'task_experts' in ast.get_source_segment(tsrc, n.test)][0]
    loop = [n for n in branch.body if isinstance(n, ast.For) and getattr(n.target, 'id', '') == 'epoch'][0]
    select = [n for n in loop.body if isinstance(n, ast.If)][0]
    assert isinstance(select, ast.If) and 'task_selection' in ast.get_source_segment(tsrc, select.test)
    lines = tsrc.splitlines()[select.lineno - 1:select.end_lineno]
    body = textwrap.indent(textwrap.dedent('\n'.join(lines)), '    ')
    ns = {}
    exec(compile('def select(task_selection, epoch, nb_tasks, rank, rollout_worker, params, p, np, MPI):\n' + body +
                 '\n    return i_policy, p\n', tpath, 'exec'), ns)

    class OneRank(object):
        class COMM_WORLD(object):
            bcast = staticmethod(lambda x, root=0: x)
    state = np.random.get_state()
    try:
        for task_selection in ('active_competence_progress', 'random'):
            np.random.seed(4)
            policy, rollout, evaluator, _ = _workers('task_experts')
            for i, cp in enumerate([0.05, 0.4, 0.0]):
                rollout[i].tracker.competence_computers[i].CP = cp
            np.random.seed(21)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                hist = train(policy, rollout, evaluator, n_epochs=15, n_test_rollouts=0, n_cycles=0, n_batches=0,
                             structure='task_experts', task_selection=task_selection)
            np.random.seed(21)
            p = 1 / 3 * np.ones([3])
            for epoch in range(15):
                i_policy, p = ns['select'](task_selection, epoch, 3, 0, rollout, {'eps_task': 0.4}, p, np, OneRank)
                assert int(i_policy) == hist[epoch]['i_policy'], (task_selection, epoch)
                assert np.array_equal(p, hist[epoch]['p'])
            if task_selection == 'active_competence_progress':
                assert len(set(h['i_policy'] for h in hist)) == 3 and np.allclose(hist[-1]['p'], 1 / 3)
                assert np.allclose(hist[-1]['proba'], 0.4 / 3 + 0.6 * np.array([0.05, 0.4, 0.0]) / 0.45)
    finally:
        np.random.set_state(state)


def _reference_rollout_worker():
    """The reference RolloutWorker class (rollout.py:13-491) compiled from its unmodified source; its module-level imports
    (mpi4py, mujoco_py, baselines.*) are replaced by a one-rank communicator, the reference's own util / queue functions and
    placeholders for what `goal_selection='random'` never touches."""
    import functools
    import importlib.util
    import inspect
    import pickle
    from collections import deque
    rpath = os.path.join(os.path.dirname(REF), 'rollout.py')
    rsrc = open(rpath).read()
    cls = [n for n in ast.parse(rsrc).body if isinstance(n, ast.ClassDef) and n.name == 'RolloutWorker'][0]
    upath = os.path.join(os.path.dirname(REF), 'util.py')
    usrc = open(upath).read()
    ns = {'np': np, 'inspect': inspect, 'functools': functools, 'deque': deque, 'pickle': pickle}
    for fn in ast.parse(usrc).body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ('store_args', 'convert_episode_to_batch_major'):
            exec(compile(ast.get_source_segment(usrc, fn), upath, 'exec'), ns)
    spec = importlib.util.spec_from_file_location('gen_golden_queue', os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), 'oracle', 'gen_golden_queue.py'))
    gq = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gq)

    class OneRank(object):
        class COMM_WORLD(object):
            Get_rank = staticmethod(lambda: 0)
            Get_size = staticmethod(lambda: 1)
            gather = staticmethod(lambda x, root=0: [x])
            bcast = staticmethod(lambda x, root=0: x)
            scatter = staticmethod(lambda x, root=0: x[0])
    ns.update(MPI=OneRank, MujocoException=KeyboardInterrupt, CompetenceQueue=gq.import_reference_queue().CompetenceQueue,
              logger=None, SAGG_RIAC=None)
    exec(compile(ast.get_source_segment(rsrc, cls), rpath, 'exec'), ns)
    return ns['RolloutWorker']


class _SeededPolicy(object):
    """get_actions stand-in: float32 actions from the host stream (like DDPG.get_actions, which returns float32)."""

    def __init__(self, dimu):
        self.dimu = dimu

    def get_actions(self, o, ag, g, task_descr=None, compute_Q=False, noise_eps=0., random_eps=0., use_target_net=False):
        n = len(np.atleast_2d(o))
        u = np.random.uniform(-1, 1, (n, self.dimu)).astype(np.float32)
        u += noise_eps * np.random.randn(n, self.dimu)
        if n == 1:
            u = u[0]
        return [u, np.full((n, 1), float(np.sum(o)), np.float32)] if compute_Q else u


@pytest.mark.parametrize('structure,eval_', [('curious', False), ('curious', True), ('flat', False), ('flat', True),
                                             ('task_experts', False), ('task_experts', True)])
def test_rollout_worker_equals_reference_class(structure, eval_):
    """curious_b200.rollout.RolloutWorker against the reference class run live on the same environments, policy stand-in
    and np.random stream: episodes (values, shapes), CP, episode counter, task / goal assignment, competence, probabilities
    and the logged statistics after every generate_rollouts call."""
    from curious_b200.envs import ModularPointEnv
    from curious_b200.rollout import RolloutWorker
    from curious_b200.train import configure_dims
    Ref = _reference_rollout_worker()
    nb, T, B = 3, 7, 2

    def make_env():
        return ModularPointEnv(nb, n_controllable=2, max_episode_steps=T)
    dims = configure_dims(make_env(), structure)
    state = np.random.get_state()
    try:
        workers = []
        for cls in (Ref, RolloutWorker):
            np.random.seed(12)
            if structure == 'task_experts' and eval_:
                policy = [_SeededPolicy(dims['u']) for _ in range(nb)]
            else:
                policy = _SeededPolicy(dims['u'])
            kw = dict(exploit=eval_, compute_Q=eval_, noise_eps=0.2, random_eps=0.3, structure=structure,
                      task_selection='active_competence_progress', queue_length=4, eval=eval_, history_len=5,
                      unique_task=None if eval_ or structure != 'task_experts' else 1)
            w = cls(make_env, policy, dims, None, T, rollout_batch_size=B, **kw)
            w.seed(3)
            workers.append(w)
        ref, mine = workers
        for call in range(25):
            outs = []
            for w in (ref, mine):
                np.random.seed(1000 + call)
                outs.append(w.generate_rollouts() + (np.random.get_state()[1].copy(),))
            (re, rcp, rn, rstate), (me, mcp, mn, mstate) = outs
            assert np.array_equal(rstate, mstate), 'np.random consumed differently'
            assert rn == mn and np.array_equal(np.asarray(rcp, np.float64), np.asarray(mcp, np.float64)), call
            assert sorted(re.keys()) == sorted(me.keys()), call
            for k in re:
                assert np.asarray(re[k]).shape == np.asarray(me[k]).shape, (call, k)
                assert np.array_equal(np.asarray(re[k], np.float64), np.asarray(me[k], np.float64)), (call, k)
            assert ref.exploit == mine.exploit
            if structure != 'flat':
                assert np.array_equal(np.asarray(ref.p, np.float64), np.asarray(mine.p, np.float64)), call
                assert np.array_equal(np.asarray(ref.get_C(), np.float64), np.asarray(mine.get_C(), np.float64))
                assert np.array_equal(np.asarray(ref.get_CP(), np.float64), np.asarray(mine.get_CP(), np.float64))
                assert [int(t) for t in ref.tasks] == [int(t) for t in mine.tasks]
                assert all(np.array_equal(a, b) for a, b in zip(ref.goals, mine.goals))
                assert [int(t) for t in ref.task_history] == [int(t) for t in mine.task_history]
            assert dict(ref.logs('x')) == dict(mine.logs('x')) and ref.additional_logs('x') == mine.additional_logs('x')
            assert ref.current_success_rate() == mine.current_success_rate()
            if eval_:
                assert ref.current_mean_Q() == mine.current_mean_Q()
    finally:
        np.random.set_state(state)


def _reference_train_functions(rows, infos, logdir):
    """train() and logs() of experiment/train.py plus mpi_average / mpi_moments / mpi_mean, compiled from their unmodified
    source; the logger is a recorder, the communicator has one rank."""
    import time

    class OneRank(object):
        SUM = 'sum'

        class COMM_WORLD(object):
            Get_rank = staticmethod(lambda: 0)
            Get_size = staticmethod(lambda: 1)
            bcast = staticmethod(lambda x, root=0: x)

            @staticmethod
            def Bcast(buf, root=0):
                pass

            @staticmethod
            def Allreduce(src, dst, op=None):
                dst[...] = src

    class Logger(object):
        row = {}
        get_dir = staticmethod(lambda: logdir)
        info = staticmethod(lambda *a: infos.append(' '.join(str(x) for x in a)))

        @classmethod
        def record_tabular(cls, k, v):
            cls.row[k] = v

        @classmethod
        def dump_tabular(cls):
            rows.append(dict(cls.row))
            cls.row = {}
    ns = {'np': np, 'os': os, 'time': time, 'MPI': OneRank, 'logger': Logger, 't0': time.time()}
    here = os.path.dirname(REF)
    for path, names in ((os.path.join(here, '..', 'common', 'mpi_moments.py'), ('mpi_mean', 'mpi_moments')),
                        (os.path.join(here, 'util.py'), ('mpi_average',)),
                        (os.path.join(here, 'experiment', 'train.py'), ('train', 'logs'))):
        src = open(path).read()
        for fn in ast.parse(src).body:
            if isinstance(fn, ast.FunctionDef) and fn.name in names:
                exec(compile(ast.get_source_segment(src, fn), path, 'exec'), ns)
    return ns['train']


@pytest.mark.parametrize('structure', ['curious', 'flat', 'task_experts'])
def test_train_loop_and_records_equal_reference_functions(structure, tmp_path):
    """curious_b200.train.train against the reference's train() + logs() run live (experiment/train.py:48-215) with the same
    workers, policy stand-ins and np.random stream: the order of rollouts / store / updates / evaluations (through the RNG
    stream and the call counters), every tabular row key by key (except Time), and the policy files written."""
    from tests.test_train_loop_cpu import _workers
    from curious_b200.train import train
    n_epochs, kw = 4, dict(n_test_rollouts=2, n_cycles=3, n_batches=2)
    results = []
    state = np.random.get_state()
    try:
        for which in ('reference', 'mine'):
            logdir = str(tmp_path / which)
            os.makedirs(logdir)
            np.random.seed(6)
            policy, rollout, evaluator, _ = _workers(structure)
            for i, w in enumerate((rollout if isinstance(rollout, list) else [rollout]) + [evaluator]):
                w.seed(70 + i)
            np.random.seed(31)
            rows, infos = [], []
            if which == 'reference':
                # NumPy >= 2 shim at the boundary, the reference code itself stays untouched: its mpi_average starts with
                # `if value == []`, which raises for an np.float64 operand today (it was False with a warning in NumPy 1),
                # so the workers hand it plain Python floats
                for w in (rollout if isinstance(rollout, list) else [rollout]) + [evaluator]:
                    w.logs = (lambda prefix, w=w, f=w.logs: [(k, float(v) if isinstance(v, np.floating) else v)
                                                             for k, v in f(prefix)])
                    w.current_success_rate = (lambda w=w, f=w.current_success_rate: float(f()))
                ref_train = _reference_train_functions(rows, infos, logdir)
                ref_train(policy, rollout, evaluator, n_epochs, policy_save_interval=2, save_policies=True, structure=structure,
                          task_selection='active_competence_progress', params={'nb_tasks': 3, 'eps_task': 0.4},
                          perturbation_study=False, **kw)
            else:
                train(policy, rollout, evaluator, n_epochs, structure=structure, logdir=logdir, policy_save_interval=2, **kw)
                lines = open(os.path.join(logdir, 'progress.csv')).read().splitlines()
                header = lines[0].split(',')
                rows = [{k: v for k, v in zip(header, line.split(',')) if v != ''} for line in lines[1:]]
            pols = policy if isinstance(policy, list) else [policy]
            results.append(dict(rows=rows, counters=[(p.trained, p.target_updates, len(p.stored)) for p in pols],
                                rng=np.random.get_state()[1].copy(),
                                files=sorted(f for f in os.listdir(logdir) if f.startswith('policy_'))))
        ref, mine = results
        assert ref['counters'] == mine['counters'] and np.array_equal(ref['rng'], mine['rng'])
        assert ref['files'] == mine['files'] and 'policy_best.pkl' in mine['files'] and 'policy_2.pkl' in mine['files']
        assert len(ref['rows']) == len(mine['rows']) == n_epochs + 1                      # epoch -1 .. 3
        for r, m in zip(ref['rows'], mine['rows']):
            assert [k for k in r if k != 'Time'] == [k for k in m if k != 'Time'], (r, m)
            for k in r:
                if k != 'Time':
                    assert str(r[k]) == m[k], (k, r[k], m[k])
    finally:
        np.random.set_state(state)


@pytest.mark.parametrize('scale_grad_by_procs', [False, True])
def test_adam_oracle_equals_reference_update(scale_grad_by_procs):
    """MpiAdam.update (common/mpi_adam.py:21-35) live against oracle.ddpg_oracle.MpiAdamOracle over 150 steps of a 3-rank
    world.  The reference ran on NumPy 1.x, whose value-based casting kept `(-a) * self.m` in float32 although `a` is an
    np.float64; NumPy >= 2 would promote that product to float64.  The one shim reproduces the old rule at the boundary:
    np.sqrt of a Python float hands back a Python float (a weak scalar today), everything else is NumPy as is."""
    from oracle.ddpg_oracle import MpiAdamOracle
    apath = os.path.join(os.path.dirname(REF), '..', 'common', 'mpi_adam.py')
    asrc = open(apath).read()
    cls = [n for n in ast.parse(asrc).body if isinstance(n, ast.ClassDef) and n.name == 'MpiAdam'][0]
    update_src = [ast.get_source_segment(asrc, fn) for fn in cls.body if isinstance(fn, ast.FunctionDef) and fn.name == 'update'][0]

    class NumPy1Scalars(object):
        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def sqrt(x):
            return float(np.sqrt(x)) if isinstance(x, float) else np.sqrt(x)
    peers = []

    class Comm(object):
        @staticmethod
        def Allreduce(src, dst, op=None):
            dst[...] = src + peers[0] + peers[1]

        @staticmethod
        def Get_size():
            return 3

    class MPI(object):
        SUM = 'sum'
    ns = {'np': NumPy1Scalars(), 'MPI': MPI}
    exec(compile(textwrap.dedent(update_src), apath, 'exec'), ns)
    rng = np.random.RandomState(2)
    n = 4097
    theta0 = rng.standard_normal(n).astype(np.float32)
    ref = _Self()
    ref.beta1, ref.beta2, ref.epsilon, ref.scale_grad_by_procs, ref.comm = 0.9, 0.999, 1e-08, scale_grad_by_procs, Comm
    ref.m, ref.v, ref.t = np.zeros(n, 'float32'), np.zeros(n, 'float32'), 0
    ref.theta = theta0.copy()
    ref.check_synced = lambda: None
    ref.getflat = lambda: ref.theta.copy()

    def setfromflat(x):
        assert x.dtype == np.float32, 'the NumPy-1 rule keeps the step in float32'
        ref.theta = np.asarray(x, np.float32)
    ref.setfromflat = setfromflat
    ora = MpiAdamOracle(theta0, scale_grad_by_procs=scale_grad_by_procs, world_size=3,
                        allreduce_sum=lambda g: g + peers[0] + peers[1])
    for step in range(150):
        grads = [(rng.standard_normal(n) * 10 ** rng.uniform(-6, 1)).astype(np.float32) for _ in range(3)]
        if step % 17 == 0:
            grads[0][::5] = 0.0                                  # exact zeros: sqrt(v) + eps carries the step
        peers[:] = grads[1:]
        ns['update'](ref, grads[0].astype(np.float64) if step % 2 else grads[0], 1e-3)
        ora.update(grads[0], 1e-3)
        assert ref.t == ora.t
        assert np.array_equal(ref.m, ora.m) and np.array_equal(ref.v, ora.v), step
        assert np.array_equal(ref.theta, ora.theta), step
    assert not np.array_equal(ref.theta, theta0)


def test_default_hyperparameters_equal_reference_config():
    """experiment/config.py:18-87 DEFAULT_PARAMS / MULTI_TASK_PARAMS (evaluated from the unmodified assignments) against
    the dictionaries curious_b200.train.make_experiment starts from."""
    from curious_b200 import train as mine
    cpath = os.path.join(os.path.dirname(REF), 'experiment', 'config.py')
    csrc = open(cpath).read()
    ns = {}
    for node in ast.parse(csrc).body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], 'id', '') in ('DEFAULT_PARAMS', 'MULTI_TASK_PARAMS'):
            exec(compile(ast.get_source_segment(csrc, node), cpath, 'exec'), ns)
    assert mine.MULTI_TASK_PARAMS == ns['MULTI_TASK_PARAMS']
    flat = dict(mine.FLAT_PARAMS)
    assert flat.pop('eps_task') == 0.4                     # not in the reference's flat defaults; unused by the flat agent
    assert flat == ns['DEFAULT_PARAMS']
