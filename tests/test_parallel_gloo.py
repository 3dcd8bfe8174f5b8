"""world_size-2 `gloo` tests (CPU) of the data-parallel plumbing in curious_b200/parallel.py.

The reference runs one MPI rank per worker (train.py:221-243); its only collectives on the hot path are
  MpiAdam.update  Allreduce(SUM) of the flat gradient, no division        mpi_adam.py:24-28, ddpg.py:452-453
  MpiAdam.sync / check_synced  Bcast from rank 0 (+ assert)               mpi_adam.py:37-50
  Normalizer._mpi_average  Allreduce(SUM) / size                          normalizer.py:84-88
The same helper functions run on NCCL tensors on the box; here they run on CPU tensors and are checked
against the oracle emulating the world in-process.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, out_dir):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        res = globals()[fn_name](rank, world)
        np.save(os.path.join(out_dir, 'r%d.npy' % rank), np.asarray(res, dtype=np.float64))
    finally:
        dist.destroy_process_group()


def _run(fn_name, tmp_path, world=2):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, fn_name, str(tmp_path)), nprocs=world, join=True)
    return [np.load(os.path.join(str(tmp_path), 'r%d.npy' % r)) for r in range(world)]


# ---------------------------------------------------------------------------------------------------
def _local_grad(rank, n=1003):
    return np.random.RandomState(100 + rank).randn(n).astype(np.float32)


def _body_grad_sum_adam(rank, world):
    """all-reduce SUM of the local flat gradients, then the oracle Adam step on every rank."""
    from curious_b200 import parallel
    from oracle.ddpg_oracle import MpiAdamOracle
    g = torch.from_numpy(_local_grad(rank))
    n = parallel.allreduce_sum_(g)
    assert n == world
    theta = np.linspace(-1, 1, g.numel()).astype(np.float32)
    adam = MpiAdamOracle(theta, scale_grad_by_procs=False)
    for _ in range(3):
        adam.update(g.numpy(), 1e-3)
    fp = torch.from_numpy(adam.theta.copy())
    parallel.assert_synced(fp)                     # every rank holds bit-identical parameters
    return adam.theta


def test_grad_allreduce_is_sum_not_mean(tmp_path):
    res = _run('_body_grad_sum_adam', tmp_path)
    from oracle.ddpg_oracle import MpiAdamOracle
    # oracle world emulated in-process: allreduce_sum callable (mpi_adam.py:24-26)
    total = _local_grad(0) + _local_grad(1)
    theta = np.linspace(-1, 1, total.size).astype(np.float32)
    adam = MpiAdamOracle(theta, scale_grad_by_procs=False, allreduce_sum=lambda x: total, world_size=2)
    for _ in range(3):
        adam.update(_local_grad(0), 1e-3)
    for r in res:
        assert np.array_equal(r.astype(np.float32), adam.theta)


def _body_bcast(rank, world):
    from curious_b200 import parallel
    t = torch.full((17,), float(rank + 1))
    parallel.broadcast_from_root_(t)
    return t.numpy()


def test_sync_broadcasts_rank0(tmp_path):
    res = _run('_body_bcast', tmp_path)
    for r in res:
        assert np.array_equal(r, np.ones(17))


def _body_desync(rank, world):
    from curious_b200 import parallel
    fp = torch.tensor([1.0 + rank])
    try:
        parallel.assert_synced(fp)
    except AssertionError:
        return [1.0]
    return [0.0]


def test_check_synced_detects_divergence(tmp_path):
    res = _run('_body_desync', tmp_path)
    assert res[0][0] == 0.0 and res[1][0] == 1.0      # rank 0 is the reference copy; rank 1 diverged


def _norm_input(rank, dim=7):
    return np.random.RandomState(7 + rank).randn(50 + 10 * rank, dim).astype(np.float32)


def _body_norm(rank, world):
    """Packed partial [sum | sumsq | count] -> SUM over ranks -> / world (normalizer.py:84-94)."""
    from curious_b200 import parallel
    v = _norm_input(rank)
    packed = torch.from_numpy(np.concatenate([v.sum(0), np.square(v).sum(0), [np.float32(v.shape[0])]])
                              .astype(np.float32))
    n = parallel.allreduce_sum_(packed)
    return (packed / np.float32(n)).numpy()


def test_normalizer_partials_are_averaged_over_ranks(tmp_path):
    res = _run('_body_norm', tmp_path)
    from oracle.ddpg_oracle import NormalizerOracle
    vs = [_norm_input(0), _norm_input(1)]

    class World:
        """the oracle's mean_over_ranks hook fed with both ranks' partials, in call order sum, sumsq, count"""
        def __init__(self):
            self.k = 0

        def __call__(self, x):
            which = self.k % 3
            self.k += 1
            parts = [(v.sum(0), np.square(v).sum(0), np.array([v.shape[0]], np.float32))[which] for v in vs]
            return ((parts[0] + parts[1]) / np.float32(2)).astype(np.float32)

    ora = NormalizerOracle(7, mean_over_ranks=World())
    ora.update(vs[0])
    ora.recompute_stats()
    dim = 7
    for r in res:
        r = r.astype(np.float32)
        assert np.array_equal((np.zeros(dim, np.float32) + r[:dim]), ora.sum)
        assert np.array_equal(r[dim:2 * dim], ora.sumsq)
        assert np.float32(1) + r[2 * dim] == ora.count[0]


def test_rank_seed_matches_reference():
    from curious_b200 import parallel
    assert parallel.rank_seed(5, 0) == 5 and parallel.rank_seed(5, 3) == 3000005      # train.py:242
    assert parallel.world(False) == (None, 1)


# ---------------------------------------------------------------------------------------------------
def _rollout_results(rank, i):
    rng = np.random.RandomState(1000 * rank + i)
    tasks = rng.randint(0, 3, size=2).tolist()
    return tasks, [float(rng.uniform() < 0.2 + 0.02 * i * (t == 1)) for t in tasks]


def _body_competence(rank, world):
    """LP pipeline (rollout.py:332-404): rank 0 gathers every rank's (task, success) pairs, updates the queues and
    broadcasts CP and p."""
    from curious_b200.queues import CompetenceTracker
    tr = CompetenceTracker(3, queue_length=6)
    for i in range(25):
        cp, p = tr.update(*_rollout_results(rank, i))
    return np.concatenate([np.asarray(cp, np.float64), np.asarray(p, np.float64)])


def test_competence_progress_is_gathered_on_rank0_and_broadcast(tmp_path):
    from curious_b200.queues import CompetenceTracker
    res = _run('_body_competence', tmp_path)
    assert np.array_equal(res[0], res[1])                       # every rank ends with rank 0's CP and p
    # emulate the world in-process: rank order of the gather is rank 0's pairs first (MPI gather order)
    tr = CompetenceTracker(3, queue_length=6, comm=False)
    for i in range(25):
        tasks, succ = [], []
        for r in range(2):
            t, s = _rollout_results(r, i)
            tasks += t
            succ += s
        cp, p = tr.update(tasks, succ)
    assert np.array_equal(res[0], np.concatenate([cp, p]))


def _body_mpi_average(rank, world):
    """mpi_average pools values and counts over ranks (her/util.py:141-146 -> mpi_moments.py:6-17); only rank 0's
    RunLog writes."""
    from curious_b200.runlog import mpi_average
    vals = [1.0, 2.0, 3.0] if rank == 0 else [10.0]
    return [mpi_average(vals), mpi_average(0.25 * (rank + 1)), mpi_average([])]


def test_mpi_average_pools_over_ranks(tmp_path):
    r0, r1 = _run('_body_mpi_average', tmp_path)
    assert np.array_equal(r0, r1)
    assert np.allclose(r0, [16.0 / 4, 0.375, 0.0])


class _StubPolicy(object):
    """Stands in for DDPG.get_actions on CPU: the rollout plumbing is what is under test."""

    def __init__(self, dimu):
        self.dimu = dimu

    def get_actions(self, o, ag, g, task_descr=None, compute_Q=False, **kw):
        u = np.random.uniform(-1, 1, (len(o), self.dimu))
        return (u, np.zeros((len(o), 1))) if compute_Q else u


def _body_rollouts(rank, world):
    """RolloutWorker on two ranks (rollout.py:118-140,316-404): rank 0 draws task and goal for every slot of every rank,
    the competence queues live on rank 0, CP and p come back to everybody."""
    from curious_b200.envs import ModularPointEnv
    from curious_b200.rollout import RolloutWorker
    from curious_b200.train import configure_dims
    np.random.seed(100 + rank)

    def make_env():
        return ModularPointEnv(3, max_episode_steps=10)
    dims = configure_dims(make_env(), 'curious')
    w = RolloutWorker(make_env, _StubPolicy(dims['u']), dims, None, T=10, rollout_batch_size=2, structure='curious',
                      task_selection='active_competence_progress', queue_length=20)
    w.seed(7 + rank)
    for _ in range(8):
        ep, cp, n = w.generate_rollouts()
    assert ep['o'].shape == (2, 11, dims['o']) and ep['task_descr'].shape == (2, 10, 3)
    mine = [int(e.unwrapped.task) for e in w.envs]
    goals = [float(np.abs(e.unwrapped.goal).sum()) for e in w.envs]
    drawn = [(-1 if t is None else int(t)) for t in w.tasks]            # filled on rank 0 only
    drawn_goals = [(-1.0 if g is None else float(np.abs(g).sum())) for g in w.goals]
    return [n] + list(np.asarray(w.p, np.float64)) + list(np.asarray(cp, np.float64)) + mine + goals + drawn + drawn_goals


def test_rollout_worker_assignments_come_from_rank0(tmp_path):
    r0, r1 = _run('_body_rollouts', tmp_path)
    B, N, world = 2, 3, 2
    assert r0[0] == r1[0] == 8 * B * world                              # n_episodes counts every rank's rollouts
    assert np.array_equal(r0[1:1 + 2 * N], r1[1:1 + 2 * N])             # p and CP identical everywhere
    assert abs(r0[1:1 + N].sum() - 1.0) < 1e-12
    k = 1 + 2 * N
    mine0, mine1 = r0[k:k + B], r1[k:k + B]
    goals0, goals1 = r0[k + B:k + 2 * B], r1[k + B:k + 2 * B]
    drawn = r0[k + 2 * B:k + 2 * B + B * world]
    drawn_goals = r0[k + 2 * B + B * world:]
    assert np.array_equal(drawn[:B], mine0) and np.array_equal(drawn[B:], mine1)          # slot = cpu * B + i
    assert np.allclose(drawn_goals[:B], goals0) and np.allclose(drawn_goals[B:], goals1)
    assert np.all(r1[k + 2 * B:k + 2 * B + B * world] == -1)           # the other rank never draws


def _body_epoch_records(rank, world):
    """train.py's per-epoch records on two ranks: tabular values are averaged over ranks (mpi_average), only rank 0 writes
    progress.csv / policies (train.py:171-206), every rank writes its own resumable checkpoint."""
    import tempfile
    from curious_b200.train import _EpochRecords

    class Worker(object):
        comm = None

        def __init__(self, rank):
            self.rank = rank

        def logs(self, prefix):
            return [(prefix + '/success_rate', 0.25 + 0.5 * self.rank), (prefix + '/episode', 40)]

        def additional_logs(self, prefix):
            return [(prefix + '/C_task0', '0.5')]

        def current_success_rate(self):
            return 0.25 + 0.5 * self.rank

        def save_policy(self, path):
            open(path, 'w').write('policy')

        def save_goal_task_history(self, path):
            pass

        def state(self):
            return {'rank': self.rank}

    class Policy(object):
        def logs(self):
            return [('stats_o/mean', 1.0 + rank)]

        def save_checkpoint(self, path):
            open(path, 'w').write('rank %d' % rank)

    box = [tempfile.mkdtemp() if rank == 0 else None]
    torch.distributed.broadcast_object_list(box, src=0)
    rec = _EpochRecords(box[0], Worker(rank), {'a': 1}, 1, True, 1, False)
    rec.epoch(0, Worker(rank), Policy())
    rec.log.close()
    torch.distributed.barrier()
    files = sorted(os.listdir(box[0]))
    want = ['checkpoint_0.pt', 'checkpoint_0_rank1.pt', 'log-rank001.txt', 'log.txt', 'params.json', 'policy_0.pkl', 'policy_best.pkl',
            'policy_latest.pkl', 'progress.csv', 'run_state.pkl', 'run_state_rank1.pkl']
    assert files == want, files
    assert open(os.path.join(box[0], 'checkpoint_0_rank1.pt')).read() == 'rank 1'
    lines = open(os.path.join(box[0], 'progress.csv')).read().splitlines()
    row = dict(zip(lines[0].split(','), lines[1].split(',')))
    return [float(row['test/success_rate']), float(row['stats_o/mean']), float(row['train/episode']), rec.best]


def test_epoch_records_on_two_ranks(tmp_path):
    r0, r1 = _run('_body_epoch_records', tmp_path)
    assert np.array_equal(r0[:3], [0.5, 1.5, 40.0]) and np.array_equal(r0[:3], r1[:3])
    assert r0[3] == 0.5 and r1[3] == -1                                    # only rank 0 tracks / saves the best policy


def _body_streams(rank, world):
    """train.py:207-212: ranks seeded alike are caught, ranks seeded with rank_seed pass."""
    from curious_b200 import parallel
    np.random.seed(parallel.rank_seed(5, rank))
    parallel.assert_rank_streams_differ()
    np.random.seed(5)                      # the mistake the check exists for
    try:
        parallel.assert_rank_streams_differ()
        caught = 0.0
    except AssertionError:
        caught = 1.0
    return [caught]


def test_identically_seeded_ranks_are_detected(tmp_path):
    r0, r1 = _run('_body_streams', tmp_path)
    assert r0[0] == 0.0 and r1[0] == 1.0          # rank 0 is the reference point, rank 1 notices


def test_excepthook_leaves_with_nonzero_status():
    """her/util.py:129-139: an uncaught exception ends the rank at once (the reference aborts MPI_COMM_WORLD)."""
    import subprocess
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from curious_b200.parallel import install_excepthook\n"
            "install_excepthook()\n"
            "import atexit; atexit.register(lambda: print('atexit ran'))\n"
            "raise ValueError('boom')\n") % ROOT
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
    assert res.returncode == 1 and 'ValueError: boom' in res.stderr and 'atexit ran' not in res.stdout


def _body_resume(rank, world):
    """train(resume=True) on two ranks: every rank restores its own checkpoint / run_state files and the run continues
    exactly like the uninterrupted one (tasks and goals keep coming from rank 0, CP / p stay identical everywhere)."""
    import tempfile
    from curious_b200 import parallel
    from curious_b200.train import train
    from tests.test_train_loop_cpu import _workers
    box = [[tempfile.mkdtemp(), tempfile.mkdtemp()] if rank == 0 else None]
    torch.distributed.broadcast_object_list(box, src=0)
    full_dir, split_dir = box[0]

    def run(logdir, n_epochs, seed, resume=False):
        np.random.seed(parallel.rank_seed(seed, rank))
        policy, rollout, evaluator, _ = _workers('curious')
        for i, w in enumerate([rollout, evaluator]):
            w.seed(50 + 10 * i + 100 * rank)
        hist = train(policy, rollout, evaluator, n_epochs=n_epochs, n_test_rollouts=2, n_cycles=2, n_batches=1,
                     structure='curious', logdir=logdir, policy_save_interval=0, checkpoint_interval=1, resume=resume)
        return hist, policy
    full, pol_full = run(full_dir, 4, seed=3)
    run(split_dir, 2, seed=3)
    tail, pol_tail = run(split_dir, 4, seed=77, resume=True)
    assert [h['epoch'] for h in tail] == [2, 3]
    for a, b in zip(full[2:], tail):
        for key in a:
            assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), (rank, a['epoch'], key)
    assert pol_full.trained == pol_tail.trained
    torch.distributed.barrier()
    files = sorted(os.listdir(split_dir))
    assert 'checkpoint_0.pt' in files and 'checkpoint_0_rank1.pt' in files and 'run_state_rank1.pkl' in files
    out = [h['test_success_rate'] for h in tail] + list(tail[-1]['p']) + list(tail[-1]['CP'])
    if rank == 0:
        rows = open(os.path.join(split_dir, 'progress.csv')).read().splitlines()
        assert [r.split(',')[0] for r in rows[1:]] == ['-1', '0', '1', '2', '3']
    return out


def test_resume_on_two_ranks(tmp_path):
    r0, r1 = _run('_body_resume', tmp_path)
    assert np.array_equal(r0[2:], r1[2:])                  # p and CP are broadcast: identical on both ranks
