"""2-rank NCCL test of the data-parallel DDPG update (needs >= 2 GPUs; skipped otherwise).

Reference semantics (mpi_adam.py:24-28 with scale_grad_by_procs=False, ddpg.py:452-453): every rank
computes gradients on ITS OWN batch, the flat gradients are SUMMED over ranks, every rank applies the same
Adam step, and the parameters stay bit-identical (check_synced, mpi_adam.py:42-50).  Replay data never
crosses ranks.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, use_graph, out_dir, grad_exchange='auto', workers_per_rank=1, workers_mode='micro',
            xchg_mode=None):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    if xchg_mode is not None:
        os.environ['CUR_XCHG_MODE'] = str(xchg_mode)
    else:
        os.environ.pop('CUR_XCHG_MODE', None)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from curious_b200 import parallel
        from tests.ddpg_util import ddpg_kwargs, episode_stream, make_gpu_agent
        parallel.ORDERED_ALLREDUCE = True      # the NCCL arm sums in rank order (the reference for > 2 ranks)
        kw, dims, ag_ids, g_ids = ddpg_kwargs(4)
        # rank 1 starts from different weights: _sync_optimizers must broadcast rank 0's (ddpg.py:466)
        agent = make_gpu_agent(kw, dims, ag_ids, g_ids, her_rng='philox', seed=rank, use_cuda_graph=use_graph,
                               device=dev, grad_exchange=grad_exchange, workers_per_rank=workers_per_rank,
                               workers_mode=workers_mode)
        assert parallel.world(agent.comm)[1] == world
        theta0 = agent.theta_main.clone()
        gathered = [torch.empty_like(theta0) for _ in range(world)]
        dist.all_gather(gathered, theta0)
        assert all(torch.equal(gathered[0], g) for g in gathered[1:]), 'init broadcast from rank 0 failed'
        np.random.seed(parallel.rank_seed(0, rank))                       # train.py:242
        n = 0
        for ep in episode_stream(dims, kw['T'], 6, seed=123 + rank):     # each rank owns its replay data
            n += 2
            agent.store_episode(ep, np.array([0.05, 0.2, 0.1, 0.0]), n)
        # normaliser statistics are averaged over ranks -> identical everywhere (normalizer.py:84-94)
        st = torch.cat([agent.o_stats.mean, agent.o_stats.std, agent.g_stats.mean, agent.g_stats.std])
        parallel.assert_synced(st)

        if not use_graph and workers_per_rank == 1:
            # one eager update, dissected: local grads -> all-gather -> expected Adam(sum) on a clone
            agent.stage_batch()
            agent._grads()
            local = agent.grads.clone()
            parts = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(parts, local)
            assert not torch.equal(parts[0], parts[1]), 'ranks must train on different batches'
            expect = parts[0].clone()
            for p in parts[1:]:
                expect += p
            agent._update(agent._view(agent.grads, 'Q'), agent._view(agent.grads, 'pi'))
            assert torch.equal(agent.grads, expect), 'all-reduce must be a plain SUM'
        for _ in range(120):                                              # crosses the every-100 check_synced
            agent.train()
        if use_graph:
            kind = 'tile' if grad_exchange == 'auto' else grad_exchange
            assert (agent._xchg is not None) == (kind == 'tile'), 'wrong gradient exchange path'
            assert (agent._peer is not None) == (kind in ('p2p', 'p2p_sharded')), 'wrong gradient exchange path'
            if agent._xchg is not None:
                assert agent._xchg.mode == (xchg_mode if xchg_mode is not None else (0 if world == 2 else 1))
                assert int(agent._xchg.error_flag.item()) == 0
            if agent._peer is not None:
                agent._peer.check()
        agent.update_target_net()
        torch.cuda.synchronize()
        fp = agent.theta_main.clone()
        parallel.assert_synced(fp)
        assert torch.isfinite(fp).all()
        assert not torch.equal(fp, theta0)
        np.save(os.path.join(out_dir, 'theta%d.npy' % rank), fp.cpu().numpy())
        np.save(os.path.join(out_dir, 'ok%d.npy' % rank), np.array([1.0]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('use_graph', [False, True])
def test_two_rank_update_sums_gradients_and_stays_synced(use_graph, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    mp.spawn(_worker, args=(2, _free_port(), use_graph, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), 'ok0.npy'))
    assert os.path.exists(os.path.join(str(tmp_path), 'ok1.npy'))


def test_peer_memory_exchange_equals_nccl_allreduce(tmp_path):
    """The fused peer-memory exchange + Adam kernels (csrc/p2p.cu; sharded and full variants) sum the ranks'
    gradients in rank order; with two ranks that is the same float32 sum as NCCL's, so 120 updates end on
    bit-identical parameters in all three modes."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    thetas = {}
    for mode in ('p2p', 'p2p_sharded', 'nccl'):
        d = tmp_path / mode
        d.mkdir()
        mp.spawn(_worker, args=(2, _free_port(), True, str(d), mode), nprocs=2, join=True)
        thetas[mode] = [np.load(os.path.join(str(d), 'theta%d.npy' % r)) for r in range(2)]
        assert np.array_equal(thetas[mode][0], thetas[mode][1])
    assert np.array_equal(thetas['p2p'][0], thetas['nccl'][0])
    assert np.array_equal(thetas['p2p_sharded'][0], thetas['nccl'][0])


@pytest.mark.parametrize('world', [2, 4, 8])
def test_tile_exchange_equals_rank_ordered_allreduce(world, tmp_path):
    """The gradient exchange inside the weight-gradient launch (csrc/ddpg_rows.cu `cur_xchg_ctx`: LL pushes over NVLink
    peer memory, sum in rank order, Adam + W^T in the same epilogue) in both modes - every rank reduces every tile /
    tile t reduced by rank t % world - against an all-gather + rank-ordered sum + Adam launch after the graph:
    120 updates, bit-identical parameters on every rank and between the three paths."""
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    thetas = {}
    for name, mode, xm in (('tile_all', 'tile', 0), ('tile_owner', 'tile', 1), ('nccl', 'nccl', None)):
        d = tmp_path / name
        d.mkdir()
        mp.spawn(_worker, args=(world, _free_port(), True, str(d), mode, 1, 'micro', xm), nprocs=world, join=True)
        thetas[name] = [np.load(os.path.join(str(d), 'theta%d.npy' % r)) for r in range(world)]
        assert all(np.array_equal(thetas[name][0], t) for t in thetas[name][1:])
    assert np.array_equal(thetas['tile_all'][0], thetas['nccl'][0])
    assert np.array_equal(thetas['tile_owner'][0], thetas['nccl'][0])


@pytest.mark.parametrize('world', [2, 4, 8])
def test_nvls_exchange_against_rank_ordered_allreduce(world, tmp_path):
    """Mode 2 of the tile exchange (`multimem.ld_reduce` through the NVSwitch on a multicast mapping of the partial slots, one
    multicast store of the stepped parameters): 120 updates end on parameters that are bit-identical on every rank; with two
    ranks the in-switch sum is the same float32 sum as the rank-ordered one (bit-equal), with more ranks the switch fixes the
    order and the result stays within a few ulp per update of the rank-ordered trajectory."""
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    thetas = {}
    for name, mode, xm in (('nvls', 'tile', 2), ('nccl', 'nccl', None)):
        d = tmp_path / name
        d.mkdir()
        mp.spawn(_worker, args=(world, _free_port(), True, str(d), mode, 1, 'micro', xm), nprocs=world, join=True)
        thetas[name] = [np.load(os.path.join(str(d), 'theta%d.npy' % r)) for r in range(world)]
        assert all(np.array_equal(thetas[name][0], t) for t in thetas[name][1:])
    if world == 2:
        assert np.array_equal(thetas['nvls'][0], thetas['nccl'][0])
    else:
        d = np.abs(thetas['nvls'][0] - thetas['nccl'][0])
        assert d.max() <= 2e-3 and (d > 1e-5).mean() <= 0.02      # 120 Adam steps apart by rounding only (cf. ref_graph_util)


def test_two_ranks_with_two_workers_each(tmp_path):
    """4-worker-equivalent on 2 GPUs: every rank sums the gradients of its 2 batch-256 workers in the
    weight-gradient epilogue, the peer-memory kernel sums the ranks and steps Adam with t = launches / 2; the NCCL
    path (all-reduce + Adam launch after the graph) must land on the same parameters bit for bit.  (The
    launch-by-launch path draws from a different Philox offset here; its equality with the graph path is covered on
    one GPU by test_workers_per_rank_sums_single_batch_gradients.)"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    thetas = {}
    for mode, graph in (('p2p', True), ('nccl', True), ('tile', True)):
        d = tmp_path / ('%s_%d' % (mode, graph))
        d.mkdir()
        mp.spawn(_worker, args=(2, _free_port(), graph, str(d), mode, 2), nprocs=2, join=True)
        thetas[(mode, graph)] = [np.load(os.path.join(str(d), 'theta%d.npy' % r)) for r in range(2)]
        assert np.array_equal(thetas[(mode, graph)][0], thetas[(mode, graph)][1])
    assert np.array_equal(thetas[('p2p', True)][0], thetas[('nccl', True)][0])
    assert np.array_equal(thetas[('tile', True)][0], thetas[('nccl', True)][0])


def test_two_ranks_wide_worker_batches(tmp_path):
    """3 workers per rank as ONE 768-row batch on the rows schedule (several waves of the stream kernel, three
    accumulating weight-gradient launches, loss seeds scaled by 1 / 256): peer-memory exchange == NCCL, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    thetas = {}
    for mode in ('p2p', 'nccl', 'tile'):
        d = tmp_path / mode
        d.mkdir()
        mp.spawn(_worker, args=(2, _free_port(), True, str(d), mode, 3, 'auto', 1 if mode == 'tile' else None),
                 nprocs=2, join=True)
        thetas[mode] = [np.load(os.path.join(str(d), 'theta%d.npy' % r)) for r in range(2)]
        assert np.array_equal(thetas[mode][0], thetas[mode][1])
    assert np.array_equal(thetas['p2p'][0], thetas['nccl'][0])
    assert np.array_equal(thetas['tile'][0], thetas['nccl'][0])
