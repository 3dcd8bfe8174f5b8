"""Per-rank update time of the multi-rank gradient exchanges (torchrun, one rank per GPU; measurement script).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scratch/xchg_bench.py [modes...]          modes: tile0 tile1 p2p p2p_sharded nccl
With TL=1 the tile modes also dump per-tile %globaltimer stamps (profiles/r02_xchg_timeline_*.json).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from curious_b200 import parallel
    from tests.ddpg_util import ddpg_kwargs, episode_stream, make_gpu_agent
    modes = sys.argv[1:] or ['tile0', 'tile1', 'p2p', 'nccl']
    workers = int(os.environ.get('WORKERS', '1'))
    kw, dims, ag_ids, g_ids = ddpg_kwargs(4)
    out = {}
    for mode in modes:
        ge = 'tile' if mode.startswith('tile') else mode
        if mode.startswith('tile') and len(mode) > 4:
            os.environ['CUR_XCHG_MODE'] = mode[4:]
        else:
            os.environ.pop('CUR_XCHG_MODE', None)
        tl_tiles = 512 if (os.environ.get('TL') == '1' and ge == 'tile') else 0
        agent = make_gpu_agent(kw, dims, ag_ids, g_ids, her_rng='philox', seed=0, device=dev, grad_exchange=ge,
                               buffer_episodes=400, workers_per_rank=workers, xchg_timeline_tiles=tl_tiles)
        np.random.seed(parallel.rank_seed(0, rank))
        n = 0
        for ep in episode_stream(dims, kw['T'], 20, seed=123 + rank):
            n += 2
            agent.store_episode(ep, np.array([0.05, 0.2, 0.1, 0.0]), n)
        for _ in range(20):
            agent.train()
        torch.cuda.synchronize()
        dist.barrier()
        reps = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(200):
                agent.train()
            e1.record()
            torch.cuda.synchronize()
            reps.append(1e3 * e0.elapsed_time(e1) / 200)
        t = torch.tensor(reps, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parallel.assert_synced(agent.theta_main.clone())
        out[mode] = {'update_us_median': float(t.median()), 'update_us_min': float(t.min()), 'reps': [float(x) for x in t.cpu()]}
        if tl_tiles:
            tl = agent._xchg.timeline.cpu().numpy().reshape(-1, 4)
            tl = tl[tl[:, 0] > 0]
            t0 = tl[:, 0].min()
            allt = [None] * world
            dist.all_gather_object(allt, (tl - t0).tolist() + [[int(t0)] * 4])
            if rank == 0:
                os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
                json.dump({'world': world, 'mode': mode, 'note': 'per tile: CTA start, tile computed (before push), reduced '
                           'tile available, CTA end; ns relative to the rank\'s first CTA start; last row = that start '
                           '(%globaltimer, comparable across GPUs only approximately)', 'ranks': allt},
                          open(os.path.join(ROOT, 'gpurun_out', 'xchg_timeline_%s_n%d.json' % (mode, world)), 'w'))
                a = np.array(allt[0][:-1])
                print('[timeline %s] rank0 tiles %d: computed-start med %.1f us | wait (reduce ready - computed) med %.1f max %.1f us | '
                      'CTA total med %.1f max %.1f us' % (mode, len(a), np.median(a[:, 1] - a[:, 0]) / 1e3,
                                                          np.median(a[:, 2] - a[:, 1]) / 1e3, np.max(a[:, 2] - a[:, 1]) / 1e3,
                                                          np.median(a[:, 3] - a[:, 0]) / 1e3, np.max(a[:, 3] - a[:, 0]) / 1e3),
                      flush=True)
        del agent
        torch.cuda.empty_cache()
        dist.barrier()
    if rank == 0:
        print(json.dumps({'world': world, 'workers_per_rank': workers, 'modes': out}), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
