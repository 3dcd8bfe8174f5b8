"""NVLS exchange (cur_xchg_ctx mode 2) next to modes 0 / 1 on N ranks: parameters after 8 updates vs the rank-ordered sum,
identical on every rank, per-rank update time.  torchrun --nproc-per-node N scratch/nvls_check.py"""
import os, sys
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import bench
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
device = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=device)
agent, *_ = bench.build_gpu_workload(device, seed=1 + rank)
out = {}
for mode in ([0, 1, 'nvls'] if world > 2 else [0, 'nvls']):
    a = agent.make_agent(grad_exchange='tile', xchg_mode=mode)
    for _ in range(8):
        a.train()
    torch.cuda.synchronize()
    th = a.theta_main.clone()
    ms = bench.time_updates(a.train, 300, torch)
    tt = torch.tensor([ms], device=device, dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out[mode] = (th, float(tt.item()))
    del a; torch.cuda.empty_cache(); dist.barrier()
base = out[0][0]
for mode, (th, ms) in out.items():
    ref = th.clone(); dist.broadcast(ref, src=0)
    same = torch.equal(ref, th)
    d = (th - base).abs().max(); dist.all_reduce(d, op=dist.ReduceOp.MAX)
    if rank == 0:
        print('mode %-5s per-rank update %.2f us   identical on rank %d: %s   max|theta - mode0| %.3g' % (mode, 1e3 * ms, rank, same, float(d)))
print(bench.nvls_exchange(agent, world, torch, dist, device)) if rank == 0 else bench.nvls_exchange(agent, world, torch, dist, device)
dist.destroy_process_group()
