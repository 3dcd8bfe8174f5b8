"""Where does the NVLS exchange (mode 2) differ from the rank-ordered sum after ONE update?  For every element the ranks'
partial gradients are gathered: a summation-order effect can only touch elements whose sum is tiny next to its partials
(cancellation); a protocol error (torn / stale word) would hit elements of any size.
torchrun --nproc-per-node N scratch/nvls_diag.py"""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
import bench
from curious_b200 import parallel
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
device = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=device)
agent, *_ = bench.build_gpu_workload(device, seed=1 + rank)
res = {}
for name, kw, ordered in (('nvls', dict(grad_exchange='tile', xchg_mode='nvls'), False), ('mode1', dict(grad_exchange='tile', xchg_mode=1), False),
                          ('ordered', dict(grad_exchange='nccl'), True)):
    parallel.ORDERED_ALLREDUCE = ordered
    a = agent.make_agent(**kw)
    th0 = a.theta_main.clone()
    a.train()
    torch.cuda.synchronize()
    res[name] = (a.theta_main.clone() - th0, a.grads.clone())
    del a; torch.cuda.empty_cache(); dist.barrier()
parallel.ORDERED_ALLREDUCE = False
# the local partial gradients of the update (mode1 agent: grads holds this rank's partial, the exchange does not touch it)
part = res['mode1'][1][:res['mode1'][0].numel()]
parts = [torch.empty_like(part) for _ in range(world)]
dist.all_gather(parts, part)
P = torch.stack(parts).double().cpu().numpy()
if rank == 0:
    d_nv, d_m1, d_or = [res[k][0].double().cpu().numpy() for k in ('nvls', 'mode1', 'ordered')]
    print('mode1 == ordered bit for bit:', np.array_equal(d_m1, d_or))
    diff = np.abs(d_nv - d_or)
    exact = P.sum(0)
    scale = np.abs(P).max(0)
    print('elements', diff.size, 'max |dtheta_nvls - dtheta_ordered| %.3g' % diff.max(), ' > 1e-7: %d   > 1e-6: %d   > 1e-5: %d' % ((diff > 1e-7).sum(), (diff > 1e-6).sum(), (diff > 1e-5).sum()))
    idx = np.argsort(-diff)[:12]
    for i in idx:
        print('  elem %7d  diff %.3g  |sum| %.3g  max|partial| %.3g  ratio %.2g  ulp(max partial) %.2g' % (i, diff[i], abs(exact[i]), scale[i], abs(exact[i]) / max(scale[i], 1e-300), np.spacing(np.float32(scale[i]))))
    big = diff > 1e-7
    if big.any():
        print('among the elements that differ by > 1e-7: largest |sum| / max|partial| = %.3g, largest |sum| = %.3g' % ((np.abs(exact[big]) / np.maximum(scale[big], 1e-300)).max(), np.abs(exact[big]).max()))
dist.destroy_process_group()
