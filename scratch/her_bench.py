import sys, os, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
torch.cuda.set_device(0)
if os.environ.get('L2GRAN'):
    import ctypes
    rt = ctypes.CDLL('libcudart.so.12')
    torch.zeros(1, device='cuda')
    v = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5); print('L2 fetch granularity before', v.value)
    print('set ->', rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ['L2GRAN']))))
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5); print('L2 fetch granularity after', v.value)
dev = torch.device('cuda', 0)
from curious_b200 import her, synth
from curious_b200.replay_buffer import ReplayBuffer
from curious_b200.reward import ModuleDistanceReward
nmod = int(os.environ.get('NMOD', '4'))
bench.N_MODULES = nmod
dims = synth.arm_dims(nmod); ag_ids, g_ids = synth.arm_task_ids(nmod)
s = her.make_sample_multi_task_her_transitions(os.environ.get('REPLAY', 'her'), int(os.environ.get('K', '4')), 'replay_task_cp_buffer', ModuleDistanceReward(ag_ids, g_ids), tasks_ag_id=ag_ids, tasks_g_id=g_ids)
s.rng = 'philox'
shapes = synth.buffer_shapes(dims, 50)
bufs = [ReplayBuffer(shapes, int(os.environ.get('BUFSIZE', '1000000')) if i > 0 else 50, 50, s, device=dev) for i in range(nmod + 1)]
nfill = min(nmod, 5)
for i in range(1, nfill + 1):
    bench.fill_buffer_on_device(bufs[i], dims, i)
rows = int(os.environ.get('ROWS', 1 << 20))
per = rows // nfill
segs = [(bufs[i].device_view(), per if i < nfill else rows - per * (nfill - 1), i - 1) for i in range(1, nfill + 1)]
want = tuple(os.environ.get('WANT', 'o,g,u,td,o_2,r').split(','))
out = {}
def step():
    s.sample_device(segs, rows, clip_obs=200.0, want=want, out=out)
for _ in range(5): step()
torch.cuda.synchronize()
n = int(os.environ.get('ITERS', 200))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
bpt = bench.algorithmic_bytes_per_transition(dims, nmod)
print('nmod %d rows %d: %.4f ms/launch  %.3f G rows/s  algorithmic %.1f GB/s (%.1f%% of 6541.8)' % (nmod, rows, ms, rows / ms / 1e6, bpt * rows / ms / 1e6, 100 * bpt * rows / ms / 1e6 / 6541.8))
