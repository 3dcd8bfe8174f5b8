import sys, os, time
sys.path.insert(0, '.')
import numpy as np, torch
from tests.ddpg_util import ddpg_kwargs, episode_stream, make_gpu_agent
graph = os.environ.get('GRAPH', '1') == '1'
B = int(os.environ.get('B', 256))
kw, dims, ag_ids, g_ids = ddpg_kwargs(4, batch_size=B)
sched = os.environ.get('SCHED', 'auto')
ag = make_gpu_agent(kw, dims, ag_ids, g_ids, her_rng='philox', use_cuda_graph=graph, buffer_episodes=2000, update_schedule=sched)
np.random.seed(0)
n = 0
for ep in episode_stream(dims, 50, 20):
    n += 2; ag.store_episode(ep, np.array([0.05, 0.2, 0.1, 0.0]), n)
for _ in range(10): ag.train()
torch.cuda.synchronize()
N = int(os.environ.get('ITERS', 300))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(N): ag.train()
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('sched=%s ' % sched + 'graph=%s B=%d: device %.1f us/update, host issue %.1f us/update, wall %.1f us/update' % (graph, B, 1e3 * e0.elapsed_time(e1) / N, 1e6 * (t1 - t0) / N, 1e6 * (t2 - t0) / N))
if os.environ.get('CUR_ROWS_TIMELINE'):
    from curious_b200 import _lib
    _lib.load().cur_rows_timeline_dump()
