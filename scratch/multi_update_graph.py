import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from tests.ddpg_util import ddpg_kwargs, episode_stream, make_gpu_agent
kw, dims, ag_ids, g_ids = ddpg_kwargs(4, batch_size=256)
ag = make_gpu_agent(kw, dims, ag_ids, g_ids, her_rng='philox', use_cuda_graph=True, buffer_episodes=2000)
np.random.seed(0); n = 0
for ep in episode_stream(dims, 50, 20):
    n += 2; ag.store_episode(ep, np.array([0.05, 0.2, 0.1, 0.0]), n)
for _ in range(10): ag.train()
torch.cuda.synchronize()
def timeit(fn, iters, per):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (iters * per)
print('1 update per graph: %.1f us/update' % timeit(ag._graph.replay, 300, 1))
for U in (2, 4, 10):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(U):
            ag._launch_sample_and_grads(keep_wT=True)
    print('%d updates per graph: %.1f us/update' % (U, timeit(g.replay, 300 // U, U)))
