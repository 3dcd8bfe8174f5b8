"""Host-side profile of DDPG.store_episode in the bench workload (measurement script): tottime per call in us."""
import cProfile
import pstats
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import torch
import bench
from curious_b200 import synth
dev = torch.device('cuda', 0)
agent, sampler, buffers, dims, ag_ids, g_ids = bench.build_gpu_workload(dev, seed=1)
rng = np.random.RandomState(99)
host_eps = [synth.make_episodes(rng, 2, bench.T, dims, change_dtype=bool) for _ in range(8)]
cp = np.array(bench.CP)
for i in range(20):
    agent.store_episode({k: v for k, v in host_eps[i % 8].items()}, cp, 2 * (i + 1))
torch.cuda.synchronize()
N = 300
t0 = time.perf_counter()
for i in range(N):
    agent.store_episode({k: v for k, v in host_eps[i % 8].items()}, cp, 2 * (i + 1))
t1 = time.perf_counter()
torch.cuda.synchronize()
print('store_episode host time %.1f us per call (wall incl. device drain %.1f us)' % (1e6 * (t1 - t0) / N, 1e6 * (time.perf_counter() - t0) / N))
pr = cProfile.Profile()
pr.enable()
for i in range(N):
    agent.store_episode({k: v for k, v in host_eps[i % 8].items()}, cp, 2 * (i + 1))
pr.disable()
st = pstats.Stats(pr)
rows = sorted(st.stats.items(), key=lambda kv: -kv[1][2])[:32]
for (fn, line, name), (cc, nc, tt, ct, callers) in rows:
    print('%8.1f us tot %8.1f us cum  x%-4.1f %s:%d %s' % (1e6 * tt / N, 1e6 * ct / N, nc / N, fn.split('/')[-1], line, name))
