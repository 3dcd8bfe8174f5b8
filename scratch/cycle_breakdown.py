import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from curious_b200 import synth
dev = torch.device('cuda', 0)
agent, sampler, buffers, dims, ag_ids, g_ids = bench.build_gpu_workload(dev, seed=1)
rng = np.random.RandomState(99)
host_eps = [synth.make_episodes(rng, 2, bench.T, dims, change_dtype=bool) for _ in range(8)]
def cycle(i, t):
    t0 = time.perf_counter()
    agent.store_episode({k: v for k, v in host_eps[i % 8].items()}, np.array(bench.CP), 2 * (i + 1))
    t1 = time.perf_counter()
    losses = [agent.train()[0] for _ in range(100)]
    t2 = time.perf_counter()
    agent.update_target_net()
    t3 = time.perf_counter()
    r = torch.stack([l.tensor for l in losses]).cpu().numpy()
    t4 = time.perf_counter()
    t.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
for i in range(3): cycle(i, [])
torch.cuda.synchronize()
t = []
for i in range(20): cycle(3 + i, t)
a = np.array(t) * 1e6
print('per cycle (us): store_episode %.0f | 100 x train issue %.0f | update_target %.0f | stack + D2H wait %.0f | total %.0f' % (*a.mean(0), a.sum(1).mean()))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(20): agent.store_episode({k: v for k, v in host_eps[i % 8].items()}, np.array(bench.CP), 2 * (i + 1))
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
