import sys, os, time
sys.path.insert(0, '.')
import numpy as np, torch
from tests.ddpg_util import ddpg_kwargs, episode_stream, make_gpu_agent
from curious_b200.experts import TaskExperts
B = int(os.environ.get('B', 256)); NE = int(os.environ.get('NE', 4)); N = int(os.environ.get('ITERS', 100))
kw, dims, ag_ids, g_ids = ddpg_kwargs(4, structure='task_experts', task_replay='replay_current_task_buffer', batch_size=B)
def build(sched):
    out = []
    for t in range(NE):
        k = dict(kw); k['t_id'] = t % 4
        a = make_gpu_agent(k, dims, ag_ids, g_ids, her_rng='philox', seed=t, update_schedule=sched, buffer_episodes=2000)
        np.random.seed(0); n = 0
        for ep in episode_stream(dims, 50, 20):
            n += 2; a.store_episode(ep, np.array([0.05, 0.2, 0.1, 0.0]), n)
        out.append(a)
    return out
def timeit(fn):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / N
for sched in (['rows', 'levels'] if B <= 1024 else ['levels']):
    seq = build(sched)
    print('B=%d experts=%d sequential (%s): %.1f us per round of %d expert updates' % (B, NE, sched, timeit(lambda: [p.train() for p in seq]), NE))
grp = TaskExperts(build('levels'))
print('B=%d experts=%d grouped: %.1f us per round of %d expert updates' % (B, NE, timeit(lambda: grp.train()), NE))
