"""train() time over the batch size: fused chain kernel (tc_chain.cu) vs the level-by-level tensor-core schedule vs the rows
schedule (measurement script).  BATCHES="512 1024 ..." selects the sizes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curious_b200 import _lib  # noqa: E402
from tests.ddpg_util import ddpg_kwargs, episode_stream, make_gpu_agent  # noqa: E402

lib = _lib.load()
batches = [int(x) for x in os.environ.get('BATCHES', '512 768 1024 1280 2048 2560 4096 4864 8192 16384').split()]
for B in batches:
    kw, dims, ag_ids, g_ids = ddpg_kwargs(4, batch_size=B)
    res = {}
    for name, sched, tc, chain in (('chain', 'levels', 1, 1), ('levels', 'levels', 1, 0), ('rows', 'rows', -1, -1)):
        if name == 'rows' and B > 1024:
            continue
        lib.cur_ddpg_set_tensor_cores(tc)
        lib.cur_ddpg_set_chain(chain)
        ag = make_gpu_agent(kw, dims, ag_ids, g_ids, her_rng='philox', buffer_episodes=2000, update_schedule=sched)
        np.random.seed(0)
        n = 0
        for ep in episode_stream(dims, 50, 20):
            n += 2
            ag.store_episode(ep, np.array([0.05, 0.2, 0.1, 0.0]), n)
        for _ in range(5):
            ag.train()
        torch.cuda.synchronize()
        N = 100 if B <= 4864 else 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(N):
            ag.train()
        e1.record()
        torch.cuda.synchronize()
        res[name] = 1e3 * e0.elapsed_time(e1) / N
        del ag
        torch.cuda.empty_cache()
    print('batch %5d: ' % B + ' | '.join('%s %.1f us' % (k, v) for k, v in res.items()), flush=True)
lib.cur_ddpg_set_tensor_cores(-1)
lib.cur_ddpg_set_chain(-1)
