"""cProfile of DDPG.get_actions on the one-launch path (measurement script)."""
import cProfile
import os
import pstats
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.ddpg_util import ddpg_kwargs, make_gpu_agent  # noqa: E402

kw, dims, ag_ids, g_ids = ddpg_kwargs(4)
rng = np.random.RandomState(0)
a = make_gpu_agent(kw, dims, ag_ids, g_ids)
n = 2
o = rng.standard_normal((n, dims['o'])).astype(np.float32)
g = rng.uniform(-0.3, 0.3, (n, dims['g'])).astype(np.float32)
td = np.eye(4, dtype=np.float32)[[0, 1]]
for _ in range(200):
    a.get_actions(o, g, g, task_descr=td, noise_eps=0.2, random_eps=0.3)
pr = cProfile.Profile()
pr.enable()
for _ in range(5000):
    a.get_actions(o, g, g, task_descr=td, noise_eps=0.2, random_eps=0.3)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
