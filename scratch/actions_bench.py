"""get_actions latency, host arrays in -> host actions out (measurement script): one-launch path vs the level kernels."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.ddpg_util import ddpg_kwargs, make_gpu_agent  # noqa: E402

kw, dims, ag_ids, g_ids = ddpg_kwargs(4)
rng = np.random.RandomState(0)
for path in ('auto', 'levels'):
    a = make_gpu_agent(kw, dims, ag_ids, g_ids, action_path=path)
    for n in (2, 38, 256):
        o = rng.standard_normal((n, dims['o'])).astype(np.float32)
        g = rng.uniform(-0.3, 0.3, (n, dims['g'])).astype(np.float32)
        ag = rng.uniform(-0.3, 0.3, (n, dims['ag'])).astype(np.float32)
        td = np.eye(4, dtype=np.float32)[rng.randint(0, 4, n)]
        for noise in ((0.2, 0.3), (0.0, 0.0)):
            for _ in range(50):
                a.get_actions(o, ag, g, task_descr=td, noise_eps=noise[0], random_eps=noise[1])
            t0 = time.perf_counter()
            for _ in range(2000):
                a.get_actions(o, ag, g, task_descr=td, noise_eps=noise[0], random_eps=noise[1])
            dt = (time.perf_counter() - t0) / 2000
            print('%-6s n=%3d noise=%s  %.1f us/call' % (path, n, noise, 1e6 * dt), flush=True)
    if path == 'auto':
        # where the time goes: kernel alone (device events, back to back)
        slot = a._action_slot(2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o = rng.standard_normal((2, dims['o'])).astype(np.float32)
        g = rng.uniform(-0.3, 0.3, (2, dims['g'])).astype(np.float32)
        td = np.eye(4, dtype=np.float32)[[0, 1]]
        torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            a.get_actions(o, g, g, task_descr=td)
        e1.record()
        torch.cuda.synchronize()
        print('device time per call incl. gaps (n=2): %.1f us' % (1e3 * e0.elapsed_time(e1) / 200))

# host-side breakdown of one call on the one-launch path (n = 2)
import ctypes as C
from curious_b200 import _lib
a = make_gpu_agent(kw, dims, ag_ids, g_ids)
o = rng.standard_normal((2, dims['o'])).astype(np.float32)
g = rng.uniform(-0.3, 0.3, (2, dims['g'])).astype(np.float32)
td = np.eye(4, dtype=np.float32)[[0, 1]]
for _ in range(200):
    a.get_actions(o, g, g, task_descr=td)
slot = a._action_slot(2)
lib = _lib.load()
t_prep = t_launch = t_wait = 0.0
N = 2000
tag = slot['tag'][:8]
for i in range(N):
    t0 = time.perf_counter()
    np.copyto(slot['o'], o.reshape(-1)); np.copyto(slot['g'], g.reshape(-1)); np.copyto(slot['td'], td.reshape(-1))
    slot['seq'] = seq = slot['seq'] % 0xFFFFFFF0 + 1
    t1 = time.perf_counter()
    lib.cur_ddpg_actions_rows(_lib.stream_ptr(), C.byref(a.net.desc), a.theta_main.data_ptr(), C.byref(a._stats), slot['d_o'],
                              None, slot['d_g'], slot['d_td'], 2, 200.0, slot['d_out'], None, seq)
    t2 = time.perf_counter()
    while tag[7] != seq or not (tag == seq).all():
        pass
    t3 = time.perf_counter()
    t_prep += t1 - t0; t_launch += t2 - t1; t_wait += t3 - t2
print('host breakdown n=2: prep %.1f us | launch call %.1f us | wait for the output words %.1f us' %
      (1e6 * t_prep / N, 1e6 * t_launch / N, 1e6 * t_wait / N))
