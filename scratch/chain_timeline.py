"""In-kernel timeline of the fused chain kernel (first actor / critic CTA), measurement script."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curious_b200 import _lib  # noqa: E402
from tests.ddpg_util import ddpg_kwargs, episode_stream, make_gpu_agent  # noqa: E402

lib = _lib.load()
B = int(os.environ.get('B', 1024))
kw, dims, ag_ids, g_ids = ddpg_kwargs(4, batch_size=B)
lib.cur_ddpg_set_tensor_cores(1)
lib.cur_ddpg_set_chain(1)
ag = make_gpu_agent(kw, dims, ag_ids, g_ids, her_rng='philox', buffer_episodes=2000, update_schedule='levels', use_cuda_graph=False)
np.random.seed(0)
n = 0
for ep in episode_stream(dims, 50, 20):
    n += 2
    ag.store_episode(ep, np.array([0.05, 0.2, 0.1, 0.0]), n)
for _ in range(5):
    ag.train()
tl = torch.zeros(1024, dtype=torch.int64, device='cuda')
lib.cur_tc_chain_timeline(tl.data_ptr())
ag.train()
torch.cuda.synchronize()
lib.cur_tc_chain_timeline(None)
t = tl.cpu().numpy().reshape(2, 512)
for role, name in ((0, 'actor'), (1, 'critic')):
    r = t[role]
    t0 = r[0]
    print('%s CTA: total %d cycles' % (name, r[1] - t0))
    mma = r[8:96]; a = r[96:184]; b = r[192:280]; p = r[288:376]
    k = int((mma > 0).sum())
    print('  k-blocks %d; MMA issue deltas (cycles): %s' % (k, np.diff(mma[:k]).tolist()))
    print('  per k-block: TMA issue -> split done -> A ready -> MMA issue (relative to start)')
    for i in range(k if os.environ.get('FULL') else min(k, 12)):
        print('   kb %2d: tma %7d split %7d A %7d mma %7d' % (i, p[i] - t0, b[i] - t0, a[i] - t0, mma[i] - t0))
    st = r[400:500]
    st = st[st > 0]
    print('  out-layer stamps: per chunk (start, acc loaded, relu+mask, stored), then (dot done, combined); deltas:')
    print('   ', np.diff(st).tolist())
