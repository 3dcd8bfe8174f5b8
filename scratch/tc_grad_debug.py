import sys, os
sys.path.insert(0, '.')
import ctypes as C
import numpy as np, torch
from curious_b200 import _lib
from tests.ddpg_util import ddpg_kwargs, episode_stream, make_gpu_agent, make_oracle_agent, rel_err
from tests.test_ddpg_gpu import _fill
B = int(os.environ.get('B', 1024))
kw, dims, ag_ids, g_ids = ddpg_kwargs(4, batch_size=B)
cp = np.linspace(0.0, 0.3, 4)
episodes = episode_stream(dims, kw['T'], 12)
ora = make_oracle_agent(kw, dims, ag_ids, g_ids)
gpu = make_gpu_agent(kw, dims, ag_ids, g_ids, her_rng='numpy', update_schedule='levels')
np.random.seed(7); _fill(ora, episodes, cp)
np.random.seed(7); _fill(gpu, episodes, cp)
lib = _lib.load()
np.random.seed(300); ob = ora.sample_batch()
np.random.seed(300); gb = gpu.sample_batch()
ref = ora.grads(ob)
# float64 reference of the same graph, if the oracle offers it
for mode in (1, 0):
    lib.cur_ddpg_set_tensor_cores(mode)
    gpu.stage_batch(gb)
    ql, qpi, gq, gp = gpu._grads()
    for name, g, r, which in (('Q', gq.cpu().numpy(), ref['Q_grad'], 'Q'), ('pi', gp.cpu().numpy(), ref['pi_grad'], 'pi')):
        shapes = gpu.net.var_shapes(which)
        k = 0
        mx = np.abs(r).max()
        print('mode', mode, name, 'overall rel_err %.2e  max|g| %.3e' % (rel_err(g, r), mx))
        for s in shapes:
            n = int(np.prod(s))
            e = np.abs(g[k:k + n] - r[k:k + n]).max()
            print('   block %-12s max abs err %.2e (/max|g| %.2e)  block max %.2e' % (s, e, e / mx, np.abs(r[k:k + n]).max()))
            k += n
