import sys, os
sys.path.insert(0, '.')
import ctypes as C
import numpy as np, torch
from curious_b200 import _lib
from tests.ddpg_util import ddpg_kwargs, episode_stream, make_gpu_agent, make_oracle_agent, rel_err
from tests.test_ddpg_gpu import _fill
for B in (1024, 4096, 16384):
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
    out = {}
    for mode in (1, 0):
        lib.cur_ddpg_set_tensor_cores(mode)
        gpu.stage_batch(gb)
        ql, qpi, gq, gp = gpu._grads()
        out[mode] = (float(ql), float(gpu._pi_loss), qpi.cpu().numpy().copy(), gq.cpu().numpy().copy(), gp.cpu().numpy().copy())
    print('B', B, 'relu_margin', ref['relu_margin'])
    for name, a, b in (('tc vs oracle', out[1], None), ('ffma vs oracle', out[0], None), ('tc vs ffma', out[1], out[0])):
        if b is None:
            b = (ref['Q_loss'], ref['pi_loss'], ref['Q_pi'], ref['Q_grad'], ref['pi_grad'])
        e = np.abs(a[2] - b[2]).reshape(-1)
        print('  %-15s Q_loss rel %.2e pi_loss rel %.2e Q_pi rel_err %.2e (argmax row %d) Q_grad %.2e pi_grad %.2e' % (
            name, abs(a[0] - b[0]) / abs(b[0]), abs(a[1] - b[1]) / abs(b[1]), rel_err(a[2], b[2]), int(e.argmax()),
            rel_err(a[3], b[3]), rel_err(a[4], b[4])))
    e = np.abs(out[1][2] - out[0][2]).reshape(-1)
    bad = np.where(e > 1e-5)[0]
    print('  rows with |tc - ffma| > 1e-5:', len(bad), bad[:20], bad[-5:] if len(bad) else '')
