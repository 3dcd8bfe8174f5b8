import sys; sys.path.insert(0,'.')
import numpy as np, torch
from tests.ddpg_util import *
for nm in (4,8):
    kw, dims, ag_ids, g_ids = ddpg_kwargs(nm)
    cp = np.linspace(0.0, 0.3, nm)
    episodes = episode_stream(dims, kw['T'], 12)
    ora = make_oracle_agent(kw, dims, ag_ids, g_ids); gpu = make_gpu_agent(kw, dims, ag_ids, g_ids, her_rng='numpy')
    for a in (ora,gpu):
        np.random.seed(2024); n=0
        for ep in episodes:
            n+=2; a.store_episode({k:v.copy() for k,v in ep.items()}, cp, n)
    for step in range(3):
        np.random.seed(100+step); ob = ora.sample_batch()
        gpu.stage_batch(ob)
        ql,qpi,gq,gp = gpu._grads()
        ref = ora.grads(ob)
        for name,g,r in (('Q',gq.cpu().numpy(),ref['Q_grad']),('pi',gp.cpu().numpy(),ref['pi_grad'])):
            err = np.abs(g-r); mx=np.abs(r).max()
            print(nm, step, name, 'max|g|',mx,'maxerr',err.max(),'argmax',err.argmax(),'n>2e-5*mx',(err>2e-5*mx).sum(),'n>2e-6*mx',(err>2e-6*mx).sum(), 'median', np.median(err))
        # fp64 reference of the oracle to see who is closer
