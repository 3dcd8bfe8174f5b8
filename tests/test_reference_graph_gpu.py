"""The CUDA drop-in against outputs of the UNMODIFIED reference learner (tests/golden/ddpg/*.npz: baselines.her.ddpg.DDPG run
over oracle/tf1_shim.py by oracle/gen_golden_ddpg.py): losses, Q_pi, gradients, parameters after several MpiAdam / polyak
updates, get_actions of both networks, and the reference's own `_weights.pkl` file - on every update schedule that takes the
shape (levels: any; rows: hidden 256 below 1024 rows; chain: hidden 256 from 1024 rows)."""
import os
import pickle

import numpy as np
import pytest

from tests import ref_graph_util as R
from tests.ddpg_util import make_gpu_agent

pytestmark = pytest.mark.gpu


class GpuAdapter(R.Adapter):
    def __init__(self, agent):
        self.a = agent

    def set_flat(self, which, flat, target):
        self.a.set_flat(which, flat, target)

    def get_flat(self, which, target):
        return self.a.get_flat(which, target)

    def set_stats(self, which, arrays):
        (self.a.o_stats if which == 'o' else self.a.g_stats).load_state_list(arrays)

    def grads(self, batch):
        self.a.stage_batch(batch)
        ql, qpi, gq, gp = self.a._grads()
        self._g = (gq, gp)
        # (gradients in the reference's flat order: a gather when the net runs zero-padded to the kernels' width)
        return dict(Q_loss=float(ql), pi_loss=float(self.a._pi_loss), Q_pi=qpi.cpu().numpy(),
                    Q_grad=self.a.ref_flat(self.a.grads, 'Q').cpu().numpy(),
                    pi_grad=self.a.ref_flat(self.a.grads, 'pi').cpu().numpy())

    def apply(self):
        self.a._update(*self._g)


def _schedules(case):
    """levels: the dependency-level kernels at the true width; rows / auto: the row kernels (below 1024 rows) or the tcgen05
    chain kernel - hidden 64 nets run on them zero-padded to 256 (DDPG pad_hidden)."""
    if case['layers'] <= 4:
        return ['levels', 'auto'] if case['batch'] >= 1024 else ['levels', 'rows']
    return ['levels']


CASES = [(n, s) for n in R.cases() for s in _schedules(R.load(n)[0]['case'])]


@pytest.mark.parametrize('name,schedule', CASES)
def test_cuda_path_against_reference_fixtures(name, schedule, tmp_path):
    meta, z = R.load(name)
    case = meta['case']
    kw, dims, ag_ids, g_ids = R.case_kwargs(case)
    gpu = make_gpu_agent(kw, dims, ag_ids, g_ids, buffer_episodes=2, her_rng='numpy', update_schedule=schedule)
    if schedule == 'rows':
        assert gpu._use_rows(case['batch'])
    assert gpu.net.padded == (schedule != 'levels' and case['hidden'] < 256)
    meta, z, worst = R.walk(name, GpuAdapter(gpu), gpu)
    R.check_actions(name, gpu, z, dims, meta['seed'])
    if 'weights_pkl' in z.files:
        # the reference's own file loads into the drop-in, and the drop-in writes the same structure back
        base = str(tmp_path / 'ref_policy')
        open(base + '_weights.pkl', 'wb').write(z['weights_pkl'].tobytes())
        fresh = make_gpu_agent(kw, dims, ag_ids, g_ids, buffer_episodes=2, her_rng='numpy', update_schedule=schedule)
        fresh.load_weights(base)
        for w in ('Q', 'pi'):
            for tgt in (False, True):
                assert np.array_equal(fresh.get_flat(w, tgt)[::meta['stride']], z['%s_%s_after' % ('target' if tgt else 'main', w)])
        R.check_actions(name, fresh, z, dims, meta['seed'], rtol=2e-5)      # now on the reference's exact parameters
        fresh.save_weights(str(tmp_path / 'mine'))
        mine = pickle.load(open(str(tmp_path / 'mine') + '_weights.pkl', 'rb'))
        theirs = pickle.loads(z['weights_pkl'].tobytes())
        assert len(mine) == len(theirs) == 6
        for a, b in zip(mine, theirs):
            assert len(a) == len(b)
            for x, y in zip(a, b):
                assert np.shape(x) == np.shape(y) and np.array_equal(np.asarray(x, np.float32), np.asarray(y, np.float32))


@pytest.mark.parametrize('schedule', ['auto', 'levels'])
@pytest.mark.parametrize('name', R.agent_cases())
def test_cuda_agent_against_reference_trajectories(name, schedule):
    """The whole hot path against the reference agent's recorded run: DDPG.store_episode (routing, normaliser update through the
    sampler) and DDPG.train() with its own sampling in her_rng='numpy' mode (np.random consumed in the reference's order) -
    statistics, buffer fill levels, LP proportions, Q_loss / Q_pi per update, parameters after, final np.random state."""
    from oracle.gen_golden_ddpg import agent_kwargs
    meta, _ = R.load(name)
    kw, dims, ag_ids, g_ids = agent_kwargs(meta['case'])
    gpu = make_gpu_agent(kw, dims, ag_ids, g_ids, her_rng='numpy', update_schedule=schedule)

    def stats_of(tag):
        s = gpu.o_stats if tag == 'o' else gpu.g_stats
        return s.mean.cpu().numpy(), s.std.cpu().numpy(), float(s.count.cpu()[0])
    R.walk_agent(name, gpu, gpu.get_flat, gpu.set_flat, stats_of)
