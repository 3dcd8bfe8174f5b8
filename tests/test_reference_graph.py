"""The learner half of the path (networks, losses, gradients, MpiAdam, polyak, normaliser statistics, get_actions) against
the UNMODIFIED reference source:

* `test_oracle_against_reference_fixtures`: oracle/ddpg_oracle.py walks tests/golden/ddpg/*.npz - outputs recorded from
  baselines.her.ddpg.DDPG itself by oracle/gen_golden_ddpg.py (the reference's own graph code over oracle/tf1_shim.py).
  Runs anywhere (CPU).  tests/test_reference_graph_gpu.py walks the CUDA drop-in through the same files.
* `test_reference_agent_live_*` (build container only): the whole reference agent - store_episode with its normaliser
  update, sample_batch through the reference HER sampler and ReplayBuffer, train, update_target_net - run live next to the
  oracle agent on the same episodes and np.random seeds.
"""
import os
import pickle

import numpy as np
import pytest

from tests import ref_graph_util as R
from tests.ddpg_util import (REFERENCE_ROOT, ddpg_kwargs, episode_stream, make_oracle_agent, rel_err)

HAVE_REF = os.path.exists(os.path.join(REFERENCE_ROOT, 'baselines', 'her', 'ddpg.py'))


class OracleAdapter(R.Adapter):
    def __init__(self, ora):
        from oracle import ddpg_oracle as D
        self.o, self.D = ora, D

    def _net(self, which, target):
        return ('target_' if target else 'main_') + which

    def set_flat(self, which, flat, target):
        shapes = self.o.ac.Q_shapes if which == 'Q' else self.o.ac.pi_shapes
        setattr(self.o, self._net(which, target), self.D.unflatten(flat, shapes))
        if not target:
            (self.o.Q_adam if which == 'Q' else self.o.pi_adam).theta = np.asarray(flat, np.float32).copy()

    def get_flat(self, which, target):
        return self.D.flatten(getattr(self.o, self._net(which, target)))

    def set_stats(self, which, arrays):
        s = self.o.o_stats if which == 'o' else self.o.g_stats
        s.sum, s.sumsq, s.count, s.mean, s.std = [np.asarray(a, np.float32) for a in arrays]

    def grads(self, batch):
        self._batch = batch
        return self.o.grads(batch)

    def apply(self):
        self.o.train(self._batch)


def test_fixture_set_is_complete():
    names = R.cases()
    assert len(names) >= 10
    structures = {R.load(n)[0]['case']['structure'] for n in names}
    assert structures == {'curious', 'flat', 'task_experts'}
    assert any(R.load(n)[0]['case']['batch'] >= 1024 for n in names)       # a chain-schedule shape
    meta, _ = R.load('arm4_h64')
    # variable creation order of the reference graph = flat order of every parameter vector (common/tf_util.py:239-246)
    assert [v.split('/')[-2] + '/' + v.split('/')[-1] for v in meta['variables_Q']] == [
        '_0_state/kernel:0', '_0_state/bias:0', '_0_goal/kernel:0', '_1/kernel:0', '_1/bias:0', '_2/kernel:0', '_2/bias:0',
        '_3/kernel:0', '_3/bias:0']
    assert [v.split('/')[-1] for v in meta['stats_variables']] == ['sum:0', 'sumsq:0', 'count:0', 'mean:0', 'std:0']


@pytest.mark.parametrize('name', R.cases())
def test_oracle_against_reference_fixtures(name):
    meta, _ = R.load(name)
    kw, dims, ag_ids, g_ids = R.case_kwargs(meta['case'])
    ora = make_oracle_agent(kw, dims, ag_ids, g_ids, buffer_episodes=2)
    meta, z, worst = R.walk(name, OracleAdapter(ora), ora)
    R.check_actions(name, ora, z, dims, meta['seed'])
    if 'weights_pkl' in z.files:
        # the file the reference's save_weights wrote: six lists in the order main/Q, main/pi, target/Q, target/pi, o_stats,
        # g_stats; the parameter lists concatenate to the flat vectors
        w = pickle.loads(z['weights_pkl'].tobytes())
        assert len(w) == 6 and len(w[4]) == 5 and len(w[5]) == 5
        assert np.array_equal(np.concatenate([np.ravel(a) for a in w[0]])[::meta['stride']], z['main_Q_after'])
        assert np.array_equal(np.concatenate([np.ravel(a) for a in w[3]])[::meta['stride']], z['target_pi_after'])


def _copy_reference_weights(ref, ora):
    from oracle import ddpg_oracle as D
    from tests.ddpg_util import reference_flat
    ora.main_Q = D.unflatten(reference_flat(ref, 'Q'), ora.ac.Q_shapes)
    ora.main_pi = D.unflatten(reference_flat(ref, 'pi'), ora.ac.pi_shapes)
    ora.target_Q = D.unflatten(reference_flat(ref, 'Q', True), ora.ac.Q_shapes)
    ora.target_pi = D.unflatten(reference_flat(ref, 'pi', True), ora.ac.pi_shapes)
    ora.Q_adam.theta, ora.pi_adam.theta = D.flatten(ora.main_Q), D.flatten(ora.main_pi)


@pytest.mark.skipif(not HAVE_REF, reason='needs the reference checkout (build container)')
@pytest.mark.parametrize('structure,n_modules,normalize_obs,relative_goals,task_replay', [
    ('curious', 4, True, False, 'replay_task_cp_buffer'),
    ('curious', 4, False, True, 'replay_task_cp_buffer'),
    ('curious', 8, True, False, 'replay_task_random_buffer'),
    ('curious', 4, True, False, 'replay_cp_task_transition'),
    ('flat', 4, True, False, ''),
    ('task_experts', 4, True, False, 'replay_current_task_buffer'),
])
def test_reference_agent_live_equals_oracle(structure, n_modules, normalize_obs, relative_goals, task_replay):
    """baselines.her.ddpg.DDPG run live (its graph over the TF1 stand-in, its ReplayBuffer and HER sampler) next to the oracle
    agent: same Xavier draw, same episodes, same np.random seeds -> normaliser statistics, losses, Q_pi of every update and
    the parameters after 6 updates + 2 target updates agree at float32 rounding level."""
    from tests.ddpg_util import make_reference_agent, reference_flat
    from oracle import ddpg_oracle as D
    kw, dims, ag_ids, g_ids = ddpg_kwargs(n_modules, structure=structure, task_replay=task_replay, hidden=64, batch_size=48,
                                          normalize_obs=normalize_obs, relative_goals=relative_goals)
    if structure == 'task_experts':
        kw['t_id'] = 2
    state = np.random.get_state()
    try:
        ref = make_reference_agent(kw, dims, ag_ids, g_ids)
        ora = make_oracle_agent(kw, dims, ag_ids, g_ids)
        _copy_reference_weights(ref, ora)
        assert np.array_equal(reference_flat(ref, 'Q'), reference_flat(ref, 'Q', True))      # _init_target_net (ddpg.py:459)
        cp = np.linspace(0.02, 0.3, n_modules)
        for agent in (ref, ora):
            np.random.seed(11)
            n = 0
            for ep in episode_stream(dims, kw['T'], 5, flat=structure == 'flat'):
                n += 2
                agent.store_episode({k: v.copy() for k, v in ep.items()}, cp, n)
        for rs, os_ in ((ref.o_stats, ora.o_stats), (ref.g_stats, ora.g_stats)):
            assert np.allclose(rs.mean.value.numpy(), os_.mean, rtol=1e-6, atol=1e-7)
            assert np.allclose(rs.std.value.numpy(), os_.std, rtol=1e-6, atol=1e-7)
            assert float(rs.count_tf.value.numpy()[0]) == float(os_.count[0])
        for step in range(6):
            np.random.seed(40 + step)
            rl, rq = ref.train()
            np.random.seed(40 + step)
            ol, oq = ora.train()
            assert abs(float(rl) - float(ol)) <= 2e-6 * abs(float(ol)) + 1e-7, (step, float(rl), float(ol))
            assert rel_err(oq, rq) <= 2e-5, step
            if step % 3 == 2:
                ref.update_target_net()
                ora.update_target_net()
        for which, main, target in (('Q', ora.main_Q, ora.target_Q), ('pi', ora.main_pi, ora.target_pi)):
            assert np.abs(D.flatten(main) - reference_flat(ref, which)).max() <= 5e-6, which
            assert np.abs(D.flatten(target) - reference_flat(ref, which, True)).max() <= 5e-6, which
        # the exploration path of get_actions consumes np.random in the same order (ddpg.py:147-155)
        o = np.random.RandomState(5).standard_normal((3, dims['o'])).astype(np.float32)
        z = np.zeros((3, dims['g']), np.float32)
        td = None if structure == 'flat' else np.eye(dims['task_descr'], dtype=np.float32)[:3]
        np.random.seed(77)
        ru = ref.get_actions(o, z, z, task_descr=td, noise_eps=0.2, random_eps=0.3)
        np.random.seed(77)
        ou = ora.get_actions(o, z, z, task_descr=td, noise_eps=0.2, random_eps=0.3)
        assert np.abs(np.asarray(ru) - np.asarray(ou)).max() <= 1e-5
    finally:
        np.random.set_state(state)


@pytest.mark.skipif(not HAVE_REF, reason='needs the reference checkout (build container)')
def test_committed_fixtures_are_what_the_reference_computes_today():
    """Re-run the generator for one small case and compare with the committed file."""
    from oracle import gen_golden_ddpg as G
    i = [c['name'] for c in G.CASES].index('flat_h64')
    rec = G.run_case(G.CASES[i], seed=7000 + 37 * i)
    _, z = R.load('flat_h64')
    for k in ('Q_loss', 'pi_loss', 'Q_pi', 'Q_grad0', 'pi_grad0', 'main_Q_after', 'target_pi_after', 'act_u_main'):
        assert np.array_equal(rec[k], z[k]), k


@pytest.mark.parametrize('name', R.agent_cases())
def test_oracle_agent_against_reference_trajectories(name):
    """Whole-agent fixtures (store_episode -> statistics; train() with the reference's own sampling): the oracle agent under
    the recorded np.random seeds."""
    from oracle.gen_golden_ddpg import agent_kwargs
    meta, _ = R.load(name)
    kw, dims, ag_ids, g_ids = agent_kwargs(meta['case'])
    ora = make_oracle_agent(kw, dims, ag_ids, g_ids)
    ad = OracleAdapter(ora)
    R.walk_agent(name, ora, ad.get_flat, ad.set_flat,
                 lambda tag: ((ora.o_stats if tag == 'o' else ora.g_stats).mean, (ora.o_stats if tag == 'o' else ora.g_stats).std,
                              (ora.o_stats if tag == 'o' else ora.g_stats).count[0]))
