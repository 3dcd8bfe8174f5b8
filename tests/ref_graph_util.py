"""Shared by the CPU and GPU tests of the reference-graph fixtures (tests/golden/ddpg/*.npz, oracle/gen_golden_ddpg.py):
rebuild the seeded inputs of a case and walk an agent through it, comparing with what the UNMODIFIED reference computed."""
import glob
import json
import os

import numpy as np

from tests.ddpg_util import rel_err, seeded_batch, seeded_net_flats, seeded_stats

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ddpg')
LOSS_RTOL = 1e-5
GRAD_RTOL = 2e-5


def cases():
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz')))
                  if not n.startswith('agent_'))


def agent_cases():
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz')))
                  if n.startswith('agent_'))


def walk_agent(name, agent, get_flat, set_flat, stats_of):
    """The whole-agent trajectory of an `agent_*` fixture on `agent` (oracle or CUDA drop-in in her_rng='numpy' mode, which
    consumes np.random in the reference's order): stores -> statistics, buffer fill levels; train() with sampling -> LP
    proportions, Q_loss, Q_pi; parameters after; the np.random state at the end."""
    from oracle.gen_golden_ddpg import agent_kwargs
    from tests.ddpg_util import episode_stream
    meta, z = load(name)
    case, seed = meta['case'], meta['seed']
    kw, dims, ag_ids, g_ids = agent_kwargs(case)
    sizes = {w: get_flat(w, False).size for w in ('Q', 'pi')}
    for (w, t), f in seeded_net_flats(seed, sizes, case['hidden']).items():
        set_flat(w, f, t)
    cp = np.array(meta['cp'])
    state = np.random.get_state()
    try:
        np.random.seed(seed + 2)
        n = 0
        for ep in episode_stream(dims, kw['T'], case['stores'], seed=seed + 1, flat=case['structure'] == 'flat'):
            n += 2
            agent.store_episode({k: v.copy() for k, v in ep.items()}, cp, n)
        for tag in ('o', 'g'):
            mean, std, count = stats_of(tag)
            assert np.allclose(mean, z['stats_%s_mean' % tag], rtol=1e-5, atol=1e-6), (name, tag)
            assert np.allclose(std, z['stats_%s_std' % tag], rtol=1e-5, atol=1e-6), (name, tag)
            assert float(count) == float(z['stats_%s_count' % tag][0]), (name, tag)
        bufs = agent.buffer if isinstance(agent.buffer, list) else [agent.buffer]
        assert [b.current_size for b in bufs] == list(z['buffer_sizes']), name
        for k in range(case['updates']):
            np.random.seed(seed + 10 + k)
            ql, qpi = agent.train()
            tol = 1e-5 if k == 0 else 2e-4                 # later updates: trajectory check (a ReLU flip is not rounding)
            assert abs(float(ql) - float(z['Q_loss'][k])) <= tol * abs(float(z['Q_loss'][k])) + 1e-7, (name, k, float(ql))
            assert rel_err(np.asarray(qpi, np.float64).reshape(-1), z['Q_pi'][k].reshape(-1)) <= 10 * tol, (name, k)
            if 'proportions' in z.files:
                assert np.array_equal(np.asarray(agent.proportions, np.int64), z['proportions'][k]), (name, k)
            if k % 2 == 1:
                agent.update_target_net()
        assert np.array_equal(np.random.get_state()[1], z['rng_after']), 'np.random consumed differently from the reference'
    finally:
        np.random.set_state(state)
    lr = max(kw['Q_lr'], kw['pi_lr'])
    for w in ('Q', 'pi'):
        for tgt in (False, True):
            got = get_flat(w, tgt)[::meta['stride']]
            want = z['%s_%s_after' % ('target' if tgt else 'main', w)]
            d = np.abs(got.astype(np.float64) - want)
            far = (d > 1e-5 * np.abs(want).max()).mean()
            assert d.max() <= 2.0 * case['updates'] * lr + 1e-6 and far <= 0.02 and d.mean() <= 0.02 * lr * case['updates'], \
                (name, w, tgt, d.max(), far, d.mean())
    return meta, z


def load(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return json.loads(str(z['meta'])), z


def case_kwargs(case):
    from oracle.gen_golden_ddpg import case_kwargs as ck
    return ck(case)


def action_inputs(seed, dims, n):
    from oracle.gen_golden_ddpg import action_inputs as ai
    return ai(seed, dims, n)


class Adapter(object):
    """What the walk needs from an agent; two implementations (oracle, CUDA drop-in)."""

    def set_flat(self, which, flat, target): raise NotImplementedError
    def get_flat(self, which, target): raise NotImplementedError
    def set_stats(self, which, arrays): raise NotImplementedError
    def grads(self, batch): raise NotImplementedError        # -> dict(Q_loss, pi_loss, Q_pi, Q_grad, pi_grad)
    def apply(self): raise NotImplementedError               # Adam step with the gradients of the last grads()


def walk(name, adapter, agent):
    """Seeded parameters / statistics in, then every update of the fixture: losses, Q_pi, first gradients, final parameters,
    get_actions of both networks."""
    meta, z = load(name)
    case, seed = meta['case'], meta['seed']
    kw, dims, ag_ids, g_ids = case_kwargs(case)
    sizes = {w: adapter.get_flat(w, False).size for w in ('Q', 'pi')}
    for (w, t), f in seeded_net_flats(seed, sizes, case['hidden']).items():
        adapter.set_flat(w, f, t)
    adapter.set_stats('o', seeded_stats(seed + 100, dims['o']))
    adapter.set_stats('g', seeded_stats(seed + 101, dims['g']))
    worst = dict(loss=0.0, grad=0.0, kink=False)
    for k in range(case['updates']):
        batch = seeded_batch(seed + 1000 + k, meta['stage_keys'], dims, case['batch'])
        out = adapter.grads(batch)
        # a ReLU unit that flips in update 0 moves the later updates by far more than rounding: from update 1 on the
        # comparison is a trajectory check at 2e-4
        tol = LOSS_RTOL if k == 0 else 2e-4
        for key in ('Q_loss', 'pi_loss'):
            err = abs(float(out[key]) - float(z[key][k])) / max(abs(float(z[key][k])), 1e-30)
            worst['loss'] = max(worst['loss'], err) if k == 0 else worst['loss']
            assert err <= tol, (name, key, k, float(out[key]), float(z[key][k]))
        assert rel_err(out['Q_pi'], z['Q_pi'][k]) <= (1e-5 if k == 0 else 2e-4), (name, 'Q_pi', k)
        if k == 0:
            margin = float(z['relu_margin0'])
            for key in ('Q_grad', 'pi_grad'):
                err = rel_err(out[key], z[key + '0'])
                worst['grad'] = max(worst['grad'], err)
                if err > GRAD_RTOL:                          # a hidden pre-activation within rounding of the ReLU kink
                    assert margin <= 2e-6 and err <= 5e-4, (name, key, err, margin)
                    worst['kink'] = True
        adapter.apply()
        if k % 2 == 1:
            agent.update_target_net()
    # parameters after all updates.  The first Adam steps are ~lr * g / (|g| + 3e-7): an element whose gradient is as small as
    # the float32 rounding of the batch sums moves by a different fraction of lr in any re-implementation, so single elements
    # may differ by up to updates * lr.  A wrong rule (step size, moments, polyak) shifts EVERY element by O(lr): bounded are
    # the largest difference, the share of elements beyond 1e-5 of the scale, and the mean difference (<= 2 % of a step).
    stride = meta['stride']
    lr = max(kw['Q_lr'], kw['pi_lr'])
    for w in ('Q', 'pi'):
        for tgt in (False, True):
            got = adapter.get_flat(w, tgt)[::stride]
            want = z['%s_%s_after' % ('target' if tgt else 'main', w)]
            d = np.abs(got.astype(np.float64) - want)
            far = (d > 1e-5 * np.abs(want).max()).mean()
            assert d.max() <= 2.0 * case['updates'] * lr + 1e-6, (name, w, tgt, d.max())
            assert far <= 0.02 and d.mean() <= 0.02 * lr * case['updates'], (name, w, tgt, far, d.mean())
            worst['far'] = max(worst.get('far', 0.0), float(far))
    return meta, z, worst


def check_actions(name, agent, z, dims, seed, rtol=1e-4):
    """get_actions(compute_Q=True), no exploration noise, after the updates of the walk (parameters differ from the
    reference's at the 1e-5 level, hence 1e-4)."""
    o, ag, g, td = action_inputs(seed + 2000, dims, 7)
    for tgt in (False, True):
        u, q = agent.get_actions(o, ag, g, task_descr=td, use_target_net=tgt, compute_Q=True)
        tag = 'target' if tgt else 'main'
        assert np.abs(np.asarray(u, np.float64) - z['act_u_' + tag]).max() <= rtol, (name, tag)
        assert rel_err(np.asarray(q, np.float64).reshape(-1), z['act_q_' + tag].reshape(-1)) <= rtol, (name, tag)
