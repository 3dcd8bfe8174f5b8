"""End-to-end driver loop on the synthetic modular environment: RolloutWorker -> DDPG.get_actions ->
store_episode -> train (CUDA graph) -> update_target_net -> CompetenceTracker -> cp / p, mirroring the reference's
train.py:125-166 / rollout.py.  The check is behavioural: the agent must actually learn the reachable module and the
learning-progress pipeline must react to it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_curious_agent_learns_reach_and_lp_follows():
    from curious_b200.train import make_experiment, train
    np.random.seed(0)
    exp = make_experiment(nb_tasks=4, n_controllable=3, structure='curious', task_selection='active_competence_progress',
                          task_replay='replay_task_cp_buffer', buffer_size=100000, n_cycles=10, n_batches=40,
                          n_test_rollouts=10, seed=0)
    # (per-epoch competence of the evaluator and no epoch -1 evaluation: the sharper signal for this behavioural check)
    hist = train(n_epochs=12, initial_evaluation=False, clear_eval_competence=True, **exp)
    assert len(hist) == 12
    C = np.array([h['C'] for h in hist])
    # module 0 (reach) is learned, the distractor (module 3: the object moves on its own) is not
    assert C[-3:, 0].mean() >= 0.8, C[:, 0]
    assert C[-3:, 3].mean() <= 0.3, C[:, 3]
    # learning progress showed up on the learned module and moved the sampling probabilities towards it
    CP = np.array([h['CP'] for h in hist])
    P = np.array([h['p'] for h in hist])
    assert CP[:, 0].max() > 0 and np.allclose(P.sum(1), 1.0)
    assert P[np.argmax(CP[:, 0]), 0] > 0.25
    # the replay buffers were filled through the per-module routing (buffer 0 stays empty, ddpg.py:191-195)
    pol = exp['policy']
    assert pol.buffer[0].current_size == 0 and pol.buffer[1].current_size > 0
    assert np.isfinite(pol.get_flat('Q')).all()


@pytest.mark.parametrize('structure,task_replay', [('task_experts', 'replay_current_task_buffer'), ('flat', '')])
def test_other_structures_run_the_same_loop(structure, task_replay, tmp_path):
    from curious_b200.train import make_experiment, train
    np.random.seed(1)
    exp = make_experiment(nb_tasks=3, structure=structure, task_selection='active_competence_progress',
                          task_replay=task_replay, buffer_size=50000, n_cycles=4, n_batches=10, n_test_rollouts=2, seed=1,
                          policy_kwargs=dict(action_noise='device') if structure == 'flat' else None)
    hist = train(n_epochs=3, logdir=str(tmp_path), policy_save_interval=2, checkpoint_interval=2, initial_evaluation=False,
                 clear_eval_competence=True, **exp)
    _check_run_records(str(tmp_path), structure, exp)
    assert len(hist) == 3 and all(0.0 <= h['test_success_rate'] <= 1.0 for h in hist)
    pols = exp['policy'] if isinstance(exp['policy'], list) else [exp['policy']]
    assert all(np.isfinite(p.get_flat('pi')).all() for p in pols)


def _check_run_records(logdir, structure, exp):
    """The files the reference's train loop leaves behind (train.py:53-55,171-206,264-266), read the way its analysis
    scripts read them (analysis/plot.py:29-75)."""
    import json
    import os
    import pickle
    import pandas as pd
    data = pd.read_csv(os.path.join(logdir, 'progress.csv'))
    assert list(data['epoch']) == [0, 1, 2]
    for col in ('test/success_rate', 'test/mean_Q', 'train/success_rate', 'train/episode', 'stats_o/mean', 'stats_g/std',
                'Time'):
        assert col in data.columns and np.isfinite(data[col]).all(), col
    if structure == 'task_experts':
        for col in ('IND_TASK_rollout', 'train/C_task0', 'train/CP_task2', 'train/%_task1', 'train/p_task0',
                    'test/C_task2'):
            assert col in data.columns, col
    params = json.load(open(os.path.join(logdir, 'params.json')))
    assert params['structure'] == structure and params['n_cycles'] == 4 and params['num_cpu'] == 1
    for name in ('policy_best.pkl', 'policy_latest.pkl', 'policy_0.pkl', 'policy_2.pkl'):
        assert os.path.exists(os.path.join(logdir, name)), name
    assert not os.path.exists(os.path.join(logdir, 'policy_1.pkl'))
    # the pickled policy plays (experiment/play.py:30-33): same actions as the live one
    with open(os.path.join(logdir, 'policy_latest.pkl'), 'rb') as f:
        loaded = pickle.load(f)
    live = exp['policy']
    loaded0, live0 = (loaded[0], live[0]) if isinstance(live, list) else (loaded, live)
    d = live0.input_dims
    rng = np.random.RandomState(3)
    o, ag, g = rng.randn(5, d['o']), rng.randn(5, d['ag']), rng.randn(5, d['g'])
    td = np.eye(d['task_descr'])[rng.randint(0, d['task_descr'], 5)] if 'task_descr' in d else None
    assert np.array_equal(loaded0.get_actions(o, ag, g, task_descr=td), live0.get_actions(o, ag, g, task_descr=td))
    assert os.path.exists(os.path.join(logdir, 'checkpoint_0.pt'))


def test_resumed_training_equals_uninterrupted(tmp_path):
    """train(resume=True) with the real agent: 2 epochs, then everything rebuilt from scratch (other seeds) and resumed from
    the files for 2 more == 4 epochs in one go - evaluation results, competence, probabilities and parameters bit for bit."""
    from curious_b200.train import make_experiment, train

    def run(logdir, n_epochs, seed, resume=False):
        np.random.seed(seed)
        exp = make_experiment(nb_tasks=3, structure='curious', task_selection='active_competence_progress',
                              task_replay='replay_task_cp_buffer', buffer_size=2000, n_cycles=3, n_batches=8,
                              n_test_rollouts=2, seed=4)
        hist = train(n_epochs=n_epochs, logdir=logdir, policy_save_interval=0, checkpoint_interval=1, resume=resume,
                     initial_evaluation=False, clear_eval_competence=True, **exp)
        return hist, exp['policy']

    full, pol_full = run(str(tmp_path / 'full'), 4, seed=3)
    run(str(tmp_path / 'split'), 2, seed=3)
    tail, pol_tail = run(str(tmp_path / 'split'), 4, seed=99, resume=True)
    assert [h['epoch'] for h in tail] == [2, 3]
    for a, b in zip(full[2:], tail):
        for key in a:
            assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), (a['epoch'], key)
    for which in ('Q', 'pi'):
        for tgt in (False, True):
            assert np.array_equal(pol_full.get_flat(which, tgt), pol_tail.get_flat(which, tgt)), (which, tgt)
    assert [b.current_size for b in pol_full.buffer] == [b.current_size for b in pol_tail.buffer]


def test_make_experiment_applies_the_reference_rank_seeding(monkeypatch):
    """train.py:241-243,328-335: rank_seed = seed + 1000000 * rank seeds np.random, the rollout workers (rank_seed, or
    rank_seed + i per expert) and the evaluator (rank_seed + 100); here it also keys the Philox HER draws and the
    device-side exploration noise.  Two ranks built from the same `seed` must therefore differ in every stream but start
    from the same weights (rank 0's are broadcast, ddpg.py:466)."""
    import curious_b200.train as tr
    built = {}
    for r in (0, 1):
        monkeypatch.setattr(tr, '_rank', lambda comm=None, r=r: r)
        exp = tr.make_experiment(nb_tasks=3, structure='curious', task_replay='replay_task_cp_buffer',
                                 buffer_size=5000, seed=7)
        pol, w, ev = exp['policy'], exp['rollout_worker'], exp['evaluator']
        built[r] = dict(np_draw=float(np.random.uniform()), env_draw=float(w.envs[0].rng.uniform()),
                        env1_draw=float(w.envs[1].rng.uniform()), eval_draw=float(ev.envs[0].rng.uniform()),
                        philox=int(pol.sample_transitions.seed), noise=int(pol.noise_seed), theta=pol.get_flat('Q'))
    a, b = built[0], built[1]
    assert (a['philox'], a['noise']) == (7, 7) and (b['philox'], b['noise']) == (1000007, 1000007)
    for key in ('np_draw', 'env_draw', 'env1_draw', 'eval_draw'):
        assert a[key] != b[key], key
    assert a['env_draw'] != a['env1_draw'] and a['env_draw'] != a['eval_draw']
    assert np.array_equal(a['theta'], b['theta'])
    # the same seed reproduces the streams (rank 0 again)
    monkeypatch.setattr(tr, '_rank', lambda comm=None: 0)
    exp = tr.make_experiment(nb_tasks=3, structure='curious', task_replay='replay_task_cp_buffer', buffer_size=5000, seed=7)
    assert float(np.random.uniform()) == a['np_draw']
    assert float(exp['rollout_worker'].envs[0].rng.uniform()) == a['env_draw']
    # experts: worker i is seeded with rank_seed + i and draws its own Philox counter range
    monkeypatch.setattr(tr, '_rank', lambda comm=None: 1)
    exp = tr.make_experiment(nb_tasks=3, structure='task_experts', task_replay='replay_current_task_buffer',
                             buffer_size=5000, seed=7)
    offs = [p.GRAPH_STREAM_OFFSET for p in exp['policy']]
    assert len(set(offs)) == 3 and [p.noise_seed for p in exp['policy']] == [1000007, 1000008, 1000009]
    draws = [float(w.envs[0].rng.uniform()) for w in exp['rollout_worker']]
    assert len(set(draws)) == 3
