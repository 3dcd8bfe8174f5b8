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
    hist = train(n_epochs=12, **exp)
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
def test_other_structures_run_the_same_loop(structure, task_replay):
    from curious_b200.train import make_experiment, train
    np.random.seed(1)
    exp = make_experiment(nb_tasks=3, structure=structure, task_selection='active_competence_progress',
                          task_replay=task_replay, buffer_size=50000, n_cycles=4, n_batches=10, n_test_rollouts=2, seed=1)
    hist = train(n_epochs=3, **exp)
    assert len(hist) == 3 and all(0.0 <= h['test_success_rate'] <= 1.0 for h in hist)
    pols = exp['policy'] if isinstance(exp['policy'], list) else [exp['policy']]
    assert all(np.isfinite(p.get_flat('pi')).all() for p in pols)
