"""Host logic of the synthetic modular environment (curious_b200/envs.py): the gym_flowers attribute contract the
reference reads (config.py:113-123,158-159,259-268; rollout.py:85-88,128-146) and the reward rule the fused kernel
evaluates from `reward_spec`."""
import numpy as np

from curious_b200.envs import ModularPointEnv
from oracle.reward_oracle import ModuleDistanceReward as OracleReward


def test_attribute_contract():
    env = ModularPointEnv(nb_tasks=4, n_controllable=3)
    assert env.unwrapped is env and env.nb_tasks == 4 and env._max_episode_steps == 50
    assert env.tasks_g_id == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10, 11]] == env.tasks_ag_id
    obs = env.reset()
    assert set(obs) == {'observation', 'achieved_goal', 'desired_goal', 'mask'}
    assert obs['observation'].shape == (24,) and obs['achieved_goal'].shape == (12,) and obs['desired_goal'].shape == (12,)
    full, mask = env._compute_goal(np.array([1.0, -1.0, 0.5]), 2)
    assert np.allclose(full[6:9], [0.15, -0.15, 0.075]) and np.count_nonzero(full) == 3 and mask.tolist() == [0, 0, 1, 0]
    obs = env.reset_task_goal(goal=np.array([0.2, 0.2, 0.2]), task=1)
    assert env.task == 1 and obs['mask'].tolist() == [0, 1, 0, 0] and np.allclose(obs['desired_goal'][3:6], 0.03)
    o2, r, done, info = env.step(env.action_space.sample())
    assert r in (-1.0, 0.0) and info['is_success'] == float(r == 0.0) and done is False
    assert (env.dim_o + env.nb_tasks) % 4 == 0            # first-layer fan-ins stay 16-byte aligned (rows schedule)


def test_dynamics_reach_carry_distractor():
    env = ModularPointEnv(nb_tasks=4, n_controllable=3)
    env.seed(3)
    env.reset()
    g0 = env.grip.copy()
    near = env.objs[0].copy()
    assert np.linalg.norm(near - g0) < env.grasp_radius          # object 1 starts within reach
    far = env.objs[1].copy()
    env.step([1.0, 0.0, 0.0, 1.0])                               # move +x with the grip closed
    assert np.allclose(env.grip - g0, [min(0.03, 0.15 - g0[0]), 0, 0])
    assert np.allclose(env.objs[0] - near, env.grip - g0)        # carried along
    assert np.allclose(env.objs[1], far)                         # out of reach: stays
    d0 = env.objs[2].copy()
    env.step([0.0, 0.0, 0.0, -1.0])
    assert not np.allclose(env.objs[2], d0)                      # the distractor moves on its own
    before = env.objs[0].copy()
    env.step([0.0, 1.0, 0.0, -1.0])                              # grip open: nothing is carried
    assert np.allclose(env.objs[0], before)


def test_reward_is_the_kernel_rule():
    rng = np.random.RandomState(0)
    env = ModularPointEnv(nb_tasks=3)
    ora = OracleReward(env.tasks_ag_id, env.tasks_g_id, threshold=env.reward_spec.threshold)
    ag = rng.uniform(-0.1, 0.1, (64, 9))
    g = ag + rng.uniform(-0.06, 0.06, (64, 9))
    td = np.eye(3)[rng.randint(0, 3, 64)]
    assert np.array_equal(env.compute_reward(ag, g, td, None), ora(ag, g, td, None))
    env.set_flat_env()
    g = ag + rng.uniform(-0.025, 0.025, (64, 9))
    assert np.array_equal(env.compute_reward(ag, g, None, None), ora(ag, g, None, None))
    assert env._compute_goal(np.ones(9), 0)[0].shape == (9,)
