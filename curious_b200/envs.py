"""A synthetic modular multi-goal environment with the gym_flowers attribute contract.

gym_flowers (MultiTaskFetchArm*-v5) and MuJoCo are not available here, so the driver loop (train.py / rollout.py of the
reference) is exercised on a NumPy stand-in that exposes exactly what the reference reads from its environments
(config.py:113-123,158-159,259-268; rollout.py:85-86,128-146,276-284,319-321; experiment/test_env.py:8-26):

  nb_tasks, tasks_g_id, tasks_ag_id, _max_episode_steps, action_space / observation layout,
  reset() -> {'observation','achieved_goal','desired_goal','mask'}, step(u) -> (obs, reward, done, info{'is_success'}),
  reset_task_goal(goal, task, directly=False, eval=False), _compute_goal(goal, task, eval=False), compute_reward(...),
  set_flat_env(), unwrapped, task, goal, last_obs, seed().

World: a point "gripper" in a 0.3 m cube and N - 1 objects.  The achieved goal is [gripper | object 1 | ... | object N-1]
(3 numbers per module, like the consecutive triples of the reference's goal layout).
  module 0            reach: bring the gripper to the target
  modules 1..n_ctrl-1 carry: an object follows the gripper while it is within `grasp_radius` and the grip action is > 0
                      (object 1 starts next to the gripper, the others anywhere: easy / hard variants)
  modules >= n_ctrl   distractors: objects that move on their own (random walk) - unlearnable, zero learning progress
                      (the role of the 4 distractor modules of MultiTaskFetchArm8, ddpg.py:104-110)
Reward: -1 if ||ag[module slice] - g[module slice]|| > 0.05 else 0 - the rule curious_b200.reward.ModuleDistanceReward
hands to the fused kernel (`reward_spec`).
"""
import numpy as np

from .reward import ModuleDistanceReward


class _Box(object):
    def __init__(self, low, high, shape):
        self.low, self.high, self.shape = low, high, shape

    def sample(self):
        return np.random.uniform(self.low, self.high, self.shape)


class ModularPointEnv(object):
    distance_threshold = 0.05
    half_extent = 0.15
    step_size = 0.03
    grasp_radius = 0.04

    def __init__(self, nb_tasks=4, n_controllable=None, max_episode_steps=50):
        self.nb_tasks = nb_tasks
        self.n_ctrl = nb_tasks if n_controllable is None else n_controllable
        self.tasks_g_id = [[3 * j, 3 * j + 1, 3 * j + 2] for j in range(nb_tasks)]
        self.tasks_ag_id = [[3 * j, 3 * j + 1, 3 * j + 2] for j in range(nb_tasks)]
        self._max_episode_steps = max_episode_steps
        self.action_space = _Box(-1.0, 1.0, (4,))
        self.dim_o = 6 + 6 * (nb_tasks - 1)          # gripper pos + vel, per object: pos + (pos - gripper)
        self.dim_g = 3 * nb_tasks
        self.reward_spec = ModuleDistanceReward(self.tasks_ag_id, self.tasks_g_id, threshold=self.distance_threshold)
        self.flat = False
        self.rng = np.random.RandomState(0)
        self.task = 0
        self.goal = np.zeros(self.dim_g)
        self.last_obs = None
        self.reset()

    # gym plumbing ---------------------------------------------------------------------------------------------
    @property
    def unwrapped(self):
        return self

    def seed(self, seed):
        self.rng = np.random.RandomState(seed)

    def set_flat_env(self):
        self.flat = True

    # dynamics -------------------------------------------------------------------------------------------------
    def _ag(self):
        return np.concatenate([self.grip] + [o for o in self.objs]).astype(np.float64)

    def _obs(self):
        parts = [self.grip, self.vel]
        for o in self.objs:
            parts += [o, o - self.grip]
        mask = np.zeros(self.nb_tasks)
        mask[self.task] = 1
        obs = dict(observation=np.concatenate(parts), achieved_goal=self._ag(), desired_goal=self.goal.copy(), mask=mask)
        self.last_obs = obs['observation'].copy()
        return obs

    def reset(self):
        h = self.half_extent
        self.grip = self.rng.uniform(-h, h, 3)
        self.vel = np.zeros(3)
        self.objs = [self.rng.uniform(-h, h, 3) for _ in range(self.nb_tasks - 1)]
        if self.nb_tasks > 1 and self.n_ctrl > 1:
            # object 1 starts in the gripper's reach (carrying it only needs the grip action), the other
            # controllable objects anywhere (they must be fetched first): a curriculum for the learning progress
            self.objs[0] = np.clip(self.grip + self.rng.uniform(-0.02, 0.02, 3), -h, h)
        return self._obs()

    def _compute_goal(self, goal, task, eval=False):
        """goal in [-1, 1]^3 for `task` -> (full goal vector with only the module's slice set, mask) like the
        reference's env (rollout.py:87-88,135: `_compute_goal(...)[0][tasks_g_id[task]]`)."""
        goal = np.clip(np.asarray(goal, np.float64), -1, 1) * self.half_extent
        if self.flat:                        # flat structure: one goal over all modules at once (rollout.py:148-152)
            return goal.copy(), np.ones(self.nb_tasks)
        full = np.zeros(self.dim_g)
        full[self.tasks_g_id[task]] = goal
        mask = np.zeros(self.nb_tasks)
        mask[task] = 1
        return full, mask

    def reset_task_goal(self, goal, task=0, directly=False, eval=False):
        self.task = int(task)
        if directly:
            self.goal = np.zeros(self.dim_g)
            self.goal[self.tasks_g_id[self.task]] = goal
        else:
            self.goal = self._compute_goal(goal, self.task, eval)[0]
        return self._obs()

    def compute_reward(self, achieved_goal, goal, task_descr=None, info=None):
        """Vectorised sparse reward, [B, 1] (config.py:158-159, ddpg.py:342).  Host version for rollouts / tests - the
        training path evaluates the same rule inside the fused kernel from `reward_spec`."""
        ag = np.atleast_2d(achieved_goal)
        g = np.atleast_2d(goal)
        if self.flat:                        # distance over every goal coordinate (the kernel's rule without task_descr)
            return np.where(np.linalg.norm(ag - g, axis=1) > self.distance_threshold, -1.0, 0.0).reshape(-1, 1)
        td = np.atleast_2d(task_descr) if task_descr is not None else np.eye(self.nb_tasks)[[self.task] * len(ag)]
        r = np.zeros((len(ag), 1))
        for i in range(len(ag)):
            m = int(np.argmax(td[i]))
            d = np.linalg.norm(ag[i, self.tasks_ag_id[m][:len(self.tasks_g_id[m])]] - g[i, self.tasks_g_id[m]])
            r[i, 0] = -1.0 if d > self.distance_threshold else 0.0
        return r

    def step(self, action):
        u = np.clip(np.asarray(action, np.float64).reshape(-1), -1, 1)
        h = self.half_extent
        new = np.clip(self.grip + self.step_size * u[:3], -h, h)
        delta = new - self.grip
        for j, o in enumerate(self.objs):
            module = j + 1
            if module < self.n_ctrl:
                if u[3] > 0 and np.linalg.norm(o - self.grip) < self.grasp_radius:
                    self.objs[j] = np.clip(o + delta, -h, h)                 # carried along
            else:
                self.objs[j] = np.clip(o + self.rng.normal(0, 0.01, 3), -h, h)   # distractor: moves on its own
        self.vel = delta
        self.grip = new
        obs = self._obs()
        mask = obs['mask']
        r = float(self.compute_reward(obs['achieved_goal'], self.goal, mask)[0, 0])
        info = dict(is_success=float(r == 0.0))
        return obs, r, False, info
