"""Restated sparse module reward (test infrastructure; see oracle/__init__.py).

PARITY UNPINNED.  The real function is `gym_flowers` `...compute_reward(achieved_goal, goal,
task_descr, info)` reached through `reward_fun` in reference
baselines/her/experiment/config.py:158-159.  gym_flowers (github.com/flowersteam/gym_flowers,
named in reference readme.md:7) is not under /root/reference and no version is pinned.

What the reference's own code pins about it, and what this restatement keeps:
  * called with keyword arguments ag_2, g, task_descr, info (her.py:56-59, 174-176);
  * vectorised over rows, returns shape (B, 1) (ddpg.py:82 stage shape, ddpg.py:342);
  * `task_descr` is one-hot per row and selects the module (test_env.py:22-25);
  * the goal is laid out as per-module slices `tasks_g_id[m]`, compared with the achieved-goal
    slice `tasks_ag_id[m][:len(tasks_g_id[m])]` (her.py:145-148, ddpg.py:181);
  * rewards live in {-1, 0} (the target is clipped to [-clip_return, 0], ddpg.py:437-438).

Restated rule (the standard gym robotics sparse reward, per module m = argmax(task_descr)):
    d = sqrt(sum_k (ag_2[ag_id_m[k]] - g[g_id_m[k]])**2)      in float64, sequential sum, no FMA
    r = -1.0 if d > threshold else 0.0
With task_descr=None (flat sampler, her.py:58) every module slice is compared at once.
"""
import numpy as np


class ModuleDistanceReward:
    """CPU restatement of the reward contract.  `kind` 0 = module L2 distance vs threshold."""

    def __init__(self, tasks_ag_id, tasks_g_id, threshold=0.05):
        self.tasks_g_id = [list(x) for x in tasks_g_id]
        self.tasks_ag_id = [list(a)[:len(g)] for a, g in zip(tasks_ag_id, tasks_g_id)]
        self.threshold = float(threshold)
        self.n_calls = 0
        self.last_kwargs = None

    def _dist(self, ag, g, ag_idx, g_idx):
        d2 = np.zeros(ag.shape[0], np.float64)
        for ka, kg in zip(ag_idx, g_idx):
            diff = ag[:, ka].astype(np.float64) - g[:, kg].astype(np.float64)
            d2 = d2 + diff * diff
        return np.sqrt(d2)

    def __call__(self, ag_2, g, task_descr, info):
        self.n_calls += 1
        self.last_kwargs = dict(ag_2=ag_2, g=g, task_descr=task_descr, info=info)
        ag_2 = np.asarray(ag_2)
        g = np.asarray(g)
        B = g.shape[0]
        r = np.zeros((B, 1), np.float64)
        if task_descr is None:
            ag_idx = sum(self.tasks_ag_id, [])
            g_idx = sum(self.tasks_g_id, [])
            d = self._dist(ag_2, g, ag_idx, g_idx)
            r[:, 0] = np.where(d > self.threshold, -1.0, 0.0)
            return r
        module = np.argmax(np.asarray(task_descr), axis=1)
        for m in range(len(self.tasks_g_id)):
            rows = np.where(module == m)[0]
            if rows.size == 0:
                continue
            d = self._dist(ag_2[rows], g[rows], self.tasks_ag_id[m], self.tasks_g_id[m])
            r[rows, 0] = np.where(d > self.threshold, -1.0, 0.0)
        return r
