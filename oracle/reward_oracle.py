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

`ModuleRewardTable` restates the other rows of the product's reward table (curious_b200/reward.py; the survey's
`(ag_idx[], g_idx[], threshold, kind)` per module) - equally unpinned:
    'pair'   d = || (ag_2[ag_id_m] - ag_2[ref_ag_id_m]) - g[g_id_m] ||   the goal is an OFFSET between two objects
    'info'   r = info[info_key_m] - 1                                     the stored success flag passes through
with one threshold per module.
"""
import numpy as np


class ModuleRewardTable:
    """CPU restatement of the reward contract, one rule per module."""

    def __init__(self, tasks_ag_id, tasks_g_id, threshold=0.05, kinds=None, ref_ag_id=None, info_keys=None,
                 flat_threshold=None):
        n = len(tasks_g_id)
        self.tasks_g_id = [list(x) for x in tasks_g_id]
        self.tasks_ag_id = [list(a)[:len(g)] for a, g in zip(tasks_ag_id, tasks_g_id)]
        self.thresholds = [float(t) for t in np.broadcast_to(np.asarray(threshold, np.float64), (n,))]
        self.threshold = self.thresholds[0] if n else float(threshold)
        self.kinds = list(kinds) if kinds is not None else ['distance'] * n
        self.ref_ag_id = [None if r is None else list(r)[:len(g)] for r, g in
                          zip(ref_ag_id if ref_ag_id is not None else [None] * n, tasks_g_id)]
        self.info_keys = list(info_keys) if info_keys is not None else [None] * n
        self.flat_threshold = float(flat_threshold) if flat_threshold is not None else self.threshold
        self.n_calls = 0
        self.last_kwargs = None

    def _d2(self, m, ag, g, d2):
        """running float64 sum of squared differences of module m, sequential, no FMA"""
        if self.kinds[m] == 'info':
            return d2
        for k, (ka, kg) in enumerate(zip(self.tasks_ag_id[m], self.tasks_g_id[m])):
            a = ag[:, ka].astype(np.float64)
            if self.kinds[m] == 'pair':
                a = a - ag[:, self.ref_ag_id[m][k]].astype(np.float64)
            diff = a - g[:, kg].astype(np.float64)
            d2 = d2 + diff * diff
        return d2

    def __call__(self, ag_2, g, task_descr, info):
        self.n_calls += 1
        self.last_kwargs = dict(ag_2=ag_2, g=g, task_descr=task_descr, info=info)
        ag_2 = np.asarray(ag_2)
        g = np.asarray(g)
        B = g.shape[0]
        r = np.zeros((B, 1), np.float64)
        if task_descr is None:
            d2 = np.zeros(B, np.float64)
            for m in range(len(self.tasks_g_id)):
                d2 = self._d2(m, ag_2, g, d2)
            r[:, 0] = np.where(np.sqrt(d2) > self.flat_threshold, -1.0, 0.0)
            return r
        module = np.argmax(np.asarray(task_descr), axis=1)
        for m in range(len(self.tasks_g_id)):
            rows = np.where(module == m)[0]
            if rows.size == 0:
                continue
            if self.kinds[m] == 'info':
                key = self.info_keys[m] or 'is_success'
                r[rows, 0] = np.asarray(info[key], np.float64).reshape(B, -1)[rows, 0] - 1.0
                continue
            d = np.sqrt(self._d2(m, ag_2[rows], g[rows], np.zeros(rows.size, np.float64)))
            r[rows, 0] = np.where(d > self.thresholds[m], -1.0, 0.0)
        return r


class ModuleDistanceReward(ModuleRewardTable):
    """`kind` 0 everywhere = module L2 distance vs one threshold (the restated default)."""

    def __init__(self, tasks_ag_id, tasks_g_id, threshold=0.05):
        super().__init__(tasks_ag_id, tasks_g_id, threshold)
