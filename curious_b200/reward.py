"""Reward specification consumed by the fused HER kernel.

The reference hands the sampler a Python callable `reward_fun(ag_2, g, task_descr, info)` that
forwards to gym_flowers' `compute_reward` (reference baselines/her/experiment/config.py:158-159).
A Python callable cannot run inside a CUDA kernel, so the drop-in takes the rule as DATA: per module
the achieved-goal columns, the goal columns and a distance threshold ("reward").
gym_flowers is absent from the reference tree and unpinned; the rule below is the restated default
    r = -1 if ||ag_2[ag_id_m] - g[g_id_m]||_2 > threshold else 0,   m = module of task_descr
evaluated in float64 inside the kernel (curious_b200/csrc/her.cu, phase C).
"""


class ModuleDistanceReward:
    """`reward_fun` for make_sample_*_her_transitions.  kind 0 = module L2 distance vs threshold."""
    kind = 0

    def __init__(self, tasks_ag_id=None, tasks_g_id=None, threshold=0.05):
        self.tasks_ag_id = tasks_ag_id
        self.tasks_g_id = tasks_g_id
        self.threshold = float(threshold)

    def __call__(self, *a, **kw):
        raise RuntimeError('ModuleDistanceReward is evaluated inside the fused CUDA HER kernel; '
                           'curious_b200 has no host reward path (no CPU fallback).')


def as_reward_spec(reward_fun, tasks_ag_id, tasks_g_id):
    if isinstance(reward_fun, ModuleDistanceReward):
        spec = reward_fun
    elif hasattr(reward_fun, 'reward_spec'):
        spec = reward_fun.reward_spec
    else:
        raise TypeError(
            'reward_fun must be a curious_b200.reward.ModuleDistanceReward (or carry one as '
            '`.reward_spec`): the reward is recomputed inside the CUDA kernel, an arbitrary Python '
            'callable cannot be used there.')
    return ModuleDistanceReward(spec.tasks_ag_id if spec.tasks_ag_id is not None else tasks_ag_id,
                                spec.tasks_g_id if spec.tasks_g_id is not None else tasks_g_id,
                                spec.threshold)
