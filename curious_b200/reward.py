"""Reward specification consumed by the fused HER kernel.

The reference hands the sampler a Python callable `reward_fun(ag_2, g, task_descr, info)` that
forwards to gym_flowers' `compute_reward` (reference baselines/her/experiment/config.py:158-159).
A Python callable cannot run inside a CUDA kernel, so the drop-in takes the rule as DATA: one row of a
table per module (`cur_task_table`, include/curious_b200.h) holding the achieved-goal columns, the goal
columns, a threshold and a rule kind:

    'distance'  r = -1 if ||ag_2[ag_id_m] - g[g_id_m]||_2 > threshold_m else 0      (the restated default)
    'pair'      r = -1 if ||(ag_2[ag_id_m] - ag_2[ref_ag_id_m]) - g[g_id_m]||_2 > threshold_m else 0
                the goal is the wanted OFFSET between two achieved-goal sub-slices (one object placed
                relative to another, the Stack-style module the survey names)
    'info'      r = info[info_key_m] - 1: the success flag stored with the transition passes through

with m = the module of `task_descr`, evaluated in float64 inside the kernel (csrc/her_device.cuh).
gym_flowers is absent from the reference tree and unpinned: every rule here is a RESTATEMENT (parity
unpinned); a different rule is a table change, not a kernel change.

`as_reward_spec` accepts, in this order: a table object; anything carrying one as `.reward_spec`; or the
reference's own closure from `configure_her` (config.py:158-159) - the environment is dug out of the
closure and the table is built from its attributes (`spec_from_env`), so `configure_her` needs no edit.
"""
import numpy as np

from . import _lib

KINDS = {'distance': _lib.REWARD_DISTANCE, 'pair': _lib.REWARD_PAIR, 'info': _lib.REWARD_INFO}


class ModuleRewardTable:
    """`reward_fun` for make_sample_*_her_transitions: one reward rule per module.

    threshold     scalar or one per module
    kinds         None (all 'distance') or one of 'distance' | 'pair' | 'info' per module
    ref_ag_id     per module the second achieved-goal slice of a 'pair' rule (None elsewhere)
    info_keys     per module the info key (without the `info_` prefix) of an 'info' rule (None elsewhere)
    flat_threshold   threshold of the flat sampler's single distance over all modules (task_descr=None, her.py:56-59)
    """

    def __init__(self, tasks_ag_id=None, tasks_g_id=None, threshold=0.05, kinds=None, ref_ag_id=None, info_keys=None,
                 flat_threshold=None):
        self.tasks_ag_id = tasks_ag_id
        self.tasks_g_id = tasks_g_id
        self.threshold = threshold if np.ndim(threshold) else float(threshold)
        self.kinds = list(kinds) if kinds is not None else None
        self.ref_ag_id = ref_ag_id
        self.info_keys = list(info_keys) if info_keys is not None else None
        self.flat_threshold = flat_threshold
        if self.kinds is not None:
            for k in self.kinds:
                if k not in KINDS:
                    raise ValueError('unknown reward kind %r (one of %s)' % (k, sorted(KINDS)))

    def needs_info(self):
        return self.kinds is not None and 'info' in self.kinds

    def task_table(self, info_layout=None):
        """The C struct.  `info_layout`: [(key, dim)] of the buffers' concatenated info row (replay_buffer.info_keys),
        needed to resolve the columns of 'info' rules."""
        n = len(self.tasks_g_id) if self.tasks_g_id is not None else 0
        kinds = [KINDS[k] for k in self.kinds] if self.kinds is not None else None
        cols = None
        if self.needs_info():
            cols = [None] * n
            offs, off = {}, 0
            for key, dim in (info_layout or []):
                offs[key[5:] if key.startswith('info_') else key] = off
                off += dim
            for m in range(n):
                if self.kinds[m] == 'info':
                    key = self.info_keys[m] if self.info_keys is not None and self.info_keys[m] is not None else 'is_success'
                    if key not in offs:
                        if info_layout is None:
                            cols[m] = 0          # resolved when the sampler first sees a buffer
                            continue
                        raise KeyError('reward rule of module %d reads info[%r], which the replay buffers do not store' % (m, key))
                    cols[m] = offs[key]
        return _lib.make_task_table(self.tasks_ag_id, self.tasks_g_id, self.threshold, kinds=kinds,
                                    ref_ag_id=self.ref_ag_id, info_cols=cols, flat_threshold=self.flat_threshold)

    def __call__(self, *a, **kw):
        raise RuntimeError('the reward table is evaluated inside the fused CUDA HER kernel; '
                           'curious_b200 has no host reward path (no CPU fallback).')


class ModuleDistanceReward(ModuleRewardTable):
    """The restated default: every module is 'distance' with one threshold."""
    kind = 0

    def __init__(self, tasks_ag_id=None, tasks_g_id=None, threshold=0.05):
        super().__init__(tasks_ag_id, tasks_g_id, threshold)


def spec_from_env(env):
    """Build the table from the gym_flowers attribute contract the reference reads (config.py:117-122, 158-159):
    `tasks_ag_id`, `tasks_g_id` and - gym robotics convention - `distance_threshold` (scalar or per module).  An
    environment whose modules follow other rules states them as `reward_kinds` / `reward_ref_ag_id` /
    `reward_info_keys` (per-module lists), or hands over a finished table as `reward_spec`."""
    env = getattr(env, 'unwrapped', env)
    spec = getattr(env, 'reward_spec', None)
    if isinstance(spec, ModuleRewardTable):
        return spec
    if not hasattr(env, 'tasks_g_id'):
        raise TypeError('%r does not expose tasks_g_id / tasks_ag_id' % (env,))
    return ModuleRewardTable(env.tasks_ag_id, env.tasks_g_id, getattr(env, 'distance_threshold', 0.05),
                             kinds=getattr(env, 'reward_kinds', None), ref_ag_id=getattr(env, 'reward_ref_ag_id', None),
                             info_keys=getattr(env, 'reward_info_keys', None),
                             flat_threshold=getattr(env, 'flat_distance_threshold', None))


def _env_of_closure(fun):
    """The environment a `reward_fun` closure forwards to (config.py:154-159: `env.unwrapped.compute_reward(...)`)."""
    owner = getattr(fun, '__self__', None)                 # a bound env.compute_reward
    cells = [owner] if owner is not None else []
    for cell in (getattr(fun, '__closure__', None) or ()):
        try:
            cells.append(cell.cell_contents)
        except ValueError:
            pass
    for obj in cells:
        for cand in (obj, getattr(obj, 'unwrapped', None)):
            if cand is not None and hasattr(cand, 'tasks_g_id') and hasattr(cand, 'compute_reward'):
                return cand
    return None


def as_reward_spec(reward_fun, tasks_ag_id, tasks_g_id):
    if isinstance(reward_fun, ModuleRewardTable):
        spec = reward_fun
    elif isinstance(getattr(reward_fun, 'reward_spec', None), ModuleRewardTable):
        spec = reward_fun.reward_spec
    else:
        env = _env_of_closure(reward_fun) if callable(reward_fun) else None
        if env is None:
            raise TypeError(
                'reward_fun must be a curious_b200.reward.ModuleRewardTable, carry one as `.reward_spec`, or be the '
                'closure over an environment that exposes tasks_g_id / tasks_ag_id / compute_reward (config.py:158-159): '
                'the reward is recomputed inside the CUDA kernel, an arbitrary Python callable cannot be used there.')
        spec = spec_from_env(env)
    return ModuleRewardTable(spec.tasks_ag_id if spec.tasks_ag_id is not None else tasks_ag_id,
                             spec.tasks_g_id if spec.tasks_g_id is not None else tasks_g_id,
                             spec.threshold, kinds=spec.kinds, ref_ag_id=spec.ref_ag_id, info_keys=spec.info_keys,
                             flat_threshold=spec.flat_threshold)
