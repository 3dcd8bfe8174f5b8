"""The reward table (one rule per module; SURVEY 8c `(ag_idx[], g_idx[], threshold, kind)`): oracle rules on hand-computed
cases, the C struct the product builds from a table, and the adapter that digs the environment out of the reference's
`configure_her` closure (config.py:154-159).  PARITY UNPINNED for every rule (gym_flowers is absent), see DESIGN.md 3.2."""
import ctypes as C

import numpy as np
import pytest

from oracle.reward_oracle import ModuleDistanceReward, ModuleRewardTable

AG_IDS = [[0, 1, 2], [3, 4, 5], [6, 7, 8]]
G_IDS = [[0, 1, 2], [3, 4, 5], [6, 7, 8]]


def _td(mods):
    return np.eye(3)[list(mods)]


def test_distance_rule_per_module_thresholds():
    r = ModuleRewardTable(AG_IDS, G_IDS, threshold=[0.05, 0.2, 0.01])
    ag = np.zeros((3, 9))
    g = np.zeros((3, 9))
    ag[0, 0] = 0.1            # module 0: d = 0.1 > 0.05
    ag[1, 3] = 0.1            # module 1: d = 0.1 <= 0.2
    ag[2, 8] = 0.02           # module 2: d = 0.02 > 0.01
    out = r(ag, g, _td([0, 1, 2]), {})
    assert out.shape == (3, 1) and out[:, 0].tolist() == [-1.0, 0.0, -1.0]
    # only the module's own slice counts
    ag[1, 0] = 5.0
    assert r(ag, g, _td([0, 1, 2]), {})[1, 0] == 0.0
    # the boundary: d == threshold is a success (strict compare)
    one = ModuleDistanceReward(AG_IDS, G_IDS, 0.05)
    a = np.zeros((1, 9)); a[0, 0] = 0.05
    assert one(a, np.zeros((1, 9)), _td([0]), {})[0, 0] == 0.0


def test_pair_rule_compares_an_offset_between_two_slices():
    """Stack-style: module 1's goal is where cube 1 (ag 3..5) should sit RELATIVE to cube 0 (ag 0..2)."""
    r = ModuleRewardTable(AG_IDS, G_IDS, threshold=0.05, kinds=['distance', 'pair', 'distance'],
                          ref_ag_id=[None, [0, 1, 2], None])
    ag = np.zeros((2, 9)); g = np.zeros((2, 9))
    ag[:, 0:3] = [0.3, -0.2, 0.1]
    ag[:, 3:6] = [0.3, -0.2, 0.15]            # 0.05 above cube 0
    g[0, 3:6] = [0.0, 0.0, 0.05]              # wanted offset reached (up to rounding of 0.15 - 0.1)
    g[1, 3:6] = [0.0, 0.0, 0.2]               # wanted offset not reached
    assert r(ag, g, _td([1, 1]), {})[:, 0].tolist() == [0.0, -1.0]
    # the same rows judged by module 0's plain distance rule
    assert r(ag, g, _td([0, 0]), {})[:, 0].tolist() == [-1.0, -1.0]


def test_info_rule_passes_the_stored_flag_through():
    r = ModuleRewardTable(AG_IDS, G_IDS, kinds=['info', 'distance', 'info'], info_keys=['is_success', None, 'done_flag'])
    info = {'is_success': np.array([[1.0], [0.0], [1.0]]), 'done_flag': np.array([[0.0], [0.0], [0.0]])}
    ag = np.ones((3, 9)); g = np.zeros((3, 9))
    out = r(ag, g, _td([0, 0, 2]), info)
    assert out[:, 0].tolist() == [0.0, -1.0, -1.0]


def test_flat_sampler_rule_is_one_distance_over_all_modules():
    r = ModuleRewardTable(AG_IDS, G_IDS, threshold=[0.05, 0.2, 0.01], flat_threshold=0.1)
    ag = np.zeros((2, 9)); g = np.zeros((2, 9))
    ag[0, [0, 3, 6]] = 0.05                   # sqrt(3) * 0.05 = 0.087 <= 0.1
    ag[1, [0, 3, 6]] = 0.06                   # 0.104 > 0.1
    assert r(ag, g, None, {})[:, 0].tolist() == [0.0, -1.0]


def test_product_table_struct_holds_the_rules():
    from curious_b200 import _lib
    from curious_b200.reward import ModuleRewardTable as Table
    t = Table(AG_IDS, G_IDS, threshold=[0.05, 0.2, 0.01], kinds=['distance', 'pair', 'info'],
              ref_ag_id=[None, [0, 1, 2], None], info_keys=[None, None, 'is_success'], flat_threshold=0.3)
    tt = t.task_table([('info_aux', 2), ('info_is_success', 1)])
    assert tt.n_tasks == 3 and [tt.kind[m] for m in range(3)] == [_lib.REWARD_DISTANCE, _lib.REWARD_PAIR, _lib.REWARD_INFO]
    assert [tt.threshold[m] for m in range(3)] == [0.05, 0.2, 0.01] and tt.flat_threshold == 0.3
    assert [tt.ref_idx[1][k] for k in range(3)] == [0, 1, 2] and tt.info_col[2] == 2
    with pytest.raises(KeyError):
        t.task_table([('info_aux', 2)])
    with pytest.raises(ValueError):
        Table(AG_IDS, G_IDS, kinds=['distance', 'nearest', 'info'])
    with pytest.raises(RuntimeError):
        t(None, None, None, None)            # no host reward path


class _Env:
    """the gym_flowers attribute contract the reference reads (config.py:117-122,158-159)"""
    nb_tasks = 3
    tasks_ag_id = AG_IDS
    tasks_g_id = G_IDS
    distance_threshold = 0.07

    @property
    def unwrapped(self):
        return self

    def compute_reward(self, achieved_goal, goal, task_descr, info):
        raise AssertionError('the host rule must not be called on the training path')


def test_configure_her_closure_needs_no_edit():
    """config.py:154-170 verbatim in shape: a closure over `env` forwarding to env.unwrapped.compute_reward is handed to the
    sampler factory; the factory finds the environment in the closure and builds the table from its attributes."""
    from curious_b200 import her
    env = _Env()

    def reward_fun(ag_2, g, task_descr, info):  # vectorized
        return env.unwrapped.compute_reward(achieved_goal=ag_2, goal=g, task_descr=task_descr, info=info)

    s = her.make_sample_multi_task_her_transitions(goal_replay='her', her_replay_k=4, task_replay='replay_task_cp_buffer',
                                                   reward_fun=reward_fun, tasks_ag_id=AG_IDS, tasks_g_id=G_IDS)
    assert [s.task_table.threshold[m] for m in range(3)] == [0.07] * 3 and s.task_table.n_tasks == 3

    class Stack(_Env):
        distance_threshold = [0.05, 0.03, 0.05]
        reward_kinds = ['distance', 'pair', 'distance']
        reward_ref_ag_id = [None, [0, 1, 2], None]

    env = Stack()
    s = her.make_sample_multi_task_her_transitions('her', 4, 'replay_task_cp_buffer', reward_fun, tasks_ag_id=AG_IDS,
                                                   tasks_g_id=G_IDS)
    from curious_b200 import _lib
    assert s.task_table.kind[1] == _lib.REWARD_PAIR and s.task_table.threshold[1] == 0.03
    # a bound method works as well; an opaque callable is refused (no CPU fallback)
    s = her.make_sample_multi_task_her_transitions('her', 4, 'replay_task_cp_buffer', env.compute_reward,
                                                   tasks_ag_id=AG_IDS, tasks_g_id=G_IDS)
    assert s.task_table.kind[1] == _lib.REWARD_PAIR
    with pytest.raises(TypeError):
        her.make_sample_multi_task_her_transitions('her', 4, 'replay_task_cp_buffer', lambda **kw: 0, tasks_ag_id=AG_IDS,
                                                   tasks_g_id=G_IDS)
