"""NumPy restatement of the two HER samplers (test infrastructure; see oracle/__init__.py).

Follows reference baselines/her/her.py:
  * flat sampler            her.py:5-68   (`make_sample_her_transitions`)
  * multi-task sampler      her.py:72-185 (`make_sample_multi_task_her_transitions`)

PINNED: tests/test_oracle_golden.py checks this file bit-for-bit against outputs of the
unmodified reference modules (fixtures made by oracle/gen_golden.py).

The restatement is split in two so the same draws can be injected into the CUDA kernel:
  `draw_stream`  consumes np.random in exactly the reference order
                 (randint(0,E,B); randint(T,size=B); uniform(B); uniform(B)  - her.py:108-116)
  `relabel`      gather, per-HER-row relabel loop (with the np.random.choice draws of the two
                 *_task_transition modes made inside the loop, her.py:139,142), reward.
The per-row Python loop of the reference (her.py:129-164) is kept as a loop on purpose: this
file is also the timed CPU baseline (`bench.py --impl reference`), and that loop is where the
reference spends its time.
"""
import numpy as np


class HerStream:
    """The random draws of one sampler call, in reference order."""
    __slots__ = ('ep', 't', 'u_her', 'u_off', 'task_choice')

    def __init__(self, ep, t, u_her, u_off, task_choice=None):
        self.ep = ep                # int64 [B]  her.py:108
        self.t = t                  # int64 [B]  her.py:109
        self.u_her = u_her          # float64 [B] her.py:115 (compared with future_p)
        self.u_off = u_off          # float64 [B] her.py:116 (scaled by T - t, truncated)
        self.task_choice = task_choice  # int64 [B], -1 where no np.random.choice was made


def future_probability(goal_replay, her_replay_k):
    # her.py:14-17 / her.py:86-89
    return 1 - (1. / (1 + her_replay_k)) if goal_replay == 'her' else 0


def draw_stream(E, T, B):
    ep = np.random.randint(0, E, B)
    t = np.random.randint(T, size=B)
    u_her = np.random.uniform(size=B)
    u_off = np.random.uniform(size=B)
    return HerStream(ep, t, u_her, u_off, np.full(B, -1, np.int64))


def _truncated_ids(tasks_ag_id, tasks_g_id):
    # her.py:145-148: the achieved-goal slice is cut to the length of the goal slice
    return [list(a)[:len(g)] for a, g in zip(tasks_ag_id, tasks_g_id)]


class FlatHerOracle:
    """her.py:5-68.  `task_replay` is accepted and ignored, as in the reference."""

    def __init__(self, goal_replay, her_replay_k, reward_fun, task_replay='', tasks_ag_id=None,
                 tasks_g_id=None):
        self.future_p = future_probability(goal_replay, her_replay_k)
        self.reward_fun = reward_fun
        self.ag_cols = sum(_truncated_ids(tasks_ag_id, tasks_g_id), [])   # her.py:43-46

    def __call__(self, episode_batch, batch_size_in_transitions, task_to_replay=None, cp_proba=None,
                 stream=None):
        T = episode_batch['u'].shape[1]
        E = episode_batch['u'].shape[0]
        B = batch_size_in_transitions
        if stream is None:
            stream = draw_stream(E, T, B)
        self.last_stream = stream
        ep, t = stream.ep, stream.t
        out = {k: v[ep, t].copy() for k, v in episode_batch.items()}
        her_rows = np.where(stream.u_her < self.future_p)[0]
        off = (stream.u_off * (T - t)).astype(int)
        future_t = (t + 1 + off)[her_rows]
        future_ag = episode_batch['ag'][ep[her_rows], future_t]
        out['g'][her_rows] = future_ag[:, self.ag_cols]
        info = {k[len('info_'):]: v for k, v in out.items() if k.startswith('info_')}
        out['r'] = self.reward_fun(ag_2=out['ag_2'], g=out['g'], task_descr=None, info=info)
        out = {k: v.reshape(B, *v.shape[1:]) for k, v in out.items()}
        assert out['u'].shape[0] == B
        return out


class MultiTaskHerOracle:
    """her.py:72-185."""

    def __init__(self, goal_replay, her_replay_k, task_replay, reward_fun, tasks_ag_id=None,
                 tasks_g_id=None):
        self.future_p = future_probability(goal_replay, her_replay_k)
        self.task_replay = task_replay
        self.reward_fun = reward_fun
        self.g_ids = [list(g) for g in tasks_g_id]
        self.ag_ids = _truncated_ids(tasks_ag_id, tasks_g_id)
        self.nb_tasks = len(tasks_ag_id)
        # her.py:94-97: with per-module buffers the module is chosen by the caller (DDPG)
        self.multiple_buffers = ('buffer' in task_replay) or task_replay == 'hand_designed'

    def __call__(self, episode_batch, batch_size_in_transitions, task_to_replay=None, cp_proba=None,
                 stream=None):
        T = episode_batch['u'].shape[1]
        E = episode_batch['u'].shape[0]
        B = batch_size_in_transitions
        injected = stream is not None
        if not injected:
            stream = draw_stream(E, T, B)
        self.last_stream = stream
        ep, t = stream.ep, stream.t
        out = {k: v[ep, t].copy() for k, v in episode_batch.items()}
        her_rows = np.where(stream.u_her < self.future_p)[0]
        off = (stream.u_off * (T - t)).astype(int)
        future_t = (t + 1 + off)[her_rows]
        future_ag = episode_batch['ag'][ep[her_rows], future_t]

        g_out, td_out = out['g'], out['task_descr']
        for i, row in enumerate(her_rows):
            if self.task_replay == 'replay_current_task_transition':
                # her.py:158-164: keep the module, overwrite only its goal slice
                m = int(np.argwhere(td_out[row] == 1).squeeze())
                g_out[row, self.g_ids[m]] = future_ag[i, self.ag_ids[m]]
                continue
            if self.multiple_buffers:
                if task_to_replay is None:
                    m = int(np.argwhere(td_out[row] == 1).squeeze())      # her.py:134
                else:
                    m = task_to_replay                                     # her.py:136
            elif injected and stream.task_choice[row] >= 0:
                m = int(stream.task_choice[row])
            elif self.task_replay == 'replay_random_task_transition':
                m = np.random.choice(range(self.nb_tasks))                 # her.py:139
                stream.task_choice[row] = m
            elif self.task_replay == 'replay_cp_task_transition':
                m = np.random.choice(range(self.nb_tasks), p=cp_proba)     # her.py:142
                stream.task_choice[row] = m
            else:
                raise NameError('replay_task')   # the reference leaves it unbound here
            g_out[row] = 0                                                 # her.py:151
            td_out[row] = 0                                                # her.py:152
            g_out[row, self.g_ids[m]] = future_ag[i, self.ag_ids[m]]      # her.py:154
            td_out[row, m] = 1                                             # her.py:155

        info = {k[len('info_'):]: v for k, v in out.items() if k.startswith('info_')}
        out['r'] = self.reward_fun(ag_2=out['ag_2'], g=out['g'], task_descr=out['task_descr'],
                                   info=info)
        out = {k: v.reshape(B, *v.shape[1:]) for k, v in out.items()}
        assert out['u'].shape[0] == B
        return out


def make_sample_her_transitions(goal_replay, her_replay_k, reward_fun, task_replay='',
                                tasks_ag_id=None, tasks_g_id=None):
    return FlatHerOracle(goal_replay, her_replay_k, reward_fun, task_replay, tasks_ag_id, tasks_g_id)


def make_sample_multi_task_her_transitions(goal_replay, her_replay_k, task_replay, reward_fun,
                                           tasks_ag_id=None, tasks_g_id=None):
    return MultiTaskHerOracle(goal_replay, her_replay_k, task_replay, reward_fun, tasks_ag_id,
                              tasks_g_id)
