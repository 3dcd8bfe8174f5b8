"""Competence / learning-progress bookkeeping that produces the `cp` vector the replay sampler consumes.

Drop-in for reference baselines/her/queues.py:7-36 (CompetenceQueue) plus the rank-0 part of
RolloutWorker.generate_rollouts that turns per-module success lists into CP and task-selection probabilities
(baselines/her/rollout.py:316-404).  Host logic only (a few floats per rollout); the gathers / broadcasts of the
reference's MPI.COMM_WORLD become torch.distributed object collectives (gloo or NCCL group alike).

The reference module itself no longer imports (`from pandas import ewma`), so this restatement is pinned by
tests/golden/competence_queue.npz, recorded from the unmodified file by oracle/gen_golden_queue.py.
"""
import itertools
from collections import deque

import numpy as np


class CompetenceQueue(object):
    def __init__(self, window=100):
        self.window = window
        self.successes = deque(maxlen=2 * self.window)
        self.CP = 0.
        self.C = 0.

    def update(self, success_list):
        """queues.py:14-23: CP = |sum(recent half) - sum(older half)| / (2 window), C = mean(recent half)."""
        for success in success_list:
            self.successes.append(success)
        if self.size > 2:
            window = min(self.size // 2, self.window)
            q1 = list(itertools.islice(self.successes, self.size - window, self.size))
            q2 = list(itertools.islice(self.successes, self.size - 2 * window, self.size - window))
            self.CP = np.abs(np.sum(q1) - np.sum(q2)) / (2 * window)
            self.C = np.sum(q1) / window

    @property
    def size(self):
        return len(self.successes)

    @property
    def full(self):
        return self.size == self.successes.maxlen

    def clear_queue(self):
        self.successes = deque(maxlen=2 * self.window)
        self.CP = 0
        self.C = 0.


def task_probabilities(CP, nb_tasks, epsilon=0.4):
    """rollout.py:381-393: epsilon-proportional task selection from competence progress, made to sum to 1."""
    CP = np.asarray(CP, np.float64)
    if CP.sum() == 0:
        p = (1 / nb_tasks) * np.ones([nb_tasks])
    else:
        p = epsilon * (1 / nb_tasks) * np.ones([nb_tasks]) + (1 - epsilon) * CP / CP.sum()
    if p.sum() > 1:
        p[np.argmax(p)] -= p.sum() - 1
    elif p.sum() < 1:
        p[-1] = 1 - p[:-1].sum()
    return p


class CompetenceTracker(object):
    """The LP pipeline of RolloutWorker (rollout.py:59-75, 316-404, 415-417, 485-491) without the environments:
    feed it the (task, success) pairs of this rank's exploit rollouts, get back (CP, p) identical on every rank."""

    def __init__(self, nb_tasks, queue_length=500, task_selection='active_competence_progress', structure='curious',
                 unique_task=None, eval=False, comm=None):
        self.nb_tasks = nb_tasks
        self.task_selection = task_selection
        self.structure = structure
        self.unique_task = unique_task
        self.eval = eval
        self.comm = comm
        self.competence_computers = [CompetenceQueue(window=queue_length) for _ in range(nb_tasks)]
        self.CP = np.zeros([nb_tasks])
        self.C = np.zeros([nb_tasks])
        self.p = 1 / nb_tasks * np.ones([nb_tasks])                    # rollout.py:61
        if structure == 'task_experts' and unique_task is not None:
            self.p = np.zeros([nb_tasks])
            self.p[unique_task] = 1

    def get_CP(self):
        return [cq.CP for cq in self.competence_computers]

    def get_C(self):
        return [cq.C for cq in self.competence_computers]

    def clear_competence_queue(self):
        for cq in self.competence_computers:
            cq.clear_queue()

    # resume support (beyond the reference, which restarts its queues from scratch)
    def state(self):
        return dict(queues=[(list(cq.successes), cq.CP, cq.C) for cq in self.competence_computers],
                    CP=np.array(self.CP, np.float64), C=np.array(self.C, np.float64), p=np.array(self.p, np.float64))

    def load_state(self, st):
        assert len(st['queues']) == self.nb_tasks
        for cq, (succ, cp, c) in zip(self.competence_computers, st['queues']):
            cq.successes = deque(succ, maxlen=2 * cq.window)
            cq.CP, cq.C = cp, c
        self.CP, self.C, self.p = st['CP'].copy(), st['C'].copy(), st['p'].copy()

    def update(self, tasks, successes):
        """tasks / successes: this rank's exploit rollouts ([] when exploration noise was used, rollout.py:318-330).
        Rank 0 gathers, updates the queues and the probabilities; p and CP are broadcast (rollout.py:332-404)."""
        from .parallel import rank as _rank, world as _world
        group, n = _world(self.comm)
        tasks, successes = list(tasks), list(successes)
        if n > 1:
            import torch.distributed as dist
            gathered = [None] * n
            dist.all_gather_object(gathered, (tasks, successes), group=group)
            tasks = sum([g[0] for g in gathered], [])
            successes = sum([g[1] for g in gathered], [])
        if _rank(self.comm) == 0:
            task_succ_list = [[] for _ in range(self.nb_tasks)]
            for succ, task in zip(successes, tasks):
                task_succ_list[int(task)].append(succ)
            for task in range(self.nb_tasks):
                self.competence_computers[task].update(task_succ_list[task])
            self.C = np.array([self.get_C()]).squeeze()
            if not self.eval:
                if self.task_selection == 'active_competence_progress' and self.structure != 'task_experts':
                    self.CP = np.array([self.get_CP()]).squeeze()
                    self.p = task_probabilities(self.CP, self.nb_tasks)
                elif self.structure == 'task_experts':
                    self.p = np.zeros([self.nb_tasks])
                    self.p[self.unique_task] = 1
        if n > 1 and not self.eval:
            import torch.distributed as dist
            box = [self.p, self.CP]
            dist.broadcast_object_list(box, src=0 if group is None else dist.get_global_rank(group, 0), group=group)
            self.p, self.CP = box
        return self.CP, self.p
