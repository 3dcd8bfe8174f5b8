"""NumPy restatement of ReplayBuffer (test infrastructure; see oracle/__init__.py).

Follows reference baselines/her/replay_buffer.py:6-109.  PINNED by tests/test_oracle_golden.py
against the unmodified reference class (fixtures from oracle/gen_golden.py).
"""
import numpy as np


class ReplayBufferOracle:
    def __init__(self, buffer_shapes, size_in_transitions, T, sample_transitions):
        self.buffer_shapes = buffer_shapes
        self.size = size_in_transitions // T            # replay_buffer.py:18 (episodes)
        self.T = T
        self.sample_transitions = sample_transitions
        # replay_buffer.py:23-24: float64 storage, uninitialised
        self.buffers = {k: np.empty([self.size, *shape]) for k, shape in buffer_shapes.items()}
        self.current_size = 0
        self.n_transitions_stored = 0

    @property
    def full(self):
        return self.current_size == self.size

    def sample(self, batch_size, task_to_replay=None, cp_proba=None, stream=None):
        assert self.current_size > 0                     # replay_buffer.py:43
        view = {k: v[:self.current_size] for k, v in self.buffers.items()}
        view['o_2'] = view['o'][:, 1:, :]                # replay_buffer.py:47-48
        view['ag_2'] = view['ag'][:, 1:, :]
        if stream is None:
            out = self.sample_transitions(view, batch_size, task_to_replay=task_to_replay,
                                          cp_proba=cp_proba)
        else:
            out = self.sample_transitions(view, batch_size, task_to_replay=task_to_replay,
                                          cp_proba=cp_proba, stream=stream)
        for key in ['r', 'o_2', 'ag_2'] + list(self.buffers.keys()):
            assert key in out, "key %s missing from transitions" % key
        return out

    def store_episode(self, episode_batch):
        sizes = [len(v) for v in episode_batch.values()]
        assert all(s == sizes[0] for s in sizes)         # replay_buffer.py:61-62
        n = sizes[0]
        idxs = self._get_storage_idx(n)
        for key in self.buffers:
            self.buffers[key][idxs] = episode_batch[key]
        self.n_transitions_stored += n * self.T
        return idxs

    def get_current_episode_size(self):
        return self.current_size

    def get_current_size(self):
        return self.current_size * self.T

    def get_transitions_stored(self):
        return self.n_transitions_stored

    def clear_buffer(self):
        self.current_size = 0

    def _get_storage_idx(self, inc=None):
        # replay_buffer.py:90-109: fill in order, then overwrite uniformly at random
        inc = inc or 1
        assert inc <= self.size, "Batch committed to replay is too large!"
        if self.current_size + inc <= self.size:
            idx = np.arange(self.current_size, self.current_size + inc)
        elif self.current_size < self.size:
            spill = inc - (self.size - self.current_size)
            head = np.arange(self.current_size, self.size)
            tail = np.random.randint(0, self.current_size, spill)
            idx = np.concatenate([head, tail])
        else:
            idx = np.random.randint(0, self.size, inc)
        self.current_size = min(self.size, self.current_size + inc)
        if inc == 1:
            idx = idx[0]
        return idx
