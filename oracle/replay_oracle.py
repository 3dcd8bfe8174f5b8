"""CPU restatement of the reference replay storage (test infrastructure; see oracle/__init__.py).

What the reference does (baselines/her/replay_buffer.py:6-109), restated as two pure rules plus a thin holder:

  slot rule      episodes fill slots 0, 1, 2, ... while there is room; a batch that only partly fits takes the free
                 tail and then random occupied slots; once full every new episode overwrites a uniformly random slot
                 (replay_buffer.py:90-109; np.random.randint is the reference's only draw here and is consumed in the
                 same order).  A single episode yields a scalar slot, not an array (replay_buffer.py:106-107).
  sampling rule  the sampler sees only the filled prefix of every key, plus the one-step-shifted views `o_2`, `ag_2`
                 (replay_buffer.py:44-48), and must return every stored key and `r` (replay_buffer.py:52-53).

Storage is float64 like the reference (replay_buffer.py:23-24).  PINNED by tests/test_oracle_golden.py against
fixtures recorded from the unmodified reference class (oracle/gen_golden.py).
"""
import numpy as np


def choose_slots(filled, capacity, count):
    """Slot rule: returns (slots [count] int array, filled slots afterwards)."""
    if count > capacity:
        raise AssertionError('Batch committed to replay is too large!')
    room = capacity - filled
    if count <= room:
        slots = filled + np.arange(count)
    elif room > 0:
        slots = np.concatenate([filled + np.arange(room), np.random.randint(0, filled, count - room)])
    else:
        slots = np.random.randint(0, capacity, count)
    return slots, min(capacity, filled + count)


def filled_views(storage, filled):
    """Sampling rule, first half: per-key views of the filled prefix and the shifted next-step views."""
    views = {key: arr[:filled] for key, arr in storage.items()}
    views['o_2'] = views['o'][:, 1:, :]
    views['ag_2'] = views['ag'][:, 1:, :]
    return views


class ReplayBufferOracle:
    def __init__(self, buffer_shapes, size_in_transitions, T, sample_transitions):
        self.buffer_shapes, self.T, self.sample_transitions = buffer_shapes, T, sample_transitions
        self.size = size_in_transitions // T                       # capacity in episodes
        self.buffers = {key: np.empty((self.size,) + tuple(shape)) for key, shape in buffer_shapes.items()}
        self.current_size = 0                                      # filled episode slots
        self.n_transitions_stored = 0

    # ---- the reference's read-only surface
    full = property(lambda self: self.current_size == self.size)

    def get_current_episode_size(self):
        return self.current_size

    def get_current_size(self):
        return self.T * self.current_size

    def get_transitions_stored(self):
        return self.n_transitions_stored

    def clear_buffer(self):
        self.current_size = 0

    # ---- store / sample
    def _get_storage_idx(self, inc=None):
        count = inc or 1
        slots, self.current_size = choose_slots(self.current_size, self.size, count)
        return slots[0] if count == 1 else slots

    def store_episode(self, episode_batch):
        counts = {len(arr) for arr in episode_batch.values()}
        assert len(counts) == 1, 'ragged episode batch'            # replay_buffer.py:61-62
        count = counts.pop()
        slots = self._get_storage_idx(count)
        for key, arr in self.buffers.items():
            arr[slots] = episode_batch[key]
        self.n_transitions_stored += count * self.T
        return slots

    def sample(self, batch_size, task_to_replay=None, cp_proba=None, stream=None):
        assert self.current_size > 0, 'sampling from an empty buffer'      # replay_buffer.py:43
        extra = {} if stream is None else {'stream': stream}
        out = self.sample_transitions(filled_views(self.buffers, self.current_size), batch_size,
                                      task_to_replay=task_to_replay, cp_proba=cp_proba, **extra)
        missing = [key for key in ('r', 'o_2', 'ag_2', *self.buffers) if key not in out]
        assert not missing, 'keys %s missing from transitions' % missing
        return out
