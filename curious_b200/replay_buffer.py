"""Device-resident replay buffer - drop-in for reference baselines/her/replay_buffer.py:6-109.

Storage is two float32 CUDA tensors of packed rows (layout in include/curious_b200.h, `cur_layout`:
"hot" transition rows [o(t) | g u task_descr ag(t+1) o(t+1)], 64-byte aligned, and "cold" rows
[change|info|ag(t)]) instead of the reference's dict of float64 host arrays (replay_buffer.py:23-24).  Everything the reference stores is
float32-representable (rollout.py:50-52,194-195 build float32 episodes; `change` is bool), so no
information is lost; inputs that are not are rounded to float32.

Slot selection (_get_storage_idx, replay_buffer.py:90-109) stays on the host and consumes the global
np.random stream exactly like the reference, so a seeded run overwrites the same slots.
"""
import ctypes as C
import threading

import numpy as np
import torch

from . import _lib
from .her import DeviceEpisodes

BASE_KEYS = ('o', 'ag', 'g', 'u')

# The device rows are float32 while the reference stores float64 (replay_buffer.py:23-24).  Everything its rollouts
# produce is float32 already (rollout.py:50-52,194-195), so nothing is lost - and this switch makes sure of it: a
# float64 episode whose values do not survive the round trip through float32 raises instead of being rounded silently.
CHECK_FLOAT32_REPRESENTABLE = True


def _as_float32(key, arr):
    arr = np.asarray(arr)
    if CHECK_FLOAT32_REPRESENTABLE and arr.dtype == np.float64:
        f = arr.astype(np.float32)
        if not np.array_equal(f.astype(np.float64), arr, equal_nan=True):
            raise ValueError('episode key %r holds float64 values that are not float32-representable (max round-off '
                             '%.3g); the device replay rows are float32' % (key, float(np.nanmax(np.abs(f - arr)))))
        return f
    return arr


def split_keys(shapes_or_batch):
    """Classify keys: returns (has_td, has_change, [info keys] sorted)."""
    keys = list(shapes_or_batch.keys())
    info = sorted(k for k in keys if k.startswith('info_'))
    known = set(BASE_KEYS) | {'task_descr', 'change', 'o_2', 'ag_2'} | set(info)
    extra = [k for k in keys if k not in known]
    if extra:
        raise KeyError('unsupported replay keys: %s' % extra)
    for k in BASE_KEYS:
        if k not in shapes_or_batch:
            raise KeyError('replay key %r is required' % k)
    return 'task_descr' in shapes_or_batch, 'change' in shapes_or_batch, info


_LAYOUTS = {}


def layout_from_shapes(buffer_shapes):
    """buffer_shapes: {key: (T or T+1, dim)} as built by config.configure_buffer (config.py:200-208).  Cached per shape
    signature (store_episode asks once per call for the normaliser's temporary episodes)."""
    sig = tuple((k, tuple(int(x) for x in v)) for k, v in buffer_shapes.items())
    hit = _LAYOUTS.get(sig)
    if hit is None:
        hit = _LAYOUTS[sig] = _layout_from_shapes(buffer_shapes)
    L, info_keys, has_td, has_change = hit
    return L, list(info_keys), has_td, has_change


def _layout_from_shapes(buffer_shapes):
    has_td, has_change, info = split_keys(buffer_shapes)
    T = buffer_shapes['u'][0]
    assert buffer_shapes['o'][0] == T + 1 and buffer_shapes['ag'][0] == T + 1
    info_keys = [(k, int(buffer_shapes[k][-1])) for k in info]
    L = _lib.make_layout(T, buffer_shapes['o'][-1], buffer_shapes['ag'][-1], buffer_shapes['g'][-1],
                         buffer_shapes['u'][-1], buffer_shapes['task_descr'][-1] if has_td else 0,
                         buffer_shapes['change'][-1] if has_change else 0, sum(d for _, d in info_keys))
    return L, info_keys, has_td, has_change


def alloc_storage(layout, n_episodes, device):
    """(hot, cold) tensors for `n_episodes` episodes: transition rows and [change | info | ag(t)] rows."""
    hot = torch.empty(n_episodes * layout.T * layout.trans_stride, dtype=torch.float32, device=device)
    cold = torch.empty(n_episodes * layout.T * layout.cold_stride, dtype=torch.float32, device=device)
    return hot, cold


class StagedEpisodes:
    """Key-major float32 episodes uploaded to the device in one blob (input of cur_store_episodes)."""

    def __init__(self, episode_batch, layout, info_keys, has_td, has_change, device):
        self.layout = layout
        n = len(episode_batch['u'])
        self.n = n
        T = layout.T
        parts = [('o', (T + 1) * layout.dimo), ('ag', (T + 1) * layout.dimag), ('g', T * layout.dimg),
                 ('u', T * layout.dimu)]
        if has_td:
            parts.append(('task_descr', T * layout.dimtd))
        if has_change:
            parts.append(('change', T * layout.dimchange))
        total = sum(sz for _, sz in parts) * n + n * T * layout.diminfo
        host = torch.empty(total, dtype=torch.float32, pin_memory=True)
        h = host.numpy()
        offs, k0 = {}, 0
        for key, sz in parts:
            arr = _as_float32(key, episode_batch[key])
            assert arr.shape[0] == n and arr[0].size == sz, 'bad shape for %s: %s' % (key, arr.shape)
            h[k0:k0 + n * sz] = arr.reshape(-1)          # casts bool / float64 to float32
            offs[key] = k0
            k0 += n * sz
        if info_keys:
            info = np.concatenate([np.asarray(_as_float32(k, episode_batch[k]), np.float32).reshape(n, T, d)
                                   for k, d in info_keys], axis=2)
            h[k0:k0 + info.size] = info.reshape(-1)
            offs['info'] = k0
            k0 += info.size
        self.blob = host.to(device, non_blocking=True)
        self._host = host
        base = self.blob.data_ptr()
        self.src = _lib.EpisodeSrc()
        self.src.o = base + 4 * offs['o']
        self.src.ag = base + 4 * offs['ag']
        self.src.g = base + 4 * offs['g']
        self.src.u = base + 4 * offs['u']
        self.src.td = base + 4 * offs['task_descr'] if has_td else None
        self.src.change = base + 4 * offs['change'] if has_change else None
        self.src.info = base + 4 * offs['info'] if info_keys else None

    @classmethod
    def from_device(cls, tensors, layout):
        """Episodes that already live on the device as key-major float32 tensors
        {o, ag, g, u[, task_descr, change, info]} (synthetic fills, device-side env stepping)."""
        self = cls.__new__(cls)
        self.layout = layout
        self.n = tensors['u'].shape[0]
        self._keep = tensors
        self.src = _lib.EpisodeSrc()
        for key, field in (('o', 'o'), ('ag', 'ag'), ('g', 'g'), ('u', 'u'), ('task_descr', 'td'),
                           ('change', 'change'), ('info', 'info')):
            t = tensors.get(key)
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
            setattr(self.src, field, None if t is None else t.data_ptr())
        return self

    def store(self, copies, stream=None):
        """copies: list of (src_episode, hot_tensor, cold_tensor_or_None, slot)."""
        # Two copies of one call may target the same slot of the same buffer (random overwrite once a buffer is full,
        # replay_buffer.py:99-102).  The reference stores one after the other, so the later episode wins as a whole;
        # the kernel writes all copies concurrently, so the earlier ones are dropped here (last writer kept).
        last = {}
        for i, c in enumerate(copies):
            last[(c[1].data_ptr(), int(c[3]))] = i
        copies = [c for i, c in enumerate(copies) if last[(c[1].data_ptr(), int(c[3]))] == i]
        n = len(copies)
        if n == 0:
            return
        src = (C.c_int32 * n)(*[int(c[0]) for c in copies])
        hot = (C.c_void_p * n)(*[c[1].data_ptr() for c in copies])
        cold = (C.c_void_p * n)(*[None if c[2] is None else c[2].data_ptr() for c in copies])
        slot = (C.c_int64 * n)(*[int(c[3]) for c in copies])
        _lib.check(_lib.load().cur_store_episodes(_lib.stream_ptr(stream), C.byref(self.layout),
                                                  C.byref(self.src), self.n, n, src, hot, cold, slot),
                   'cur_store_episodes')


def default_device():
    if not torch.cuda.is_available():
        raise _lib.CuriousLibError('curious_b200 needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def episodes_to_device(episode_batch, device=None, staged=None):
    """Pack a host episode batch {key: [n, T(+1), dim]} into a temporary device buffer and return the
    DeviceEpisodes view the sampler consumes (used for the normaliser path, ddpg.py:209-215).  `staged`: the
    StagedEpisodes of the SAME batch already uploaded by the caller (store_episode) - reused instead of a second
    host->device copy."""
    device = device or default_device()
    batch = {k: v for k, v in episode_batch.items() if k not in ('o_2', 'ag_2')}
    shapes = {k: np.asarray(v).shape[1:] for k, v in batch.items()}
    L, info_keys, has_td, has_change = layout_from_shapes(shapes)
    n = len(batch['u'])
    hot, cold = alloc_storage(L, n, device)
    if staged is None or staged.n != n or bytes(staged.layout) != bytes(L):
        staged = StagedEpisodes(batch, L, info_keys, has_td, has_change, device)
    staged.store([(e, hot, cold, e) for e in range(n)])
    epi = DeviceEpisodes(hot, cold, n, L, info_keys, has_td, has_change)
    epi._staged = staged
    return epi


class ReplayBuffer:
    def __init__(self, buffer_shapes, size_in_transitions, T, sample_transitions, device=None):
        """Same arguments as the reference (replay_buffer.py:7-16) + optional `device`."""
        self.buffer_shapes = buffer_shapes
        self.size = size_in_transitions // T                     # replay_buffer.py:18
        self.T = T
        self.sample_transitions = sample_transitions
        self.device = device or default_device()
        self.layout, self.info_keys, self.has_td, self.has_change = layout_from_shapes(buffer_shapes)
        assert self.layout.T == T
        self.storage, self.cold = alloc_storage(self.layout, self.size, self.device)
        self.current_size = 0
        self.n_transitions_stored = 0
        self.lock = threading.Lock()

    @property
    def full(self):
        with self.lock:
            return self.current_size == self.size

    def device_view(self):
        return DeviceEpisodes(self.storage, self.cold, self.current_size, self.layout, self.info_keys,
                              self.has_td, self.has_change)

    def sample(self, batch_size, task_to_replay=None, cp_proba=None):
        """Returns a dict {key: array(batch_size x shapes[key])} (replay_buffer.py:37-55)."""
        with self.lock:
            assert self.current_size > 0
            view = self.device_view()
        transitions = self.sample_transitions(view, batch_size, task_to_replay=task_to_replay,
                                              cp_proba=cp_proba)
        for key in (['r', 'o_2', 'ag_2'] + list(self.buffer_shapes.keys())):
            assert key in transitions, "key %s missing from transitions" % key
        return transitions

    def store_episode(self, episode_batch):
        """episode_batch: array(batch_size x (T or T+1) x dim_key) (replay_buffer.py:57-72)."""
        batch_sizes = [len(episode_batch[key]) for key in episode_batch.keys()]
        assert np.all(np.array(batch_sizes) == batch_sizes[0])
        batch_size = batch_sizes[0]
        with self.lock:
            idxs = np.atleast_1d(self._get_storage_idx(batch_size))
            staged = StagedEpisodes({k: episode_batch[k] for k in self.buffer_shapes.keys()}, self.layout,
                                    self.info_keys, self.has_td, self.has_change, self.device)
            staged.store([(e, self.storage, self.cold, int(idxs[e])) for e in range(batch_size)])
            self._last_staged = staged       # keep the pinned/device blobs alive until the copy ran
            self.n_transitions_stored += batch_size * self.T

    def store_staged(self, staged, src_episode):
        """Reserve a slot for episode `src_episode` of an already uploaded batch and return the copy
        descriptor (DDPG.store_episode duplicates one episode into several module buffers,
        ddpg.py:194-195, with a single upload and a single kernel launch)."""
        with self.lock:
            idx = self._get_storage_idx(1)
            self.n_transitions_stored += self.T
        return (src_episode, self.storage, self.cold, int(idx))

    def get_current_episode_size(self):
        with self.lock:
            return self.current_size

    def get_current_size(self):
        with self.lock:
            return self.current_size * self.T

    def get_transitions_stored(self):
        with self.lock:
            return self.n_transitions_stored

    def clear_buffer(self):
        with self.lock:
            self.current_size = 0

    def _get_storage_idx(self, inc=None):
        inc = inc or 1   # size increment
        assert inc <= self.size, "Batch committed to replay is too large!"
        # consecutive until the end is hit, then uniformly random slots (replay_buffer.py:94-102)
        if self.current_size + inc <= self.size:
            idx = np.arange(self.current_size, self.current_size + inc)
        elif self.current_size < self.size:
            overflow = inc - (self.size - self.current_size)
            idx = np.concatenate([np.arange(self.current_size, self.size),
                                  np.random.randint(0, self.current_size, overflow)])
        elif inc == 1:
            # one slot of a full buffer (every per-module copy of store_staged): the scalar form draws the same value from the
            # same stream position as np.random.randint(0, size, 1)[0] at a third of the cost
            # (tests/test_host_logic.py::test_scalar_randint_is_the_size_one_draw)
            return np.random.randint(0, self.size)
        else:
            idx = np.random.randint(0, self.size, inc)
        self.current_size = min(self.size, self.current_size + inc)
        if inc == 1:
            idx = idx[0]
        return idx

    @property
    def buffers(self):
        """Host copy in the reference's layout {key: float64 [size, T(+1), dim]} (debug / export only)."""
        L, T = self.layout, self.T
        n = self.current_size              # only the filled slots are copied; the rest reads as zeros (np.empty in the reference)
        rows = np.zeros((self.size, T, L.trans_stride), np.float64)
        rows[:n] = self.storage[:n * T * L.trans_stride].view(n, T, L.trans_stride).cpu().numpy()
        cold = np.zeros((self.size, T, L.cold_stride), np.float64)
        cold[:n] = self.cold[:n * T * L.cold_stride].view(n, T, L.cold_stride).cpu().numpy()
        b0 = (L.dimo + 3) // 4 * 4                       # the step block follows o(t)
        step = lambda off, dim: rows[:, :, b0 + off:b0 + off + dim]
        # o(t) heads transition t, o(T) is the o(t+1) block of the last transition; ag(t) lives in the cold rows
        out = {'o': np.concatenate([rows[:, :, :L.dimo], step(L.off_o, L.dimo)[:, -1:]], axis=1),
               'ag': np.concatenate([cold[:, :, L.off_agc:L.off_agc + L.dimag], step(L.off_ag, L.dimag)[:, -1:]], axis=1),
               'g': step(L.off_g, L.dimg), 'u': step(L.off_u, L.dimu)}
        if self.has_td:
            out['task_descr'] = step(L.off_td, L.dimtd)
        if self.has_change:
            out['change'] = cold[:, :, L.off_change:L.off_change + L.dimchange]
        k0 = L.off_info
        for key, d in self.info_keys:
            out[key] = cold[:, :, k0:k0 + d]
            k0 += d
        return out
