"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in
the CPU tests).  Replaces the reference's mpi4py calls on the hot path:

  MpiAdam.update    comm.Allreduce(localg, globalg, op=MPI.SUM)     baselines/common/mpi_adam.py:24-26
  MpiAdam.sync      comm.Bcast(theta, root=0)                        baselines/common/mpi_adam.py:37-40
  MpiAdam.check_synced  Bcast + assert equal                         baselines/common/mpi_adam.py:42-50
  Normalizer._mpi_average  Allreduce(SUM) then / size                baselines/her/normalizer.py:84-88
  rank seeding      rank_seed = seed + 1000000 * rank                baselines/her/experiment/train.py:242

The replay data never crosses ranks (each rank owns its buffers, train.py:242-243 / config.py:210-214), so
the HER path has no collective at all.  These helpers work on any tensor the backend supports, which is what
lets the world_size-2 gloo tests exercise the exact same code on CPU tensors.
"""
import torch


def world(comm=None):
    """(group, world_size).  comm=False forces single-process behaviour."""
    import torch.distributed as dist
    if comm is False:
        return None, 1
    if dist.is_available() and dist.is_initialized():
        group = comm if comm is not None else dist.group.WORLD
        return group, dist.get_world_size(group)
    return None, 1


def rank(comm=None):
    import torch.distributed as dist
    group, n = world(comm)
    return dist.get_rank(group) if n > 1 else 0


def _root(group):
    import torch.distributed as dist
    return dist.get_global_rank(group, 0) if group is not None and group is not dist.group.WORLD else 0


# True: sums over ranks are evaluated in RANK ORDER ((x0 + x1) + x2 ...) from an all-gather instead of NCCL's
# all-reduce, whose summation order depends on its algorithm.  With more than two ranks this is the reference result
# the peer-memory exchanges (csrc/p2p.cu, the tile exchange of csrc/ddpg_rows.cu) reproduce bit for bit; the parity
# checks of tests/test_multi_gpu.py and bench.py switch it on for the NCCL arm (CUR_ORDERED_ALLREDUCE=1 does the same).
import os as _os
ORDERED_ALLREDUCE = _os.environ.get('CUR_ORDERED_ALLREDUCE', '0') == '1'


def allreduce_sum_(t, comm=None):
    """In-place SUM over ranks - gradients are summed, NOT averaged (scale_grad_by_procs=False,
    ddpg.py:452-453).  Returns the world size."""
    group, n = world(comm)
    if n > 1:
        import torch.distributed as dist
        if ORDERED_ALLREDUCE:
            parts = [torch.empty_like(t) for _ in range(n)]
            dist.all_gather(parts, t.contiguous(), group=group)
            acc = parts[0].clone()
            for p in parts[1:]:
                acc += p
            t.copy_(acc)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return n


def broadcast_from_root_(t, comm=None):
    group, n = world(comm)
    if n > 1:
        import torch.distributed as dist
        dist.broadcast(t, src=_root(group), group=group)
    return t


def assert_synced(fingerprint, comm=None):
    """check_synced: every rank's fingerprint (any tensor that is a function of the parameters - the full
    vector as in the reference, or a 64-bit checksum of its bit patterns) must equal rank 0's."""
    group, n = world(comm)
    if n <= 1:
        return
    ref = fingerprint.clone()
    broadcast_from_root_(ref, comm)
    assert bool(torch.equal(ref, fingerprint)), 'parameters diverged from rank 0 (mpi_adam.py:50)'


def rank_seed(seed, r):
    """train.py:242: every worker draws from its own NumPy stream."""
    return seed + 1000000 * r


def bcast_object(obj, comm=None):
    """MPI.COMM_WORLD.bcast(obj, root=0) (train.py:104, rollout.py:403-404): rank 0's picklable object on every rank."""
    group, n = world(comm)
    if n <= 1:
        return obj
    import torch.distributed as dist
    box = [obj]
    dist.broadcast_object_list(box, src=_root(group), group=group)
    return box[0]


def assert_rank_streams_differ(comm=None):
    """train.py:207-212: once per epoch every rank draws one uniform from its np.random stream and compares with rank 0's -
    ranks that were seeded alike (identical exploration, identical replay slots) fail here instead of silently
    training on duplicated data.  Consumes one draw on every rank, like the reference."""
    import numpy as np
    local = float(np.random.uniform(size=(1,))[0])
    group, n = world(comm)
    if n > 1:
        import torch.distributed as dist
        box = [local]
        dist.broadcast_object_list(box, src=_root(group), group=group)
        if rank(comm) != 0:
            assert local != box[0], 'this rank draws the same np.random stream as rank 0 (train.py:211-212)'
    return local


def install_excepthook():
    """her/util.py:129-139 (`install_mpi_excepthook`): an uncaught exception on one rank must not leave the others
    waiting in a collective.  The reference calls MPI.COMM_WORLD.Abort(); under torchrun the equivalent is to print the
    traceback and leave at once with a non-zero status - the launcher then tears the whole group down."""
    import os
    import sys
    previous = sys.excepthook

    def hook(kind, value, tb):
        previous(kind, value, tb)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
    sys.excepthook = hook
    return hook


class PeerGradExchange(object):
    """Gradient exchange over NVLink peer memory, fused with Adam (csrc/p2p.cu; replaces the Allreduce + Adam of
    mpi_adam.py:24-35 on the CUDA-graph path).  Every rank allocates a region [flags | grads 0 | grads 1],
    ships its CUDA-IPC handle through torch.distributed and maps the peers' regions.  Needs one process per GPU
    on one NVLink domain (cudaIpcOpenMemHandle fails otherwise -> the caller falls back to NCCL explicitly)."""

    FLAG_BYTES = 256

    def __init__(self, arena_floats, comm=None, sharded=True):
        import ctypes as C
        self.sharded = sharded
        import torch.distributed as dist
        from . import _lib
        lib = _lib.load()
        group, n = world(comm)
        assert 1 < n <= _lib.CUR_MAX_RANKS, 'peer exchange needs 2..%d ranks' % _lib.CUR_MAX_RANKS
        self.lib = lib
        self.world = n
        self.rank = dist.get_rank(group)
        self.arena = int(arena_floats)
        nbytes = lib.cur_p2p_region_bytes(self.arena)
        assert nbytes > 0, 'gradient arena must be a positive multiple of 4 floats'
        self.nbytes = nbytes
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(lib.cur_p2p_alloc(nbytes, C.byref(own), handle), 'cur_p2p_alloc')
        self.own = own.value
        self.opened = []
        handles = [None] * n
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ctx = _lib.P2PCtx()
        self.ctx.rank, self.ctx.world, self.ctx.arena = self.rank, n, self.arena
        for r in range(n):
            if r == self.rank:
                self.ctx.region[r] = self.own
                continue
            p = C.c_void_p()
            _lib.check(lib.cur_p2p_open(handles[r], C.byref(p)), 'cur_p2p_open')
            self.opened.append(p.value)
            self.ctx.region[r] = p.value
        self.error_flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        dist.barrier(group=group)                       # every region is mapped before anyone signals

    def grads_ptr(self, parity=0):
        """Device address of gradient buffer `parity` in this rank's own region."""
        return self.own + self.FLAG_BYTES + 4 * self.arena * parity

    def allreduce_adam(self, stream, theta, m, v, neg_a_table, table_len, step_counter, beta1, beta2, eps,
                       transposes=None):
        import ctypes as C
        from . import _lib
        if transposes is not None and not self.sharded:
            _lib.check(self.lib.cur_p2p_allreduce_adam_t(
                stream, C.byref(self.ctx), theta.data_ptr(), m.data_ptr(), v.data_ptr(), neg_a_table.data_ptr(),
                int(table_len), step_counter.data_ptr(), beta1, beta2, eps, self.error_flag.data_ptr(),
                C.byref(transposes)), 'cur_p2p_allreduce_adam_t')
            return
        fn = self.lib.cur_p2p_sharded_adam if self.sharded else self.lib.cur_p2p_allreduce_adam
        _lib.check(fn(
            stream, C.byref(self.ctx), theta.data_ptr(), m.data_ptr(), v.data_ptr(), neg_a_table.data_ptr(),
            int(table_len), step_counter.data_ptr(), beta1, beta2, eps, self.error_flag.data_ptr()),
            'cur_p2p_allreduce_adam')

    def check(self):
        """Raises if a peer did not show up within the kernel's spin budget (a rank died or diverged)."""
        if int(self.error_flag.item()) != 0:
            raise RuntimeError('peer gradient exchange timed out waiting for another rank')

    def close(self):
        for p in self.opened:
            self.lib.cur_p2p_close(p)
        self.opened = []
        if self.own:
            self.lib.cur_p2p_free(self.own)
            self.own = None


class TileGradExchange(object):
    """Gradient exchange INSIDE the weight-gradient launch of the rows schedule (csrc/ddpg_rows.cu, `cur_xchg_ctx`;
    replaces the Allreduce(SUM) of mpi_adam.py:24-28): every CTA pushes its tile of dW to the reducing rank(s) as
    8-byte {value, update number} words over NVLink peer memory, the reducer sums the world's tiles in rank order and
    applies Adam in the same epilogue.  mode 0: every rank reduces every tile (one hop, full Adam state everywhere);
    mode 1: tile t is reduced by rank t % world, which pushes the stepped parameters back (two hops, 1/W of the bytes
    per rank at large W, Adam moments only on the owner); mode 2 / 'nvls': as mode 1 with the NVSwitch doing the sum
    (`multimem.ld_reduce` on a multicast mapping of the partial slots) and the broadcast (one store to the multicast
    mapping of the result slots) - parameters identical on every rank, summation order the hardware's.
    'auto' = 0 for two ranks, 1 above.
    One process per GPU on one NVLink domain (cudaIpcOpenMemHandle)."""

    def __init__(self, arena_floats, comm=None, mode='auto', timeline_tiles=0):
        import ctypes as C
        import os
        import torch.distributed as dist
        from . import _lib
        lib = _lib.load()
        group, n = world(comm)
        assert 1 < n <= _lib.CUR_MAX_RANKS, 'the tile exchange needs 2..%d ranks' % _lib.CUR_MAX_RANKS
        self.lib, self.group, self.world = lib, group, n
        self.rank = dist.get_rank(group)
        self.arena = int(arena_floats)
        env = os.environ.get('CUR_XCHG_MODE')
        if env in ('0', '1', '2'):
            mode = int(env)
        if mode == 'nvls':
            mode = 2
        auto = mode == 'auto'
        if auto:
            # measured per-rank update (us), modes 0 / 1 / 2: 2 GPUs 68.0 / 71.5 / 70.8, 8 GPUs 99.6 / 74.3 / 71.7.  The NVLS
            # form (2) is the fastest from 4 ranks but stays opt-in (xchg_mode='nvls', CUR_XCHG_MODE=2, CUR_XCHG_AUTO_NVLS=1):
            # its sum is not in rank order, and its setup depends on the system's multicast support
            mode = 0 if n == 2 else (2 if os.environ.get('CUR_XCHG_AUTO_NVLS') == '1' else 1)
        assert mode in (0, 1, 2)
        self.mode = mode
        self.ctx = _lib.XchgCtx()
        self.ctx.rank, self.ctx.world, self.ctx.mode, self.ctx.arena = self.rank, n, mode, self.arena
        self.opened = []
        self.own = None
        self._symm = None
        if mode == 2:
            # NVLS: one symmetric allocation per rank with peer AND multicast mappings (torch's symmetric memory does the
            # cuMem / cuMulticast plumbing): [partial slots: arena x 8 bytes | result slots: arena x 8 bytes]
            import torch.distributed._symmetric_memory as symm_mem
            self.nbytes = 2 * 8 * self.arena
            self._symm = symm_mem.empty(self.nbytes // 4, dtype=torch.float32, device='cuda')
            self._symm.zero_()
            torch.cuda.synchronize()
            hdl, why = None, ''
            try:
                hdl = symm_mem.rendezvous(self._symm, (group or dist.group.WORLD).group_name)
                if not hdl.multicast_ptr:
                    hdl, why = None, 'no multicast mapping (NVSwitch) on this system'
            except Exception as e:
                why = '%s: %s' % (type(e).__name__, e)
            ok = torch.tensor([1 if hdl is not None else 0], dtype=torch.int32, device='cuda')
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)          # every rank takes the same path
            if int(ok.item()) == 0:
                if not auto:
                    raise RuntimeError('exchange mode 2 (NVLS) is not available: %s' % (why or 'a peer could not map it'))
                self._symm, mode = None, 1                                   # 'auto' falls back to the peer-memory form
                self.mode = self.ctx.mode = mode
            else:
                self._hdl = hdl
                for r in range(n):
                    self.ctx.region[r] = self._symm.data_ptr() if r == self.rank else hdl.buffer_ptrs[r]
                self.ctx.mc_region = hdl.multicast_ptr
        if mode != 2:
            self.nbytes = lib.cur_xchg_region_bytes(self.arena, n)
            assert self.nbytes > 0
            own = C.c_void_p()
            handle = C.create_string_buffer(64)
            _lib.check(lib.cur_p2p_alloc(self.nbytes, C.byref(own), handle), 'cur_p2p_alloc')
            self.own = own.value
            handles = [None] * n
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            for r in range(n):
                if r == self.rank:
                    self.ctx.region[r] = self.own
                    continue
                p = C.c_void_p()
                _lib.check(lib.cur_p2p_open(handles[r], C.byref(p)), 'cur_p2p_open')
                self.opened.append(p.value)
                self.ctx.region[r] = p.value
        self.error_flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        self.ctx.error_flag = self.error_flag.data_ptr()
        self.timeline = None
        if timeline_tiles > 0:
            self.timeline = torch.zeros(4 * int(timeline_tiles), dtype=torch.int64, device='cuda')
            self.ctx.timeline = self.timeline.data_ptr()
        torch.cuda.synchronize()
        dist.barrier(group=group)                       # every region is mapped and zeroed before anyone pushes

    def reset(self):
        """Collective: forget every word in flight (needed before the device step counter moves backwards, e.g. when a
        checkpoint is loaded into a running agent - the update number travelling with the data must never repeat)."""
        import torch.distributed as dist
        from . import _lib
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        if self._symm is not None:
            self._symm.zero_()
        else:
            _lib.check(self.lib.cur_p2p_zero(_lib.stream_ptr(), self.own, self.nbytes), 'cur_p2p_zero')
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def close(self):
        for p in self.opened:
            self.lib.cur_p2p_close(p)
        self.opened = []
        if self.own:
            self.lib.cur_p2p_free(self.own)
            self.own = None
