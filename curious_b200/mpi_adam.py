"""Flat-vector data-parallel Adam - drop-in for reference baselines/common/mpi_adam.py:6-50.

`var_list` is a list of float32 CUDA tensors that are contiguous views of one flat vector (in
GetFlat order, tf_util.py:221-244) or a single flat tensor.  The MPI collectives become NCCL
collectives on the device buffer (torch.distributed, one process per GPU): Allreduce(SUM) of the flat
gradient (mpi_adam.py:24-26), Bcast from rank 0 in sync() (mpi_adam.py:37-40) and the periodic
equality check of check_synced() (mpi_adam.py:42-50, every 100 updates).
"""
import numpy as np
import torch

from . import _lib
from .parallel import allreduce_sum_, assert_synced, broadcast_from_root_, world as _world


def adam_step_scale(stepsize, beta1, beta2, t):
    """mpi_adam.py:31 evaluated in float64 exactly like the reference."""
    return stepsize * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)


def flat_view(var_list):
    """Return the single flat tensor the variables are views of (they must tile it contiguously)."""
    if torch.is_tensor(var_list):
        return var_list.reshape(-1)
    if len(var_list) == 1:
        return var_list[0].reshape(-1)
    base = var_list[0]
    start = base.data_ptr()
    n = 0
    for v in var_list:
        assert v.is_contiguous() and v.dtype == torch.float32
        assert v.data_ptr() == start + 4 * n, 'var_list must be contiguous views of one flat vector'
        n += v.numel()
    storage_off = base.storage_offset()
    return torch.as_strided(base, (n,), (1,), storage_off)


class MpiAdam(object):
    def __init__(self, var_list, *, beta1=0.9, beta2=0.999, epsilon=1e-08, scale_grad_by_procs=True, comm=None,
                 m=None, v=None):
        self.var_list = var_list
        self.beta1 = beta1
        self.beta2 = beta2
        self.epsilon = epsilon
        self.scale_grad_by_procs = scale_grad_by_procs
        self.theta = flat_view(var_list)
        assert self.theta.is_cuda, 'MpiAdam works on device vectors (no CPU fallback)'
        assert self.theta.data_ptr() % 16 == 0, 'flat parameter vector must be 16-byte aligned'
        size = self.theta.numel()
        # optional caller-provided state views (DDPG keeps Q and pi moments in one arena so that a single
        # fused launch can step both nets)
        self.m = m if m is not None else torch.zeros(size, dtype=torch.float32, device=self.theta.device)
        self.v = v if v is not None else torch.zeros(size, dtype=torch.float32, device=self.theta.device)
        assert self.m.numel() == size and self.v.numel() == size
        self.t = 0
        self.comm = comm
        self.on_change = None      # optional callback: the owner (DDPG) invalidates state derived from theta

    # reference helpers (tf_util.GetFlat / SetFromFlat)
    def getflat(self):
        return self.theta.detach().cpu().numpy().copy()

    def setfromflat(self, theta):
        self.theta.copy_(torch.from_numpy(np.ascontiguousarray(theta, dtype=np.float32)).to(self.theta.device))
        if self.on_change:
            self.on_change()

    def _grad_tensor(self, localg):
        if torch.is_tensor(localg):
            g = localg.reshape(-1)
            if g.dtype != torch.float32 or not g.is_cuda:
                g = g.to(self.theta.device, torch.float32)
            return g
        return torch.from_numpy(np.ascontiguousarray(localg, dtype=np.float32).reshape(-1)).to(self.theta.device)

    def update(self, localg, stepsize):
        if self.t % 100 == 0:
            self.check_synced()
        g = self._grad_tensor(localg)                       # localg.astype('float32')
        group, world = _world(self.comm)
        if world > 1:
            if g.data_ptr() == (localg.data_ptr() if torch.is_tensor(localg) else 0):
                g = g.clone()                               # the reference leaves localg untouched
            allreduce_sum_(g, self.comm)                    # mpi_adam.py:26
        self.t += 1
        a = adam_step_scale(stepsize, self.beta1, self.beta2, self.t)
        grad_div = float(world) if self.scale_grad_by_procs else 1.0  # mpi_adam.py:27-28
        _lib.check(_lib.load().cur_adam_step(_lib.stream_ptr(), self.theta.data_ptr(), g.data_ptr(),
                                             self.m.data_ptr(), self.v.data_ptr(), self.theta.numel(),
                                             float(np.float32(-a)), self.beta1, self.beta2, self.epsilon, grad_div),
                   'cur_adam_step')
        if self.on_change:
            self.on_change()

    def sync(self):
        broadcast_from_root_(self.theta, self.comm)
        if self.on_change:
            self.on_change()

    def checksum(self):
        out = torch.zeros(1, dtype=torch.int64, device=self.theta.device)
        _lib.check(_lib.load().cur_checksum(_lib.stream_ptr(), self.theta.data_ptr(), self.theta.numel(),
                                            out.data_ptr()), 'cur_checksum')
        return out

    def check_synced(self):
        """All ranks must hold bit-identical parameters (mpi_adam.py:42-50).  Instead of broadcasting the
        whole vector, a 64-bit order-independent checksum of the bit patterns is compared with rank 0's."""
        group, world = _world(self.comm)
        if world <= 1:
            return
        assert_synced(self.checksum(), self.comm)
