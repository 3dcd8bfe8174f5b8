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


def allreduce_sum_(t, comm=None):
    """In-place SUM over ranks - gradients are summed, NOT averaged (scale_grad_by_procs=False,
    ddpg.py:452-453).  Returns the world size."""
    group, n = world(comm)
    if n > 1:
        import torch.distributed as dist
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
