"""Thin glue kept from reference baselines/her/util.py (only what the hot path's callers use)."""
import contextlib
import functools
import gc
import importlib
import inspect

import numpy as np


def store_args(method):
    """Stores provided method args as instance attributes (reference util.py:13-37)."""
    argspec = inspect.getfullargspec(method)
    defaults = {}
    if argspec.defaults is not None:
        defaults = dict(zip(argspec.args[-len(argspec.defaults):], argspec.defaults))
    if argspec.kwonlydefaults is not None:
        defaults.update(argspec.kwonlydefaults)
    arg_names = argspec.args[1:]

    @functools.wraps(method)
    def wrapper(*positional_args, **keyword_args):
        self = positional_args[0]
        args = defaults.copy()
        for name, value in zip(arg_names, positional_args[1:]):
            args[name] = value
        args.update(keyword_args)
        self.__dict__.update(args)
        return method(*positional_args, **keyword_args)

    return wrapper


# the reference's plugin strings resolve to the B200 implementations (config.py:60,84)
_ALIASES = {
    'baselines.her.actor_critic': 'curious_b200.actor_critic',
    'baselines.her.her': 'curious_b200.her',
}


def import_function(spec):
    """Import a function identified by a string like "pkg.module:fn_name" (reference util.py:40-46).
    Reference module paths are mapped onto their curious_b200 drop-ins."""
    mod_name, fn_name = spec.split(':')
    mod_name = _ALIASES.get(mod_name, mod_name)
    module = importlib.import_module(mod_name)
    return getattr(module, fn_name)


def convert_episode_to_batch_major(episode):
    """Time-major lists -> batch-major arrays (reference util.py:174-184)."""
    episode_batch = {}
    for key in episode.keys():
        val = np.array(episode[key]).copy()
        episode_batch[key] = val.swapaxes(0, 1)
    return episode_batch


def transitions_in_episode_batch(episode_batch):
    """Number of transitions in a given episode batch (reference util.py:187-191)."""
    shape = episode_batch['u'].shape
    return shape[0] * shape[1]


def dims_to_shapes(input_dims):
    return {key: tuple([val]) if val > 0 else tuple() for key, val in input_dims.items()}


class LazyHost:
    """A device result that turns into a host value only when somebody looks at it.

    `DDPG.train()` returns (critic_loss, actor_loss) like the reference (ddpg.py:368-373); the reference's
    own loop ignores them (train.py:152-153), so forcing a device sync per update would only cost time.
    float(x), np.asarray(x), x.item() and arithmetic all work and synchronise on demand."""

    def __init__(self, tensor, still_valid=None):
        self._tensor = tensor
        # optional callable: False once the device buffer behind `tensor` has been reused by a later update (the CUDA-graph
        # path of DDPG.train hands out views of its fixed output buffers); reading then raises instead of returning
        # another update's numbers
        self._still_valid = still_valid

    @property
    def tensor(self):
        if self._still_valid is not None and not self._still_valid():
            raise RuntimeError('this train() result was overwritten by a later update: read it before the next '
                               'train() call, or build the agent with own_train_outputs=True')
        return self._tensor

    def numpy(self):
        return self.tensor.detach().cpu().numpy()

    def item(self):
        return self.tensor.item()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __float__(self):
        return float(self.tensor.item())

    def __repr__(self):
        return 'LazyHost(%r)' % (self.numpy(),)

    def __add__(self, o): return self.numpy() + o
    def __radd__(self, o): return o + self.numpy()
    def __sub__(self, o): return self.numpy() - o
    def __rsub__(self, o): return o - self.numpy()
    def __mul__(self, o): return self.numpy() * o
    def __rmul__(self, o): return o * self.numpy()
    def __truediv__(self, o): return self.numpy() / o

    @property
    def shape(self):
        return tuple(self.tensor.shape)


@contextlib.contextmanager
def capture_graph(graph):
    """torch.cuda.graph(graph) with the Python garbage collector held off: a CUDAGraph of a dropped agent that the
    collector frees while this capture is open calls cudaGraphExecDestroy on the capturing thread, which CUDA answers
    by invalidating the capture (cudaErrorStreamCaptureInvalidated)."""
    import torch
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph):
            yield
    finally:
        if was_enabled:
            gc.enable()
