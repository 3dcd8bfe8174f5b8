"""A stand-in `tensorflow` (1.x graph API subset) that lets the UNMODIFIED reference graph code run - TEST INFRASTRUCTURE.

The reference builds its networks, losses, gradients, Adam plumbing, normaliser statistics and target-network updates as a
TF1 graph (baselines/her/ddpg.py:362-462, actor_critic.py, util.py:56-107, normalizer.py:33-60, common/tf_util.py:200-246,
common/mpi_adam.py).  TensorFlow 1.x cannot be installed here, so that source could only be RESTATED (oracle/ddpg_oracle.py).
This module closes the gap from the other side: it implements the few dozen TF1 entry points those files call - lazily
evaluated graph nodes, variables and variable scopes with TF's naming and reuse rules, `tf.layers.dense`, collections,
`tf.gradients`, sessions with feed dicts, `StagingArea` - on top of torch CPU tensors, so that the reference's OWN source
builds and runs its graph.  What is emulated is the semantics of TF primitives (matmul, relu, tanh, mean, clip, concat,
assign, reverse-mode gradients), not anything of the reference: which tensors are concatenated in which order, which
scopes share variables, the loss formulas, the clip range, the gradient flattening order and the polyak rule all come from the
reference files executed as they lie.

`install()` puts the stand-in (plus stub `mpi4py`, `gym`) into sys.modules and returns it; after that
`import baselines.her.ddpg` works from /root/reference.  Arithmetic runs in float64 by default (DTYPE) - the reference
graph evaluated more precisely than any float32 implementation, which is what a tolerance test wants as its centre; variables
hold float32-representable values like TF's float32 variables do, and `Session.run` returns float32 arrays like TF
(`Session.run64` returns the float64 values for the fixture generator).

Only tests/ and oracle/gen_golden_ddpg.py use this file; the product never imports it.
"""
import contextlib
import re
import sys
import types

import numpy as np
import torch

DTYPE = torch.float64


# ------------------------------------------------------------------------------------------------ dtypes / shapes
class _DType(object):
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return 'tf.' + self.name


float32 = _DType('float32')


class TensorShape(object):
    def __init__(self, dims):
        self.dims = None if dims is None else list(dims)

    def as_list(self):
        if self.dims is None:
            raise ValueError('as_list() is not defined on an unknown TensorShape')
        return list(self.dims)

    def __len__(self):
        if self.dims is None:
            raise ValueError('unknown rank')
        return len(self.dims)

    def __iter__(self):
        return iter(self.as_list())

    def __repr__(self):
        return 'TensorShape(%r)' % (self.dims,)


# ------------------------------------------------------------------------------------------------ graph state
class _Graph(object):
    def __init__(self):
        self.variables = []          # creation order = collection order
        self.by_name = {}
        self.scopes = []             # stack of _Scope
        self.init_rng = np.random.RandomState(0)


_graph = _Graph()
_default_session = [None]


def reset_default_graph():
    global _graph
    _graph = _Graph()
    _default_session[0] = None


def set_random_seed(seed):
    _graph.init_rng = np.random.RandomState(seed)


class _Scope(object):
    def __init__(self, name, reuse=False):
        self.name = name
        self.reuse = reuse

    def reuse_variables(self):
        self.reuse = True


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    parent = _graph.scopes[-1] if _graph.scopes else None
    full = (parent.name + '/' + name) if parent and parent.name else name
    sc = _Scope(full, bool(reuse) or (parent.reuse if parent else False))
    _graph.scopes.append(sc)
    try:
        yield sc
    finally:
        _graph.scopes.pop()


def _scope_name():
    return _graph.scopes[-1].name if _graph.scopes else ''


def _scope_reuse():
    return _graph.scopes[-1].reuse if _graph.scopes else False


# ------------------------------------------------------------------------------------------------ evaluation context
class _Ctx(object):
    def __init__(self, feed, override=None):
        self.feed = feed                  # {Tensor: torch tensor}
        self.override = override or {}    # {Variable: torch tensor} (leaves of a gradient evaluation)
        self.memo = {}


def _to_torch(x):
    if isinstance(x, torch.Tensor):
        return x.to(DTYPE)
    return torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=DTYPE)


# ------------------------------------------------------------------------------------------------ nodes
class Tensor(object):
    """A lazily evaluated graph node."""
    dtype = float32

    def __init__(self, fn, shape=None, name=None):
        self._fn = fn
        self._shape = None if shape is None else list(shape)
        self.name = name

    def _eval(self, ctx):
        if self in ctx.feed:
            return ctx.feed[self]
        k = id(self)
        if k not in ctx.memo:
            ctx.memo[k] = self._fn(ctx)
        return ctx.memo[k]

    def get_shape(self):
        return TensorShape(self._shape)

    @property
    def shape(self):
        return TensorShape(self._shape)

    def eval(self, feed_dict=None, session=None):
        return (session or get_default_session()).run(self, feed_dict)

    # arithmetic
    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return subtract(self, o)
    def __rsub__(self, o): return subtract(o, self)
    def __mul__(self, o): return multiply(self, o)
    def __rmul__(self, o): return multiply(o, self)
    def __truediv__(self, o): return divide(self, o)
    def __rtruediv__(self, o): return divide(o, self)
    def __neg__(self): return _unary(lambda x: -x, self)

    def __getitem__(self, idx):
        return Tensor(lambda c: self._eval(c)[idx])

    __hash__ = object.__hash__


def _wrap(x):
    if isinstance(x, Tensor):
        return x
    t = _to_torch(x)
    return Tensor(lambda c: t, shape=list(t.shape))


def _bshape(a, b):
    sa, sb = a._shape, b._shape
    if sa is None or sb is None:
        return sa if sb is None else sb if sa is None else None
    return sa if len(sa) >= len(sb) else sb


def _binary(f, a, b):
    a, b = _wrap(a), _wrap(b)
    return Tensor(lambda c: f(a._eval(c), b._eval(c)), shape=_bshape(a, b))


def _unary(f, a):
    a = _wrap(a)
    return Tensor(lambda c: f(a._eval(c)), shape=a._shape)


def add(a, b, name=None): return _binary(lambda x, y: x + y, a, b)
def subtract(a, b, name=None): return _binary(lambda x, y: x - y, a, b)
def multiply(a, b, name=None): return _binary(lambda x, y: x * y, a, b)
def divide(a, b, name=None): return _binary(lambda x, y: x / y, a, b)
def maximum(a, b, name=None): return _binary(torch.maximum, a, b)
def minimum(a, b, name=None): return _binary(torch.minimum, a, b)
def square(a, name=None): return _unary(lambda x: x * x, a)
def sqrt(a, name=None): return _unary(torch.sqrt, a)
def tanh(a, name=None): return _unary(torch.tanh, a)
def sin(a, name=None): return _unary(torch.sin, a)
def stop_gradient(a, name=None): return _unary(lambda x: x.detach(), a)
def zeros_like(a, name=None): return _unary(torch.zeros_like, a)
def cast(a, dtype, name=None): return _wrap(a) if isinstance(a, Tensor) else _wrap(a)


def constant(value, dtype=None, shape=None, name=None):
    return _wrap(value)


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    lo, hi = float(clip_value_min), float(clip_value_max)
    return _unary(lambda x: torch.clamp(x, min=lo, max=hi), t)


def reduce_mean(t, axis=None, name=None):
    t = _wrap(t)
    if axis is None:
        return Tensor(lambda c: t._eval(c).mean(), shape=[])
    return Tensor(lambda c: t._eval(c).mean(dim=axis))


def reduce_sum(t, axis=None, name=None):
    t = _wrap(t)
    if axis is None:
        return Tensor(lambda c: t._eval(c).sum(), shape=[])
    return Tensor(lambda c: t._eval(c).sum(dim=axis))


def reshape(t, shape, name=None):
    t = _wrap(t)
    shape = [int(s) for s in shape]
    return Tensor(lambda c: t._eval(c).reshape(shape), shape=[None if s < 0 else s for s in shape])


def concat(values, axis, name=None):
    if isinstance(values, int):                  # tf.concat(axis=1, values=[...]) is always called by keyword; be safe
        values, axis = axis, values
    vals = [_wrap(v) for v in values]
    shape = None
    if all(v._shape is not None for v in vals):
        shape = list(vals[0]._shape)
        dims = [v._shape[axis] for v in vals]
        shape[axis] = None if any(d is None for d in dims) else sum(dims)
    return Tensor(lambda c: torch.cat([v._eval(c) for v in vals], dim=axis), shape=shape)


class _NN(object):
    @staticmethod
    def relu(t, name=None):
        return _unary(torch.relu, t)


nn = _NN()


# ------------------------------------------------------------------------------------------------ variables
def _f32(t):
    """TF variables are float32: round whatever is assigned."""
    return t.to(torch.float32).to(DTYPE)


class Operation(object):
    def __init__(self, fn):
        self._fn = fn            # ctx -> list of (variable, new value) | None

    def _prepare(self, ctx):
        return self._fn(ctx) or []

    def run(self, feed_dict=None, session=None):
        (session or get_default_session()).run(self, feed_dict)


class Variable(Tensor):
    def __init__(self, value, name, trainable=True):
        self.value = _f32(_to_torch(value))
        Tensor.__init__(self, None, shape=list(self.value.shape), name=name + ':0')
        self.trainable = trainable
        self._initial = self.value.clone()
        self.initializer = Operation(lambda c: [(self, self._initial)])
        if self.name in _graph.by_name:
            raise ValueError('Variable %s already exists, disallowed. Did you mean to set reuse=True?' % name)
        _graph.variables.append(self)
        _graph.by_name[self.name] = self

    def _eval(self, ctx):
        if self in ctx.feed:
            return ctx.feed[self]
        return ctx.override.get(self, self.value)

    def assign(self, value):
        value = _wrap(value)
        return Operation(lambda c: [(self, value._eval(c))])

    def assign_add(self, delta):
        delta = _wrap(delta)
        return Operation(lambda c: [(self, self._eval(c) + delta._eval(c))])

    def load(self, value, session=None):
        v = _to_torch(value)
        assert list(v.shape) == list(self.value.shape), (self.name, v.shape, self.value.shape)
        self.value = _f32(v)


def assign(ref, value, name=None):
    return ref.assign(value)


def group(*ops, **kw):
    return Operation(lambda c: [p for op in ops for p in op._prepare(c)])


def variables_initializer(var_list, name=None):
    var_list = list(var_list)
    return Operation(lambda c: [(v, v._initial) for v in var_list])


def zeros_initializer():
    return lambda shape: np.zeros(shape, np.float32)


def ones_initializer():
    return lambda shape: np.ones(shape, np.float32)


def _xavier_initializer(uniform=True, seed=None, dtype=None):
    """tf.contrib.layers.xavier_initializer: U(-l, l), l = sqrt(6 / (fan_in + fan_out))."""
    def init(shape):
        fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return _graph.init_rng.uniform(-lim, lim, size=shape).astype(np.float32)
    return init


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
    scope = _scope_name()
    full = (scope + '/' + name) if scope else name
    if _scope_reuse():
        if full + ':0' not in _graph.by_name:
            raise ValueError('Variable %s does not exist, or was not created with tf.get_variable()' % full)
        return _graph.by_name[full + ':0']
    shape = [int(s) for s in shape]
    init = initializer if initializer is not None else _xavier_initializer()
    return Variable(init(shape), full, trainable=trainable)


class GraphKeys(object):
    TRAINABLE_VARIABLES = 'trainable_variables'
    GLOBAL_VARIABLES = 'variables'


def get_collection(key, scope=None):
    vs = [v for v in _graph.variables if key == GraphKeys.GLOBAL_VARIABLES or v.trainable]
    if scope is None:
        return vs
    rx = re.compile(scope)
    return [v for v in vs if rx.match(v.name)]


def global_variables():
    return list(_graph.variables)


def trainable_variables():
    return [v for v in _graph.variables if v.trainable]


# ------------------------------------------------------------------------------------------------ layers
class _Layers(object):
    @staticmethod
    def dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, reuse=None, name=None, **kw):
        """tf.layers.dense: variables `<scope>/<name>/kernel` [in, units] then `<scope>/<name>/bias` [units]."""
        inputs = _wrap(inputs)
        in_dim = inputs._shape[-1]
        assert in_dim is not None, 'dense needs a static input width'
        with variable_scope(name, reuse=reuse):
            kernel = get_variable('kernel', shape=[in_dim, units], initializer=kernel_initializer)
            bias = get_variable('bias', shape=[units], initializer=zeros_initializer()) if use_bias else None

        def fn(c):
            y = inputs._eval(c) @ kernel._eval(c)
            return y + bias._eval(c) if bias is not None else y
        out = Tensor(fn, shape=list(inputs._shape[:-1]) + [units])
        return activation(out) if activation else out


layers = _Layers()


# ------------------------------------------------------------------------------------------------ gradients
def gradients(ys, xs, name=None):
    """Reverse-mode gradients of a scalar node with respect to variables (torch autograd over a re-evaluation of the
    node with the variables as leaves)."""
    xs = list(xs)
    key = ('grad', id(ys), tuple(id(x) for x in xs))

    def all_grads(c):
        if key not in c.memo:
            leaves = {x: x._eval(c).detach().clone().requires_grad_(True) for x in xs}
            ov = dict(c.override)
            ov.update(leaves)
            sub = _Ctx(c.feed, ov)
            y = ys._eval(sub)
            g = torch.autograd.grad(y, [leaves[x] for x in xs], allow_unused=True)
            c.memo[key] = [None if gi is None else gi.detach() for gi in g]
        return c.memo[key]

    out = []
    for i, x in enumerate(xs):
        def fn(c, i=i, x=x):
            g = all_grads(c)[i]
            return torch.zeros_like(x._eval(c)) if g is None else g
        out.append(Tensor(fn, shape=x._shape))
    return out


# ------------------------------------------------------------------------------------------------ placeholders, staging
def placeholder(dtype, shape=None, name=None):
    def fn(c):
        raise ValueError('You must feed a value for placeholder tensor %r' % (name,))
    return Tensor(fn, shape=shape, name=name)


class StagingArea(object):
    """tensorflow.contrib.staging.StagingArea as the reference uses it: one put, one get, values held in between."""

    def __init__(self, dtypes, shapes=None, **kw):
        self._shapes = list(shapes)
        self._held = None

    def put(self, values):
        values = [_wrap(v) for v in values]

        def fn(c):
            self._held = [v._eval(c) for v in values]
        return Operation(fn)

    def get(self):
        def make(i):
            def fn(c):
                assert self._held is not None, 'StagingArea.get() before put()'
                return self._held[i]
            return Tensor(fn, shape=self._shapes[i])
        return [make(i) for i in range(len(self._shapes))]


# ------------------------------------------------------------------------------------------------ sessions
class Session(object):
    def __init__(self, *a, **kw):
        pass

    def __enter__(self):
        self._prev = _default_session[0]
        _default_session[0] = self
        return self

    def __exit__(self, *a):
        _default_session[0] = self._prev

    def _run(self, fetches, feed_dict, np_dtype):
        ctx = _Ctx({k: _to_torch(v) for k, v in (feed_dict or {}).items()})
        pending = []

        def fetch(f):
            if isinstance(f, (list, tuple)):
                return [fetch(x) for x in f]
            if isinstance(f, Operation):
                pending.extend(f._prepare(ctx))
                return None
            return f._eval(ctx).detach().cpu().numpy().astype(np_dtype)
        out = fetch(fetches)
        for var, val in pending:                 # every value was computed from the variables as they were
            var.value = _f32(val.detach().reshape(var.value.shape))
        return out

    def run(self, fetches, feed_dict=None):
        return self._run(fetches, feed_dict, np.float32)

    def run64(self, fetches, feed_dict=None):
        """Shim-only: the float64 values (fixture generation)."""
        return self._run(fetches, feed_dict, np.float64)

    def close(self):
        pass


class InteractiveSession(Session):
    def __init__(self, *a, **kw):
        Session.__init__(self)
        _default_session[0] = self


def get_default_session():
    return _default_session[0]


# ------------------------------------------------------------------------------------------------ installation
class _FakeComm(object):
    """MPI.COMM_WORLD of a one-process job."""

    def Get_rank(self): return 0
    def Get_size(self): return 1

    def Allreduce(self, src, dst, op=None):
        dst[...] = src

    def Bcast(self, buf, root=0):
        pass

    def Abort(self):
        raise SystemExit(1)


def install(reference_root='/root/reference'):
    """Put this module into sys.modules as `tensorflow` (+ contrib.staging / contrib.layers), stub `mpi4py` and `gym`,
    make the reference importable.  Returns this module."""
    me = sys.modules[__name__]
    contrib = types.ModuleType('tensorflow.contrib')
    staging = types.ModuleType('tensorflow.contrib.staging')
    staging.StagingArea = StagingArea
    clayers = types.ModuleType('tensorflow.contrib.layers')
    clayers.xavier_initializer = _xavier_initializer
    contrib.staging, contrib.layers = staging, clayers
    me.contrib = contrib
    sys.modules['tensorflow'] = me
    sys.modules['tensorflow.contrib'] = contrib
    sys.modules['tensorflow.contrib.staging'] = staging
    sys.modules['tensorflow.contrib.layers'] = clayers
    mpi = types.ModuleType('mpi4py')
    mpi.MPI = types.SimpleNamespace(COMM_WORLD=_FakeComm(), SUM='sum')
    sys.modules['mpi4py'] = mpi
    if 'gym' not in sys.modules:
        gym = types.ModuleType('gym')                    # baselines/common/misc_util.py imports it at module level
        gym.Env = type('Env', (), {})
        gym.Wrapper = type('Wrapper', (), {})
        sys.modules['gym'] = gym
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return me
