"""Network descriptors - drop-in names for reference baselines/her/actor_critic.py.

In the reference these classes build TF sub-graphs.  Here they only DESCRIBE the network (which input
layout, which flat parameter order); the computation is csrc/ddpg.cu.  `DDPG` resolves them through
`network_class` strings exactly like the reference (ddpg.py:63, config.py:60).
"""
import ctypes as C

from . import _lib


class _NetSpec:
    modular = False

    def __init__(self, dimo, dimg, dimu, max_u, hidden, layers, dimtd=0, normalize_obs=True, norm_clip=5.0,
                 kernel_hidden=None, **kwargs):
        """kernel_hidden: width the kernels run at (>= hidden).  The hand-written row / chain / action kernels exist for
        256 hidden units; a narrower network runs on them ZERO-PADDED: every arena (parameters, gradients, Adam moments)
        is laid out 256 wide, the columns / rows beyond `hidden` hold zeros and stay zero (their activations are
        relu(0) = 0, their gradients x * 0 = 0, Adam steps 0 / (0 + eps) = 0), and every sum only gains exact zeros - the
        result is the `hidden`-wide network's.  `var_shapes` / `ref_index` translate to the reference's flat order."""
        self.dimo, self.dimg, self.dimu, self.dimtd = dimo, dimg, dimu, (dimtd if self.modular else 0)
        self.max_u, self.hidden, self.layers = max_u, hidden, layers
        self.kernel_hidden = int(kernel_hidden) if kernel_hidden else hidden
        assert self.kernel_hidden >= hidden
        self.padded = self.kernel_hidden != hidden
        self.normalize_obs = normalize_obs
        self.desc = _lib.NetDesc(1 if self.modular else 0, dimo, dimg, dimu, self.dimtd, self.kernel_hidden, layers,
                                 float(max_u), 1 if normalize_obs else 0, float(norm_clip))
        lib = _lib.load()
        self.n_Q = lib.cur_net_param_count(C.byref(self.desc), 0)
        self.n_pi = lib.cur_net_param_count(C.byref(self.desc), 1)
        if self.n_Q < 0 or self.n_pi < 0:
            raise ValueError('unsupported network dimensions (hidden must be a multiple of 4, layers 1..8)')
        total = C.c_int64()
        self.pi_offset = lib.cur_theta_pi_offset(C.byref(self.desc), C.byref(total))
        self.arena = total.value

    def ref_index(self, which):
        """int64 array: position of every element of the reference's flat vector (GetFlat order, `hidden` wide) inside
        the kernel-side flat vector of the net (`kernel_hidden` wide)."""
        import numpy as np
        idx, base = [], 0
        for (s, ks) in zip(self.var_shapes(which), self.var_shapes(which, self.kernel_hidden)):
            if len(s) == 2:
                r, c = np.meshgrid(np.arange(s[0]), np.arange(s[1]), indexing='ij')
                idx.append((base + r * ks[1] + c).reshape(-1))
            else:
                idx.append(base + np.arange(s[0]))
            base += int(np.prod(ks))
        return np.concatenate(idx).astype(np.int64)

    def var_shapes(self, which, hidden=None):
        """Variable shapes in TF creation order == GetFlat order (util.py:56-107, tf_util.py:221-244)."""
        act = self.dimu if which == 'Q' else 0
        out = 1 if which == 'Q' else self.dimu
        H = self.hidden if hidden is None else hidden
        if self.modular:
            shapes = [(self.dimo + self.dimtd + act, H), (H,), (self.dimg, H)]
        else:
            shapes = [(self.dimo + self.dimg + act, H), (H,)]
        for _ in range(self.layers - 1):
            shapes += [(H, H), (H,)]
        shapes += [(H, out), (out,)]
        return shapes


class ActorCritic(_NetSpec):
    """Flat UVFA: pi([o|g]), Q([o|g|u/max_u])  (reference actor_critic.py:5-48)."""
    modular = False


class MultiTaskActorCritic(_NetSpec):
    """Modular UVFA: state branch [o|task_descr(|u/max_u)] + bias-free goal branch
    (reference actor_critic.py:51-98, util.py:73-107)."""
    modular = True
