"""Network descriptors - drop-in names for reference baselines/her/actor_critic.py.

In the reference these classes build TF sub-graphs.  Here they only DESCRIBE the network (which input
layout, which flat parameter order); the computation is csrc/ddpg.cu.  `DDPG` resolves them through
`network_class` strings exactly like the reference (ddpg.py:63, config.py:60).
"""
import ctypes as C

from . import _lib


class _NetSpec:
    modular = False

    def __init__(self, dimo, dimg, dimu, max_u, hidden, layers, dimtd=0, normalize_obs=True, norm_clip=5.0, **kwargs):
        self.dimo, self.dimg, self.dimu, self.dimtd = dimo, dimg, dimu, (dimtd if self.modular else 0)
        self.max_u, self.hidden, self.layers = max_u, hidden, layers
        self.normalize_obs = normalize_obs
        self.desc = _lib.NetDesc(1 if self.modular else 0, dimo, dimg, dimu, self.dimtd, hidden, layers, float(max_u),
                                 1 if normalize_obs else 0, float(norm_clip))
        lib = _lib.load()
        self.n_Q = lib.cur_net_param_count(C.byref(self.desc), 0)
        self.n_pi = lib.cur_net_param_count(C.byref(self.desc), 1)
        if self.n_Q < 0 or self.n_pi < 0:
            raise ValueError('unsupported network dimensions (hidden must be a multiple of 4, layers 1..8)')
        total = C.c_int64()
        self.pi_offset = lib.cur_theta_pi_offset(C.byref(self.desc), C.byref(total))
        self.arena = total.value

    def var_shapes(self, which):
        """Variable shapes in TF creation order == GetFlat order (util.py:56-107, tf_util.py:221-244)."""
        act = self.dimu if which == 'Q' else 0
        out = 1 if which == 'Q' else self.dimu
        H = self.hidden
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
