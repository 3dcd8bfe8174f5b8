"""NumPy float32 restatement of Normalizer / networks / DDPG graph / MpiAdam
(test infrastructure; see oracle/__init__.py).

Pinning (the reference has no tests for baselines/her/, and TensorFlow 1.x / mpi4py cannot be installed here):
  * PINNED against the reference's own code executed live (tests/test_reference_live.py cuts the NumPy-only methods out of
    the unmodified ddpg.py / normalizer.py and runs them on recording stubs): sample_batch incl. the LP apportioning,
    concatenation and shuffle, store_episode routing and its normaliser batch, get_actions post-processing,
    _preprocess_og, Normalizer.update / synchronize / snapshot-and-reset, MpiAdam.update (with NumPy-1 scalar casting
    reproduced at the boundary);
  * PINNED against the reference's own graph code (round 2): the unmodified ddpg.py / actor_critic.py / util.py /
    normalizer.py / tf_util.py / mpi_adam.py build and run their TF1 graph over oracle/tf1_shim.py (a stand-in `tensorflow`
    whose primitives are torch CPU ops); oracle/gen_golden_ddpg.py records that agent's losses, gradients, parameters after
    MpiAdam + polyak updates, get_actions outputs, weight files and whole-agent trajectories as tests/golden/ddpg/*.npz;
    tests/test_reference_graph.py walks this file through them and runs the reference agent live beside it.
    tests/test_host_logic.py additionally cross-checks the hand-written backward pass against torch autograd and Adam
    against the reference's `test_MpiAdam` problem definition (mpi_adam.py:54-63).
Each function cites the reference lines it restates.

Conventions
-----------
* every array that is a TF float32 tensor/variable in the reference is np.float32 here;
* a "net" is the list of variables in TF creation order, which is also the GetFlat order
  (reference util.py:49-53, tf_util.py:221-244):
    modular (nn_modular_her, util.py:73-107):  W0s[in_s,H], b0[H], W0g[dimg,H], W1, b1, ..., Wout, bout
    flat    (nn,             util.py:56-71):   W0[in,H], b0, W1, b1, ..., Wout, bout
* collectives are passed in as callables so a world of ranks can be emulated in-process.
"""
import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------------------
# Normalizer  (reference baselines/her/normalizer.py:10-118)
# --------------------------------------------------------------------------------------
class NormalizerOracle:
    def __init__(self, size, eps=1e-2, default_clip_range=np.inf, mean_over_ranks=None):
        self.size, self.eps, self.default_clip_range = size, eps, default_clip_range
        self.local_sum = np.zeros(size, f32)           # normalizer.py:27-29
        self.local_sumsq = np.zeros(size, f32)
        self.local_count = np.zeros(1, f32)
        self.sum = np.zeros(size, f32)                 # normalizer.py:31-45 (count starts at ONE)
        self.sumsq = np.zeros(size, f32)
        self.count = np.ones(1, f32)
        self.mean = np.zeros(size, f32)
        self.std = np.ones(size, f32)
        # normalizer.py:84-94: Allreduce(SUM) / world_size of each partial
        self.mean_over_ranks = mean_over_ranks or (lambda x: x)

    def update(self, v):
        v = v.reshape(-1, self.size)                   # normalizer.py:64-70
        self.local_sum += v.sum(axis=0)
        self.local_sumsq += (np.square(v)).sum(axis=0)
        self.local_count[0] += v.shape[0]

    def recompute_stats(self):
        lc, ls, lq = self.local_count.copy(), self.local_sum.copy(), self.local_sumsq.copy()
        self.local_count[...] = 0
        self.local_sum[...] = 0
        self.local_sumsq[...] = 0
        ls = self.mean_over_ranks(ls)
        lq = self.mean_over_ranks(lq)
        lc = self.mean_over_ranks(lc)
        self.count = (self.count + lc).astype(f32)     # normalizer.py:50-54
        self.sum = (self.sum + ls).astype(f32)
        self.sumsq = (self.sumsq + lq).astype(f32)
        self.mean = (self.sum / self.count).astype(f32)            # normalizer.py:55-61
        var = self.sumsq / self.count - np.square(self.sum / self.count)
        self.std = np.sqrt(np.maximum(np.square(f32(self.eps)), var)).astype(f32)

    def normalize(self, v, clip_range=None):
        if clip_range is None:
            clip_range = self.default_clip_range
        v = np.asarray(v, f32)                         # normalizer.py:72-77
        return np.clip((v - self.mean) / self.std, -f32(clip_range), f32(clip_range)).astype(f32)

    def denormalize(self, v):
        return (self.mean + np.asarray(v, f32) * self.std).astype(f32)


# --------------------------------------------------------------------------------------
# Networks  (reference baselines/her/util.py:56-107, actor_critic.py:5-98)
# --------------------------------------------------------------------------------------
def net_shapes(modular, in_state, in_goal, hidden, layers, out):
    """Variable shapes in creation (= flat) order."""
    shapes = []
    if modular:
        shapes += [(in_state, hidden), (hidden,), (in_goal, hidden)]   # _0_state k,b ; _0_goal k (no bias)
    else:
        shapes += [(in_state + in_goal, hidden), (hidden,)]
    for _ in range(layers - 1):
        shapes += [(hidden, hidden), (hidden,)]
    shapes += [(hidden, out), (out,)]
    return shapes


def xavier_uniform_net(rng, shapes):
    """tf.contrib.layers.xavier_initializer() (uniform, +-sqrt(6/(fan_in+fan_out))), zero biases
    (util.py:63).  Drawn from `rng` (np.random.RandomState) in flat order."""
    net = []
    for s in shapes:
        if len(s) == 2:
            lim = np.sqrt(6.0 / (s[0] + s[1]))
            net.append(rng.uniform(-lim, lim, s).astype(f32))
        else:
            net.append(np.zeros(s, f32))
    return net


def flatten(net):
    return np.concatenate([v.reshape(-1) for v in net]).astype(f32)


def unflatten(flat, shapes):
    out, k = [], 0
    for s in shapes:
        n = int(np.prod(s))
        out.append(np.asarray(flat[k:k + n], f32).reshape(s).copy())
        k += n
    assert k == flat.size
    return out


def mlp_forward(net, modular, x_state, x_goal):
    """Returns (output, cache).  relu on all but the last layer."""
    if modular:
        W0s, b0, W0g = net[0], net[1], net[2]
        pre = x_state @ W0s + b0 + x_goal @ W0g        # util.py:79-91 (goal branch has no bias)
        rest = net[3:]
        x0 = (x_state, x_goal)
    else:
        x = np.concatenate([x_state, x_goal], axis=1)
        pre = x @ net[0] + net[1]
        rest = net[2:]
        x0 = (x,)
    acts = []
    margin = float(np.abs(pre).min())      # distance of the closest hidden pre-activation to the ReLU kink
    h = np.maximum(pre, 0).astype(f32)
    acts.append(h)
    n_rest = len(rest) // 2
    for i in range(n_rest):
        W, b = rest[2 * i], rest[2 * i + 1]
        pre = h @ W + b
        if i < n_rest - 1:
            margin = min(margin, float(np.abs(pre).min()))
            h = np.maximum(pre, 0).astype(f32)
            acts.append(h)
        else:
            h = pre.astype(f32)
    return h, (x0, acts, margin)


def mlp_backward(net, modular, cache, dout, want_param_grads=True):
    """Backprop `dout` (dL/d output).  Returns (grads in flat order or None, dx_state, dx_goal)."""
    x0, acts = cache[0], cache[1]
    rest = net[3:] if modular else net[2:]
    n_rest = len(rest) // 2
    grads_rest = [None] * len(rest)
    d = dout.astype(f32)
    for i in reversed(range(n_rest)):
        W = rest[2 * i]
        h_in = acts[i]
        if want_param_grads:
            grads_rest[2 * i] = (h_in.T @ d).astype(f32)
            grads_rest[2 * i + 1] = d.sum(axis=0).astype(f32)
        d = (d @ W.T).astype(f32)
        d = (d * (h_in > 0)).astype(f32)
    if modular:
        xs, xg = x0
        g0 = [(xs.T @ d).astype(f32), d.sum(axis=0).astype(f32), (xg.T @ d).astype(f32)] \
            if want_param_grads else None
        dxs = (d @ net[0].T).astype(f32)
        dxg = (d @ net[2].T).astype(f32)
    else:
        (x,) = x0
        g0 = [(x.T @ d).astype(f32), d.sum(axis=0).astype(f32)] if want_param_grads else None
        dx = (d @ net[0].T).astype(f32)
        dxs, dxg = dx, None
    grads = (g0 + grads_rest) if want_param_grads else None
    return grads, dxs, dxg


class ActorCriticOracle:
    """ActorCritic (flat, actor_critic.py:5-48) / MultiTaskActorCritic (actor_critic.py:51-98)."""

    def __init__(self, modular, dimo, dimg, dimu, dimtd, hidden, layers, max_u, normalize_obs,
                 o_stats, g_stats):
        self.modular, self.dimo, self.dimg, self.dimu, self.dimtd = modular, dimo, dimg, dimu, dimtd
        self.hidden, self.layers, self.max_u = hidden, layers, f32(max_u)
        self.normalize_obs, self.o_stats, self.g_stats = normalize_obs, o_stats, g_stats
        if modular:
            self.pi_shapes = net_shapes(True, dimo + dimtd, dimg, hidden, layers, dimu)
            self.Q_shapes = net_shapes(True, dimo + dimtd + dimu, dimg, hidden, layers, 1)
        else:
            self.pi_shapes = net_shapes(False, dimo, dimg, hidden, layers, dimu)
            self.Q_shapes = net_shapes(False, dimo + dimg + dimu, 0, hidden, layers, 1)

    def inputs(self, o, g, td):
        o = np.asarray(o, f32)
        g = np.asarray(g, f32)
        if self.normalize_obs:                         # actor_critic.py:31-36 / 76-83
            o = self.o_stats.normalize(o)
            g = self.g_stats.normalize(g)
        return o, g, (None if td is None else np.asarray(td, f32))

    def pi(self, pi_net, o, g, td):
        """Returns (pi, cache) with pi = max_u * tanh(net)  (actor_critic.py:41-42 / 88-90)."""
        if self.modular:
            xs, xg = np.concatenate([o, td], axis=1), g
        else:
            xs, xg = o, g
        y, cache = mlp_forward(pi_net, self.modular, xs, xg)
        th = np.tanh(y).astype(f32)
        return (self.max_u * th).astype(f32), (cache, th)

    def Q(self, Q_net, o, g, td, u):
        """Q(o, g, u / max_u)  (actor_critic.py:44-48 / 92-98)."""
        a = (np.asarray(u, f32) / self.max_u).astype(f32)
        if self.modular:
            xs, xg = np.concatenate([o, td, a], axis=1), g
        else:
            xs, xg = np.concatenate([o, g, a], axis=1), np.zeros((o.shape[0], 0), f32)
        return mlp_forward(Q_net, self.modular, xs, xg)


def ddpg_losses_and_grads(ac, main_Q, main_pi, target_Q, target_pi, batch, gamma, clip_return,
                          clip_pos_returns, action_l2):
    """One evaluation of DDPG._grads (reference ddpg.py:235-243, graph at ddpg.py:412-449).

    batch: dict with o, g, u, task_descr (or None), o_2, g_2, r  - float32 [B, .]
    Returns dict(Q_loss, pi_loss, Q_pi, Q, target, Q_grad(flat), pi_grad(flat)).
    """
    B = batch['o'].shape[0]
    td = batch.get('task_descr')
    o, g, td = ac.inputs(batch['o'], batch['g'], td)
    o2, g2, _ = ac.inputs(batch['o_2'], batch['g_2'], td)
    u = np.asarray(batch['u'], f32)
    r = np.asarray(batch['r'], f32).reshape(-1, 1)

    # main network (ddpg.py:417-421)
    pi, (pi_cache, th) = ac.pi(main_pi, o, g, td)
    Q_pi, Qpi_cache = ac.Q(main_Q, o, g, td, pi)
    Q, Q_cache = ac.Q(main_Q, o, g, td, u)
    # target network on (o_2, g_2), same u and td (ddpg.py:422-431); only Q_pi is used
    pi_t, _ = ac.pi(target_pi, o2, g2, td)
    Q_pi_t, _ = ac.Q(target_Q, o2, g2, td, pi_t)

    hi = f32(0.) if clip_pos_returns else f32(np.inf)
    target = np.clip(r + f32(gamma) * Q_pi_t, -f32(clip_return), hi).astype(f32)   # ddpg.py:436-438
    diff = (target - Q).astype(f32)
    Q_loss = np.mean(np.square(diff), dtype=f32)                                   # ddpg.py:439
    pi_loss = -np.mean(Q_pi, dtype=f32) + f32(action_l2) * np.mean(np.square(pi / ac.max_u), dtype=f32)

    # critic gradient wrt main/Q only (ddpg.py:442)
    dQ = (f32(-2.0 / B) * diff).astype(f32)
    Q_grads, _, _ = mlp_backward(main_Q, ac.modular, Q_cache, dQ)
    # actor gradient wrt main/pi only; flows through Q's weights via Q_pi (ddpg.py:440-443)
    dQpi = np.full((B, 1), f32(-1.0 / B), f32)
    _, dxs, _ = mlp_backward(main_Q, ac.modular, Qpi_cache, dQpi, want_param_grads=False)
    d_a = dxs[:, -ac.dimu:]                       # d pi_loss / d (pi / max_u)
    # pi/max_u = tanh(y);  L2 term: action_l2 * mean((pi/max_u)^2) over B*dimu elements (ddpg.py:441)
    d_th = (d_a + f32(action_l2) * f32(2.0 / (B * ac.dimu)) * th).astype(f32)
    dy = (d_th * (f32(1.0) - th * th)).astype(f32)
    pi_grads, _, _ = mlp_backward(main_pi, ac.modular, pi_cache, dy)
    # smallest |pre-activation| over the three evaluations that are differentiated: a float32
    # re-implementation may put such a unit on the other side of the ReLU kink (discontinuous gradient)
    relu_margin = min(pi_cache[2], Qpi_cache[2], Q_cache[2])
    return dict(Q_loss=Q_loss, pi_loss=pi_loss, Q_pi=Q_pi, Q=Q, target=target, pi=pi,
                Q_grad=flatten(Q_grads), pi_grad=flatten(pi_grads), relu_margin=relu_margin)


# --------------------------------------------------------------------------------------
# MpiAdam  (reference baselines/common/mpi_adam.py:6-50)
# --------------------------------------------------------------------------------------
class MpiAdamOracle:
    """Flat-vector Adam on float32 arrays.

    The reference ran under NumPy 1.x value-based casting: `a` is np.float64 but
    `(-a) * self.m` stays float32 (mpi_adam.py:31-35).  NumPy >= 2 would promote to float64, so the
    float32 casts are made explicit here.
    """

    def __init__(self, theta, beta1=0.9, beta2=0.999, epsilon=1e-08, scale_grad_by_procs=True,
                 allreduce_sum=None, world_size=1):
        self.theta = np.asarray(theta, f32).copy()
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.scale_grad_by_procs = scale_grad_by_procs
        self.m = np.zeros_like(self.theta)
        self.v = np.zeros_like(self.theta)
        self.t = 0
        self.allreduce_sum = allreduce_sum or (lambda x: x)
        self.world_size = world_size

    @staticmethod
    def step_scale(stepsize, beta1, beta2, t):
        return stepsize * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)     # mpi_adam.py:31 (float64)

    def update(self, localg, stepsize):
        g = self.allreduce_sum(np.asarray(localg).astype(f32))           # mpi_adam.py:24-26
        if self.scale_grad_by_procs:
            g = (g / f32(self.world_size)).astype(f32)
        self.t += 1
        a = self.step_scale(stepsize, self.beta1, self.beta2, self.t)
        b1, b2 = f32(self.beta1), f32(self.beta2)
        omb1, omb2 = f32(1 - self.beta1), f32(1 - self.beta2)
        self.m = (b1 * self.m + omb1 * g).astype(f32)                    # mpi_adam.py:32
        self.v = (b2 * self.v + omb2 * (g * g)).astype(f32)              # mpi_adam.py:33
        step = (f32(-a) * self.m / (np.sqrt(self.v) + f32(self.epsilon))).astype(f32)   # :34
        self.theta = (self.theta + step).astype(f32)                     # mpi_adam.py:35
        return self.theta


def polyak_update(target_flat, main_flat, polyak):
    """ddpg.py:461-462: target <- polyak*target + (1-polyak)*main, float32."""
    return (f32(polyak) * target_flat + f32(1. - polyak) * main_flat).astype(f32)


# --------------------------------------------------------------------------------------
# sample_batch apportioning, preprocess, store routing  (reference baselines/her/ddpg.py)
# --------------------------------------------------------------------------------------
def preprocess_og(o, ag, g, clip_obs, relative_goals=False):
    """ddpg.py:118-127 with subtract_goals = simple_goal_subtract (config.py:179-181)."""
    if relative_goals:
        g = g - ag
    return np.clip(o, -clip_obs, clip_obs), np.clip(g, -clip_obs, clip_obs)


def cp_probabilities(cp, eps_task):
    """ddpg.py:273-278 / 289-295: epsilon-mixture of uniform and CP-proportional."""
    cp = np.asarray(cp, np.float64)
    n = cp.size
    if cp.sum() == 0:
        p = (1 / n) * np.ones([n])
    else:
        p = eps_task * (1 / n) * np.ones([n]) + (1 - eps_task) * cp / cp.sum()
    p[-1] = 1 - p[:-1].sum()
    return p


def apportion_curious(buffer_episode_sizes, T, batch_size, task_replay, cp, eps_task):
    """ddpg.py:255-286 (structure='curious', per-module buffers).  Returns int proportions[N+1]."""
    sizes = np.array([e * T for e in buffer_episode_sizes])
    prop = np.zeros([len(sizes)])
    if sizes[1:].sum() < T:
        # ddpg.py:260-263: the reference divides by buffers_sizes.sum(); buffer[0] is never
        # written (ddpg.py:191-192), so this is 0/0 -> NaN -> undefined ints.  Not restated.
        raise ValueError('no module buffer holds an episode (reference path is undefined)')
    valid = np.argwhere(sizes[1:] > 0).reshape(-1)
    n_valid = len(valid)
    if task_replay == 'replay_task_random_buffer':
        p = 1 / valid.size * np.ones([n_valid])
    elif task_replay == 'replay_task_cp_buffer':
        p = cp_probabilities(np.asarray(cp)[valid], eps_task)
    else:
        raise NameError('proba')            # e.g. 'hand_designed': unbound in the reference
    prop[valid + 1] = p * batch_size
    prop = prop.astype(int)                 # ddpg.py:282 (np.int truncation)
    remain = batch_size - prop.sum()
    for i in range(remain):                 # ddpg.py:284-285 round-robin remainder
        prop[valid[i % n_valid] + 1] += 1
    return prop


def apportion_task_expert(buffer_episode_sizes, T, batch_size, t_id):
    """ddpg.py:302-318 (structure='task_experts', 'replay_current_task_buffer')."""
    sizes = np.array([e * T for e in buffer_episode_sizes])
    valid = np.argwhere(sizes > 0).reshape(-1)
    n_valid = len(valid)
    prop = np.zeros([len(sizes)])
    if sizes[t_id + 1] > 0:
        prop[t_id + 1] = 1
    else:
        prop[valid] = 1 / len(valid)
    prop *= batch_size
    prop = prop.astype(int)
    remain = batch_size - prop.sum()
    for i in range(remain):
        prop[valid[i % n_valid]] += 1
    return prop


def active_modules(change_last, tasks_ag_id, tasks_g_id):
    """ddpg.py:178-184: modules whose achieved-goal slice moved by the last step; j<5 cap."""
    nb = len(tasks_g_id)
    act = []
    for j in range(nb):
        cols = list(tasks_ag_id[j])[:len(tasks_g_id[j])]
        if any(change_last[cols]):
            if nb < 5 or j < 5:
                act.append(j)
    return act


def stage_keys(input_dims):
    """ddpg.py:73-83: order of the staged batch list."""
    keys = [k for k in sorted(input_dims.keys()) if not k.startswith('info_')]
    return keys + ['o_2', 'g_2', 'r']


class DDPGOracle:
    """The DDPG surface restated end to end on CPU (ddpg.py:18-537), float64 buffers + float32 nets.

    `buffers` are ReplayBufferOracle objects (list for *_buffer task_replay modes).
    `weights_rng` seeds the Xavier init in flat order: main/Q, main/pi  (target <- main, ddpg.py:459).
    """

    def __init__(self, input_dims, hidden, layers, polyak, batch_size, Q_lr, pi_lr, norm_eps, norm_clip,
                 max_u, action_l2, clip_obs, T, rollout_batch_size, relative_goals, clip_pos_returns,
                 clip_return, normalize_obs, sample_transitions, gamma, buffers, structure,
                 tasks_ag_id=None, tasks_g_id=None, task_replay='', t_id=None, eps_task=None,
                 weights_rng=None, allreduce_sum=None, mean_over_ranks=None, world_size=1):
        self.__dict__.update({k: v for k, v in locals().items() if k != 'self'})
        if self.clip_return is None:
            self.clip_return = np.inf
        self.dimo, self.dimg = input_dims['o'], input_dims['g']
        self.dimag, self.dimu = input_dims['ag'], input_dims['u']
        self.modular = structure in ('curious', 'task_experts')
        self.dimtd = input_dims['task_descr'] if self.modular else 0
        self.nb_tasks = len(tasks_g_id) if tasks_g_id is not None else 0
        self.stage_keys = stage_keys(input_dims)
        self.buffer = buffers
        if isinstance(self.buffer, list) and len(self.buffer) > 5:      # ddpg.py:104-110
            for i in range(6, len(self.buffer)):
                self.buffer[i] = self.buffer[5]
        self.o_stats = NormalizerOracle(self.dimo, norm_eps, norm_clip, mean_over_ranks)
        self.g_stats = NormalizerOracle(self.dimg, norm_eps, norm_clip, mean_over_ranks)
        self.ac = ActorCriticOracle(self.modular, self.dimo, self.dimg, self.dimu, self.dimtd, hidden,
                                    layers, max_u, normalize_obs, self.o_stats, self.g_stats)
        rng = weights_rng or np.random.RandomState(0)
        self.main_Q = xavier_uniform_net(rng, self.ac.Q_shapes)
        self.main_pi = xavier_uniform_net(rng, self.ac.pi_shapes)
        self.target_Q = [v.copy() for v in self.main_Q]                  # ddpg.py:459-460
        self.target_pi = [v.copy() for v in self.main_pi]
        kw = dict(scale_grad_by_procs=False, allreduce_sum=allreduce_sum, world_size=world_size)
        self.Q_adam = MpiAdamOracle(flatten(self.main_Q), **kw)          # ddpg.py:452-453
        self.pi_adam = MpiAdamOracle(flatten(self.main_pi), **kw)
        self.cp = None

    # ---- data path -------------------------------------------------------------------
    def store_episode(self, episode_batch, cp, n_ep, update_stats=True):
        n = episode_batch['ag'].shape[0]
        self.cp, self.n_episodes = cp, n_ep
        multi = ('buffer' in self.task_replay) or self.task_replay == 'hand_designed'
        if self.structure in ('curious', 'task_experts'):
            for b in range(n):
                act = active_modules(episode_batch['change'][b, -1], self.tasks_ag_id, self.tasks_g_id)
                ep = {k: v[b].reshape([1, v.shape[1], v.shape[2]]) for k, v in episode_batch.items()}
                if multi:
                    for task in act:                                     # ddpg.py:194-195
                        self.buffer[task + 1].store_episode(ep)
                else:
                    self.buffer.store_episode(ep)
        else:
            for b in range(n):
                ep = {k: v[b].reshape([1, v.shape[1], v.shape[2]]) for k, v in episode_batch.items()}
                self.buffer.store_episode(ep)
        if update_stats:                                                 # ddpg.py:206-223
            episode_batch['o_2'] = episode_batch['o'][:, 1:, :]
            episode_batch['ag_2'] = episode_batch['ag'][:, 1:, :]
            num = episode_batch['u'].shape[0] * episode_batch['u'].shape[1]
            if self.modular:
                tr = self.sample_transitions(episode_batch, num, task_to_replay=None)
            else:
                tr = self.sample_transitions(episode_batch, num)
            o, g = preprocess_og(tr['o'], tr['ag'], tr['g'], self.clip_obs, self.relative_goals)
            self.last_stats_batch = (o, g)
            self.o_stats.update(o)
            self.g_stats.update(g)
            self.o_stats.recompute_stats()
            self.g_stats.recompute_stats()

    def sample_batch(self):
        multi = ('buffer' in self.task_replay) or self.task_replay == 'hand_designed'
        if self.modular and multi:
            sizes = [self.buffer[i].current_size for i in range(self.nb_tasks + 1)]
            if self.structure == 'curious':
                prop = apportion_curious(sizes, self.T, self.batch_size, self.task_replay, self.cp,
                                         self.eps_task)
            else:
                prop = apportion_task_expert(sizes, self.T, self.batch_size, self.t_id)
            self.proportions = prop
            assert prop.sum() == self.batch_size                         # ddpg.py:323
            parts = []
            for i in range(self.nb_tasks + 1):
                if prop[i] > 0:
                    if self.structure == 'curious':
                        ttr = i - 1 if i > 0 else None
                    else:
                        ttr = self.t_id
                    parts.append(self.buffer[i].sample(int(prop[i]), task_to_replay=ttr))
            perm = np.arange(self.batch_size)
            np.random.shuffle(perm)                                      # ddpg.py:338-339
            self.last_perm = perm
            tr = {}
            for key in parts[0].keys():
                cat = np.concatenate([np.array([]).reshape([0, parts[0][key].shape[1]])] +
                                     [p[key] for p in parts])
                tr[key] = cat[perm, :]
        elif self.modular and self.structure == 'curious' and self.task_replay == 'replay_cp_task_transition':
            proba = cp_probabilities(self.cp, self.eps_task)             # ddpg.py:288-296
            tr = self.buffer.sample(self.batch_size, task_to_replay=None, cp_proba=proba)
        elif self.modular:
            tr = self.buffer.sample(self.batch_size, task_to_replay=None, cp_proba=None)
        else:
            tr = self.buffer.sample(self.batch_size)
        o, o_2, g, ag, ag_2 = tr['o'], tr['o_2'], tr['g'], tr['ag'], tr['ag_2']
        tr['o'], tr['g'] = preprocess_og(o, ag, g, self.clip_obs, self.relative_goals)
        tr['o_2'], tr['g_2'] = preprocess_og(o_2, ag_2, g, self.clip_obs, self.relative_goals)
        self.last_transitions = tr
        return [tr[k] for k in self.stage_keys]

    # ---- learner ---------------------------------------------------------------------
    def grads(self, batch_list):
        batch = {k: np.asarray(v, f32) for k, v in zip(self.stage_keys, batch_list)}   # TF feed cast
        return ddpg_losses_and_grads(self.ac, self.main_Q, self.main_pi, self.target_Q, self.target_pi,
                                     batch, self.gamma, self.clip_return, self.clip_pos_returns,
                                     self.action_l2)

    def train(self, batch_list=None):
        if batch_list is None:
            batch_list = self.sample_batch()
        out = self.grads(batch_list)
        self.main_Q = unflatten(self.Q_adam.update(out['Q_grad'], self.Q_lr), self.ac.Q_shapes)
        self.main_pi = unflatten(self.pi_adam.update(out['pi_grad'], self.pi_lr), self.ac.pi_shapes)
        self.last = out
        return out['Q_loss'], out['Q_pi']                                # ddpg.py:237-243,373

    def update_target_net(self):
        tq = polyak_update(flatten(self.target_Q), flatten(self.main_Q), self.polyak)
        tp = polyak_update(flatten(self.target_pi), flatten(self.main_pi), self.polyak)
        self.target_Q = unflatten(tq, self.ac.Q_shapes)
        self.target_pi = unflatten(tp, self.ac.pi_shapes)

    def get_actions(self, o, ag, g, task_descr=None, noise_eps=0., random_eps=0., use_target_net=False,
                    compute_Q=False):
        """ddpg.py:129-161 (host RNG order: randn, binomial, uniform)."""
        o, g = preprocess_og(np.asarray(o), np.asarray(ag), np.asarray(g), self.clip_obs,
                             self.relative_goals)
        pi_net, Q_net = (self.target_pi, self.target_Q) if use_target_net else (self.main_pi, self.main_Q)
        o_in, g_in, td = self.ac.inputs(o.reshape(-1, self.dimo), g.reshape(-1, self.dimg),
                                        None if task_descr is None else
                                        np.asarray(task_descr).reshape(-1, self.dimtd))
        u, _ = self.ac.pi(pi_net, o_in, g_in, td)
        ret = [u]
        if compute_Q:
            q, _ = self.ac.Q(Q_net, o_in, g_in, td, u)
            ret.append(q)
        u = ret[0].copy()
        noise = noise_eps * self.max_u * np.random.randn(*u.shape)
        u += noise                                   # float64 noise added in place to the fp32 output
        u = np.clip(u, -self.max_u, self.max_u)
        u += np.random.binomial(1, random_eps, u.shape[0]).reshape(-1, 1) * \
            (np.random.uniform(low=-self.max_u, high=self.max_u, size=(u.shape[0], self.dimu)) - u)
        if u.shape[0] == 1:
            u = u[0]
        ret[0] = u.copy()
        return ret[0] if len(ret) == 1 else ret
