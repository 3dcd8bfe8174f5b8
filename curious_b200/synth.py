"""Synthetic MultiTaskFetchArm-shaped episodes.

gym_flowers / MuJoCo are not available, so benches and tests run on episodes
that have the *shape and value structure* of what `RolloutWorker.generate_rollouts`
hands to `DDPG.store_episode` (reference baselines/her/rollout.py:290-303,406):
float32 `o [n,T+1,dimo]`, `ag [n,T+1,dimag]`, `g [n,T,dimg]`, `u [n,T,dimu]`,
`task_descr [n,T,N]` one-hot and constant over the episode (rollout.py:283),
`change [n,T,dimag]` = |ag[0] - ag[t+1]| > 1e-3 (rollout.py:284) and
`info_is_success [n,T,1]`.

This module only *generates inputs*; it performs no part of the hot path.
"""
import numpy as np


def arm_dims(n_modules=4, dimo=None):
    """Dims dict in the shape of config.configure_dims (reference config.py:257-275)."""
    if dimo is None:
        dimo = {4: 40, 8: 64}.get(n_modules, 10 + 15 * ((n_modules + 1) // 2))
    return {
        'o': dimo, 'u': 4, 'g': 3 * n_modules, 'ag': 3 * n_modules,
        'task_descr': n_modules, 'info_is_success': 1,
    }


def arm_task_ids(n_modules=4):
    """tasks_g_id / tasks_ag_id: consecutive triples (reference experiment/test_env.py:14)."""
    ids = [[3 * j, 3 * j + 1, 3 * j + 2] for j in range(n_modules)]
    return [list(x) for x in ids], [list(x) for x in ids]


def buffer_shapes(dims, T):
    """Per-episode shapes as built by config.configure_buffer (reference config.py:200-208)."""
    shapes = {}
    for key, val in dims.items():
        shapes[key] = (T + 1 if key == 'o' else T, val)
    shapes['ag'] = (T + 1, dims['ag'])
    shapes['change'] = (T, dims['ag'])
    return shapes


def make_episodes(rng, n_ep, T, dims, walk_sigma=0.02, goal_span=0.15, still_prob=0.3,
                  change_dtype=np.float32):
    """Return an episode_batch dict {key: float32 [n_ep, T(+1), dim]}.

    `ag` is a per-coordinate random walk so that ||ag_2 - g|| straddles the 0.05
    reward threshold once HER substitutes future achieved goals; a fraction
    `still_prob` of (episode, module) pairs does not move at all so that the
    `change`-mask routing of store_episode (reference ddpg.py:181) sees inactive modules.
    """
    N = dims.get('task_descr', 0)
    dimo, dimu, dimg, dimag = dims['o'], dims['u'], dims['g'], dims['ag']
    o = np.clip(rng.standard_normal((n_ep, T + 1, dimo)), -5, 5).astype(np.float32)
    ag0 = rng.uniform(-goal_span, goal_span, (n_ep, 1, dimag))
    steps = rng.standard_normal((n_ep, T, dimag)) * walk_sigma
    per_mod = dimag // N if N > 0 else dimag
    if N > 0 and per_mod > 0:
        moving = (rng.uniform(size=(n_ep, N)) >= still_prob).astype(np.float64)
        mask = np.ones((n_ep, dimag))
        rep = np.repeat(moving, per_mod, axis=1)[:, :dimag]
        mask[:, :rep.shape[1]] = rep
        steps = steps * mask[:, None, :]
    ag = np.concatenate([ag0, ag0 + np.cumsum(steps, axis=1)], axis=1).astype(np.float32)
    task = rng.randint(0, max(N, 1), n_ep)
    td = np.zeros((n_ep, T, N), np.float32)
    g = np.zeros((n_ep, T, dimg), np.float32)
    gvals = rng.uniform(-goal_span, goal_span, (n_ep, dimg)).astype(np.float32)
    per_g = dimg // N if N > 0 else dimg
    for e in range(n_ep):
        if N > 0:
            td[e, :, task[e]] = 1.0
            sl = slice(task[e] * per_g, (task[e] + 1) * per_g)
            g[e, :, sl] = gvals[e, sl]
        else:
            g[e, :, :] = gvals[e]
    u = rng.uniform(-1, 1, (n_ep, T, dimu)).astype(np.float32)
    change = (np.abs(ag[:, :1, :] - ag[:, 1:, :]) > 1e-3).astype(change_dtype)
    ep = {'o': o, 'u': u, 'g': g, 'ag': ag}
    if N > 0:
        ep['task_descr'] = td
        ep['change'] = change
    for key, val in dims.items():
        if key.startswith('info_'):
            ep[key] = (rng.uniform(size=(n_ep, T, val)) < 0.25).astype(np.float32)
    return ep
