"""Host arithmetic of DDPG.sample_batch's batch apportioning (reference ddpg.py:255-318).

At most nb_tasks+1 integers per update: it stays on the host and feeds the fused kernel's segment table.
"""
import numpy as np


def cp_probabilities(cp, eps_task):
    """eps-mixture of uniform and CP-proportional probabilities (ddpg.py:273-278, 289-295)."""
    cp = np.asarray(cp, np.float64)
    n = cp.size
    if cp.sum() == 0:
        proba = (1 / n) * np.ones([n])
    else:
        proba = eps_task * (1 / n) * np.ones([n]) + (1 - eps_task) * cp / cp.sum()
    proba[-1] = 1 - proba[:-1].sum()
    return proba


def proportions_curious(episode_sizes, T, batch_size, task_replay, cp, eps_task):
    """structure='curious' with per-module buffers (ddpg.py:255-286)."""
    sizes = np.array([e * T for e in episode_sizes])
    prop = np.zeros([len(sizes)])
    if sizes[1:].sum() < T:
        raise RuntimeError('no module buffer holds an episode yet: the reference divides 0/0 here '
                           '(ddpg.py:260-263) and never terminates; store an active episode first')
    valid = np.argwhere(sizes[1:] > 0).reshape(-1)
    n_valid = len(valid)
    if task_replay == 'replay_task_random_buffer':
        proba = 1 / valid.size * np.ones([n_valid])
    elif task_replay == 'replay_task_cp_buffer':
        proba = cp_probabilities(np.asarray(cp)[valid], eps_task)
    else:
        raise NameError("task_replay %r defines no buffer probabilities (unbound `proba` in the reference, "
                        "ddpg.py:279)" % (task_replay,))
    prop[valid + 1] = proba * batch_size
    prop = prop.astype(int)                       # truncation (ddpg.py:282)
    remain = batch_size - prop.sum()
    for i in range(remain):                       # round-robin remainder (ddpg.py:284-285)
        prop[valid[i % n_valid] + 1] += 1
    assert prop.sum() == batch_size               # ddpg.py:323
    return prop


def proportions_task_expert(episode_sizes, T, batch_size, t_id):
    """structure='task_experts', task_replay='replay_current_task_buffer' (ddpg.py:302-318)."""
    sizes = np.array([e * T for e in episode_sizes])
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
    assert prop.sum() == batch_size
    return prop


def active_modules(change_last, tasks_ag_id, tasks_g_id):
    """Modules whose achieved-goal slice moved by the last step; only j<5 when nb_tasks>=5 (ddpg.py:178-184)."""
    nb = len(tasks_g_id)
    active = []
    moved = np.asarray(change_last).tolist()          # plain Python truth values: a dozen scalar look-ups beat fancy indexing
    for j in range(nb):
        cols = list(tasks_ag_id[j])[:len(tasks_g_id[j])]
        if any(moved[c] for c in cols):
            if nb < 5 or j < 5:
                active.append(j)
    return active
