"""Helpers to load tests/golden fixtures (made by oracle/gen_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def sampler_cases():
    names = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, '*.npz'))):
        n = os.path.basename(p)[:-4]
        if not n.startswith('storage_') and n != 'competence_queue':
            names.append(n)
    return names


def load_case(name, directory=None):
    z = np.load(os.path.join(directory or GOLDEN, name + '.npz'))
    meta = json.loads(str(z['meta']))
    eps = {k[3:]: z[k] for k in z.files if k.startswith('in_')}
    out = {k[4:]: z[k] for k in z.files if k.startswith('out_')}
    stream = {k: z[k] for k in z.files if k.startswith('s_')}
    return meta, eps, stream, out


def future_p(meta):
    return 1 - (1. / (1 + meta['her_replay_k'])) if meta['goal_replay'] == 'her' else 0


def per_row_choices(meta, stream):
    """Map the sequential np.random.choice results onto HER rows (her.py:129-142)."""
    B = meta['B']
    choice = np.full(B, -1, np.int64)
    seq = stream['s_choice_seq']
    if seq.size:
        her_rows = np.where(stream['s_uher'] < future_p(meta))[0]
        assert her_rows.size == seq.size
        choice[her_rows] = seq
    return choice
