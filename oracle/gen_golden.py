"""Generate tests/golden/*.npz from the UNMODIFIED reference modules (test infrastructure).

Run in the build container only (it needs /root/reference, which does not exist on the GPU
box):      python oracle/gen_golden.py

It imports reference `baselines/her/her.py` and `baselines/her/replay_buffer.py` as they lie
(a stub `mpi4py` is put in sys.modules because replay_buffer.py imports MPI without using it),
runs them on small seeded episode batches and records, per case:
    in_<key>     the episode arrays handed to the reference
    s_*          the np.random draws the reference made, in order (ep, t, u_her, u_off, choices)
    out_<key>    the transitions the reference returned
    meta         JSON: sampler arguments, seed, dims
The reward callable handed to the reference is oracle.reward_oracle.ModuleDistanceReward
(gym_flowers is absent, see that file); the fixtures therefore pin the reference's *call* of
the reward (which arrays, after which relabelling), not gym_flowers itself.

Nothing from the reference is copied into the repo: only its outputs.
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = '/root/reference'


def import_reference():
    if 'mpi4py' not in sys.modules:
        stub = types.ModuleType('mpi4py')
        stub.MPI = types.SimpleNamespace()
        sys.modules['mpi4py'] = stub
    sys.path.insert(0, REF)
    from baselines.her import her as ref_her
    from baselines.her import replay_buffer as ref_rb
    return ref_her, ref_rb


class Recorder:
    """Wraps the np.random entry points the reference uses and logs what they returned."""

    def __init__(self):
        self.log = []
        self._orig = {}

    def __enter__(self):
        for name in ('randint', 'uniform', 'choice', 'shuffle'):
            self._orig[name] = getattr(np.random, name)
            setattr(np.random, name, self._wrap(name))
        return self

    def __exit__(self, *a):
        for name, fn in self._orig.items():
            setattr(np.random, name, fn)

    def _wrap(self, name):
        orig = self._orig[name]

        def f(*args, **kw):
            out = orig(*args, **kw)
            if name == 'shuffle':
                self.log.append((name, np.array(args[0]).copy()))
            else:
                self.log.append((name, np.array(out).copy()))
            return out
        return f


def small_dims(n_modules, dimo, extra_ag=0):
    return {'o': dimo, 'u': 4, 'g': 3 * n_modules, 'ag': 3 * n_modules + extra_ag,
            'task_descr': n_modules, 'info_is_success': 1}


def task_ids(n_modules, longer_ag=False):
    g_ids = [[3 * j, 3 * j + 1, 3 * j + 2] for j in range(n_modules)]
    if longer_ag:
        # achieved-goal slices longer than the goal slices (truncated by her.py:147-148);
        # the last module's slice reaches into the extra ag columns
        ag_ids = [[3 * j, 3 * j + 1, 3 * j + 2, 3 * j + 3] for j in range(n_modules)]
    else:
        ag_ids = [list(x) for x in g_ids]
    return ag_ids, g_ids


def stream_from_log(log, B, her_rows=None):
    """First four draws are ep, t, u_her, u_off (her.py:108-116); then per-row choices."""
    names = [n for n, _ in log]
    assert names[:4] == ['randint', 'randint', 'uniform', 'uniform'], names[:6]
    s = dict(s_ep=log[0][1].astype(np.int64), s_t=log[1][1].astype(np.int64),
             s_uher=log[2][1].astype(np.float64), s_uoff=log[3][1].astype(np.float64))
    choices = [int(v) for n, v in log[4:] if n == 'choice']
    s['s_choice_seq'] = np.array(choices, np.int64)
    return s


def run_sampler_case(name, ref_her, ref_rb, *, n_modules, dimo, E, T, B, seed, goal_replay='her',
                     task_replay='replay_task_cp_buffer', task_to_replay=None, cp_proba=None,
                     flat=False, longer_ag=False, via_buffer=True, data_seed=0, outdir=None):
    from curious_b200 import synth
    from oracle.reward_oracle import ModuleDistanceReward
    dims = small_dims(n_modules, dimo, extra_ag=1 if longer_ag else 0)
    ag_ids, g_ids = task_ids(n_modules, longer_ag)
    rng = np.random.RandomState(data_seed)
    eps = synth.make_episodes(rng, E, T, dims, change_dtype=np.float32 if via_buffer else bool)
    reward = ModuleDistanceReward(ag_ids, g_ids, 0.05)
    if flat:
        sampler = ref_her.make_sample_her_transitions(goal_replay, 4, reward, task_replay,
                                                      tasks_ag_id=ag_ids, tasks_g_id=g_ids)
    else:
        sampler = ref_her.make_sample_multi_task_her_transitions(goal_replay, 4, task_replay, reward,
                                                                 tasks_ag_id=ag_ids, tasks_g_id=g_ids)
    np.random.seed(seed)
    with Recorder() as rec:
        if via_buffer:
            shapes = {k: v.shape[1:] for k, v in eps.items()}
            buf = ref_rb.ReplayBuffer(shapes, (E + 3) * T, T, sampler)
            buf.store_episode({k: v.copy() for k, v in eps.items()})
            rec.log.clear()
            if flat:
                out = buf.sample(B)
            else:
                out = buf.sample(B, task_to_replay=task_to_replay, cp_proba=cp_proba)
        else:
            batch = {k: v.copy() for k, v in eps.items()}
            batch['o_2'] = batch['o'][:, 1:, :]
            batch['ag_2'] = batch['ag'][:, 1:, :]
            if flat:
                out = sampler(batch, B)
            else:
                out = sampler(batch, B, task_to_replay=task_to_replay, cp_proba=cp_proba)
    arrays = {'in_' + k: v for k, v in eps.items()}
    arrays.update(stream_from_log(rec.log, B))
    arrays.update({'out_' + k: np.asarray(v) for k, v in out.items()})
    meta = dict(name=name, n_modules=n_modules, dimo=dimo, E=E, T=T, B=B, seed=seed,
                goal_replay=goal_replay, task_replay=task_replay, task_to_replay=task_to_replay,
                cp_proba=None if cp_proba is None else list(map(float, cp_proba)), flat=flat,
                longer_ag=longer_ag, via_buffer=via_buffer, dims=dims, tasks_ag_id=ag_ids,
                tasks_g_id=g_ids, threshold=0.05, her_replay_k=4,
                reward_calls=reward.n_calls,
                reward_td_none=reward.last_kwargs['task_descr'] is None)
    arrays['meta'] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(outdir, name + '.npz'), **arrays)
    return meta


def run_storage_case(name, ref_rb, *, size_ep, T, increments, seed, outdir):
    """ReplayBuffer._get_storage_idx sequence (replay_buffer.py:90-109)."""
    shapes = {'o': (T + 1, 2), 'u': (T, 1)}
    buf = ref_rb.ReplayBuffer(shapes, size_ep * T, T, None)
    np.random.seed(seed)
    idxs, sizes, stored = [], [], []
    orig = buf._get_storage_idx

    def recording_idx(inc=None):
        out = orig(inc)
        idxs.append(np.atleast_1d(out).astype(np.int64).copy())
        return out
    buf._get_storage_idx = recording_idx
    for k, inc in enumerate(increments):
        ep = {'o': np.full((inc, T + 1, 2), float(k)), 'u': np.full((inc, T, 1), float(k))}
        buf.store_episode(ep)
        sizes.append(buf.get_current_episode_size())
        stored.append(buf.get_transitions_stored())
    arrays = {'idx_%d' % k: np.asarray(v, np.int64) for k, v in enumerate(idxs)}
    arrays['sizes'] = np.array(sizes, np.int64)
    arrays['stored'] = np.array(stored, np.int64)
    arrays['final_u'] = buf.buffers['u'][:buf.current_size, 0, 0].copy()
    arrays['meta'] = np.array(json.dumps(dict(name=name, size_ep=size_ep, T=T,
                                              increments=list(increments), seed=seed)))
    np.savez_compressed(os.path.join(outdir, name + '.npz'), **arrays)


def random_cases(kw, n=24, seed=777):
    """Randomly drawn scenarios (sizes, modules, every task_replay mode, with / without HER, flat, longer ag slices,
    buffer / direct calls) on top of the hand-picked ones: rnd_00 .. rnd_23."""
    rng = np.random.RandomState(seed)
    modes = ['replay_task_cp_buffer', 'replay_task_random_buffer', 'replay_random_task_transition',
             'replay_cp_task_transition', 'replay_current_task_transition']
    out = []
    for i in range(n):
        flat = bool(rng.rand() < 0.2)
        n_modules = int(rng.randint(1, 7))
        mode = '' if flat else modes[int(rng.randint(len(modes)))]
        case = dict(n_modules=n_modules, dimo=int(rng.randint(1, 12)), E=int(rng.randint(1, 9)), T=int(rng.randint(1, 14)),
                    B=int(rng.randint(1, 130)), seed=int(rng.randint(1 << 30)), data_seed=int(rng.randint(1 << 30)),
                    goal_replay='her' if rng.rand() < 0.8 else 'none', task_replay=mode, flat=flat,
                    longer_ag=bool(rng.rand() < 0.3) and not flat, via_buffer=bool(rng.rand() < 0.7))
        if not flat:
            if 'buffer' in mode and rng.rand() < 0.7:
                case['task_to_replay'] = int(rng.randint(n_modules))
            if mode == 'replay_cp_task_transition':
                p = rng.rand(n_modules) + 0.05
                case['cp_proba'] = list(p / p.sum())
        out.append(run_sampler_case('rnd_%02d' % i, **case, **kw))
    return out


def main(outdir=None):
    outdir = outdir or os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(outdir, exist_ok=True)
    ref_her, ref_rb = import_reference()
    kw = dict(ref_her=ref_her, ref_rb=ref_rb, outdir=outdir)
    cases = []
    # per-module buffers: the module is dictated by the buffer (her.py:131-136)
    for ttr in (None, 0, 2):
        cases.append(run_sampler_case('mt_buffer_ttr%s' % ('N' if ttr is None else ttr),
                                      n_modules=3, dimo=7, E=6, T=10, B=96, seed=11 + (ttr or 0),
                                      task_to_replay=ttr, **kw))
    # Arm4-shaped (T=50, dimo=40, N=4)
    cases.append(run_sampler_case('mt_buffer_arm4', n_modules=4, dimo=40, E=5, T=50, B=128, seed=21,
                                  task_to_replay=1, **kw))
    # single buffer modes with np.random.choice inside the loop (her.py:138-142)
    cases.append(run_sampler_case('mt_random_task', n_modules=3, dimo=7, E=6, T=10, B=96, seed=31,
                                  task_replay='replay_random_task_transition', **kw))
    cases.append(run_sampler_case('mt_cp_task', n_modules=3, dimo=7, E=6, T=10, B=96, seed=32,
                                  task_replay='replay_cp_task_transition',
                                  cp_proba=[0.5, 0.2, 0.3], **kw))
    cases.append(run_sampler_case('mt_current_task', n_modules=3, dimo=7, E=6, T=10, B=96, seed=33,
                                  task_replay='replay_current_task_transition', **kw))
    # no HER (her.py:88-89): nothing relabelled, reward still recomputed
    cases.append(run_sampler_case('mt_no_her', n_modules=3, dimo=7, E=6, T=10, B=64, seed=34,
                                  goal_replay='none', task_to_replay=1, **kw))
    # achieved-goal slices longer than goal slices (her.py:147-148)
    cases.append(run_sampler_case('mt_longer_ag', n_modules=3, dimo=5, E=4, T=8, B=64, seed=35,
                                  longer_ag=True, task_to_replay=2, **kw))
    # the normaliser path of DDPG.store_episode: sampler called directly on the float32 rollout
    # batch with a bool `change` (ddpg.py:209-215)
    cases.append(run_sampler_case('mt_stats_path', n_modules=4, dimo=9, E=2, T=50, B=100, seed=36,
                                  via_buffer=False, **kw))
    # single-episode buffer, single timestep edge
    cases.append(run_sampler_case('mt_one_episode', n_modules=2, dimo=3, E=1, T=1, B=16, seed=37,
                                  task_to_replay=0, **kw))
    # flat sampler (her.py:5-68)
    cases.append(run_sampler_case('flat_her', n_modules=3, dimo=7, E=6, T=10, B=96, seed=41, flat=True,
                                  task_replay='', **kw))
    cases.append(run_sampler_case('flat_no_her', n_modules=3, dimo=7, E=6, T=10, B=32, seed=42,
                                  flat=True, goal_replay='none', task_replay='', **kw))
    cases.extend(random_cases(kw))
    # storage index policy
    run_storage_case('storage_single', ref_rb, size_ep=5, T=3, increments=[1] * 12, seed=51,
                     outdir=outdir)
    run_storage_case('storage_batched', ref_rb, size_ep=7, T=3, increments=[2, 2, 2, 2, 3, 1, 2],
                     seed=52, outdir=outdir)
    with open(os.path.join(outdir, 'MANIFEST.json'), 'w') as f:
        json.dump(dict(generator='oracle/gen_golden.py', reference='/root/reference (flowersteam/curious)',
                       numpy=np.__version__, cases=[c['name'] for c in cases] +
                       ['storage_single', 'storage_batched']), f, indent=1)
    print('wrote', len(cases) + 2, 'fixtures to', outdir)


if __name__ == '__main__':
    main()
