"""The driver loop (curious_b200/train.py, reference experiment/train.py:48-206) with a stub policy: control flow, expert
selection arithmetic and run records need no GPU.  (The real agent runs the same loop in tests/test_train_loop_gpu.py.)"""
import json
import os

import numpy as np
import pytest

from curious_b200.envs import ModularPointEnv
from curious_b200.rollout import RolloutWorker
from curious_b200.train import configure_dims, train


class StubPolicy(object):
    def __init__(self, dimu, name='p'):
        self.dimu, self.name = dimu, name
        self.stored, self.trained, self.target_updates = [], 0, 0

    def get_actions(self, o, ag, g, task_descr=None, compute_Q=False, **kw):
        u = np.random.uniform(-1, 1, (len(o), self.dimu))
        return (u, np.full((len(o), 1), -3.0)) if compute_Q else u

    def store_episode(self, episode, cp, n_ep):
        self.stored.append((episode['o'].shape, np.array(cp, np.float64).copy(), n_ep))

    def train(self):
        self.trained += 1

    def update_target_net(self):
        self.target_updates += 1

    def logs(self, prefix=''):
        return [('stats_o/mean', 0.25), ('stats_g/std', 1.0)]

    def save_checkpoint(self, path):
        """Like DDPG.save_checkpoint: the agent's own state plus the host np.random state."""
        import pickle
        with open(path, 'wb') as f:
            pickle.dump(dict(name=self.name, trained=self.trained, target_updates=self.target_updates,
                             stored=len(self.stored), numpy_rng=np.random.get_state()), f)

    def load_checkpoint(self, path):
        import pickle
        with open(path, 'rb') as f:
            st = pickle.load(f)
        assert st['name'] == self.name
        self.trained, self.target_updates = st['trained'], st['target_updates']
        self.stored = [None] * st['stored']
        np.random.set_state(st['numpy_rng'])


def _workers(structure, nb_tasks=3, T=8):
    def make_env():
        return ModularPointEnv(nb_tasks, max_episode_steps=T)
    dims = configure_dims(make_env(), structure)
    kw = dict(dims=dims, logger=None, T=T, rollout_batch_size=2, structure=structure,
              task_selection='active_competence_progress', queue_length=10)
    if structure == 'task_experts':
        policy = [StubPolicy(dims['u'], 'p%d' % i) for i in range(nb_tasks)]
        rollout = [RolloutWorker(make_env, policy[i], unique_task=i, **kw) for i in range(nb_tasks)]
    else:
        policy = StubPolicy(dims['u'])
        rollout = RolloutWorker(make_env, policy, **kw)
    evaluator = RolloutWorker(make_env, policy, exploit=True, compute_Q=True, eval=True, **kw)
    return policy, rollout, evaluator, dims


@pytest.mark.parametrize('structure', ['curious', 'flat'])
def test_loop_counts_and_records(structure, tmp_path):
    np.random.seed(0)
    policy, rollout, evaluator, dims = _workers(structure)
    hist = train(policy, rollout, evaluator, n_epochs=3, n_test_rollouts=2, n_cycles=4, n_batches=5, structure=structure,
                 logdir=str(tmp_path), params={'structure': structure, 'n_cycles': 4}, policy_save_interval=2,
                 checkpoint_interval=1)
    assert len(hist) == 3
    # train.py:148-155: per cycle one store_episode, n_batches updates, one target update
    assert len(policy.stored) == 12 and policy.trained == 60 and policy.target_updates == 12
    assert policy.stored[0][0] == (2, 9, dims['o']) and policy.stored[-1][2] == 12 * 2
    lines = open(str(tmp_path / 'progress.csv')).read().splitlines()
    header = lines[0].split(',')
    assert len(lines) == 5 and header[0] == 'epoch' and header[-1] == 'Time'          # epoch -1 (untrained policy), 0, 1, 2
    first = dict(zip(header, lines[1].split(',')))
    assert first['epoch'] == '-1' and first['train/success_rate'] == 'nan' and first['train/episode'] == '0'
    for col in ('test/success_rate', 'test/mean_Q', 'train/success_rate', 'train/episode', 'stats_o/mean'):
        assert col in header, col
    row = dict(zip(header, lines[-1].split(',')))
    assert row['epoch'] == '2' and row['test/mean_Q'] == '-3' and row['train/episode'] == '24' and row['stats_o/mean'] == '0.25'
    if structure == 'curious':
        for col in ('train/C_task0', 'train/CP_task2', 'train/%_task1', 'train/p_task0', 'test/C_task2'):
            assert col in header, col
        assert all('CP' in h and 'p' in h for h in hist)
    else:
        assert not any('task' in h for h in header)
    assert json.load(open(str(tmp_path / 'params.json'))) == {'structure': structure, 'n_cycles': 4}
    files = set(os.listdir(str(tmp_path)))
    assert {'policy_best.pkl', 'policy_latest.pkl', 'policy_0.pkl', 'policy_2.pkl', 'checkpoint_0.pt', 'log.txt'} <= files
    assert 'policy_1.pkl' not in files


def test_task_experts_selection_like_the_reference(tmp_path):
    """train.py:79-104: the reference computes proba = eps / N + (1 - eps) CP / sum(CP) over the experts' own competence
    progress but draws the expert from `p`, which it never updates - uniform selection.  The drop-in behaves the same by
    default and draws from proba with experts_follow_cp=True."""
    np.random.seed(1)
    policy, rollout, evaluator, dims = _workers('task_experts')
    hist = train(policy, rollout, evaluator, n_epochs=4, n_test_rollouts=1, n_cycles=2, n_batches=3,
                 structure='task_experts', logdir=str(tmp_path), checkpoint_interval=1)
    assert len(hist) == 4
    for h in hist:
        assert np.allclose(h['p'], 1.0 / 3)
    chosen = [h['i_policy'] for h in hist]
    for i, pol in enumerate(policy):
        assert pol.trained == 6 * chosen.count(i) and pol.target_updates == 2 * chosen.count(i)
    header = open(str(tmp_path / 'progress.csv')).read().splitlines()[0].split(',')
    assert 'IND_TASK_rollout' in header
    assert {'checkpoint_0.pt', 'checkpoint_1.pt', 'checkpoint_2.pt'} <= set(os.listdir(str(tmp_path)))
    # an expert that progresses: proba follows it, the draw does not (reference behaviour) unless asked to
    rollout[1].tracker.competence_computers[1].CP = 0.5
    want = [0.4 / 3, 0.4 / 3 + 0.6, 0.4 / 3]
    np.random.seed(2)
    hist = train(policy, rollout, evaluator, n_epochs=1, n_test_rollouts=1, n_cycles=1, n_batches=1, structure='task_experts')
    assert np.allclose(hist[0]['proba'], want) and np.allclose(hist[0]['p'], 1.0 / 3)
    np.random.seed(2)
    hist = train(policy, rollout, evaluator, n_epochs=1, n_test_rollouts=1, n_cycles=1, n_batches=1, structure='task_experts',
                 experts_follow_cp=True)
    assert np.allclose(hist[0]['proba'], want) and np.allclose(hist[0]['p'], want)
    # task_selection='random': round robin over the experts (train.py:79-81)
    hist = train(policy, rollout, evaluator, n_epochs=4, n_test_rollouts=1, n_cycles=1, n_batches=1, structure='task_experts',
                 task_selection='random')
    assert [h['i_policy'] for h in hist] == [0, 1, 2, 0]


@pytest.mark.parametrize('structure', ['curious', 'task_experts'])
def test_resumed_run_equals_uninterrupted_run(structure, tmp_path):
    """train(resume=True): 2 epochs, a fresh process' worth of new objects, then 3 more epochs == 5 epochs in one go -
    same evaluation results, competence, probabilities, expert choices and progress.csv (except the Time column)."""
    def run(logdir, n_epochs, seed, resume=False):
        np.random.seed(seed)
        policy, rollout, evaluator, _ = _workers(structure)
        for i, w in enumerate((rollout if isinstance(rollout, list) else [rollout]) + [evaluator]):
            w.seed(50 + 10 * i)
        hist = train(policy, rollout, evaluator, n_epochs=n_epochs, n_test_rollouts=2, n_cycles=3, n_batches=2,
                     structure=structure, logdir=logdir, policy_save_interval=0, checkpoint_interval=1, resume=resume)
        return hist, policy

    full, pol_full = run(str(tmp_path / 'full'), 5, seed=3)
    run(str(tmp_path / 'split'), 2, seed=3)
    tail, pol_tail = run(str(tmp_path / 'split'), 5, seed=99, resume=True)        # another seed: everything comes from the files
    assert [h['epoch'] for h in tail] == [2, 3, 4]
    for a, b in zip(full[2:], tail):
        assert set(a) == set(b)
        for key in a:
            assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key
    pf, pt = (pol_full, pol_tail) if isinstance(pol_full, list) else ([pol_full], [pol_tail])
    assert [p.trained for p in pf] == [p.trained for p in pt]

    def table(path):
        lines = open(path).read().splitlines()
        t = lines[0].split(',').index('Time')
        return [[v for i, v in enumerate(line.split(',')) if i != t] for line in lines]
    assert table(str(tmp_path / 'full' / 'progress.csv')) == table(str(tmp_path / 'split' / 'progress.csv'))
    assert 'Resuming after epoch 1' in open(str(tmp_path / 'split' / 'log.txt')).read()
    # resume=True without checkpoints starts from scratch
    fresh, _ = run(str(tmp_path / 'empty'), 1, seed=3, resume=True)
    assert [h['epoch'] for h in fresh] == [0]


def test_resume_from_the_epoch_minus_one_checkpoint(tmp_path):
    """With checkpoint_interval=1 the evaluation of the untrained policy (epoch -1, train.py:62-75) is checkpointed too; a run
    killed right after it and resumed must not evaluate the untrained policy a second time: one epoch -1 row, evaluator
    queues and all later epochs equal to the uninterrupted run."""
    def run(logdir, n_epochs, seed, resume=False):
        np.random.seed(seed)
        policy, rollout, evaluator, _ = _workers('curious')
        for i, w in enumerate([rollout, evaluator]):
            w.seed(70 + 10 * i)
        return train(policy, rollout, evaluator, n_epochs=n_epochs, n_test_rollouts=2, n_cycles=2, n_batches=2,
                     structure='curious', logdir=logdir, policy_save_interval=0, checkpoint_interval=1, resume=resume,
                     initial_evaluation=True)

    full = run(str(tmp_path / 'full'), 2, seed=5)
    first = run(str(tmp_path / 'split'), 0, seed=5)                       # only the epoch -1 evaluation
    assert first == []                                                      # (epoch -1 goes to the run records only)
    tail = run(str(tmp_path / 'split'), 2, seed=77, resume=True)
    assert [h['epoch'] for h in full] == [0, 1] and [h['epoch'] for h in tail] == [0, 1]
    for a, b in zip(full, tail):
        for key in a:
            assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key

    def epochs(path):
        lines = open(path).read().splitlines()
        e = lines[0].split(',').index('epoch')
        return [line.split(',')[e] for line in lines[1:]]
    assert epochs(str(tmp_path / 'split' / 'progress.csv')) == epochs(str(tmp_path / 'full' / 'progress.csv')) == ['-1', '0', '1']

