"""CompetenceQueue / LP pipeline (curious_b200/queues.py) against vectors recorded from the unmodified reference
baselines/her/queues.py (oracle/gen_golden_queue.py) and the probability rule of rollout.py:381-393."""
import os

import numpy as np

from curious_b200.queues import CompetenceQueue, CompetenceTracker, task_probabilities

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'competence_queue.npz')


def test_competence_queue_matches_reference_bit_for_bit():
    g = np.load(GOLD)
    for c in range(int(g['n_cases'])):
        q = CompetenceQueue(window=int(g['c%d_window' % c]))
        succ, k = g['c%d_succ' % c], 0
        for i, n in enumerate(g['c%d_lens' % c]):
            q.update(succ[k:k + n].tolist())
            k += n
            assert float(q.CP) == g['c%d_CP' % c][i] and float(q.C) == g['c%d_C' % c][i], (c, i)
            assert q.size == g['c%d_size' % c][i] and q.full == bool(g['c%d_full' % c][i])
        q.clear_queue()
        assert [float(q.CP), float(q.C), float(q.size)] == g['c%d_after_clear' % c].tolist()


def test_task_probabilities_rule():
    p = task_probabilities(np.zeros(4), 4)
    assert np.array_equal(p, 0.25 * np.ones(4))                         # no progress anywhere: uniform
    cp = np.array([0.05, 0.2, 0.1, 0.0])
    p = task_probabilities(cp, 4)
    want = 0.4 * 0.25 + 0.6 * cp / cp.sum()
    assert np.allclose(p, want, atol=1e-15) and abs(p.sum() - 1.0) < 1e-15
    assert p[3] >= 0.1 - 1e-12                                          # epsilon floor for a module with CP = 0


def test_tracker_single_rank_pipeline():
    tr = CompetenceTracker(3, queue_length=4)
    rng = np.random.RandomState(0)
    seen = []
    for i in range(30):
        tasks = rng.randint(0, 3, size=2).tolist()
        succ = [float(t == 1 and i > 15) for t in tasks]                # module 1 starts succeeding late
        cp, p = tr.update(tasks, succ)
        assert cp.shape == (3,) and abs(p.sum() - 1) < 1e-12
        seen.append((cp.copy(), p.copy()))
    # while module 1 is improving it is the only one with learning progress and gets the largest share
    best = max(seen, key=lambda x: x[0][1])
    assert best[0][1] > 0 and best[0][0] == 0 and best[0][2] == 0 and best[1][1] == best[1].max() > 0.4
    # once both halves of its window are all successes the progress is back to 0: uniform again
    assert np.array_equal(seen[-1][0], np.zeros(3)) and np.allclose(seen[-1][1], 1 / 3)
    ev = CompetenceTracker(3, eval=True)
    cp, p = ev.update([0, 1], [1.0, 0.0])
    assert np.array_equal(p, np.ones(3) / 3)                            # evaluators never touch p (rollout.py:369)
    ex = CompetenceTracker(3, structure='task_experts', unique_task=2)
    cp, p = ex.update([2], [1.0])
    assert np.array_equal(p, np.array([0., 0., 1.]))
