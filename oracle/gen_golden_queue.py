"""Generate tests/golden/competence_queue.npz from the UNMODIFIED reference baselines/her/queues.py
(test infrastructure; build container only - needs /root/reference).

queues.py does `from pandas import ewma` (unused; removed from pandas long ago) - a stub attribute is put on a
stand-in `pandas` module for the duration of the import, nothing else is touched.  Seeded success streams are
pushed through CompetenceQueue.update in ragged chunks (including empty ones) and CP / C / size are recorded
after every call, plus one clear_queue().
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


def import_reference_queue():
    real = sys.modules.get('pandas')
    stub = types.ModuleType('pandas')
    stub.ewma = None
    sys.modules['pandas'] = stub
    sys.path.insert(0, REF)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location('ref_queues', os.path.join(REF, 'baselines/her/queues.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if real is not None:
            sys.modules['pandas'] = real
        else:
            del sys.modules['pandas']
    return mod


def main(outdir=None):
    ref = import_reference_queue()
    rng = np.random.RandomState(2024)
    arrays = {}
    for case, (window, n_calls, p_start, p_end) in enumerate([(100, 60, 0.05, 0.9), (5, 40, 0.5, 0.5), (500, 80, 0.0, 0.3)]):
        q = ref.CompetenceQueue(window=window)
        chunks, cps, cs, sizes, fulls = [], [], [], [], []
        for i in range(n_calls):
            n = int(rng.randint(0, 7)) if i % 9 else 0            # some empty updates
            p = p_start + (p_end - p_start) * i / max(n_calls - 1, 1)
            succ = (rng.uniform(size=n) < p).astype(np.float64)
            q.update(succ.tolist())
            chunks.append(succ)
            cps.append(float(q.CP)); cs.append(float(q.C)); sizes.append(q.size); fulls.append(bool(q.full))
        q.clear_queue()
        arrays['c%d_window' % case] = np.array(window)
        arrays['c%d_lens' % case] = np.array([len(c) for c in chunks])
        arrays['c%d_succ' % case] = np.concatenate(chunks) if chunks else np.zeros(0)
        arrays['c%d_CP' % case] = np.array(cps)
        arrays['c%d_C' % case] = np.array(cs)
        arrays['c%d_size' % case] = np.array(sizes)
        arrays['c%d_full' % case] = np.array(fulls)
        arrays['c%d_after_clear' % case] = np.array([float(q.CP), float(q.C), float(q.size)])
    arrays['n_cases'] = np.array(3)
    out = os.path.join(outdir or os.path.join(ROOT, 'tests', 'golden'), 'competence_queue.npz')
    np.savez_compressed(out, **arrays)
    print('wrote', out)


if __name__ == '__main__':
    main()
