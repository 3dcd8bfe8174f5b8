import sys
sys.path.insert(0, '.')
import numpy as np, torch
from tests.test_tc_gemm_gpu import _run, _ref
rng = np.random.RandomState(0)
for a_trans in (0, 1):
    for b_trans in (0, 1):
        for (M, K) in ((256, 32), (256, 256), (1024, 64)):
            a = rng.randn(M, K).astype(np.float32)
            b = rng.randn(K, 256).astype(np.float32)
            A = np.ascontiguousarray(a.T) if a_trans else a
            B = np.ascontiguousarray(b.T) if b_trans else b
            got = _run(A, B, a_trans, b_trans)
            ref, mag = _ref(A, B, a_trans, b_trans)
            err = np.abs(got - ref) / mag
            bad_rows = np.where(err.max(1) > 1e-5)[0]
            bad_cols = np.where(err.max(0) > 1e-5)[0]
            print('a_trans=%d b_trans=%d M=%4d K=%3d max err/mag %.2e  bad rows %d (first %s) bad cols %d (first %s)' % (
                a_trans, b_trans, M, K, err.max(), len(bad_rows), bad_rows[:4], len(bad_cols), bad_cols[:4]), flush=True)
