import sys
sys.path.insert(0, '.')
import numpy as np, torch
from tests.test_tc_gemm_gpu import _run, _ref
np.set_printoptions(linewidth=200, precision=4, suppress=True)
rng = np.random.RandomState(0)
for a_trans in (0, 1):
    for b_trans in (0, 1):
        for K in (32, 64, 256):
            M = 128
            for ks in ('all', 0, 1, 2, 3):
                a = rng.randn(M, K).astype(np.float32)
                b = rng.randn(K, 256).astype(np.float32)
                if ks != 'all':
                    if K != 32:
                        continue
                    mask = np.zeros(K, bool); mask[8 * ks:8 * ks + 8] = True
                    a[:, ~mask] = 0
                A = np.ascontiguousarray(a.T) if a_trans else a
                B = np.ascontiguousarray(b.T) if b_trans else b
                got = _run(A, B, a_trans, b_trans)
                ref, mag = _ref(A, B, a_trans, b_trans)
                err = np.abs(got - ref) / mag
                bad_rows = np.where(err.max(1) > 1e-5)[0]
                bad_cols = np.where(err.max(0) > 1e-5)[0]
                print('a_trans=%d b_trans=%d K=%3d ks=%-3s max err/mag %.2e  bad rows %d (first %s) bad cols %d (first %s)' % (
                    a_trans, b_trans, K, ks, err.max(), len(bad_rows), bad_rows[:6], len(bad_cols), bad_cols[:6]))
# plain single-pass check: is the result close to single TF32 (1e-3)?  which term is missing?
