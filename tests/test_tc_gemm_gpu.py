"""tcgen05 layer GEMM (csrc/tc_gemm.cu) against float64 NumPy: the three operand orientations the DDPG graph
uses at large batch - forward (util.py:56-107), dX = dY W^T and dW = X^T dY (tf.gradients, ddpg.py:443-449).

Tolerance (stated): error-compensated 3xTF32 must be at fp32 level - max |err| <= 2e-6 * sum_k |a||b| per element
(a plain fp32 FMA chain of length K is bounded by ~K * 6e-8 of the same quantity); a single TF32 pass would sit
at ~5e-4 and fail this by two orders of magnitude."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(A, B, a_trans, b_trans, bias=None, aux=None, epi=0):
    from curious_b200 import _lib
    lib = _lib.load()
    dev = torch.device('cuda', 0)
    M = A.shape[1] if a_trans else A.shape[0]
    K = A.shape[0] if a_trans else A.shape[1]
    N = B.shape[0] if b_trans else B.shape[1]
    dA, dB = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    dC = torch.full((M, N), float('nan'), dtype=torch.float32, device=dev)
    dbias = torch.from_numpy(bias).to(dev) if bias is not None else None
    daux = torch.from_numpy(aux).to(dev) if aux is not None else None
    assert lib.cur_tc_gemm_supported(M, N, K) == 1
    wsf = lib.cur_tc_gemm_workspace_floats(M, N, K, int(a_trans))
    ws = torch.empty(max(int(wsf), 4), dtype=torch.float32, device=dev)
    _lib.check(lib.cur_tc_gemm(_lib.stream_ptr(), dA.data_ptr(), A.shape[1], int(a_trans), dB.data_ptr(), B.shape[1],
                               int(b_trans), dC.data_ptr(), N, M, N, K, dbias.data_ptr() if dbias is not None else None,
                               daux.data_ptr() if daux is not None else None, N, epi, ws.data_ptr()), 'cur_tc_gemm')
    torch.cuda.synchronize()
    return dC.cpu().numpy()


def _ref(A, B, a_trans, b_trans, bias=None, aux=None, epi=0):
    a = A.astype(np.float64).T if a_trans else A.astype(np.float64)
    b = B.astype(np.float64).T if b_trans else B.astype(np.float64)
    c = a @ b
    mag = np.abs(a) @ np.abs(b)
    if bias is not None:
        c = c + bias.astype(np.float64)
        mag = mag + np.abs(bias)
    if epi == 1:
        c = np.maximum(c, 0.0)
    elif epi == 2:
        c = np.where(aux > 0, c, 0.0)
    return c, mag


def _check(got, ref, mag, epi=0):
    assert np.isfinite(got).all()
    err = np.abs(got - ref)
    if epi == 1:
        # a ReLU whose pre-activation is within rounding of 0 may land on either side
        err = np.where(np.abs(ref) <= 2e-6 * mag, 0.0, err)
    assert (err <= 2e-6 * mag + 1e-30).all(), 'max err/mag %.3e' % float((err / (mag + 1e-30)).max())


@pytest.mark.parametrize('M', [128, 1024, 4096])
def test_forward_layer(M):
    rng = np.random.RandomState(M)
    X = rng.randn(M, 256).astype(np.float32)
    W = (rng.uniform(-1, 1, (256, 256)) * 0.108).astype(np.float32)
    b = rng.randn(256).astype(np.float32) * 0.1
    got = _run(X, W, 0, 0, bias=b, epi=1)
    ref, mag = _ref(X, W, 0, 0, bias=b, epi=1)
    _check(got, ref, mag, epi=1)


@pytest.mark.parametrize('M', [128, 2048])
def test_backward_data(M):
    rng = np.random.RandomState(M + 1)
    dY = rng.randn(M, 256).astype(np.float32) * 1e-3
    W = (rng.uniform(-1, 1, (256, 256)) * 0.108).astype(np.float32)
    act = np.maximum(rng.randn(M, 256), 0).astype(np.float32)
    got = _run(dY, W, 0, 1, aux=act, epi=2)            # dX = dY W^T, W stored [in][out]
    ref, mag = _ref(dY, W, 0, 1, aux=act, epi=2)
    _check(got, ref, mag, epi=2)


@pytest.mark.parametrize('n', [1024, 16384])
def test_weight_gradient_split_k(n):
    rng = np.random.RandomState(n + 2)
    X = np.maximum(rng.randn(n, 256), 0).astype(np.float32)
    dY = rng.randn(n, 256).astype(np.float32) * 1e-3
    got = _run(X, dY, 1, 0)                            # dW = X^T dY
    ref, mag = _ref(X, dY, 1, 0)
    _check(got, ref, mag)


def test_k_tail_and_column_order():
    """Distinct values per (row, column): catches transposed / permuted tiles that random data would also catch,
    but with a readable failure; K = 96 exercises an odd number of ring rounds."""
    M, K = 256, 96
    A = (np.arange(M)[:, None] * 0.01 + np.arange(K)[None, :] * 0.001).astype(np.float32)
    B = np.zeros((K, 256), np.float32)
    B[np.arange(K), (np.arange(K) * 7) % 256] = 1.0
    got = _run(A, B, 0, 0)
    ref, mag = _ref(A, B, 0, 0)
    _check(got, ref, mag)


def test_first_layer_shapes_ride_on_tma_zero_fill():
    """K = 44 (first layer of main.pi: dimo + N) and M = 48 / 12 (first-layer weight gradients): the tails of the
    128 x 256 x 32 tiles are out-of-bounds boxes that the TMA unit zero-fills."""
    rng = np.random.RandomState(5)
    X = rng.randn(1024, 44).astype(np.float32)
    W = (rng.uniform(-1, 1, (44, 256)) * 0.14).astype(np.float32)
    b = rng.randn(256).astype(np.float32) * 0.1
    got = _run(X, W, 0, 0, bias=b, epi=1)
    ref, mag = _ref(X, W, 0, 0, bias=b, epi=1)
    _check(got, ref, mag, epi=1)
    for m in (48, 12):
        Xs = rng.randn(2048, m).astype(np.float32)
        dY = rng.randn(2048, 256).astype(np.float32) * 1e-3
        got = _run(Xs, dY, 1, 0)                       # dW0 = Xs^T dY, M = 48 / 12 rows of a 128-row tile
        ref, mag = _ref(Xs, dY, 1, 0)
        assert got.shape == (m, 256)
        _check(got, ref, mag)


def test_cta_pair_variant_in_a_subprocess():
    """CUR_TC_PAIR=1 launches the cta_group::2 variant (a cluster of two CTAs computes a 256 x 256 tile, each CTA stages
    only half of the B tile, rank 0 issues tcgen05.mma.cta_group::2 for both).  Kept as an opt-in experiment (measured
    slower than the single-CTA kernel, profiles/README.md); it must stay correct for all four operand orientations."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from tests.test_tc_gemm_gpu import _run, _ref, _check\n"
        "rng = np.random.RandomState(1)\n"
        "for a_t in (0, 1):\n"
        "    for b_t in (0, 1):\n"
        "        a = rng.randn(512, 96).astype(np.float32); b = rng.randn(96, 256).astype(np.float32)\n"
        "        A = np.ascontiguousarray(a.T) if a_t else a; B = np.ascontiguousarray(b.T) if b_t else b\n"
        "        got = _run(A, B, a_t, b_t); ref, mag = _ref(A, B, a_t, b_t); _check(got, ref, mag)\n"
        "print('pair ok')\n" % root)
    env = dict(os.environ, CUR_TC_PAIR='1')
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and 'pair ok' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
