import sys
sys.path.insert(0, '.')
import numpy as np, torch
from curious_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda', 0)
def run(M, K, a_trans, b_trans, epi, dbg=0):
    A = torch.randn((K, M) if a_trans else (M, K), device=dev)
    Bm = torch.randn((256, K) if b_trans else (K, 256), device=dev)
    Cm = torch.empty((M, 256), device=dev)
    aux = torch.randn((M, 256), device=dev)
    bias = torch.randn(256, device=dev)
    ws = torch.empty(max(4, lib.cur_tc_gemm_workspace_floats(M, 256, K, a_trans)), device=dev)
    tl = torch.zeros(128, dtype=torch.int64, device=dev)
    def call():
        _lib.check(lib.cur_tc_gemm(_lib.stream_ptr(), A.data_ptr(), A.shape[1], a_trans, Bm.data_ptr(), Bm.shape[1], b_trans,
                                   Cm.data_ptr(), 256, M, 256, K, bias.data_ptr() if epi == 1 else None,
                                   aux.data_ptr() if epi == 2 else None, 256, epi, ws.data_ptr()), 'tc')
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): call()
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 20
    flops = 2.0 * M * 256 * K
    print('M=%d K=%d a_trans=%d b_trans=%d epi=%d: %.1f us  %.1f TFLOP/s useful (x3 on the tensor pipe)' % (M, K, a_trans, b_trans, epi, us, flops / us / 1e6))
    tl[120] = dbg
    lib.cur_tc_gemm_timeline(tl.data_ptr())
    call(); torch.cuda.synchronize()
    lib.cur_tc_gemm_timeline(None)
    t = tl.cpu().numpy()
    t0 = t[0]
    nkb = min(16, (K if not a_trans or M >= 1024 else 256) // 32)
    print('  setup %d  total %d cycles' % (t[1] - t0, t[2] - t0))
    print('  producer issue :', [int(t[8 + k] - t0) for k in range(nkb)])
    print('  split wait/done:', [(int(t[24 + 2 * k] - t0), int(t[25 + 2 * k] - t0)) for k in range(nkb)])
    print('  mma wait/issued:', [(int(t[56 + 2 * k] - t0), int(t[57 + 2 * k] - t0)) for k in range(nkb)])
    print('  epilogue acc ready %d, chunks %s' % (t[88] - t0, [int(t[89 + c] - t0) for c in range(4)]))
run(16384, 256, 0, 0, 1)
run(16384, 256, 0, 1, 2)
run(256, 16384, 1, 0, 0)
run(128, 256, 0, 0, 1)
