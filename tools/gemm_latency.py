#!/usr/bin/env python
"""Latency model of the tcgen05 GEMM building block: one 128 x N tile, K = 32*s.
The slope over s is the per-stage time of the serial load->stage->MMA chain, the intercept
the fixed cost (launch, barrier init, TMEM alloc, epilogue)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mms_answer_selection_b200 import _lib

h = _lib.Handle()
h.set_stream(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())

def t_us(M, N, K, a_mn=0, b_mn=0, iters=20):
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda")
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda")
    C = torch.empty((M, N), device="cuda")
    lda = M if a_mn else K; ldb = N if b_mn else K
    f = lambda: _lib.check(_lib.lib().mms_tc_gemm_f32(h.ptr, p(A), lda, a_mn, p(B), ldb, b_mn, p(C), N, M, N, K, 1, MODE))
    for _ in range(3): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    return best

only = sys.argv[1] if len(sys.argv) > 1 else None
MODE = 0x100 if 'tma' in sys.argv else 0
if only == "one":
    print("one tile 128x256x2048: %.1f us" % t_us(128, 256, 2048, iters=3))
    sys.exit(0)
for (M, N) in [(128, 256), (128, 64), (128 * 148, 256)]:
    for s in (1, 2, 4, 8, 16, 64, 256):
        print("M=%6d N=%3d K=%5d (%3d stages): %8.1f us" % (M, N, 32 * s, s, t_us(M, N, 32 * s)), flush=True)
