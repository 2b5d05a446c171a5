#!/usr/bin/env python
"""Times the tcgen05 TF32 GEMM building block (mms_tc_gemm_f32) against cuBLAS TF32
(torch.matmul with allow_tf32) on the contraction shapes of the MMS path, and measures
the TF32 peak the same way MEASURED_PEAKS.json measured bf16 (8192^3, best of 10).

    python tools/gemm_bench.py [--json out.json]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mms_answer_selection_b200 import _lib  # noqa: E402


def time_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    tot = 0.0
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        best = min(best, t); tot += t
    return best, tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = True
    h = _lib.Handle()
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    out = {}
    # (name, M, N, K, a_mn, b_mn)
    shapes = [
        ("peak 8192^3 (A K-major, B K-major)", 8192, 8192, 8192, 0, 0),
        ("peak 8192^3 (A K-major, B MN-major)", 8192, 8192, 8192, 0, 1),
        ("C3 T=Q M_k (163840x300x300)", 163840, 300, 300, 0, 1),
        ("C5 T=qW (16384x1024x1024)", 16384, 1024, 1024, 0, 1),
        ("C5 dW (1024x1024x16384, both MN-major)", 1024, 1024, 16384, 1, 1),
        ("C4 slab (1000x131072x1024)", 1000, 131072, 1024, 0, 0),
    ]
    for name, M, N, K, a_mn, b_mn in shapes:
        A = torch.randn((K, M) if a_mn else (M, K), device="cuda")
        B = torch.randn((K, N) if b_mn else (N, K), device="cuda")
        C = torch.empty((M, N), device="cuda")
        lda = M if a_mn else K
        ldb = N if b_mn else K
        ksplit = 1
        mode = 0
        if M * N <= 1024 * 1024 and K >= 8192:
            ksplit, mode = 8, 2

        def ours():
            _lib.check(_lib.lib().mms_tc_gemm_f32(h.ptr, p(A), lda, a_mn, p(B), ldb, b_mn, p(C), N, M, N, K, ksplit, mode))

        def ours_tma():
            _lib.check(_lib.lib().mms_tc_gemm_f32(h.ptr, p(A), lda, a_mn, p(B), ldb, b_mn, p(C), N, M, N, K, ksplit,
                                                  mode | 0x100))

        Am = A.t() if a_mn else A
        Bm = B if b_mn else B.t()

        def cublas():
            torch.matmul(Am, Bm, out=C)

        fl = 2.0 * M * N * K
        bo, ao = time_ms(ours)
        bt, at = time_ms(ours_tma)
        bc, ac = time_ms(cublas)
        out[name] = {"ours_tflops_best": fl / bo / 1e9, "ours_tflops_avg": fl / ao / 1e9,
                     "cublas_tf32_tflops_best": fl / bc / 1e9, "cublas_tf32_tflops_avg": fl / ac / 1e9,
                     "ours_ms": bo, "cublas_ms": bc,
                     "ours_tma_tflops_best": fl / bt / 1e9, "ours_tma_ms": bt}
        print("%-45s staged %7.1f TF/s (%.3f ms)  tma %7.1f TF/s (%.3f ms)  cuBLAS tf32 %7.1f TF/s (%.3f ms)"
              % (name, fl / bo / 1e9, bo, fl / bt / 1e9, bt, fl / bc / 1e9, bc), flush=True)
        del A, B, C
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
