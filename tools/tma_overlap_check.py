#!/usr/bin/env python
"""Does the TMA-fed GEMM accept operand views whose rows OVERLAP (leading dimension smaller than the row length)?
A(m, k) = buf[m*lda + k] with lda < K (K-major) and B(n, k) = buf[k*ldb + n] with ldb < N (MN-major)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mms_answer_selection_b200 import _lib

def tf32(a):
    b = a.astype(np.float32).view(np.uint32)
    b = (b + 0x1000) & 0xFFFFE000
    return b.view(np.float32)

h = _lib.Handle()
h.set_stream(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())
rng = np.random.default_rng(0)
for name, (M, N, K, lda) in {"K-major A, lda<K (conv dx: 520 over rows of 104)": (4000, 300, 520, 104),
                             "K-major A, lda<K (conv fwd: 1500 over rows of 300)": (4000, 100, 1500, 300)}.items():
    buf = tf32(rng.uniform(-1, 1, (M + 8) * lda))
    Bm = tf32(rng.uniform(-1, 1, (N, K)))
    A = np.lib.stride_tricks.as_strided(buf, (M, K), (lda * 4, 4))
    ref = A.astype(np.float64) @ Bm.astype(np.float64).T
    tb, tB, tC = torch.from_numpy(buf).cuda(), torch.from_numpy(Bm).cuda(), torch.zeros((M, N), device="cuda")
    rc = _lib.lib().mms_tc_gemm_f32(h.ptr, p(tb), lda, 0, p(tB), K, 0, p(tC), N, M, N, K, 1, 0x100)
    torch.cuda.synchronize()
    err = np.abs(tC.cpu().numpy() - ref).max() / np.abs(ref).max()
    print(name, "rc", rc, "err %.2e" % err, _lib.lib().mms_last_error() if rc else "")
# MN-major B with ldb < N: dW[c][n] = sum_r G[r][c] x[r*D + n], n < kh*D
R, C, D, kh = 6000, 100, 300, 5
x = tf32(rng.uniform(-1, 1, (R + kh) * D)); G = tf32(rng.uniform(-1, 1, (R, 104)))
G[:, 100:] = 0
Bv = np.lib.stride_tricks.as_strided(x, (R, kh * D), (D * 4, 4))
ref = G[:, :C].astype(np.float64).T @ Bv.astype(np.float64)
tx, tG, tC = torch.from_numpy(x).cuda(), torch.from_numpy(G).cuda(), torch.zeros((C, kh * D), device="cuda")
rc = _lib.lib().mms_tc_gemm_f32(h.ptr, p(tG), 104, 1, p(tx), D, 1, p(tC), kh * D, C, kh * D, R, 4, 0x102)
torch.cuda.synchronize()
print("MN-major B, ldb<N (conv dW merged)", "rc", rc, "err %.2e" % (np.abs(tC.cpu().numpy() - ref).max() / np.abs(ref).max()),
      _lib.lib().mms_last_error() if rc else "")
