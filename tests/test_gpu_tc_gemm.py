"""GPU tests of the tcgen05 TF32 GEMM building block (mms_tc_gemm_f32) against a float64
reference, for every operand majorness, ragged sizes, unaligned leading dimensions,
split-K with the atomic epilogue and the accumulate epilogue.  Tolerance: 1e-3 of the
output's largest magnitude (TF32 operands rounded to nearest, fp32 accumulation)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from mms_answer_selection_b200 import _lib   # noqa: E402

TOL = 1e-3


def tf32_exact(x):
    """Round to the nearest TF32 value (10-bit mantissa), ties away from zero, like cvt.rna."""
    i = x.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def run(M, N, K, a_mn, b_mn, ksplit=1, mode=0, pad=0, seed=0, tma=False):
    g = torch.Generator(device="cuda").manual_seed(seed + M + 7 * N + 13 * K)
    lda = (M if a_mn else K) + pad
    ldb = (N if b_mn else K) + pad
    A = torch.rand(((K if a_mn else M), lda), device="cuda", generator=g) - 0.5
    B = torch.rand(((K if b_mn else N), ldb), device="cuda", generator=g) - 0.5
    if tma:
        A, B = tf32_exact(A), tf32_exact(B)
        mode_flag = 0x100
    else:
        mode_flag = 0
    C0 = torch.rand((M, N + pad), device="cuda", generator=g) - 0.5
    C = C0.clone() if mode else torch.full((M, N + pad), 7.0, device="cuda")
    if mode == 2:
        C = C0.clone()
    h = _lib.Handle()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().mms_tc_gemm_f32(h.ptr, p(A), lda, a_mn, p(B), ldb, b_mn, p(C), N + pad, M, N, K,
                                          ksplit, mode | mode_flag))
    torch.cuda.synchronize()
    Am = (A[:, :M].T if a_mn else A[:, :K]).double()
    Bm = (B[:, :N].T if b_mn else B[:, :K]).double()
    ref = Am @ Bm.T
    if mode:
        ref = ref + C0[:, :N].double()
    got = C[:, :N].double()
    err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
    if pad and not mode:
        assert (C[:, N:] == 7.0).all()          # nothing written outside the N columns
    return err


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("shape", [(128, 64, 32), (128, 256, 64), (200, 300, 300), (40, 40, 50), (1, 1, 1),
                                   (333, 130, 77), (1024, 1024, 1024)])
def test_tc_gemm_majorness(shape, a_mn, b_mn):
    M, N, K = shape
    assert run(M, N, K, a_mn, b_mn) <= TOL


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (0, 1)])
def test_tc_gemm_unaligned_leading_dims(a_mn, b_mn):
    assert run(150, 90, 50, a_mn, b_mn, pad=1) <= TOL
    assert run(150, 90, 50, a_mn, b_mn, pad=2) <= TOL


def test_tc_gemm_split_k_atomic_and_accumulate():
    assert run(300, 300, 20000, 1, 1, ksplit=8, mode=2) <= TOL
    assert run(300, 300, 4096, 0, 0, ksplit=1, mode=1) <= TOL
    assert run(128, 128, 10000, 1, 0, ksplit=7, mode=2) <= TOL


# ---- the TMA-fed kernel: operands are TF32-exact, so the result must match fp64 to fp32 rounding
@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("shape", [(128, 64, 32), (128, 256, 64), (200, 300, 300), (40, 40, 52), (4, 4, 4),
                                   (333, 132, 76), (1024, 1024, 1024)])
def test_tc_gemm_tma_majorness(shape, a_mn, b_mn):
    M, N, K = shape
    assert run(M, N, K, a_mn, b_mn, tma=True) <= 1e-5


def test_tc_gemm_tma_split_k_and_accumulate():
    assert run(300, 300, 20000, 1, 1, ksplit=8, mode=2, tma=True) <= 1e-4
    assert run(300, 300, 4096, 0, 0, ksplit=1, mode=1, tma=True) <= 1e-4
    assert run(128, 128, 10000, 1, 0, ksplit=7, mode=2, tma=True) <= 1e-4


def test_tc_gemm_tma_unaligned_falls_back_to_staged_kernel():
    # leading dimensions that are not 16-byte multiples cannot be described to TMA
    assert run(150, 90, 50, 0, 0, pad=1, tma=True) <= 1e-5


# ---- the CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 tiles): taken when M > 128, N >= 256 and there are at
# ---- least 74 such tiles; TF32-exact operands, so the result must match fp64 to fp32 rounding
@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("shape", [(2304, 2304, 160), (2000, 2500, 100), (1000, 20000, 1024), (129 + 256 * 3, 256 * 20, 36)])
def test_tc_gemm_pair_majorness(shape, a_mn, b_mn):
    M, N, K = shape
    assert run(M, N, K, a_mn, b_mn, tma=True) <= 1e-5


def test_tc_gemm_pair_split_k_and_accumulate():
    assert run(1024, 1024, 4096, 1, 1, ksplit=5, mode=2, tma=True) <= 1e-4
    assert run(1024, 1024, 4100, 0, 0, ksplit=6, mode=2, tma=True) <= 1e-4
    assert run(2304, 2304, 200, 0, 1, ksplit=1, mode=1, tma=True) <= 1e-4
