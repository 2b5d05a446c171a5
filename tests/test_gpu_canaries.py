"""Out-of-bounds evidence without compute-sanitizer (closed on this pool): every buffer a kernel may write is carved out
of one arena with 4 KB guard bands of a sentinel pattern on both sides; after the hot path has run over ragged and
full-size shapes the guard bands must be untouched.  Covers the kernels that write with red.global / vector stores /
TMEM epilogues: Embed forward (incl. the staged copy) and both backward modes, SimCross mode 2 forward / backward (fused
tcgen05 kernels, blocked U export, dM), the 2-D convolution trio, top-k candidate scoring and the gradient exchange."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mms_answer_selection_b200 as mms  # noqa: E402
from mms_answer_selection_b200 import _lib, synth  # noqa: E402

GUARD = 1024            # floats (4 KB) on each side
SENTINEL = -1.2345678e30


class Arena(object):
    def __init__(self, floats):
        self.buf = torch.full((floats,), SENTINEL, device="cuda")
        self.off = 0
        self.guards = []

    def take(self, *shape, fill=None):
        n = int(np.prod(shape))
        start = self.off + GUARD
        start = (start + 63) // 64 * 64                       # 256-byte aligned payload
        self.guards.append((self.off, start))
        t = self.buf[start:start + n].view(*shape)
        if fill is not None:
            t.copy_(fill) if torch.is_tensor(fill) else t.fill_(fill)
        self.off = start + n
        self.guards.append((self.off, self.off + GUARD))
        return t

    def check(self, what):
        torch.cuda.synchronize()
        for a, b in self.guards:
            g = self.buf[a:b]
            assert bool((g == SENTINEL).all().item()), "%s wrote outside its buffers (guard band %d..%d)" % (what, a, b)


p = lambda t: ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("N,L,D,mc,V", [(7, 40, 300, 4, 500), (515, 40, 300, 4, 3000), (5, 13, 52, 3, 100)])
def test_hot_path_stays_inside_its_buffers(N, L, D, mc, V):
    d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
    ar = Arena(4 * N * L * D + 2 * N * mc * L * L + 2 * V * D + 2 * mc * D * D + 2 * mc * L * L + 64 * GUARD + 4096)
    idx_q = torch.from_numpy(d["idx_q"]).cuda(); idx_a = torch.from_numpy(d["idx_a"]).cuda()
    W = torch.from_numpy(d["W"]).cuda(); b = torch.from_numpy(d["b"]).cuda()
    Mw = torch.from_numpy(d["M"]).cuda(); B = torch.from_numpy(d["B"]).cuda(); dS = torch.from_numpy(d["dS"]).cuda()
    q, a = ar.take(N, L, D), ar.take(N, L, D)
    S = ar.take(N, mc, L, L)
    dq, da = ar.take(N, L, D), ar.take(N, L, D)
    dM, dB = ar.take(mc, D, D, fill=0.0), ar.take(mc, L, L, fill=0.0)
    dW, db = ar.take(V, D, fill=0.0), ar.take(D, fill=0.0)
    L_ = _lib.lib()
    he, hs = _lib.Handle(), _lib.Handle()
    he.set_option(_lib.MMS_OPT_STAGE_TF32, 1)
    hs.set_option(_lib.MMS_OPT_REUSE_FORWARD, 1)
    for det in (0, 1):
        he.set_option(_lib.MMS_OPT_EMBED_DETERMINISTIC, det)
        _lib.check(L_.mms_embed_forward_f32(he.ptr, p(idx_q), p(W), p(b), p(q), N * L, D, V))
        _lib.check(L_.mms_embed_forward_f32(he.ptr, p(idx_a), p(W), p(b), p(a), N * L, D, V))
        _lib.check(L_.mms_simcross_forward_f32(hs.ptr, 2, p(q), p(a), p(Mw), p(B), p(S), None, None, N, L, L, D, mc))
        _lib.check(L_.mms_simcross_backward_f32(hs.ptr, 2, p(q), p(a), p(Mw), p(S), p(dS), None, None, p(dq), p(da), p(dM),
                                                p(dB), N, L, L, D, mc, 1, 1))
        _lib.check(L_.mms_embed_backward_f32(he.ptr, p(idx_q), p(dq), p(dW), p(db), N * L, D, V))
        _lib.check(L_.mms_embed_backward_f32(he.ptr, p(idx_a), p(da), p(dW), p(db), N * L, D, V))
        ar.check("Embed / SimCross (deterministic=%d)" % det)
    assert torch.isfinite(S).all() and torch.isfinite(dW).all()


@pytest.mark.parametrize("N,C,H,W,Co,k", [(5, 4, 40, 40, 32, 5), (37, 32, 9, 9, 64, 5), (3, 3, 11, 7, 5, 3)])
def test_conv2d_stays_inside_its_buffers(N, C, H, W, Co, k):
    OH, OW = H - k + 1, W - k + 1
    ar = Arena(2 * N * C * H * W + 2 * N * Co * OH * OW + 2 * Co * C * k * k + 2 * Co + 32 * GUARD + 4096)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = ar.take(N, C, H, W, fill=torch.rand((N, C, H, W), device="cuda", generator=g) - 0.5)
    Wt = ar.take(Co, C, k, k, fill=torch.rand((Co, C, k, k), device="cuda", generator=g) - 0.5)
    b = ar.take(Co, fill=0.1)
    y = ar.take(N, Co, OH, OW)
    dy = ar.take(N, Co, OH, OW, fill=torch.rand((N, Co, OH, OW), device="cuda", generator=g) - 0.5)
    dW, db, dx = ar.take(Co, C, k, k, fill=0.0), ar.take(Co, fill=0.0), ar.take(N, C, H, W)
    h = _lib.Handle()
    _lib.check(_lib.lib().mms_conv2d_forward_f32(h.ptr, p(x), p(Wt), p(b), p(y), N, C, H, W, Co, k, k))
    _lib.check(_lib.lib().mms_conv2d_backward_f32(h.ptr, p(x), p(Wt), p(dy), p(dW), p(db), p(dx), N, C, H, W, Co, k, k))
    ar.check("conv2d")
    assert torch.isfinite(y).all() and torch.isfinite(dx).all() and torch.isfinite(dW).all()


def test_topk_and_exchange_stay_inside_their_buffers():
    Nq, Nc, K, k = 9, 70_001, 96, 37
    Qn, Cn, Wn = synth.make_rerank(Nq, Nc, K, seed=5)
    ar = Arena(Nq * K + Nq * k * 3 + Nq * K + 16 * GUARD + 4096)
    Q = ar.take(Nq, K, fill=torch.from_numpy(Qn).cuda())
    QW = ar.take(Nq, K)
    top_s = ar.take(Nq, k)
    top_i_f = ar.take(Nq, 2 * k)                               # int64 lists inside the float arena
    C, W = torch.from_numpy(Cn).cuda(), torch.from_numpy(Wn).cuda()
    h = _lib.Handle()
    _lib.check(_lib.lib().mms_rerank_topk_f32(h.ptr, p(Q), p(C), p(W), p(QW), p(top_s), p(top_i_f), Nq, Nc, K, K, k, 0))
    ar.check("top-k candidate scoring")
    idx = top_i_f.view(torch.int64).reshape(Nq, k)
    assert int(idx.min()) >= 0 and int(idx.max()) < Nc
