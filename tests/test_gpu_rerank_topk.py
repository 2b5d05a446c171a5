"""Per-query top-k candidate scoring (mms_rerank_topk_f32, mms_topk_merge_f32, rerank.Reranker): the lists must be
exactly the top of the library's own full score matrix (same kernels, so bit-equal scores; ties by index), must agree
with an fp64 ranking wherever score gaps exceed the TF32 tolerance (north_star: "ranking order bit-exact wherever score
gaps exceed tolerance"), and must not depend on how the candidates are sharded."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mms_answer_selection_b200 as mms  # noqa: E402
from mms_answer_selection_b200 import _lib, synth  # noqa: E402
from mms_answer_selection_b200.rerank import Reranker  # noqa: E402

p = lambda t: ctypes.c_void_p(t.data_ptr())


def _full_scores(Q, C, W):
    h = _lib.Handle()
    Nq, K = Q.shape
    QW = torch.empty((Nq, W.shape[1]), device="cuda"); sc = torch.empty((Nq, C.shape[0]), device="cuda")
    _lib.check(_lib.lib().mms_rerank_scores_f32(h.ptr, p(Q), p(C), p(W), p(QW), p(sc), Nq, C.shape[0], K, W.shape[1]))
    torch.cuda.synchronize()
    return sc


def _reference_topk(sc, k, base=0):
    """(score desc, index asc) top-k of every row of a score matrix, on the host."""
    s = sc.cpu().numpy()
    out_s, out_i = [], []
    for row in s:
        order = np.lexsort((np.arange(row.size), -row.astype(np.float64)))[:k]
        out_s.append(row[order]); out_i.append(order + base)
    return np.stack(out_s), np.stack(out_i)


@pytest.mark.parametrize("Nq,Nc,K,k", [(37, 50_000, 128, 10), (8, 70_001, 300, 100), (5, 3_000, 64, 1000), (3, 40, 32, 64)])
def test_topk_is_the_top_of_the_full_score_matrix(Nq, Nc, K, k):
    Q, C, W = (torch.from_numpy(x).cuda() for x in synth.make_rerank(Nq, Nc, K, seed=Nq))
    C[Nc // 2] = C[3]; C[Nc - 1] = C[3]                     # exact ties: equal candidates at different indices
    sc = _full_scores(Q, C, W)
    rr = Reranker(W, k=k)
    top_s, top_i = rr.local_topk(Q, C, idx_base=1000)
    torch.cuda.synchronize()
    ref_s, ref_i = _reference_topk(sc, min(k, Nc), base=1000)
    kk = min(k, Nc)
    np.testing.assert_array_equal(top_s.cpu().numpy()[:, :kk], ref_s)
    np.testing.assert_array_equal(top_i.cpu().numpy()[:, :kk], ref_i)
    if k > Nc:                                               # unused slots
        assert torch.isinf(top_s[:, Nc:]).all() and (top_i[:, Nc:] == 2 ** 63 - 1).all()
    # prepared candidates give the same lists
    rr.prepare(C)
    ps, pi = rr.local_topk(Q, None, idx_base=1000)
    assert torch.equal(ps, top_s) and torch.equal(pi, top_i)


def test_topk_ranking_vs_fp64_where_gaps_exceed_tolerance():
    Nq, Nc, K, k = 16, 20_000, 256, 20
    Qn, Cn, Wn = synth.make_rerank(Nq, Nc, K, seed=9)
    Q, C, W = (torch.from_numpy(x).cuda() for x in (Qn, Cn, Wn))
    top_s, top_i = Reranker(W, k=k).local_topk(Q, C)
    ref = (Qn.astype(np.float64) @ Wn.astype(np.float64)) @ Cn.astype(np.float64).T
    tol = 1e-3 * np.abs(ref).max()
    got_i = top_i.cpu().numpy()
    for q in range(Nq):
        order = np.argsort(-ref[q], kind="stable")
        for r in range(k):
            # rank r is pinned when its fp64 score is separated from both neighbours by more than the tolerance
            lo = ref[q, order[r]] - ref[q, order[r + 1]]
            hi = ref[q, order[r - 1]] - ref[q, order[r]] if r else np.inf
            if lo > 2 * tol and hi > 2 * tol:
                assert got_i[q, r] == order[r], (q, r)
        assert abs(top_s[q, 0].item() - ref[q, order[0]]) <= tol


def test_merge_does_not_depend_on_the_sharding():
    """One GPU, the candidate set cut into 1, 3 and 7 shards: merged lists are identical."""
    Nq, Nc, K, k = 21, 30_000, 96, 50
    Q, C, W = (torch.from_numpy(x).cuda() for x in synth.make_rerank(Nq, Nc, K, seed=4))
    rr = Reranker(W, k=k)
    whole_s, whole_i = rr.local_topk(Q, C)
    for shards in (3, 7):
        bounds = np.linspace(0, Nc, shards + 1).astype(int)
        parts = [rr.local_topk(Q, C[a:b], idx_base=int(a)) for a, b in zip(bounds[:-1], bounds[1:])]
        rows_s = torch.cat([s for s, _ in parts], dim=1).contiguous()
        rows_i = torch.cat([i for _, i in parts], dim=1).contiguous()
        ms, mi = rr.merge_rows(rows_s, rows_i)
        assert torch.equal(ms, whole_s) and torch.equal(mi, whole_i), shards


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        Nq, Nc, K, k = 19, 40_000, 128, 25
        Q, C, W = (torch.from_numpy(x).cuda() for x in synth.make_rerank(Nq, Nc, K, seed=2))
        n = Nc // world
        s, i = Reranker(W, k=k).topk(Q, C[rank * n:(rank + 1) * n].contiguous())
        torch.cuda.synchronize()
        out[rank] = (s.cpu().numpy(), i.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_multigpu_topk_merge_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    Q, C, W = (torch.from_numpy(x).cuda() for x in synth.make_rerank(19, 40_000, 128, seed=2))
    s, i = Reranker(W, k=25).local_topk(Q, C)
    for r in (0, 1):
        np.testing.assert_array_equal(out[r][0], s.cpu().numpy())
        np.testing.assert_array_equal(out[r][1], i.cpu().numpy())


def test_topk_fused_epilogue_and_its_overflow_fallback():
    """Large enough for the threshold-epilogue form (scores never written): lists equal the top of the full score matrix;
    and an adversarial candidate order -- every query's scores rise with the candidate index, so every new candidate
    beats the current threshold and the candidate buffers overflow -- still gives the exact lists (stored fallback)."""
    Nq, Nc, K, k = 24, 120_000, 64, 50
    Q, C, W = (torch.from_numpy(x).cuda() for x in synth.make_rerank(Nq, Nc, K, seed=11))
    rr = Reranker(W, k=k)
    s1, i1 = rr.local_topk(Q, C, idx_base=7)
    ref_s, ref_i = _reference_topk(_full_scores(Q, C, W), k, base=7)
    np.testing.assert_array_equal(s1.cpu().numpy(), ref_s); np.testing.assert_array_equal(i1.cpu().numpy(), ref_i)
    # rising scores: q > 0, W = identity-like positive, candidates scaled by their index
    Qp = torch.rand((Nq, K), device="cuda") + 0.5
    Wp = torch.eye(K, device="cuda")
    ramp = torch.arange(Nc, device="cuda", dtype=torch.float32).reshape(-1, 1) / Nc + 0.1
    Cp = (torch.ones((Nc, K), device="cuda") * ramp).contiguous()
    rr2 = Reranker(Wp, k=k)
    s2, i2 = rr2.local_topk(Qp, Cp)
    ref_s2, ref_i2 = _reference_topk(_full_scores(Qp, Cp, Wp), k)
    np.testing.assert_array_equal(s2.cpu().numpy(), ref_s2); np.testing.assert_array_equal(i2.cpu().numpy(), ref_i2)
