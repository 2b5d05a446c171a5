"""Candidate scoring with per-query top-k, candidates sharded over the ranks (BASELINE.json configs[3]):

    every rank:   scores = (Q W) C_shard^T  ->  local top-k per query, global candidate indices   (mms_rerank_topk_f32)
    all ranks:    all-gather of the (Nq, k) lists (torch.distributed, 12 k Nq bytes per rank)     (plumbing)
    every rank:   merge of the world x k entries per query, ties by candidate index                (mms_topk_merge_f32)

The result is the same whatever the sharding: the order is total (score descending, candidate index ascending).
The consumer in the reference ranks each query's candidates by score (do_trec_qa_clean.py:617-650, map_layer.cpp:41-100).
"""
import ctypes

import torch
import torch.distributed as dist

from ._lib import Handle, c_p, check, lib


class Reranker(object):
    def __init__(self, W, k=10, group=None):
        """``W`` (K1, K2) float32 CUDA tensor; ``k`` entries per query."""
        if not W.is_cuda or W.dtype != torch.float32:
            raise RuntimeError("candidate scoring runs on the GPU in float32 only (no CPU fallback)")
        self.W, self.k, self.group = W.contiguous(), int(k), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.handle = Handle()
        self._prepared = None

    def prepare(self, C):
        """A static candidate shard: round it to the tensor-core operand format once (mms_rerank_prepare_f32)."""
        Nc, K2 = C.shape
        Cr = torch.empty((Nc, (K2 + 3) // 4 * 4), device=C.device, dtype=torch.float32)
        self.handle.set_stream(torch.cuda.current_stream().cuda_stream)
        check(lib().mms_rerank_prepare_f32(self.handle.ptr, c_p(C.data_ptr()), c_p(Cr.data_ptr()), Nc, K2))
        self._prepared = (Cr, Nc, K2)
        return Cr

    def local_topk(self, Q, C=None, idx_base=0):
        """(scores (Nq, k), idx (Nq, k) int64) of this rank's candidates; ``C=None`` scores the prepared shard."""
        Nq, K1 = Q.shape
        K2 = self.W.shape[1]
        QW = torch.empty((Nq, K2), device=Q.device, dtype=torch.float32)
        top_s = torch.empty((Nq, self.k), device=Q.device, dtype=torch.float32)
        top_i = torch.empty((Nq, self.k), device=Q.device, dtype=torch.int64)
        self.handle.set_stream(torch.cuda.current_stream().cuda_stream)
        if C is None:
            Cr, Nc, _ = self._prepared
            fn, cand = lib().mms_rerank_topk_prepared_f32, Cr
        else:
            Nc = C.shape[0]
            fn, cand = lib().mms_rerank_topk_f32, C.contiguous()
        check(fn(self.handle.ptr, c_p(Q.data_ptr()), c_p(cand.data_ptr()), c_p(self.W.data_ptr()), c_p(QW.data_ptr()),
                 c_p(top_s.data_ptr()), c_p(top_i.data_ptr()), Nq, Nc, K1, K2, self.k, int(idx_base)))
        return top_s, top_i

    def merge(self, top_s, top_i):
        """All-gather the per-rank lists and merge them per query (every rank ends with the same global lists)."""
        if self.world == 1:
            return top_s, top_i
        Nq, k = top_s.shape
        all_s = torch.empty((self.world, Nq, k), device=top_s.device, dtype=torch.float32)
        all_i = torch.empty((self.world, Nq, k), device=top_s.device, dtype=torch.int64)
        dist.all_gather_into_tensor(all_s, top_s.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_i, top_i.contiguous(), group=self.group)
        rows_s = all_s.permute(1, 0, 2).reshape(Nq, self.world * k).contiguous()
        rows_i = all_i.permute(1, 0, 2).reshape(Nq, self.world * k).contiguous()
        return self.merge_rows(rows_s, rows_i)

    def merge_rows(self, rows_s, rows_i):
        Nq, n = rows_s.shape
        out_s = torch.empty((Nq, self.k), device=rows_s.device, dtype=torch.float32)
        out_i = torch.empty((Nq, self.k), device=rows_s.device, dtype=torch.int64)
        self.handle.set_stream(torch.cuda.current_stream().cuda_stream)
        check(lib().mms_topk_merge_f32(self.handle.ptr, c_p(rows_s.data_ptr()), c_p(rows_i.data_ptr()), n, n,
                                       c_p(out_s.data_ptr()), c_p(out_i.data_ptr()), Nq, self.k))
        return out_s, out_i

    def topk(self, Q, C=None, idx_base=None):
        """Global top-k of every query over all ranks' candidates.  ``idx_base`` defaults to rank * local count
        (equal shards)."""
        Nc = self._prepared[1] if C is None else C.shape[0]
        base = self.rank * Nc if idx_base is None else idx_base
        return self.merge(*self.local_topk(Q, C, base))
