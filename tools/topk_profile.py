#!/usr/bin/env python
"""Where does per-query top-k candidate scoring spend its time?  Per-kernel device times (CUDA events around every
launch) of mms_rerank_topk_f32 at C4 (1000 queries x 1M candidates, K = 1024), wall time of the call, and the same for
the full-score call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mms_answer_selection_b200 import _lib, synth
from mms_answer_selection_b200.rerank import Reranker
Nq, Nc, K = 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 1024
g = torch.Generator(device="cuda").manual_seed(1)
Q = torch.randn((Nq, K), device="cuda", generator=g) / K ** 0.5
C = torch.randn((Nc, K), device="cuda", generator=g) / K ** 0.5
W = (torch.rand((K, K), device="cuda", generator=g) * 2 - 1) * (3.0 / K) ** 0.5
rr = Reranker(W, k=100)
for prepared in (False, True):
    if prepared:
        rr.prepare(C)
    cand = None if prepared else C
    rr.local_topk(Q, cand); torch.cuda.synchronize()
    t0 = time.perf_counter(); rr.local_topk(Q, cand); t_issue = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    rr.handle.profile_enable(True)
    rr.local_topk(Q, cand); torch.cuda.synchronize()
    rep = rr.handle.profile_report()
    rr.handle.profile_enable(False)
    print("prepared" if prepared else "raw", "host issue %.2f ms, wall %.2f ms" % (t_issue * 1e3, t_all * 1e3))
    for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
        print("   %-28s x%-4d %.3f ms" % (k, n, ms))
