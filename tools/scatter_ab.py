#!/usr/bin/env python
"""Embed scatter-add of a QA step, per-layer atomic kernels (q and a on two streams, as MMSNet runs them) against the
grouped pair call (plan timed separately: MMSNet hides it beside the forward), for several per-GPU batch sizes."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mms_answer_selection_b200 import _lib, synth

L, D, V = 40, 300, 60002
p = lambda t: ctypes.c_void_p(t.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def time_ms(fn, iters=10):
    best = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best.append(e0.elapsed_time(e1))
    return float(np.median(best[2:]))


for N in (512, 1024, 2048, 4096):
    rng = np.random.default_rng(1)
    iq = torch.from_numpy(synth.make_indices(rng, N, L, V, 3, 20).reshape(-1)).cuda()
    ia = torch.from_numpy(synth.make_indices(rng, N, L, V, 5, 40).reshape(-1)).cuda()
    M = N * L
    gq = torch.randn((M, D), device="cuda") * 1e-6; ga = torch.randn((M, D), device="cuda") * 1e-6
    dW = torch.zeros((V, D), device="cuda"); db = torch.zeros(D, device="cuda")
    hq, ha = _lib.Handle(), _lib.Handle()
    s2 = torch.cuda.Stream()
    Lb = _lib.lib()

    def per_layer():
        main = torch.cuda.current_stream()
        s2.wait_stream(main)
        ha.set_stream(s2.cuda_stream); hq.set_stream(main.cuda_stream)
        _lib.check(Lb.mms_embed_backward_f32(ha.ptr, p(ia), p(ga), p(dW), p(db), M, D, V))
        _lib.check(Lb.mms_embed_backward_f32(hq.ptr, p(iq), p(gq), p(dW), p(db), M, D, V))
        main.wait_stream(s2)

    def plan():
        hq.set_stream(torch.cuda.current_stream().cuda_stream)
        _lib.check(Lb.mms_embed_plan_pair_f32(hq.ptr, p(iq), M, p(ia), M, V))

    def grouped():
        _lib.check(Lb.mms_embed_backward_pair_f32(hq.ptr, p(iq), p(gq), M, p(ia), p(ga), M, p(dW), p(db), D, V))

    def grouped_after_plan():
        plan(); grouped()

    t_pl = time_ms(per_layer); t_plan = time_ms(plan); t_both = time_ms(grouped_after_plan)
    print("pairs %5d (rows %6d): per-layer on two streams %.1f us | plan %.1f us + grouped reduce %.1f us" %
          (N, 2 * M, 1e3 * t_pl, 1e3 * t_plan, 1e3 * (t_both - t_plan)))
