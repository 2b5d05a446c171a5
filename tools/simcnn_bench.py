#!/usr/bin/env python
"""network_v4 up to Flatten (Embed x2 -> SimCross -> Dropout -> (Conv5x5 + BN -> AvePool -> TanH) x 2) at N pairs per step:
step time as one CUDA graph and per-kernel device times (CUDA events around every launch of an eager step).
    python tools/simcnn_bench.py [N] [iters]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mms_answer_selection_b200 import synth
from mms_answer_selection_b200.simcnn import SimCNNNet


def run(N=4096, iters=5):
    c3 = synth.CONFIGS["c3"]
    L, D, mc, V = c3["L"], c3["D"], c3["mc"], c3["V"]
    d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
    net = SimCNNNet(N, L, D, mc, V)
    net.embed_q.blobs[0].set_cpu_data(d["W"])
    net.sim.blobs[0].set_cpu_data((np.random.default_rng(1).uniform(-1, 1, d["M"].shape) * 3).astype(np.float32))
    net.set_inputs(d["idx_q"], d["idx_a"])
    net.Forward()
    net.feat.diff.fill_(1.0 / N)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    net.capture()
    evs = []
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net.replay(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / iters
    handles = [l.handle for l in net.layers()]
    for h in handles:
        h.profile_enable(True)
    for _ in range(iters):
        for _ in range(4):
            flush.fill_(1)
        net.ClearParamDiffs(); net.ForwardBackward(); torch.cuda.synchronize()
    prof = {}
    for l in net.layers():
        for k, (n, t) in l.handle.profile_report().items():
            key = "%s:%s" % (l.layer_param_.name, k)
            prof[key] = round(prof.get(key, 0.0) + t / iters, 5)
        l.handle.profile_enable(False)
    conv_flops = 3 * 2.0 * N * (32 * 36 * 36 * 4 * 25 + 64 * 5 * 5 * 32 * 25)
    return {"workload": "network_v4 up to Flatten, %d QA pairs/step: Embed x2 -> SimCross(mode 2, mc 4) -> Dropout -> "
                        "Conv5x5(32)+BN -> AvePool4 -> TanH -> Conv5x5(64)+BN -> AvePool5 -> TanH, fwd+bwd+ClearParamDiffs, one CUDA graph" % N,
            "ms_per_step": ms, "qa_pairs_per_sec": N / (ms / 1e3), "kernels_ms_per_step": dict(sorted(prof.items(), key=lambda kv: -kv[1])),
            "kernel_ms_sum": sum(prof.values()), "conv_algorithmic_gflop_per_step": conv_flops / 1e9}


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    print(json.dumps(run(N, iters), indent=1))
