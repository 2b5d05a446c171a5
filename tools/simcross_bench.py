#!/usr/bin/env python
"""Times SimCross (mode 2) forward and backward alone, eagerly, with CUDA events, and prints the
per-kernel device times the library's own profiler records.

    python tools/simcross_bench.py [c2|c3|N] [iters]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import mms_answer_selection_b200 as mms

arg = sys.argv[1] if len(sys.argv) > 1 else "c2"
N = {"c2": 50, "c3": 4096}.get(arg) or int(arg)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
L, D, mc = 40, 300, 4
gen = torch.Generator(device="cuda").manual_seed(22)
lay = mms.SimCrossLayer(mms.LayerParameter("SimCross", sim_cross_param=dict(dist_mode=2, mesure_count=mc)))
bq, ba, top = mms.Blob((N, L, D)), mms.Blob((N, L, D)), mms.Blob(())
lay.SetUp([bq, ba], [top])
bq.data.copy_((torch.rand((N, L, D), device="cuda", generator=gen) - 0.5) * 0.16)
ba.data.copy_((torch.rand((N, L, D), device="cuda", generator=gen) - 0.5) * 0.16)
lay.blobs[0].data.copy_((torch.rand((mc, D, D), device="cuda", generator=gen) - 0.5) * 0.2)
lay.Forward([bq, ba], [top])
top.diff.copy_(torch.rand(top.shape, device="cuda", generator=gen) - 0.5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


fwd = timed(lambda: lay.Forward([bq, ba], [top]))
bwd = timed(lambda: lay.Backward([top], [True, True], [bq, ba]))
f_fwd = N * mc * 2.0 * L * D * (D + L)
f_all = N * mc * (6.0 * L * D * D + 8.0 * L * L * D)
print("N=%d fused=%s  fwd %.4f ms (%.1f TF/s alg)  bwd %.4f ms (%.1f TF/s alg)  fwd+bwd %.1f TF/s alg" % (
    N, "off (needs a -DMMS_DEV_KNOBS build)" if os.environ.get("MMS_NO_FUSED") else "on", fwd, f_fwd / fwd * 1e-9, bwd,
    (f_all - f_fwd) / bwd * 1e-9, f_all / (fwd + bwd) * 1e-9))
h = lay.handle
h.profile_enable(True)
for _ in range(5):
    lay.Forward([bq, ba], [top]); lay.Backward([top], [True, True], [bq, ba])
torch.cuda.synchronize()
print({k: (n, round(ms / 5, 4)) for k, (n, ms) in h.profile_report().items()})
