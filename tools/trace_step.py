#!/usr/bin/env python
"""Runs one eager C2 step with MMS_TC_TRACE=1 so that every TMA GEMM prints its per-CTA timeline."""
import os, sys
if "notrace" not in sys.argv: os.environ["MMS_TC_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mms_answer_selection_b200 as mms
from mms_answer_selection_b200 import synth
cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
N, L, D, mc, V = cfg["N"], cfg["L"], cfg["D"], cfg["mc"], cfg["V"]
d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
net = mms.MMSNet(N, L, D, mc, V, keep_embed_tops=False)      # the configuration bench.py times
net.set_params(d["W"], d["b"], d["M"], d["B"]); net.set_inputs(d["idx_q"], d["idx_a"]); net.set_upstream_gradient(d["dS"])
for i in range(3):
    sys.stderr.write("---- step %d\n" % i)
    net.ClearParamDiffs(); net.ForwardBackward()
torch.cuda.synchronize()
