#!/usr/bin/env python
"""Debug helper: SimCross mode 2 backward at (N, L, D) against a torch fp64 contraction; prints where dM differs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mms_answer_selection_b200 as mms
N, L, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mc = 4
gen = torch.Generator(device="cuda").manual_seed(22)
lay = mms.SimCrossLayer(mms.LayerParameter("SimCross", sim_cross_param=dict(dist_mode=2, mesure_count=mc)))
bq, ba, top = mms.Blob((N, L, D)), mms.Blob((N, L, D)), mms.Blob(())
lay.SetUp([bq, ba], [top])
bq.data.copy_((torch.rand((N, L, D), device="cuda", generator=gen) - 0.5) * 0.16)
ba.data.copy_((torch.rand((N, L, D), device="cuda", generator=gen) - 0.5) * 0.16)
lay.blobs[0].data.copy_((torch.rand((mc, D, D), device="cuda", generator=gen) - 0.5) * 0.2)
lay.Forward([bq, ba], [top])
top.diff.copy_(torch.rand(top.shape, device="cuda", generator=gen) - 0.5)
lay.Backward([top], [True, True], [bq, ba])
torch.cuda.synchronize()
q, a, G, M = bq.data.double(), ba.data.double(), top.diff.double(), lay.blobs[0].data.double()
U = torch.einsum("nkij,njd->nkid", G, a)
dM = torch.einsum("nid,nkie->kde", q, U)
dq = torch.einsum("nkie,kde->nid", U, M)
got = lay.blobs[0].diff.double()
scale = dM.abs().max()
err = (got - dM).abs() / scale
print("dM max scaled err %.3e   dq err %.3e" % (err.max().item(), ((bq.diff.double() - dq).abs().max() / dq.abs().max()).item()))
for k in range(mc):
    e = err[k]
    rb = [e[i:i + 32].max().item() for i in range(0, D, 32)]
    cb = [e[:, i:i + 32].max().item() for i in range(0, D, 32)]
    print("k=%d rows/32: %s" % (k, " ".join("%.0e" % x for x in rb)))
    print("     cols/32: %s" % " ".join("%.0e" % x for x in cb))
print("ratio got/ref sample:", (got[0, :4, :4] / dM[0, :4, :4]))
