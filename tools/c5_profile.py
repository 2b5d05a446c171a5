#!/usr/bin/env python
"""Per-kernel device times of one SimMatrix modality of C5 (N=16384, 1024x1024), forward + backward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mms_answer_selection_b200 import _lib, layers, synth
from mms_answer_selection_b200.blob import Blob
N5, K1, K2 = 16384, 1024, 1024
qv, av, Wv = synth.make_sentence_vectors(N5, K1, K2, seed=1)
lay = layers.SimMatrixLayer(layers.LayerParameter("SimMatrix"))
bq, ba, top = Blob((N5, K1)), Blob((N5, K2)), Blob(())
bq.set_cpu_data(qv); ba.set_cpu_data(av)
lay.SetUp([bq, ba], [top]); lay.handle.set_option(_lib.MMS_OPT_REUSE_FORWARD, 1)
lay.blobs[0].set_cpu_data(Wv); top.diff.fill_(1.0 / N5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def step():
    lay.Forward([bq, ba], [top]); lay.Backward([top], [True, True], [bq, ba])
step(); step(); torch.cuda.synchronize()
lay.handle.profile_enable(True)
for _ in range(5):
    flush.fill_(1); flush.fill_(1); step(); torch.cuda.synchronize()
tot = 0
for k, (n, ms) in lay.handle.profile_report().items():
    print("%-32s x%d  %.4f ms/step" % (k, n // 5, ms / 5)); tot += ms / 5
print("sum %.4f ms; 6*N*K1*K2 = %.1f GFLOP -> %.0f TFLOP/s" % (tot, 6.0 * N5 * K1 * K2 / 1e9, 6.0 * N5 * K1 * K2 / tot / 1e9))
