#!/usr/bin/env python
"""Generates tests/golden/conv2d_golden.npz from the REFERENCE's own ConvolutionLayer compiled in place
(oracle/_ref/libmms_ref.so: conv_layer.cpp + base_conv_layer.cpp + im2col.cpp), for the two geometries of the CNN over
the similarity tensor (examples/trec_qa_w2v_mms/do_trec_qa_clean.py:470-477) at small batch sizes: forward, and
backward onto parameter diffs that already hold 0.25 (they accumulate).  Run where /root/reference exists:
    python tests/golden/make_conv2d_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402

CASES = {"conv0": (1, 4, 40, 40, 32, 5, 5), "conv1": (3, 32, 9, 9, 64, 5, 5), "odd": (2, 3, 11, 7, 5, 3, 2)}


def main():
    refbind.build(ref=True, oracle=False, dropin=False)
    out = {}
    rng = np.random.default_rng(22)
    for name, (N, C, H, W, Co, kh, kw) in CASES.items():
        for dtype in ((np.float32, np.float64) if name == "odd" else (np.float32,)):
            k = "%s/%s/" % (name, np.dtype(dtype).name)
            x = rng.uniform(-1, 1, (N, C, H, W)).astype(dtype)
            conv = refbind.RefLayer("Convolution", [x], {"conv.num_output": Co, "conv.kernel_h": kh, "conv.kernel_w": kw,
                                                         "weight_filler.type": "xavier", "bias_filler.type": "constant",
                                                         "bias_filler.value": 0.1}, dtype=dtype)
            conv.forward()
            y = conv.read("top", 0)
            dy = rng.uniform(-1, 1, y.shape).astype(dtype)
            conv.write("top", 0, dy, diff=True)
            for i in range(2):
                conv.write("blob", i, np.full(conv.shape("blob", i), 0.25), diff=True)
            conv.backward([True])
            out.update({k + "x": x, k + "W": conv.read("blob", 0), k + "b": conv.read("blob", 1), k + "top": y, k + "dtop": dy,
                        k + "dW": conv.read("blob", 0, diff=True), k + "db": conv.read("blob", 1, diff=True),
                        k + "dx": conv.read("bottom", 0, diff=True)})
    path = os.path.join(ROOT, "tests", "golden", "conv2d_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
