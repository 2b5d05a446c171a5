#!/usr/bin/env python
"""Generates tests/golden/sentenc_golden.npz from the REFERENCE's own Convolution / BN / Pooling / TanH layers compiled
in place (oracle/_ref, `make -C oracle ref`; needs /root/reference).  Run here, commit the .npz; the GPU box only reads it.

    python tests/golden/make_sentenc_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402

CONV = {"conv.num_output": 10, "conv.kernel_h": 5, "conv.kernel_w": 16, "weight_filler.type": "xavier",
        "bias_filler.type": "uniform", "bias_filler.min": -0.1, "bias_filler.max": 0.1}
BN = {"scale_filler.type": "uniform", "scale_filler.min": 0.5, "scale_filler.max": 1.5,
      "shift_filler.type": "uniform", "shift_filler.min": -0.2, "shift_filler.max": 0.2, "bn_memory": 0.9}


def main():
    out = {}
    for dtype, tag in ((np.float32, "f32"), (np.float64, "f64")):
        rng = np.random.default_rng(77)
        k = tag + "/"
        N, L, D = 6, 12, 16
        x = rng.uniform(-1, 1, (N, 1, L, D)).astype(dtype)
        # ---- Convolution (kernel 5 x D): forward, then backward onto param diffs that already hold 0.5 (they accumulate)
        conv = refbind.RefLayer("Convolution", [x], CONV, dtype=dtype)
        conv.forward()
        y = conv.read("top", 0)
        dy = rng.uniform(-1, 1, y.shape).astype(dtype)
        conv.write("top", 0, dy, diff=True)
        for i in range(2):
            conv.write("blob", i, np.full(conv.shape("blob", i), 0.5), diff=True)
        conv.backward([True])
        out.update({k + "conv/x": x, k + "conv/W": conv.read("blob", 0), k + "conv/b": conv.read("blob", 1), k + "conv/top": y,
                    k + "conv/dtop": dy, k + "conv/dW": conv.read("blob", 0, diff=True),
                    k + "conv/db": conv.read("blob", 1, diff=True), k + "conv/dx": conv.read("bottom", 0, diff=True)})
        # ---- BN, TRAIN phase: two forwards (the running statistics blend twice), then backward; then a TEST-phase forward
        bn = refbind.RefLayer("BN", [y], BN, dtype=dtype)
        bn.forward()
        y2 = (y * 1.5 + 0.25).astype(dtype)
        bn.write("bottom", 0, y2)
        bn.forward()
        z = bn.read("top", 0)
        dz = rng.uniform(-1, 1, z.shape).astype(dtype)
        bn.write("top", 0, dz, diff=True)
        bn.backward([True])
        out.update({k + "bn/x0": y, k + "bn/x1": y2, k + "bn/scale": bn.read("blob", 0), k + "bn/shift": bn.read("blob", 1),
                    k + "bn/run_mean": bn.read("blob", 2), k + "bn/run_var": bn.read("blob", 3), k + "bn/top": z,
                    k + "bn/dtop": dz, k + "bn/dscale": bn.read("blob", 0, diff=True),
                    k + "bn/dshift": bn.read("blob", 1, diff=True), k + "bn/dx": bn.read("bottom", 0, diff=True)})
        bnt = refbind.RefLayer("BN", [y], dict(BN, phase=1), dtype=dtype)
        for i in range(4):
            bnt.write("blob", i, bn.read("blob", i))
        bnt.forward()
        out[k + "bn/top_test"] = bnt.read("top", 0)
        # ---- Pooling: MAX over time (kernel (L-4) x 1) with ties, and a padded / strided 2-D AVE and MAX
        zq = np.round(z * 4) / 4                                  # quantised: equal maxima, the first one must win
        for name, src, params in (
                ("pool_time", zq.astype(dtype), {"pool.method": 0, "pool.kernel_h": L - 4, "pool.kernel_w": 1}),
                ("pool_ave2d", rng.uniform(-1, 1, (3, 4, 9, 7)).astype(dtype),
                 {"pool.method": 1, "pool.kernel_h": 4, "pool.kernel_w": 3, "pool.stride_h": 2, "pool.stride_w": 2}),
                ("pool_max2d", np.round(rng.uniform(-1, 1, (3, 4, 9, 7)) * 3).astype(dtype) / 3,
                 {"pool.method": 0, "pool.kernel_h": 3, "pool.kernel_w": 3, "pool.stride_h": 2, "pool.stride_w": 1})):
            pool = refbind.RefLayer("Pooling", [src], params, dtype=dtype)
            pool.forward()
            p = pool.read("top", 0)
            dp = rng.uniform(-1, 1, p.shape).astype(dtype)
            pool.write("top", 0, dp, diff=True)
            pool.backward([True])
            out.update({k + name + "/x": src, k + name + "/top": p, k + name + "/dtop": dp,
                        k + name + "/dx": pool.read("bottom", 0, diff=True)})
        # ---- TanH
        th = refbind.RefLayer("TanH", [out[k + "pool_time/top"]], {}, dtype=dtype)
        th.forward()
        t = th.read("top", 0)
        dt = rng.uniform(-1, 1, t.shape).astype(dtype)
        th.write("top", 0, dt, diff=True)
        th.backward([True])
        out.update({k + "tanh/top": t, k + "tanh/dtop": dt, k + "tanh/dx": th.read("bottom", 0, diff=True)})
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sentenc_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
