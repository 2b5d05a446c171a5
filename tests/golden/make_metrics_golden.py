#!/usr/bin/env python
"""Generates tests/golden/metrics_golden.npz from the REFERENCE's own MAP / MRR / AUC / RankAccuracy layers compiled in
place (oracle/_ref, `make -C oracle ref`; needs /root/reference).  Run here, commit the .npz; the GPU box only reads it.

    python tests/golden/make_metrics_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402


def run(typ, bottoms, params, dtype):
    lay = refbind.RefLayer(typ, bottoms, params, dtype=dtype)
    lay.forward()
    return lay.read("top", 0).reshape(-1)[0]


def main():
    rng = np.random.default_rng(2024)
    out = {}
    cases = [("trec", 1517, 68, 0.17), ("small", 37, 5, 0.4), ("one_group", 300, 1, 0.1), ("no_pos_groups", 400, 90, 0.02),
             ("three_class", 500, 25, 0.3)]
    for name, n, ngroups, ppos in cases:
        for dtype, tag in ((np.float32, "f32"), (np.float64, "f64")):
            C = 3 if name == "three_class" else 2
            fa = C - 1
            # distinct scores (ties have unspecified order in the reference): a random permutation of a grid
            prob = np.zeros((n, C), dtype)
            prob[:, fa] = (rng.permutation(n) + 0.5) / n
            prob[:, 0] = 1 - prob[:, fa]
            label = (rng.uniform(0, 1, n) < ppos).astype(dtype)
            if name == "three_class":
                label[rng.uniform(0, 1, n) < 0.1] = 2          # MAP: a negative; MRR: neither positive nor negative
            group = rng.integers(-2, ngroups - 2, n).astype(dtype)     # unsorted, negative ids included
            key = "%s_%s/" % (name, tag)
            out[key + "prob"], out[key + "label"], out[key + "group"] = prob, label, group
            out[key + "map"] = run("MAP", [prob, label, group], {"map.fixed_axis": fa}, dtype)
            out[key + "mrr"] = run("MRR", [prob, label, group], {"mrr.fixed_axis": fa}, dtype)
            if C == 2:
                out[key + "auc"] = run("AUC", [prob, label], {"auc.fixed_axis": fa}, dtype)
                out[key + "auc_ignore"] = run("AUC", [prob, label], {"auc.fixed_axis": fa, "auc.ignore_label": 0}, dtype)
            a, b = prob[:, fa].copy(), rng.uniform(0, 1, n).astype(dtype)
            y = np.where(rng.uniform(0, 1, n) < 0.5, 1.0, -1.0).astype(dtype)
            out[key + "ra_b"], out[key + "ra_y"] = b, y
            out[key + "rank_accuracy"] = run("RankAccuracy", [a, b, y], {}, dtype)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
