"""Generate tests/golden/mms_golden.npz from the REFERENCE ITSELF.

Runs the reference's own layer code (oracle/_ref/libmms_ref.so, built in place from
/root/reference by oracle/Makefile) on small seeded inputs and stores inputs and
outputs.  Needs /root/reference (this container only); the fixture it writes is
committed so that the GPU box -- which has no /root/reference -- can check both the
C oracle and the CUDA path against genuine reference outputs.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.refbind import RefLayer, build  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mms_golden.npz")


def gen(dtype, tag, store):
    rng = np.random.default_rng(20161018 + (0 if dtype == np.float32 else 1))

    def put(name, arr):
        store["%s/%s" % (tag, name)] = np.asarray(arr)

    # ---- Embed: duplicates (the reference test's {4,2,2,3} pattern), bias, accumulation
    V, D = 7, 5
    idx = np.array([[4, 2, 2, 3], [6, 0, 2, 6], [1, 1, 1, 5]], dtype=dtype)
    for bias_term in (False, True):
        lay = RefLayer("Embed", [idx], {"num_output": D, "input_dim": V, "embed.bias_term": bias_term,
                                        "weight_filler.type": "uniform", "weight_filler.min": -10.0,
                                        "weight_filler.max": 10.0, "bias_filler.type": "uniform",
                                        "bias_filler.min": -10.0, "bias_filler.max": 10.0}, dtype=dtype)
        W = lay.read("blob", 0)
        lay.forward()
        top = lay.read("top", 0)
        dtop = rng.uniform(-1, 1, top.shape).astype(dtype)
        dW0 = rng.uniform(-1, 1, W.shape).astype(dtype)  # pre-existing diff: backward accumulates
        lay.write("top", 0, dtop, diff=True)
        lay.write("blob", 0, dW0, diff=True)
        if bias_term:
            b = lay.read("blob", 1)
            db0 = rng.uniform(-1, 1, b.shape).astype(dtype)
            lay.write("blob", 1, db0, diff=True)
        lay.backward([False])
        k = "embed_b%d" % int(bias_term)
        put(k + "/idx", idx); put(k + "/W", W); put(k + "/top", top); put(k + "/dtop", dtop)
        put(k + "/dW0", dW0); put(k + "/dW", lay.read("blob", 0, diff=True))
        if bias_term:
            put(k + "/b", b); put(k + "/db0", db0); put(k + "/db", lay.read("blob", 1, diff=True))

    # ---- SimCross modes 0, 1, 2 (ragged Lq != La, odd D)
    N, Lq, La, D, mc = 3, 5, 7, 6, 2
    q = rng.uniform(-0.5, 0.5, (N, Lq, D)).astype(dtype)
    a = rng.uniform(-0.5, 0.5, (N, La, D)).astype(dtype)
    for mode in (0, 1, 2):
        params = {"dist_mode": mode, "mesure_count": mc, "sim_cross.bias_term": True,
                  "weight_filler.type": "uniform", "weight_filler.min": -0.3, "weight_filler.max": 0.3,
                  "bias_filler.type": "uniform", "bias_filler.min": -0.2, "bias_filler.max": 0.2}
        lay = RefLayer("SimCross", [q, a], params, dtype=dtype)
        lay.forward()
        S = lay.read("top", 0)
        dS = rng.uniform(-1, 1, S.shape).astype(dtype)
        lay.write("top", 0, dS, diff=True)
        k = "simcross_m%d" % mode
        put(k + "/q", q); put(k + "/a", a); put(k + "/S", S); put(k + "/dS", dS)
        if mode == 2:
            Mw, B = lay.read("blob", 0), lay.read("blob", 1)
            dM0 = rng.uniform(-1, 1, Mw.shape).astype(dtype)   # must be discarded (layer zeroes dM)
            dB0 = rng.uniform(-1, 1, B.shape).astype(dtype)    # must be kept (layer accumulates dB)
            lay.write("blob", 0, dM0, diff=True)
            lay.write("blob", 1, dB0, diff=True)
            put(k + "/M", Mw); put(k + "/B", B); put(k + "/dM0", dM0); put(k + "/dB0", dB0)
        lay.backward([True, True])
        put(k + "/dq", lay.read("bottom", 0, diff=True))
        put(k + "/da", lay.read("bottom", 1, diff=True))
        if mode == 2:
            put(k + "/dM", lay.read("blob", 0, diff=True))
            put(k + "/dB", lay.read("blob", 1, diff=True))

    # ---- SimMatrix
    N, K1, K2 = 4, 6, 5
    q = rng.uniform(-1, 1, (N, K1)).astype(dtype)
    a = rng.uniform(-1, 1, (N, K2)).astype(dtype)
    lay = RefLayer("SimMatrix", [q, a], {"weight_filler.type": "uniform", "weight_filler.min": -0.5,
                                         "weight_filler.max": 0.5}, dtype=dtype)
    W = lay.read("blob", 0)
    lay.forward()
    s = lay.read("top", 0)
    T = lay.read("bottom", 1, diff=True)   # the reference's forward scratch (sim_matrix_layer.cpp:58)
    ds = rng.uniform(-1, 1, s.shape).astype(dtype)
    dW0 = rng.uniform(-1, 1, W.shape).astype(dtype)
    lay.write("top", 0, ds, diff=True)
    lay.write("blob", 0, dW0, diff=True)
    lay.backward([True, True])
    k = "simmatrix"
    put(k + "/q", q); put(k + "/a", a); put(k + "/W", W); put(k + "/s", s); put(k + "/T", T)
    put(k + "/ds", ds); put(k + "/dW0", dW0); put(k + "/dW", lay.read("blob", 0, diff=True))
    put(k + "/dq", lay.read("bottom", 0, diff=True)); put(k + "/da", lay.read("bottom", 1, diff=True))

    # ---- PairRankLoss: labels 1 / 0 / -1, one element exactly on the hinge (ordered == 0),
    #      one exactly similar (a == b with y = 0)
    sa = np.array([2.0, 0.5, -1.0, 0.25, 3.0, 3.0, -0.5, 1.0, 0.0, 1.5, -2.0, 0.75], dtype=dtype).reshape(12, 1)
    sb = np.array([1.0, 1.5, -1.0, 0.75, 1.0, 3.0, 0.5, 0.5, 0.0, -1.5, -1.0, 0.75], dtype=dtype).reshape(12, 1)
    yy = np.array([1, 1, 0, 0, 1, 0, -1, -1, 1, 0, 1, 0], dtype=dtype).reshape(12, 1)
    for margin in (1.0, 0.5):
        lay = RefLayer("PairRankLoss", [sa, sb, yy], {"margin": margin, "loss_weight": 1.0}, dtype=dtype)
        loss = lay.forward()
        lay.write("top", 0, np.array([2.0], dtype=dtype), diff=True)  # loss weight 2
        lay.backward([True, True, False])
        k = "pairrank_m%g" % margin
        put(k + "/a", sa); put(k + "/b", sb); put(k + "/y", yy)
        put(k + "/loss", np.array([lay.read("top", 0).reshape(-1)[0]], dtype=dtype))
        put(k + "/da", lay.read("bottom", 0, diff=True)); put(k + "/db", lay.read("bottom", 1, diff=True))

    # ---- FM (4 modalities)
    N, C, Dm = 3, 4, 5
    x = rng.uniform(-1, 1, (N, C, Dm)).astype(dtype)
    for bias_term in (False, True):
        lay = RefLayer("FM", [x], {"fm.bias_term": bias_term}, dtype=dtype)
        if bias_term:
            lay.write("blob", 0, np.array([0.375], dtype=dtype))
        lay.forward()
        y = lay.read("top", 0)
        dy = rng.uniform(-1, 1, y.shape).astype(dtype)
        lay.write("top", 0, dy, diff=True)
        if bias_term:
            lay.write("blob", 0, np.array([5.0], dtype=dtype), diff=True)  # must be overwritten
        lay.backward([True])
        k = "fm_b%d" % int(bias_term)
        put(k + "/x", x); put(k + "/y", y); put(k + "/dy", dy); put(k + "/dx", lay.read("bottom", 0, diff=True))
        if bias_term:
            put(k + "/db", lay.read("blob", 0, diff=True))


def main():
    build(ref=True, oracle=False)
    store = {}
    gen(np.float32, "f32", store)
    gen(np.float64, "f64", store)
    np.savez_compressed(OUT, **store)
    print("wrote %s: %d arrays, %d bytes" % (OUT, len(store), os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
