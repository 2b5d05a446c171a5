"""CPU tests: pin the plain-C oracle (oracle/mms_oracle.c).

Three anchors, per the parity contract:
 1. the reference's only known-answer tests for this path,
    src/caffe/test/test_embed_layer.cpp:54-176, restated;
 2. the committed golden fixtures tests/golden/mms_golden.npz (outputs of the
    reference's own layer code, see tests/golden/make_golden.py);
 3. the reference itself (oracle/_ref) on fresh random inputs, when it is present.
plus GradientChecker-style finite differences (test_gradient_check_util.hpp:148-175).
"""
import os

import numpy as np
import pytest

from conftest import scaled_err
from oracle import cport, refbind

DTYPES = [np.float32, np.float64]
TAG = {np.float32: "f32", np.float64: "f64"}
# contractions go through BLAS in the reference: summation order unspecified
TOL = {np.float32: 2e-5, np.float64: 1e-12}


def g(golden, dtype, key):
    return golden["%s/%s" % (TAG[dtype], key)]


# ---------------------------------------------------------------- 1. reference KATs
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bias", [False, True])
def test_embed_forward_kat(dtype, bias):
    # test_embed_layer.cpp:54-135: bottom (4,1,1,1) random ids in [0,5), W 5x10 ~ U(-10,10)
    rng = np.random.default_rng(1701)
    W = rng.uniform(-10, 10, (5, 10)).astype(dtype)
    b = rng.uniform(-10, 10, 10).astype(dtype) if bias else None
    idx = rng.integers(0, 5, size=(4, 1, 1, 1)).astype(dtype)
    top = cport.embed_forward(idx, W, b)
    assert top.shape == (4, 1, 1, 1, 10)              # TestSetUp :38-52
    for i in range(4):
        want = W[int(idx.reshape(-1)[i])] + (b if bias else 0)
        assert np.array_equal(top.reshape(4, 10)[i], want.astype(dtype))   # EXPECT_EQ :84,125


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bias", [False, True])
def test_embed_gradient_kat(dtype, bias):
    # test_embed_layer.cpp:137-176: ids {4,2,2,3}; duplicate id 2 exercises accumulation.
    # Embed is linear in W and b, so central differences are exact up to rounding:
    # objective = sum(top * R)  =>  dW[v] = sum_{n: idx[n]==v} R[n], db = sum_n R[n].
    rng = np.random.default_rng(7)
    idx = np.array([4, 2, 2, 3], dtype=dtype).reshape(4, 1, 1, 1)
    R = rng.uniform(-1, 1, (4, 10)).astype(dtype)
    dW = np.zeros((5, 10), dtype=dtype)
    db = np.zeros(10, dtype=dtype) if bias else None
    cport.embed_backward(idx, R, dW, db)
    want = np.zeros((5, 10), dtype=np.float64)
    for n, v in enumerate([4, 2, 2, 3]):
        want[v] += R[n]
    assert scaled_err(dW, want, 1.0) <= 1e-3           # GradientChecker(1e-2, 1e-3) :147
    if bias:
        assert scaled_err(db, R.astype(np.float64).sum(0), 1.0) <= 1e-3


# ---------------------------------------------------------------- 2. golden fixtures
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bias", [False, True])
def test_embed_golden(golden, dtype, bias):
    k = "embed_b%d" % int(bias)
    idx, W = g(golden, dtype, k + "/idx"), g(golden, dtype, k + "/W")
    b = g(golden, dtype, k + "/b") if bias else None
    assert np.array_equal(cport.embed_forward(idx, W, b), g(golden, dtype, k + "/top"))
    dW = g(golden, dtype, k + "/dW0").copy()
    db = g(golden, dtype, k + "/db0").copy() if bias else None
    cport.embed_backward(idx, g(golden, dtype, k + "/dtop"), dW, db)
    assert np.array_equal(dW, g(golden, dtype, k + "/dW"))       # same axpy order => bit-exact
    if bias:
        assert scaled_err(db, g(golden, dtype, k + "/db")) <= TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_simcross_golden(golden, dtype, mode):
    k = "simcross_m%d" % mode
    q, a = g(golden, dtype, k + "/q"), g(golden, dtype, k + "/a")
    Mw = g(golden, dtype, k + "/M") if mode == 2 else None
    B = g(golden, dtype, k + "/B") if mode == 2 else None
    S, n0, n1 = cport.simcross_forward(mode, q, a, Mw, B)
    assert S.shape == g(golden, dtype, k + "/S").shape
    assert scaled_err(S, g(golden, dtype, k + "/S")) <= TOL[dtype]
    dB = g(golden, dtype, k + "/dB0").copy() if mode == 2 else None
    dq, da, dM, dB = cport.simcross_backward(mode, q, a, Mw if mode == 2 else q, S,
                                             g(golden, dtype, k + "/dS"), n0, n1, dB)
    assert scaled_err(dq, g(golden, dtype, k + "/dq")) <= 10 * TOL[dtype]
    assert scaled_err(da, g(golden, dtype, k + "/da")) <= 10 * TOL[dtype]
    if mode == 2:
        # dM is zeroed by the layer (dM0 discarded), dB accumulates on top of dB0
        assert scaled_err(dM, g(golden, dtype, k + "/dM")) <= 10 * TOL[dtype]
        assert scaled_err(dB, g(golden, dtype, k + "/dB")) <= TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_simmatrix_golden(golden, dtype):
    k = "simmatrix"
    q, a, W = (g(golden, dtype, k + "/" + n) for n in ("q", "a", "W"))
    s, T = cport.simmatrix_forward(q, a, W)
    assert scaled_err(s, g(golden, dtype, k + "/s")) <= TOL[dtype]
    assert scaled_err(T, g(golden, dtype, k + "/T")) <= TOL[dtype]     # bottom[1].diff scratch
    dW = g(golden, dtype, k + "/dW0").copy()
    dW, dq, da = cport.simmatrix_backward(q, a, W, g(golden, dtype, k + "/ds"), dW)
    assert scaled_err(dW, g(golden, dtype, k + "/dW")) <= TOL[dtype]
    assert scaled_err(dq, g(golden, dtype, k + "/dq")) <= TOL[dtype]
    assert scaled_err(da, g(golden, dtype, k + "/da")) <= TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("margin", [1.0, 0.5])
def test_pairrankloss_golden(golden, dtype, margin):
    k = "pairrank_m%g" % margin
    a, b, y = (g(golden, dtype, k + "/" + n) for n in ("a", "b", "y"))
    loss, ordered, similar = cport.pairrankloss_forward(a, b, y, margin)
    assert loss == g(golden, dtype, k + "/loss")[0]                    # explicit loops: bit-exact
    da, db = cport.pairrankloss_backward(y, ordered, similar, top_diff=2.0, ge=False)
    assert np.array_equal(da, g(golden, dtype, k + "/da"))
    assert np.array_equal(db, g(golden, dtype, k + "/db"))
    # the reference's GPU kernel differs only where ordered == 0 (>= instead of >)
    da_ge, _ = cport.pairrankloss_backward(y, ordered, similar, top_diff=2.0, ge=True)
    diff = (da_ge != da).reshape(-1)
    assert np.array_equal(diff, (ordered.reshape(-1) == 0) & (y.reshape(-1) != 0))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bias", [False, True])
def test_fm_golden(golden, dtype, bias):
    k = "fm_b%d" % int(bias)
    x = g(golden, dtype, k + "/x")
    b = np.array([0.375], dtype=dtype) if bias else None
    assert np.array_equal(cport.fm_forward(x, b), g(golden, dtype, k + "/y"))
    dx, db = cport.fm_backward(x, g(golden, dtype, k + "/dy"), bias_term=bias)
    assert np.array_equal(dx, g(golden, dtype, k + "/dx"))
    if bias:
        assert np.array_equal(db, g(golden, dtype, k + "/db"))


# ---------------------------------------------------------------- 3. the reference itself
needs_ref = pytest.mark.skipif(not refbind.ref_available(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 3, 4, 5, 1), (4, 40, 40, 50, 4), (3, 7, 11, 13, 3)])
def test_simcross_mode2_vs_reference(dtype, shape):
    N, Lq, La, D, mc = shape
    rng = np.random.default_rng(sum(shape))
    q = rng.uniform(-0.08, 0.08, (N, Lq, D)).astype(dtype)
    a = rng.uniform(-0.08, 0.08, (N, La, D)).astype(dtype)
    lay = refbind.RefLayer("SimCross", [q, a], {"dist_mode": 2, "mesure_count": mc,
                           "weight_filler.type": "uniform", "weight_filler.min": -0.1,
                           "weight_filler.max": 0.1}, dtype=dtype)
    Mw, B = lay.read("blob", 0), lay.read("blob", 1)
    assert Mw.shape == (mc, D, D) and B.shape == (mc, Lq, La)
    lay.forward()
    S_ref = lay.read("top", 0)
    S, _, _ = cport.simcross_forward(2, q, a, Mw, B)
    assert scaled_err(S, S_ref) <= TOL[dtype]
    dS = rng.uniform(-1, 1, S.shape).astype(dtype)
    lay.write("top", 0, dS, diff=True)
    lay.backward([True, True])
    dq, da, dM, dB = cport.simcross_backward(2, q, a, Mw, S, dS)
    assert scaled_err(dq, lay.read("bottom", 0, diff=True)) <= 10 * TOL[dtype]
    assert scaled_err(da, lay.read("bottom", 1, diff=True)) <= 10 * TOL[dtype]
    assert scaled_err(dM, lay.read("blob", 0, diff=True)) <= 10 * TOL[dtype]
    assert scaled_err(dB, lay.read("blob", 1, diff=True)) <= TOL[dtype]
    # second backward: dM is re-zeroed, dB keeps accumulating (sim_cross_layer.cpp:256,301-304)
    lay.backward([True, True])
    assert scaled_err(2 * dB, lay.read("blob", 1, diff=True)) <= TOL[dtype]
    assert scaled_err(dM, lay.read("blob", 0, diff=True)) <= 10 * TOL[dtype]


@needs_ref
def test_simcross_default_filler_is_zero():
    # caffe.proto:43-47: default filler constant 0 => iteration-0 scores equal B (= 0)
    rng = np.random.default_rng(3)
    q = rng.uniform(-1, 1, (2, 4, 6)).astype(np.float32)
    a = rng.uniform(-1, 1, (2, 5, 6)).astype(np.float32)
    lay = refbind.RefLayer("SimCross", [q, a], {"dist_mode": 2, "mesure_count": 3})
    lay.forward()
    assert not lay.read("top", 0).any()
    S, _, _ = cport.simcross_forward(2, q, a, lay.read("blob", 0), lay.read("blob", 1))
    assert not S.any()


@needs_ref
def test_reference_error_behaviour():
    e = refbind.RefLayer("Embed", [np.zeros((2, 2), np.float32)], {"num_output": 3, "input_dim": 4})
    e.forward()
    with pytest.raises(refbind.RefError, match="Can't backpropagate to EmbedLayer input"):
        e.backward([True])                                   # embed_layer.cpp:158
    z = np.zeros((3, 1), np.float32)
    p = refbind.RefLayer("PairRankLoss", [z, z, z], {"loss_weight": 1.0})
    p.forward()
    with pytest.raises(refbind.RefError, match="cannot backpropagate to label"):
        p.backward([True, True, True])                       # pair_rank_loss_layer.cpp:58-61
    with pytest.raises(refbind.RefError):
        refbind.RefLayer("SimCross", [np.zeros((2, 3, 4), np.float32), np.zeros((2, 3, 5), np.float32)],
                         {"dist_mode": 2})                   # height mismatch, sim_cross_layer.cpp:14


# ---------------------------------------------------------------- finite differences
def _fd(f, x, h):
    gnum = np.zeros_like(x, dtype=np.float64)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        old = x[i]
        x[i] = old + h; fp = f()
        x[i] = old - h; fm = f()
        x[i] = old
        gnum[i] = (fp - fm) / (2 * h)
    return gnum


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_simcross_finite_differences(mode):
    rng = np.random.default_rng(11 + mode)
    N, Lq, La, D, mc = 2, 3, 4, 5, 2
    q = rng.uniform(-0.5, 0.5, (N, Lq, D))
    a = rng.uniform(-0.5, 0.5, (N, La, D))
    Mw = rng.uniform(-0.3, 0.3, (mc, D, D))
    B = rng.uniform(-0.3, 0.3, (mc, Lq, La))
    R = rng.uniform(-1, 1, (N, mc if mode == 2 else 1, Lq, La))

    def obj():
        S, _, _ = cport.simcross_forward(mode, q, a, Mw, B)
        return float((S * R).sum())

    S, n0, n1 = cport.simcross_forward(mode, q, a, Mw, B)
    dq, da, dM, dB = cport.simcross_backward(mode, q, a, Mw, S, R, n0, n1)
    assert scaled_err(dq, _fd(obj, q, 1e-5), 1.0) < 1e-6
    assert scaled_err(da, _fd(obj, a, 1e-5), 1.0) < 1e-6
    if mode == 2:
        assert scaled_err(dM, _fd(obj, Mw, 1e-5), 1.0) < 1e-6
        assert scaled_err(dB, _fd(obj, B, 1e-5), 1.0) < 1e-6


def test_simmatrix_fm_finite_differences():
    rng = np.random.default_rng(5)
    q, a, W = rng.uniform(-1, 1, (3, 4)), rng.uniform(-1, 1, (3, 5)), rng.uniform(-1, 1, (4, 5))
    R = rng.uniform(-1, 1, (3, 1))
    obj = lambda: float((cport.simmatrix_forward(q, a, W)[0] * R).sum())
    dW, dq, da = cport.simmatrix_backward(q, a, W, R, np.zeros_like(W))
    assert scaled_err(dq, _fd(obj, q, 1e-5), 1.0) < 1e-7
    assert scaled_err(da, _fd(obj, a, 1e-5), 1.0) < 1e-7
    assert scaled_err(dW, _fd(obj, W, 1e-5), 1.0) < 1e-7
    x = rng.uniform(-1, 1, (3, 4, 5)); Ry = rng.uniform(-1, 1, (3, 1))
    objf = lambda: float((cport.fm_forward(x, np.array([0.1])) * Ry).sum())
    dx, db = cport.fm_backward(x, Ry)
    assert scaled_err(dx, _fd(objf, x, 1e-5), 1.0) < 1e-7
    assert abs(db[0] - Ry.sum()) < 1e-12


def test_pairrankloss_label_table():
    # y=1 ordered hinge, y=0 similar |a-b| + margin, y=-1 reversed (+2|a-b|)
    for y, a, b, want in [(1, 2.0, 0.0, 0.0), (1, 0.0, 2.0, 3.0), (0, 2.0, 0.5, 1.0 + 1.5),
                          (-1, 0.0, 2.0, 0.0 + 2 * 2.0), (-1, 2.0, 0.0, 3.0 + 4.0)]:
        loss, _, _ = cport.pairrankloss_forward(np.array([a]), np.array([b]), np.array([float(y)]), 1.0)
        assert loss == pytest.approx(want)


# ---------------------------------------------------------------- AdaDelta step
def _adadelta_numpy(w, g, h, h2, scale, decay, mom, delta, rate):
    """The update as the reference's GPU kernel states it (adadelta_solver.cu:7-16), in float64, with the passes
    SGDSolver::ApplyUpdate makes around it."""
    g = g * scale + decay * w
    h = mom * h + (1 - mom) * g * g
    g = g * np.sqrt((h2 + delta) / (h + delta))
    h2 = mom * h2 + (1 - mom) * g * g
    g = rate * g
    return w - g, g, h, h2


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_adadelta_oracle_matches_the_kernel_formula(dtype):
    """The oracle follows the CPU branch pass by pass (adadelta_solver.cpp:33-94); the reference's GPU branch is the
    closed form above -- the two must agree (the reference's own test_gradient_based_solver.cpp checks its solvers
    against such a re-derivation, not against golden vectors)."""
    rng = np.random.default_rng(3)
    n = 1000
    w = rng.uniform(-0.1, 0.1, n)
    hist = [np.zeros(n), np.zeros(n)]
    wo = w.astype(dtype); ho = [np.zeros(n, dtype), np.zeros(n, dtype)]
    for step in range(4):
        g = rng.normal(0, 1e-2, n)
        w, upd, hist[0], hist[1] = _adadelta_numpy(w, g, hist[0], hist[1], 0.5, 5e-4, 0.95, 5e-7, 1.0)
        go = g.astype(dtype)
        cport.adadelta_step(wo, go, ho[0], ho[1], grad_scale=0.5, local_decay=5e-4, momentum=0.95, delta=5e-7,
                            local_rate=1.0)
        tol = 1e-5 if dtype == np.float32 else 1e-12
        assert scaled_err(go, upd) <= tol and scaled_err(wo, w) <= tol
        assert scaled_err(ho[0], hist[0]) <= tol and scaled_err(ho[1], hist[1]) <= tol


def test_adadelta_oracle_known_answer():
    """First iteration from zero histories, g = 1, no decay: h = 0.05, update = sqrt(delta / (0.05 + delta)),
    h2 = 0.05 * update^2 (hand-computed)."""
    g = np.ones(3); w = np.zeros(3); h = np.zeros(3); h2 = np.zeros(3)
    cport.adadelta_step(w, g, h, h2, momentum=0.95, delta=5e-7, local_rate=2.0)
    u = np.sqrt(5e-7 / (0.05 + 5e-7))
    assert np.allclose(h, 0.05, rtol=1e-12) and np.allclose(h2, 0.05 * u * u, rtol=1e-12)
    assert np.allclose(g, 2.0 * u, rtol=1e-12) and np.allclose(w, -2.0 * u, rtol=1e-12)
    # data = None: the bare adadelta_update (no Net::Update, no decay)
    g2 = np.ones(3); h = np.zeros(3); h2 = np.zeros(3)
    cport.adadelta_step(None, g2, h, h2, local_decay=1.0, momentum=0.95, delta=5e-7, local_rate=2.0)
    assert np.allclose(g2, 2.0 * u, rtol=1e-12)


# ---------------------------------------------------------------- ranking metrics
METRIC_CASES = ["trec", "small", "one_group", "no_pos_groups", "three_class"]


def _metrics_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_golden.npz"))


@pytest.mark.parametrize("tag,dtype", [("f32", np.float32), ("f64", np.float64)])
@pytest.mark.parametrize("case", METRIC_CASES)
def test_metric_oracles_match_reference_fixtures(case, tag, dtype):
    """The plain-C restatements of MAP / MRR / AUC / RankAccuracy against fixtures produced by the reference's own
    layers (tests/golden/make_metrics_golden.py): bit-exact -- same order of operations, same accumulation type."""
    g = _metrics_golden()
    k = "%s_%s/" % (case, tag)
    prob, label, group = g[k + "prob"], g[k + "label"], g[k + "group"]
    fa = prob.shape[1] - 1
    m, r = cport.map_mrr(prob, label, group, fixed_axis=fa)
    assert m == g[k + "map"] or (np.isnan(m) and np.isnan(g[k + "map"]))
    assert r == g[k + "mrr"] or (np.isnan(r) and np.isnan(g[k + "mrr"]))
    if prob.shape[1] == 2:
        same = lambda x, y: x == y or (np.isnan(x) and np.isnan(y))
        assert same(cport.auc(prob, label, fixed_axis=fa), g[k + "auc"])
        # ignoring label 0 leaves no negative: 0/0, in the reference and here
        assert same(cport.auc(prob, label, fixed_axis=fa, ignore_label=0), g[k + "auc_ignore"])
    a = np.ascontiguousarray(prob[:, fa])
    assert cport.rank_accuracy(a, g[k + "ra_b"], g[k + "ra_y"]) == g[k + "rank_accuracy"]


@pytest.mark.skipif(not os.path.exists(refbind.REF_SO), reason="compiled reference not present")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_metric_oracles_match_compiled_reference(dtype):
    rng = np.random.default_rng(8)
    for _ in range(10):
        n = int(rng.integers(2, 600))
        prob = rng.uniform(0, 1, (n, 2)).astype(dtype)
        label = (rng.uniform(0, 1, n) < 0.3).astype(dtype)
        group = rng.integers(-3, 30, n).astype(dtype)
        m, r = cport.map_mrr(prob, label, group)
        for typ, params, bots, ours in (("MAP", {"map.fixed_axis": 1}, [prob, label, group], m),
                                        ("MRR", {"mrr.fixed_axis": 1}, [prob, label, group], r),
                                        ("AUC", {"auc.fixed_axis": 1}, [prob, label], cport.auc(prob, label))):
            lay = refbind.RefLayer(typ, bots, params, dtype=dtype)
            lay.forward()
            ref = lay.read("top", 0).reshape(-1)[0]
            assert ours == ref or (np.isnan(ours) and np.isnan(ref)), typ


# ---------------------------------------------------------------- sentence encoder (conv 5xD / BN / pooling / TanH)
def _sentenc_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sentenc_golden.npz"))


@pytest.mark.parametrize("tag,tol", [("f32", 2e-5), ("f64", 1e-12)])
def test_sentence_encoder_restatement_matches_reference_fixtures(tag, tol):
    """oracle/sentenc_np.py against fixtures produced by the reference's own Convolution / BN / Pooling / TanH layers
    (tests/golden/make_sentenc_golden.py).  Summation order differs from BLAS: 2e-5 / 1e-12 of the largest magnitude."""
    from oracle import sentenc_np as snp
    g = _sentenc_golden()
    k = tag + "/"
    err = lambda got, ref: float(np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))
    x, W, b = g[k + "conv/x"], g[k + "conv/W"], g[k + "conv/b"]
    assert err(snp.conv_forward(x, W, b), g[k + "conv/top"]) <= tol
    dW, db, dx = snp.conv_backward(x.astype(np.float64), W.astype(np.float64), g[k + "conv/dtop"].astype(np.float64))
    assert err(dW + 0.5, g[k + "conv/dW"]) <= tol and err(db + 0.5, g[k + "conv/db"]) <= tol      # param diffs accumulate
    assert err(dx, g[k + "conv/dx"]) <= tol
    sc, sh = g[k + "bn/scale"], g[k + "bn/shift"]
    f64 = lambda a: a.astype(np.float64)
    mem = float(np.float32(0.9))                              # BNParameter.bn_memory is a float field (caffe.proto:485)
    _, _, _, rm, rv = snp.bn_forward(f64(g[k + "bn/x0"]), f64(sc), f64(sh), np.zeros(sc.size), np.zeros(sc.size), memory=mem)
    top, xn, std, rm, rv = snp.bn_forward(f64(g[k + "bn/x1"]), f64(sc), f64(sh), rm, rv, memory=mem)
    assert err(top, g[k + "bn/top"]) <= 10 * tol
    assert err(rm, g[k + "bn/run_mean"].reshape(-1)) <= 10 * tol and err(rv, g[k + "bn/run_var"].reshape(-1)) <= 10 * tol
    dsc, dsh, dxb = snp.bn_backward(f64(g[k + "bn/dtop"]), xn, f64(sc), std)
    assert err(dsc, g[k + "bn/dscale"].reshape(-1)) <= 10 * tol and err(dsh, g[k + "bn/dshift"].reshape(-1)) <= 10 * tol
    assert err(dxb, g[k + "bn/dx"]) <= 20 * tol
    ttop, _, _, _, _ = snp.bn_forward(f64(g[k + "bn/x0"]), f64(sc), f64(sh), rm, rv, train=False)
    assert err(ttop, g[k + "bn/top_test"]) <= 10 * tol
    L = x.shape[2]
    for name, geom in (("pool_time", dict(kh=L - 4, kw=1, method="MAX")),
                       ("pool_ave2d", dict(kh=4, kw=3, sh=2, sw=2, method="AVE")),
                       ("pool_max2d", dict(kh=3, kw=3, sh=2, sw=1, method="MAX"))):
        src = g[k + name + "/x"]
        top, mask = snp.pool_forward(src, **geom)
        assert top.shape == g[k + name + "/top"].shape and err(top, g[k + name + "/top"]) <= tol
        assert err(snp.pool_backward(g[k + name + "/dtop"], mask, src.shape, **geom), g[k + name + "/dx"]) <= tol
    assert err(snp.tanh_forward(g[k + "pool_time/top"]), g[k + "tanh/top"]) <= tol
    assert err(snp.tanh_backward(g[k + "tanh/top"], g[k + "tanh/dtop"]), g[k + "tanh/dx"]) <= tol


def test_conv2d_numpy_restatement_matches_the_reference_fixtures():
    """oracle/conv2d_np.py against tests/golden/conv2d_golden.npz (the reference's ConvolutionLayer compiled in place)."""
    import os
    from oracle import conv2d_np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conv2d_golden.npz"))
    cases = sorted({k.rsplit("/", 1)[0] for k in g.files})
    assert len(cases) == 4
    for c in cases:
        x, W, b = g[c + "/x"], g[c + "/W"], g[c + "/b"]
        tol = 2e-5 if x.dtype == np.float32 else 1e-12
        y = conv2d_np.conv2d_forward(x, W, b)
        assert np.abs(y - g[c + "/top"]).max() <= tol * np.abs(g[c + "/top"]).max()
        dW = np.full_like(W, 0.25); db = np.full_like(b, 0.25)
        dW, db, dx = conv2d_np.conv2d_backward(x, W, g[c + "/dtop"], dW, db)
        for got, name in ((dW, "dW"), (db, "db"), (dx, "dx")):
            ref = g[c + "/" + name]
            assert np.abs(got - ref).max() <= tol * np.abs(ref).max(), (c, name)
