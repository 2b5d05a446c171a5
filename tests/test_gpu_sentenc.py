"""GPU parity tests (-m gpu) of the sentence encoder: Convolution(kernel kh x D) / BN / Pooling / TanH through the C-ABI
(host mirror layers) against fixtures produced by the reference's own layers (tests/golden/sentenc_golden.npz) and the
numpy restatement (oracle/sentenc_np.py) on seeded inputs.

Tolerances (|got - ref| / max|ref|): float TF32 contractions (the convolution) 1e-3, the tolerance north_star states;
float SIMT convolution and the elementwise / reduction layers 2e-5 (BN's variance E[x^2]-E[x]^2: 1e-4); double 1e-11;
pooling and its argmax routing bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import mms_answer_selection_b200 as mms                      # noqa: E402
from mms_answer_selection_b200 import _lib                    # noqa: E402
from oracle import sentenc_np as snp                          # noqa: E402

TAG = {np.float32: "f32", np.float64: "f64"}


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sentenc_golden.npz"))


def err(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def blob(arr, dtype):
    b = mms.Blob(arr.shape, dtype=dtype)
    b.set_cpu_data(arr)
    return b


def conv_layer(dtype, C, kh, D, W, b, math=None):
    lay = mms.create_layer(mms.LayerParameter("Convolution", dtype=dtype, convolution_param=dict(
        num_output=C, kernel_h=kh, kernel_w=D, bias_term=b is not None)))
    return lay


def run_conv(dtype, x, W, b, dtop, math=None, dW0=0.5, reuse=False):
    C, _, kh, D = W.shape
    lay = conv_layer(dtype, C, kh, D, W, b)
    bottom, top = blob(x, dtype), mms.Blob((), dtype=dtype)
    lay.SetUp([bottom], [top])
    if reuse:                                                # backward reads the rounded x the forward left behind
        lay.handle.set_option(_lib.MMS_OPT_REUSE_FORWARD, 1)
    if math is not None:
        lay.set_math(math)
    lay.blobs[0].set_cpu_data(W)
    if b is not None:
        lay.blobs[1].set_cpu_data(b)
    lay.Forward([bottom], [top])
    y = top.cpu_data().copy()
    top.set_cpu_diff(dtop)
    for p in lay.blobs:
        p.diff.fill_(dW0)                                    # param diffs accumulate
    lay.Backward([top], [True], [bottom])
    return y, lay.blobs[0].cpu_diff(), (lay.blobs[1].cpu_diff() if b is not None else None), bottom.cpu_diff()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_conv_golden(gold, dtype):
    k = TAG[dtype] + "/conv/"
    y, dW, db, dx = run_conv(dtype, gold[k + "x"], gold[k + "W"], gold[k + "b"], gold[k + "dtop"])
    tol = 1e-3 if dtype == np.float32 else 1e-11
    assert y.shape == gold[k + "top"].shape
    assert err(y, gold[k + "top"]) <= tol and err(dx, gold[k + "dx"]) <= tol
    assert err(dW, gold[k + "dW"]) <= tol and err(db, gold[k + "db"]) <= (2e-5 if dtype == np.float32 else 1e-11)


@pytest.mark.parametrize("N,L,D,C,kh,math,tol", [
    (33, 40, 300, 100, 5, _lib.MMS_MATH_TF32, 1e-3),     # the reference's sentence convolution, tcgen05 path
    (7, 40, 300, 100, 5, _lib.MMS_MATH_FP32, 2e-5),      # same on the SIMT GEMM
    (9, 23, 50, 17, 3, _lib.MMS_MATH_TF32, 2e-5),        # D % 4 != 0: no 16-byte rows for TMA -> SIMT
    (5, 12, 64, 130, 12, _lib.MMS_MATH_TF32, 1e-3),      # kernel as tall as the sentence (T = 1), C > 128
    (1, 5, 8, 1, 5, _lib.MMS_MATH_TF32, 1e-3),           # one sentence, one window, one channel
    (70, 30, 128, 128, 3, _lib.MMS_MATH_TF32, 1e-3),     # dedicated kernels: 128 filters, 3 kernel rows
    (513, 40, 300, 64, 5, _lib.MMS_MATH_TF32, 1e-3),     # dedicated kernels: many row tiles, ragged last tile
    (3, 19, 36, 7, 8, _lib.MMS_MATH_TF32, 1e-3),         # 8 kernel rows: forward / dx dedicated, dW on the generic engine
    (40, 40, 300, 120, 5, _lib.MMS_MATH_TF32, 1e-3),     # 120 filters: filter blocks of 120 rows still fit two stages
])
def test_conv_vs_restatement(N, L, D, C, kh, math, tol):
    rng = np.random.default_rng(N * 1000 + D)
    x = rng.uniform(-1, 1, (N, 1, L, D)).astype(np.float32)
    W = (rng.uniform(-1, 1, (C, 1, kh, D)) * np.sqrt(3.0 / (kh * D))).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, C).astype(np.float32)
    dtop = rng.uniform(-1, 1, (N, C, L - kh + 1, 1)).astype(np.float32)
    y, dW, db, dx = run_conv(np.float32, x, W, b, dtop, math=math, dW0=0.0)
    f64 = lambda a: a.astype(np.float64)
    assert err(y, snp.conv_forward(f64(x), f64(W), f64(b))) <= tol
    rW, rb, rx = snp.conv_backward(f64(x), f64(W), f64(dtop))
    assert err(dW, rW) <= tol and err(dx, rx) <= tol and err(db, rb) <= 2e-5


def test_conv_backward_reusing_the_forward_copy_is_identical():
    rng = np.random.default_rng(12)
    N, L, D, C, kh = 19, 40, 300, 100, 5
    x = rng.uniform(-1, 1, (N, 1, L, D)).astype(np.float32)
    W = (rng.uniform(-1, 1, (C, 1, kh, D)) * 0.05).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, C).astype(np.float32)
    dtop = rng.uniform(-1, 1, (N, C, L - kh + 1, 1)).astype(np.float32)
    plain = run_conv(np.float32, x, W, b, dtop, dW0=0.0)
    reused = run_conv(np.float32, x, W, b, dtop, dW0=0.0, reuse=True)
    assert np.array_equal(plain[0], reused[0]) and np.array_equal(plain[3], reused[3])
    assert err(reused[1], plain[1]) <= 1e-6                   # split-K atomics: order only


def test_conv_without_bias_and_partial_propagation():
    rng = np.random.default_rng(3)
    N, L, D, C, kh = 4, 10, 16, 6, 5
    x = rng.uniform(-1, 1, (N, 1, L, D))
    W = rng.uniform(-0.2, 0.2, (C, 1, kh, D))
    lay = conv_layer(np.float64, C, kh, D, W, None)
    bottom, top = blob(x, np.float64), mms.Blob((), dtype=np.float64)
    lay.SetUp([bottom], [top])
    assert len(lay.blobs) == 1
    lay.blobs[0].set_cpu_data(W)
    lay.Forward([bottom], [top])
    assert err(top.cpu_data(), snp.conv_forward(x, W, None)) <= 1e-12
    top.set_cpu_diff(np.ones(top.shape))
    bottom.diff.fill_(7.0)
    lay.Backward([top], [False], [bottom])                    # propagate_down false: the bottom diff is left alone
    assert np.all(bottom.cpu_diff() == 7.0)
    # strided / padded / grouped convolutions are the stock layer's business
    with pytest.raises(mms.layers.CheckError, match="only stride 1, pad 0, group 1"):
        bad = mms.create_layer(mms.LayerParameter("Convolution", convolution_param=dict(num_output=3, kernel_size=5, stride=2)))
        bad.SetUp([blob(np.zeros((2, 4, 9, 9), np.float32), np.float32)], [mms.Blob(())])


def test_conv_adjoint_identity_at_full_size():
    """C3-sized batch (4096 sentences of 40 tokens, 300-d, 100 filters): <conv(x; W), G> = <x, dx(G)> = <W, dW(G)>
    (bias 0) -- a size-independent check of forward against both gradients."""
    N, L, D, C, kh = 4096, 40, 300, 100, 5
    g = torch.Generator(device="cuda").manual_seed(5)
    lay = mms.create_layer(mms.LayerParameter("Convolution", convolution_param=dict(
        num_output=C, kernel_h=kh, kernel_w=D, bias_term=False, weight_filler=dict(type="xavier"))))
    bottom, top = mms.Blob((N, 1, L, D)), mms.Blob(())
    bottom.data.copy_(torch.rand((N, 1, L, D), device="cuda", generator=g) * 2 - 1)
    lay.SetUp([bottom], [top])
    lay.Forward([bottom], [top])
    top.diff.copy_(torch.rand(top.data.shape, device="cuda", generator=g) * 2 - 1)
    lay.blobs[0].diff.zero_()
    lay.Backward([top], [True], [bottom])
    lhs = float((top.data.double() * top.diff.double()).sum())
    via_x = float((bottom.data.double() * bottom.diff.double()).sum())
    via_w = float((lay.blobs[0].data.double() * lay.blobs[0].diff.double()).sum())
    scale = float(top.data.double().abs().mean() * top.diff.double().abs().mean() * top.data.numel()) ** 0.5 + abs(lhs)
    assert abs(lhs - via_x) <= 1e-3 * scale and abs(lhs - via_w) <= 1e-3 * scale


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bn_golden(gold, dtype):
    k = TAG[dtype] + "/bn/"
    tol = 1e-4 if dtype == np.float32 else 1e-10
    par = dict(bn_param=dict(bn_memory=0.9))
    lay = mms.create_layer(mms.LayerParameter("BN", dtype=dtype, **par))
    bottom, top = blob(gold[k + "x0"], dtype), mms.Blob((), dtype=dtype)
    lay.SetUp([bottom], [top])
    assert [b.shape for b in lay.blobs] == [(1, gold[k + "x0"].shape[1], 1, 1)] * 4
    lay.blobs[0].set_cpu_data(gold[k + "scale"]); lay.blobs[1].set_cpu_data(gold[k + "shift"])
    lay.Forward([bottom], [top])
    bottom.set_cpu_data(gold[k + "x1"])
    lay.Forward([bottom], [top])                              # the running statistics have now blended twice
    assert err(top.cpu_data(), gold[k + "top"]) <= tol
    assert err(lay.blobs[2].cpu_data(), gold[k + "run_mean"]) <= tol and err(lay.blobs[3].cpu_data(), gold[k + "run_var"]) <= tol
    top.set_cpu_diff(gold[k + "dtop"])
    lay.blobs[0].diff.fill_(3.0); lay.blobs[1].diff.fill_(3.0)    # overwritten, not accumulated (gemv beta 0)
    lay.Backward([top], [True], [bottom])
    assert err(lay.blobs[0].cpu_diff(), gold[k + "dscale"]) <= tol and err(lay.blobs[1].cpu_diff(), gold[k + "dshift"]) <= tol
    assert err(bottom.cpu_diff(), gold[k + "dx"]) <= 10 * tol
    test = mms.create_layer(mms.LayerParameter("BN", dtype=dtype, phase="TEST", **par))
    tb, tt = blob(gold[k + "x0"], dtype), mms.Blob((), dtype=dtype)
    test.SetUp([tb], [tt])
    for i in range(4):
        test.blobs[i].set_cpu_data(lay.blobs[i].cpu_data())
    test.Forward([tb], [tt])
    assert err(tt.cpu_data(), gold[k + "top_test"]) <= tol
    with pytest.raises(mms.layers.CheckError, match="in-place"):
        mms.create_layer(mms.LayerParameter("BN", dtype=dtype)).SetUp([tb], [tb])


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-4), (np.float64, 1e-10)])
@pytest.mark.parametrize("shape", [(5, 7, 17, 1), (4, 3, 7, 1), (3, 5, 6, 6), (300, 100, 36, 1), (2, 700, 5, 5)])
def test_bn_vs_restatement(dtype, tol, shape):
    """Every kernel variant (vectorised planes, warp per plane, per element; coalesced and strided statistics)."""
    rng = np.random.default_rng(sum(shape))
    C = shape[1]
    x = (rng.standard_normal(shape) * rng.uniform(0.5, 2.0, (1, C, 1, 1)) + rng.uniform(-1, 1, (1, C, 1, 1))).astype(dtype)
    sc, sh = rng.uniform(0.5, 1.5, C).astype(dtype), rng.uniform(-0.5, 0.5, C).astype(dtype)
    dtop = rng.uniform(-1, 1, shape).astype(dtype)
    lay = mms.create_layer(mms.LayerParameter("BN", dtype=dtype))
    bottom, top = blob(x, dtype), mms.Blob((), dtype=dtype)
    lay.SetUp([bottom], [top])
    lay.blobs[0].set_cpu_data(sc.reshape(1, C, 1, 1)); lay.blobs[1].set_cpu_data(sh.reshape(1, C, 1, 1))
    lay.Forward([bottom], [top])
    f64 = lambda a: a.astype(np.float64)
    rt, xn, std, rm, rv = snp.bn_forward(f64(x), f64(sc), f64(sh), np.zeros(C), np.zeros(C), memory=float(np.float32(0.9)))
    assert err(top.cpu_data(), rt) <= tol and err(lay.blobs[2].cpu_data().reshape(-1), rm) <= tol
    assert err(lay.blobs[3].cpu_data().reshape(-1), rv) <= tol
    top.set_cpu_diff(dtop)
    lay.Backward([top], [True], [bottom])
    dsc, dsh, dx = snp.bn_backward(f64(dtop), xn, f64(sc), std)
    assert err(lay.blobs[0].cpu_diff().reshape(-1), dsc) <= tol and err(lay.blobs[1].cpu_diff().reshape(-1), dsh) <= tol
    assert err(bottom.cpu_diff(), dx) <= 10 * tol


POOLS = {"pool_time": dict(pool="MAX", kernel_w=1), "pool_ave2d": dict(pool="AVE", kernel_h=4, kernel_w=3, stride=2),
         "pool_max2d": dict(pool="MAX", kernel_size=3, stride_h=2, stride_w=1)}


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", sorted(POOLS))
def test_pooling_golden(gold, dtype, name):
    k = TAG[dtype] + "/" + name + "/"
    x = gold[k + "x"]
    pp = dict(POOLS[name])
    if name == "pool_time":
        pp["kernel_h"] = x.shape[2]                           # max over the whole time axis
    lay = mms.create_layer(mms.LayerParameter("Pooling", dtype=dtype, pooling_param=pp))
    bottom, top = blob(x, dtype), mms.Blob((), dtype=dtype)
    lay.SetUp([bottom], [top])
    lay.Forward([bottom], [top])
    assert top.shape == gold[k + "top"].shape
    top.set_cpu_diff(gold[k + "dtop"])
    lay.Backward([top], [True], [bottom])
    if pp["pool"] == "MAX":                                   # selection and routing: exact (ties go to the first maximum)
        assert np.array_equal(top.cpu_data(), gold[k + "top"]) and np.array_equal(bottom.cpu_diff(), gold[k + "dx"])
    else:
        tol = 2e-6 if dtype == np.float32 else 1e-14
        assert err(top.cpu_data(), gold[k + "top"]) <= tol and err(bottom.cpu_diff(), gold[k + "dx"]) <= tol


def test_pooling_padded_vs_restatement():
    rng = np.random.default_rng(8)
    x = rng.uniform(-1, 1, (2, 3, 11, 10))
    for method in ("MAX", "AVE"):
        lay = mms.create_layer(mms.LayerParameter("Pooling", dtype=np.float64, pooling_param=dict(
            pool=method, kernel_size=3, stride=2, pad=1)))
        bottom, top = blob(x, np.float64), mms.Blob((), dtype=np.float64)
        lay.SetUp([bottom], [top])
        lay.Forward([bottom], [top])
        ref, mask = snp.pool_forward(x, 3, 3, 2, 2, 1, 1, method)
        assert top.shape == ref.shape and err(top.cpu_data(), ref) <= 1e-14
        dtop = rng.uniform(-1, 1, ref.shape)
        top.set_cpu_diff(dtop)
        lay.Backward([top], [True], [bottom])
        assert err(bottom.cpu_diff(), snp.pool_backward(dtop, mask, x.shape, 3, 3, 2, 2, 1, 1, method)) <= 1e-14


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("H,W", [(7, 1), (36, 1), (33, 1), (5, 3)])
def test_global_max_pooling_with_ties(dtype, H, W):
    """One window over the whole plane (max over time), quantised values: the first maximum wins, in the warp-per-plane
    kernel (odd plane sizes) and in the vectorised thread-per-plane kernel alike."""
    rng = np.random.default_rng(H * 10 + W)
    x = (np.round(rng.uniform(-1, 1, (37, 11, H, W)) * 3) / 3).astype(dtype)
    lay = mms.create_layer(mms.LayerParameter("Pooling", dtype=dtype, pooling_param=dict(pool="MAX", kernel_h=H, kernel_w=W)))
    bottom, top = blob(x, dtype), mms.Blob((), dtype=dtype)
    lay.SetUp([bottom], [top])
    lay.Forward([bottom], [top])
    ref, mask = snp.pool_forward(x, H, W)
    assert np.array_equal(top.cpu_data(), ref)
    assert np.array_equal(lay.max_idx_.cpu().numpy().astype(np.int64), mask)
    dtop = rng.uniform(-1, 1, ref.shape).astype(dtype)
    top.set_cpu_diff(dtop)
    lay.Backward([top], [True], [bottom])
    assert np.array_equal(bottom.cpu_diff(), snp.pool_backward(dtop, mask, x.shape, H, W))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_tanh_golden(gold, dtype):
    k = TAG[dtype] + "/"
    tol = 1e-6 if dtype == np.float32 else 1e-14
    lay = mms.create_layer(mms.LayerParameter("TanH", dtype=dtype))
    b = blob(gold[k + "pool_time/top"], dtype)
    lay.SetUp([b], [b])                                       # in place, as the reference net runs it
    lay.Forward([b], [b])
    assert err(b.cpu_data(), gold[k + "tanh/top"]) <= tol
    b.set_cpu_diff(gold[k + "tanh/dtop"])
    lay.Backward([b], [True], [b])
    assert err(b.cpu_diff(), gold[k + "tanh/dx"]) <= tol


def test_sentence_encoder_chain_feeds_simmatrix():
    """Embed-shaped input -> Convolution(5 x D) -> BN -> Pooling(MAX over time) -> TanH for q and a, SimMatrix on the two
    sentence vectors; forward and the whole backward chain against the restatements (float64, SIMT path: exact)."""
    from oracle import cport
    rng = np.random.default_rng(21)
    N, L, D, C, kh = 6, 14, 12, 9, 5
    dt = np.float64
    outs, chains = [], []
    for s in range(2):
        x = rng.uniform(-1, 1, (N, 1, L, D))
        W = rng.uniform(-0.3, 0.3, (C, 1, kh, D)); b = rng.uniform(-0.1, 0.1, C)
        conv = mms.create_layer(mms.LayerParameter("Convolution", dtype=dt, convolution_param=dict(num_output=C, kernel_h=kh, kernel_w=D)))
        bn = mms.create_layer(mms.LayerParameter("BN", dtype=dt, bn_param=dict(scale_filler=dict(type="constant", value=1.0),
                                                                          shift_filler=dict(type="constant", value=1e-3))))
        pool = mms.create_layer(mms.LayerParameter("Pooling", dtype=dt, pooling_param=dict(pool="MAX", kernel_h=L - kh + 1, kernel_w=1)))
        tanh = mms.create_layer(mms.LayerParameter("TanH", dtype=dt))
        bx, by, bz, bp = blob(x, dt), mms.Blob((), dtype=dt), mms.Blob((), dtype=dt), mms.Blob((), dtype=dt)
        conv.SetUp([bx], [by]); conv.blobs[0].set_cpu_data(W); conv.blobs[1].set_cpu_data(b)
        conv.Forward([bx], [by])
        bn.SetUp([by], [bz]); bn.Forward([by], [bz])
        pool.SetUp([bz], [bp]); pool.Forward([bz], [bp])
        tanh.SetUp([bp], [bp]); tanh.Forward([bp], [bp])
        y = snp.conv_forward(x, W, b)
        z, xn, std, _, _ = snp.bn_forward(y, np.ones(C), np.full(C, 1e-3), np.zeros(C), np.zeros(C), memory=float(np.float32(0.9)))
        p, mask = snp.pool_forward(z, L - kh + 1, 1)
        v = np.tanh(p)
        assert err(bp.cpu_data(), v) <= 1e-11
        outs.append((bp, v))
        chains.append((conv, bn, pool, tanh, bx, by, bz, bp, x, W, xn, std, mask, z.shape))
    sim = mms.create_layer(mms.LayerParameter("SimMatrix", dtype=dt, sim_matrix_param=dict(weight_filler=dict(type="xavier"))))
    top = mms.Blob((), dtype=dt)
    sim.SetUp([outs[0][0], outs[1][0]], [top])
    Wm = sim.blobs[0].cpu_data()
    sim.Forward([outs[0][0], outs[1][0]], [top])
    qv, av = outs[0][1].reshape(N, C), outs[1][1].reshape(N, C)
    s_ref, _ = cport.simmatrix_forward(qv, av, Wm)
    assert err(top.cpu_data(), s_ref.reshape(top.shape)) <= 1e-11
    ds = rng.uniform(-1, 1, top.shape)
    top.set_cpu_diff(ds)
    sim.blobs[0].diff.zero_()
    sim.Backward([top], [True, True], [outs[0][0], outs[1][0]])
    dq_ref = ds.reshape(N, 1) * (av @ Wm.T)
    da_ref = ds.reshape(N, 1) * (qv @ Wm)
    for (conv, bn, pool, tanh, bx, by, bz, bp, x, W, xn, std, mask, zshape), dv, v in zip(chains, (dq_ref, da_ref), (qv, av)):
        assert err(bp.cpu_diff().reshape(N, C), dv) <= 1e-10
        tanh.Backward([bp], [True], [bp])
        pool.Backward([bp], [True], [bz])
        bn.Backward([bz], [True], [by])
        conv.blobs[0].diff.zero_(); conv.blobs[1].diff.zero_()
        conv.Backward([by], [True], [bx])
        dp = snp.tanh_backward(v.reshape(N, C, 1, 1), dv.reshape(N, C, 1, 1))
        dz = snp.pool_backward(dp, mask, zshape, zshape[2], 1)
        _, _, dy = snp.bn_backward(dz, xn, np.ones(C), std)
        rW, rb, rx = snp.conv_backward(x, W, dy)
        assert err(conv.blobs[0].cpu_diff(), rW) <= 1e-9
        # the bias gradient behind a BN layer is a sum that cancels to zero: absolute, against the size of its terms
        assert np.abs(conv.blobs[1].cpu_diff() - rb).max() <= 1e-12 * np.abs(dy).sum()
        assert err(bx.cpu_diff(), rx) <= 1e-9


# ---------------------------------------------------------------- the whole sentence-vector variant of the net
def _sentence_net_reference(idx_q, idx_a, label, P, margin, L, D, C, kh):
    """The step on the CPU restatements (float64): Embed -> Convolution -> BN -> MAX over time -> TanH per branch
    (shared parameters, running statistics blended by the question branch first), SimMatrix, PairRankLoss on the two
    halves of the score column; backward in the reverse layer order."""
    from oracle import cport
    N = idx_q.shape[0]
    h = N // 2
    mem = float(np.float32(0.9))
    rm, rv = np.zeros(C), np.zeros(C)
    br = []
    for idx in (idx_q, idx_a):
        e = cport.embed_forward(idx.astype(np.float64), P["W"], P["b"]).reshape(N, 1, L, D)
        y = snp.conv_forward(e, P["cW"], P["cb"])
        z, xn, std, rm, rv = snp.bn_forward(y, P["scale"].reshape(-1), P["shift"].reshape(-1), rm, rv, memory=mem)
        p, mask = snp.pool_forward(z, L - kh + 1, 1)
        br.append(dict(idx=idx, e=e, y=y, xn=xn, std=std, mask=mask, v=np.tanh(p), zshape=z.shape))
    vq, va = br[0]["v"].reshape(N, C), br[1]["v"].reshape(N, C)
    s, _ = cport.simmatrix_forward(vq, va, P["Wm"])
    s = s.reshape(N, 1)
    loss, ordered, similar = cport.pairrankloss_forward(s[:h].copy(), s[h:].copy(), label.reshape(h, 1).astype(np.float64), margin)
    da, db = cport.pairrankloss_backward(label.reshape(h, 1).astype(np.float64), ordered, similar, 1.0)
    ds = np.concatenate([da, db], axis=0)
    dWm, dvq, dva = cport.simmatrix_backward(vq, va, P["Wm"], ds, np.zeros_like(P["Wm"]))
    out = dict(loss=float(loss), dWm=dWm, dcW=np.zeros_like(P["cW"]), dcb=np.zeros_like(P["cb"]),
               dW=np.zeros_like(P["W"]), db=np.zeros_like(P["b"]), run_mean=rm, run_var=rv)
    for b, dv in ((br[1], dva), (br[0], dvq)):                # answer branch first: the question branch's BN diffs survive
        dp = snp.tanh_backward(b["v"], dv.reshape(N, C, 1, 1))
        dz = snp.pool_backward(dp, b["mask"], b["zshape"], L - kh + 1, 1)
        out["dscale"], out["dshift"], dy = snp.bn_backward(dz, b["xn"], P["scale"].reshape(-1), b["std"])
        dW_, db_, dx = snp.conv_backward(b["e"], P["cW"], dy)
        out["dcW"] += dW_; out["dcb"] += db_
        cport.embed_backward(b["idx"].astype(np.float64), dx.reshape(N * L, D), out["dW"], out["db"])
    return out


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-9), (np.float32, 2e-3)])
def test_sentence_vector_net_step(dtype, tol):
    """The north-star path as one net (Embed -> sentence encoder -> SimMatrix -> PairRankLoss) against the chained CPU
    restatements; float runs the TF32 contractions (1e-3 per contraction, 2e-3 through the chain), double is exact to
    summation order.  The CUDA-graph replay of the step reproduces the eager step."""
    N, L, D, C, kh, V, margin = 12, 16, 20, 9, 5, 50, 0.7
    rng = np.random.default_rng(31)
    net = mms.SentenceVectorNet(N, L, D, C, kh, V, dtype=dtype, margin=margin)
    bq = net.branches[0]
    P = dict(W=rng.uniform(-0.5, 0.5, (V, D)), b=rng.uniform(-0.1, 0.1, D), cW=rng.uniform(-0.3, 0.3, (C, 1, kh, D)),
             cb=rng.uniform(-0.1, 0.1, C), scale=rng.uniform(0.5, 1.5, (1, C, 1, 1)), shift=rng.uniform(-0.2, 0.2, (1, C, 1, 1)),
             Wm=rng.uniform(-0.5, 0.5, (C, C)))
    for blob_, key in ((bq["embed"].blobs[0], "W"), (bq["embed"].blobs[1], "b"), (bq["conv"].blobs[0], "cW"),
                       (bq["conv"].blobs[1], "cb"), (bq["bn"].blobs[0], "scale"), (bq["bn"].blobs[1], "shift"),
                       (net.sim.blobs[0], "Wm")):
        blob_.set_cpu_data(P[key])
        P[key] = blob_.cpu_data().astype(np.float64)          # what the net holds (float32-rounded in the float run)
    idx_q = rng.integers(0, V, (N, L)); idx_a = rng.integers(0, V, (N, L))
    idx_q[:, :3] = V - 1                                      # pad ids, as centre-padded sentences have
    label = rng.choice([1.0, 0.0, -1.0], N // 2)
    net.set_inputs(idx_q, idx_a, label)
    net.ClearParamDiffs()
    loss = net.ForwardBackward()
    ref = _sentence_net_reference(idx_q, idx_a, label, P, margin, L, D, C, kh)
    assert abs(loss - ref["loss"]) <= tol * max(abs(ref["loss"]), 1.0)
    got = dict(dWm=net.sim.blobs[0], dcW=bq["conv"].blobs[0], dcb=bq["conv"].blobs[1], dW=bq["embed"].blobs[0],
               db=bq["embed"].blobs[1], dscale=bq["bn"].blobs[0], dshift=bq["bn"].blobs[1])
    for key, blob_ in got.items():
        r = ref[key].reshape(blob_.shape)
        scale = max(np.abs(r).max(), 1e-6 if dtype == np.float64 else 1e-3)
        if key == "db":       # behind BN a constant added to every token cancels: the sum is ~0, judge it by its terms
            scale = np.abs(ref["dW"]).sum(axis=0).max()
        assert np.abs(blob_.cpu_diff() - r).max() <= tol * scale, key
    assert err(bq["bn"].blobs[2].cpu_data().reshape(-1), ref["run_mean"]) <= max(tol, 1e-4)
    # the step as one CUDA graph: identical parameter gradients (the running statistics blend once more)
    eager = {k: b_.cpu_diff().copy() for k, b_ in got.items()}
    net.capture(clear_diffs=True)
    net.replay()
    torch.cuda.synchronize()
    assert abs(net.loss_value() - ref["loss"]) <= 5 * tol * max(abs(ref["loss"]), 1.0)
    for key, blob_ in got.items():
        scale = np.abs(eager["dW"]).sum(axis=0).max() if key == "db" else max(np.abs(eager[key]).max(), 1e-3)
        assert np.abs(blob_.cpu_diff() - eager[key]).max() <= 5 * tol * scale, key


def test_empty_and_degenerate_calls_through_the_c_abi():
    """Zero-sized batches are no-ops that return 0; bad arguments are refused with MMS_E_INVALID and a message."""
    import ctypes
    L_ = _lib.lib()
    h = _lib.Handle()
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    x = torch.zeros(64, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    null = ctypes.c_void_p(0)
    assert L_.mms_sentconv_forward_f32(h.ptr, p(x), p(x), null, p(x), 0, 8, 4, 2, 5) == 0
    assert L_.mms_sentconv_backward_f32(h.ptr, p(x), p(x), p(x), p(x), null, p(x), 0, 8, 4, 2, 5) == 0
    assert L_.mms_pool_forward_f32(h.ptr, p(x), p(x), p(x), 0, 4, 4, 1, 1, 4, 4, 1, 1, 0, 0, 0) == 0
    assert L_.mms_tanh_forward_f32(h.ptr, null, null, 0) == 0
    assert L_.mms_sentconv_forward_f32(h.ptr, p(x), p(x), null, p(x), 2, 4, 4, 2, 5) == _lib.MMS_E_INVALID   # kernel taller than L
    assert b"bad size" in L_.mms_last_error()
    assert L_.mms_pool_forward_f32(h.ptr, p(x), p(x), null, 2, 4, 4, 1, 1, 4, 4, 1, 1, 0, 0, 0) == _lib.MMS_E_INVALID  # MAX needs a mask
    assert L_.mms_bn_forward_f32(h.ptr, p(x), p(x), p(x), p(x), p(x), p(x), p(x), p(x), p(x), 0, 2, 4, 1,
                                 ctypes.c_float(0.9), ctypes.c_float(1e-9)) == _lib.MMS_E_INVALID
    out = torch.zeros(2, device="cuda")
    assert L_.mms_rank_map_mrr_f32(h.ptr, p(x), 2, 1, p(x), p(x), 0, p(out), ctypes.c_void_p(out.data_ptr() + 4)) == 0
    torch.cuda.synchronize()
    assert torch.isnan(out).all()                             # no group qualifies: 0/0, as in the reference
