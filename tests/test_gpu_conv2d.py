"""The CNN over the similarity tensor (do_trec_qa_clean.py:470-477): mms_conv2d_* (implicit GEMMs on tcgen05), Dropout,
and the net that chains SimCross -> Dropout -> (Conv5x5 + BN -> AvePool -> TanH) x 2.
Oracle: fixtures from the reference's own ConvolutionLayer compiled in place (tests/golden/conv2d_golden.npz), the numpy
restatement pinned to them (oracle/conv2d_np.py) at larger sizes, and for the net the chained oracles of every layer.
Tolerance: 1e-3 of the tensor's largest magnitude for the TF32 contractions (GradientChecker's scale rule), 2e-5 on the
SIMT path (MMS_MATH_FP32), 1e-11 for double; Dropout is exact."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import scaled_err

pytestmark = pytest.mark.gpu

import mms_answer_selection_b200 as mms  # noqa: E402
from mms_answer_selection_b200 import _lib, synth  # noqa: E402
from mms_answer_selection_b200.simcnn import SimCNNNet  # noqa: E402
from oracle import conv2d_np, cport, sentenc_np  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conv2d_golden.npz"))


def _conv_layer(x, W, b, dtype, math=None):
    N, C, H, Wd = x.shape
    Co, _, kh, kw = W.shape
    lay = mms.ConvolutionLayer(mms.LayerParameter("Convolution", dtype=dtype, convolution_param=dict(
        num_output=Co, kernel_h=kh, kernel_w=kw)))
    bx, top = mms.Blob(x.shape, dtype=dtype), mms.Blob((), dtype=dtype)
    bx.set_cpu_data(x)
    lay.SetUp([bx], [top])
    if math is not None:
        lay.set_math(math)
    lay.blobs[0].set_cpu_data(W); lay.blobs[1].set_cpu_data(b)
    return lay, bx, top


@pytest.mark.parametrize("case", sorted({k.rsplit("/", 1)[0] for k in GOLD.files}))
@pytest.mark.parametrize("math", ["tf32", "fp32"])
def test_conv2d_vs_reference_fixtures(case, math):
    g = {k.rsplit("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(case + "/")}
    dtype = g["x"].dtype
    if dtype == np.float64 and math == "tf32":
        pytest.skip("double blobs always take the direct path")
    lay, bx, top = _conv_layer(g["x"], g["W"], g["b"], dtype, _lib.MMS_MATH_FP32 if math == "fp32" else None)
    lay.Forward([bx], [top])
    tol = 1e-11 if dtype == np.float64 else (1e-3 if math == "tf32" else 2e-5)
    assert top.shape == g["top"].shape
    assert scaled_err(top.cpu_data(), g["top"]) <= tol
    top.set_cpu_diff(g["dtop"])
    for b in lay.blobs:
        b.diff.fill_(0.25)                                   # parameter diffs accumulate
    lay.Backward([top], [True], [bx])
    assert scaled_err(lay.blobs[0].cpu_diff(), g["dW"]) <= tol
    assert scaled_err(lay.blobs[1].cpu_diff(), g["db"]) <= (1e-11 if dtype == np.float64 else 2e-5)
    assert scaled_err(bx.cpu_diff(), g["dx"]) <= tol


@pytest.mark.parametrize("N,C,H,W,Co,kh,kw", [(64, 4, 40, 40, 32, 5, 5), (300, 32, 9, 9, 64, 5, 5), (7, 5, 13, 21, 24, 3, 7),
                                               (2, 1, 6, 6, 128, 6, 6), (33, 8, 12, 12, 16, 1, 1)])
def test_conv2d_vs_numpy_oracle(N, C, H, W, Co, kh, kw):
    rng = np.random.default_rng(N)
    x = rng.uniform(-1, 1, (N, C, H, W)).astype(np.float32)
    Wt = rng.uniform(-0.2, 0.2, (Co, C, kh, kw)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, Co).astype(np.float32)
    lay, bx, top = _conv_layer(x, Wt, b, np.float32)
    lay.Forward([bx], [top])
    y = conv2d_np.conv2d_forward(x.astype(np.float64), Wt.astype(np.float64), b.astype(np.float64))
    assert scaled_err(top.cpu_data(), y) <= 1e-3
    dy = rng.uniform(-1, 1, y.shape).astype(np.float32)
    top.set_cpu_diff(dy)
    lay.blobs[0].diff.zero_(); lay.blobs[1].diff.zero_()
    lay.Backward([top], [True], [bx])
    dW, db, dx = conv2d_np.conv2d_backward(x.astype(np.float64), Wt.astype(np.float64), dy.astype(np.float64),
                                           np.zeros(Wt.shape), np.zeros(Co))
    assert scaled_err(lay.blobs[0].cpu_diff(), dW) <= 1e-3
    assert scaled_err(lay.blobs[1].cpu_diff(), db) <= 2e-5
    assert scaled_err(bx.cpu_diff(), dx) <= 1e-3
    # adjoint identities at this size: <conv(x), dy> = <x, dx> = <W, dW> (+ bias part)
    lhs = float((y * dy).sum())
    assert abs(float((x.astype(np.float64) * bx.cpu_diff()).sum()) + float((b * db).sum()) - lhs) <= 2e-3 * abs(lhs) + 1e-6


def test_conv2d_full_size_adjoint_property():
    """C3 size (4096 x 4 x 40 x 40 -> 32 channels): too big for the host oracle in seconds; linearity and the adjoint
    identities <conv(x; W), G> = <x, dx(G)> = <W, dW(G)> hold on the device, in float64 reductions."""
    N, C, H, W, Co, k = 4096, 4, 40, 40, 32, 5
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand((N, C, H, W), device="cuda", generator=g) - 0.5
    Wt = (torch.rand((Co, C, k, k), device="cuda", generator=g) - 0.5) * 0.2
    lay = mms.ConvolutionLayer(mms.LayerParameter("Convolution", convolution_param=dict(num_output=Co, kernel_h=k, kernel_w=k,
                                                                                       bias_term=False)))
    bx, top = mms.Blob(x.shape), mms.Blob(())
    bx.data.copy_(x)
    lay.SetUp([bx], [top])
    lay.blobs[0].data.copy_(Wt)
    lay.Forward([bx], [top])
    G = torch.rand(top.shape, device="cuda", generator=g) - 0.5
    top.diff.copy_(G)
    lay.blobs[0].diff.zero_()
    lay.Backward([top], [True], [bx])
    torch.cuda.synchronize()
    lhs = (top.data.double() * G.double()).sum().item()
    assert abs((x.double() * bx.diff.double()).sum().item() - lhs) <= 2e-3 * abs(lhs)
    assert abs((Wt.double() * lay.blobs[0].diff.double()).sum().item() - lhs) <= 2e-3 * abs(lhs)
    # a slice of the batch against the host oracle
    y = conv2d_np.conv2d_forward(x[:4].cpu().numpy().astype(np.float64), Wt.cpu().numpy().astype(np.float64))
    assert scaled_err(top.data[:4].cpu().numpy(), y) <= 1e-3


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_dropout_matches_the_reference_formula(dtype):
    rng = np.random.default_rng(1)
    x = rng.normal(0, 1, (6, 4, 9, 5)).astype(dtype)
    words = rng.integers(0, 2 ** 32, x.size, dtype=np.uint64).astype(np.uint32)
    lay = mms.DropoutLayer(mms.LayerParameter("Dropout", dtype=dtype, dropout_param=dict(dropout_ratio=0.1)))
    bx, top = mms.Blob(x.shape, dtype=dtype), mms.Blob((), dtype=dtype)
    bx.set_cpu_data(x)
    lay.SetUp([bx], [top])
    lay.set_mask(words)
    lay.Forward([bx], [top])
    np.testing.assert_array_equal(top.cpu_data(), conv2d_np.dropout(x, words, 0.1))
    dy = rng.normal(0, 1, x.shape).astype(dtype)
    top.set_cpu_diff(dy)
    lay.Backward([top], [True], [bx])
    np.testing.assert_array_equal(bx.cpu_diff(), conv2d_np.dropout(dy, words, 0.1))
    # drawn masks: about 10 % dropped, a different mask on every forward, the same stream for the same seed
    lay2 = mms.DropoutLayer(mms.LayerParameter("Dropout", dtype=dtype, dropout_param=dict(dropout_ratio=0.1)))
    big = mms.Blob((1 << 20,), dtype=dtype); big.data.fill_(1.0)
    t2 = mms.Blob((), dtype=dtype)
    lay2.SetUp([big], [t2])
    lay2.Forward([big], [t2])
    first = t2.cpu_data().copy()
    assert abs(float((first == 0).mean()) - 0.1) < 0.003
    lay2.Forward([big], [t2])
    assert (t2.cpu_data() != first).any()
    # TEST phase: identity
    lay3 = mms.DropoutLayer(mms.LayerParameter("Dropout", dtype=dtype, phase="TEST", dropout_param=dict(dropout_ratio=0.1)))
    t3 = mms.Blob((), dtype=dtype)
    lay3.SetUp([bx], [t3]); lay3.Forward([bx], [t3])
    np.testing.assert_array_equal(t3.cpu_data(), x)


def test_simcnn_net_vs_chained_oracles():
    """network_v4 up to Flatten on a small batch against the oracle chain: cport (Embed, SimCross), conv2d_np (Dropout,
    Convolution), sentenc_np (the fork's BN, Pooling, TanH)."""
    N, L, D, mc, V = 6, 40, 52, 4, 300
    d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
    rng = np.random.default_rng(8)
    net = SimCNNNet(N, L, D, mc, V)
    net.embed_q.blobs[0].set_cpu_data(d["W"]); net.embed_q.blobs[1].set_cpu_data(d["b"])
    M = (rng.uniform(-1, 1, d["M"].shape) * 3.0).astype(np.float32)          # scores of order 1 so that BN sees a signal
    net.sim.blobs[0].set_cpu_data(M); net.sim.blobs[1].set_cpu_data(d["B"])
    net.set_inputs(d["idx_q"], d["idx_a"])
    words = rng.integers(0, 2 ** 32, N * mc * L * L, dtype=np.uint64).astype(np.uint32)
    net.drop.Reshape([net.S], [net.Sd]) if net.drop.rand_vec_ is None else None
    net.drop.set_mask(words)
    dfeat = rng.normal(0, 1, (N, 64, 1, 1)).astype(np.float32)
    net.ClearParamDiffs()
    net.Forward()
    net.set_upstream_gradient(dfeat)
    net.Backward()
    torch.cuda.synchronize()
    # ---- oracle chain (float64 where the restatements allow it)
    q = cport.embed_forward(d["idx_q"], d["W"], d["b"]); a = cport.embed_forward(d["idx_a"], d["W"], d["b"])
    S, _, _ = cport.simcross_forward(2, q, a, M, d["B"])
    assert scaled_err(net.S.cpu_data(), S) <= 1e-3
    W0, b0 = net.conv0.blobs[0].cpu_data(), net.conv0.blobs[1].cpu_data()
    W1, b1 = net.conv1.blobs[0].cpu_data(), net.conv1.blobs[1].cpu_data()
    Sd = conv2d_np.dropout(net.S.cpu_data(), words, 0.1)       # from the device's S: the mask is exact, S within tolerance
    np.testing.assert_array_equal(net.Sd.cpu_data(), Sd)
    c0 = conv2d_np.conv2d_forward(Sd.astype(np.float64), W0.astype(np.float64), b0.astype(np.float64))
    assert scaled_err(net.c0.cpu_data(), c0) <= 1e-3
    # BN / pooling / tanh: the device layers were checked one by one in test_gpu_sentenc.py; here the composition's end
    assert net.feat.shape == (N, 64, 1, 1) and np.isfinite(net.feat.cpu_data()).all()
    assert np.abs(net.feat.cpu_data()).max() <= 1.0
    # gradient chain: central finite differences of <feat, dfeat> w.r.t. a few entries of conv0's weights and of M
    def loss_of():
        net.Forward()
        return float((net.feat.data.double() * torch.from_numpy(dfeat).cuda().double().reshape(net.feat.shape)).sum().item())
    for blob, grad, idxs, h in ((net.conv1.blobs[0], net.conv1.blobs[0].cpu_diff().copy(), [(3, 5, 2, 2), (60, 31, 4, 0)], 2e-2),
                                (net.conv0.blobs[0], net.conv0.blobs[0].cpu_diff().copy(), [(0, 0, 0, 0), (17, 3, 4, 2)], 2e-2)):
        for ix in idxs:
            w0 = float(blob.data[ix].item())
            blob.data[ix] = w0 + h; lp = loss_of()
            blob.data[ix] = w0 - h; lm = loss_of()
            blob.data[ix] = w0
            fd = (lp - lm) / (2 * h)
            assert abs(fd - float(grad[ix])) <= 0.05 * max(abs(fd), abs(float(grad[ix]))) + 2e-3, (ix, fd, float(grad[ix]))


def test_bn_and_tiled_ave_pool_at_the_cnn_geometry():
    """The fork's BN on 36 x 36 planes (plane-streaming statistics kernel) and AVE 4 x 4 / stride 4 pooling (tiled fast
    path) against the numpy restatements pinned to the reference's own layers (oracle/sentenc_np.py)."""
    rng = np.random.default_rng(12)
    N, C, H, W = 9, 32, 36, 36
    x = (rng.normal(0.3, 1.5, (N, C, H, W)) * rng.uniform(0.5, 2, (1, C, 1, 1))).astype(np.float32)
    bn = mms.BNLayer(mms.LayerParameter("BN", bn_param=dict(scale_filler=dict(type="constant", value=1.0),
                                                              shift_filler=dict(type="constant", value=1e-3))))
    bx, bt = mms.Blob(x.shape), mms.Blob(())
    bx.set_cpu_data(x)
    bn.SetUp([bx], [bt])
    scale = rng.uniform(0.5, 1.5, (1, C, 1, 1)).astype(np.float32); shift = rng.uniform(-0.2, 0.2, (1, C, 1, 1)).astype(np.float32)
    bn.blobs[0].set_cpu_data(scale); bn.blobs[1].set_cpu_data(shift)
    bn.Forward([bx], [bt])
    top, xn, std, rm, rv = sentenc_np.bn_forward(x.astype(np.float64), scale.astype(np.float64), shift.astype(np.float64),
                                                 np.zeros(C), np.zeros(C))
    assert scaled_err(bt.cpu_data(), top) <= 1e-4
    assert scaled_err(bn.blobs[2].cpu_data().reshape(-1), rm) <= 1e-5 and scaled_err(bn.blobs[3].cpu_data().reshape(-1), rv) <= 1e-4
    dy = rng.normal(0, 1, x.shape).astype(np.float32)
    bt.set_cpu_diff(dy)
    bn.Backward([bt], [True], [bx])
    dscale, dshift, dx = sentenc_np.bn_backward(dy.astype(np.float64), xn, scale.astype(np.float64), std)
    assert scaled_err(bn.blobs[0].cpu_diff().reshape(-1), dscale) <= 1e-4
    assert scaled_err(bn.blobs[1].cpu_diff().reshape(-1), dshift) <= 1e-4
    assert scaled_err(bx.cpu_diff(), dx) <= 1e-4
    for dtype in (np.float32, np.float64):
        pool = mms.PoolingLayer(mms.LayerParameter("Pooling", dtype=dtype, pooling_param=dict(
            pool="AVE", kernel_h=4, kernel_w=4, stride_h=4, stride_w=4)))
        px, pt = mms.Blob(x.shape, dtype=dtype), mms.Blob((), dtype=dtype)
        px.set_cpu_data(x.astype(dtype))
        pool.SetUp([px], [pt])
        pool.Forward([px], [pt])
        ref, _ = sentenc_np.pool_forward(x.astype(np.float64), 4, 4, 4, 4, method="AVE")
        assert pt.shape == (N, C, 9, 9) and scaled_err(pt.cpu_data(), ref) <= (1e-6 if dtype == np.float32 else 1e-14)
        g = rng.normal(0, 1, ref.shape).astype(dtype)
        pt.set_cpu_diff(g)
        pool.Backward([pt], [True], [px])
        dref = sentenc_np.pool_backward(g.astype(np.float64), None, x.shape, 4, 4, 4, 4, method="AVE")
        assert scaled_err(px.cpu_diff(), dref) <= (1e-6 if dtype == np.float32 else 1e-14)
