"""The drop-in boundary itself: the product's C++ layer classes
(mms_answer_selection_b200/caffe_layers/*.cpp, compiled against the reference's unmodified
headers) are created through the reference's LayerRegistry::CreateLayer and driven through
Layer::SetUp / Forward / Backward with the reference's own Blob + SyncedMemory in Caffe::GPU mode
(oracle/_ref/libmms_dropin.so), side by side with the reference's own layers on the CPU
(oracle/_ref/libmms_ref.so), on identical inputs and parameters.

Tolerances: Embed gather bit-exact; fp32 elementwise layers 1e-5; TF32 contractions 1e-3 of the
tensor's largest magnitude (GradientChecker's scale rule, test_gradient_check_util.hpp:172-175)."""
import numpy as np
import pytest

from conftest import scaled_err
from oracle import cport, refbind

needs_dropin = pytest.mark.skipif(not (refbind.dropin_available() and refbind.ref_available()),
                                  reason="oracle/_ref/libmms_{ref,dropin}.so not built (needs /root/reference at build time)")


def pair(type_, bottoms, params, dtype, num_top=1):
    ref = refbind.RefLayer(type_, bottoms, params, dtype=dtype, num_top=num_top)
    refbind.dropin_lib().mmsref_set_mode(1)
    new = refbind.DropinLayer(type_, bottoms, params, dtype=dtype, num_top=num_top)
    assert new.num_blobs() == ref.num_blobs()
    for i in range(ref.num_blobs()):
        assert new.shape("blob", i) == ref.shape("blob", i)
        new.write("blob", i, ref.read("blob", i))          # identical parameters
    for i in range(num_top):
        assert new.shape("top", i) == ref.shape("top", i)
    return ref, new


@needs_dropin
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("bias", [True, False])
def test_embed(dtype, bias):
    rng = np.random.default_rng(3)
    idx = rng.integers(0, 97, size=(6, 11)).astype(dtype)
    idx[:, :3] = 96                                        # a run of pad ids, like centre-padded sentences
    params = {"num_output": 20, "input_dim": 97, "embed.bias_term": int(bias), "weight_filler.type": "uniform",
              "weight_filler.min": -0.08, "weight_filler.max": 0.08, "bias_filler.type": "gaussian"}
    ref, new = pair("Embed", [idx], params, dtype)
    ref.forward(); new.forward()
    assert np.array_equal(ref.read("top", 0), new.read("top", 0))          # bit-exact gather (+bias)
    dtop = rng.standard_normal(ref.shape("top", 0)).astype(dtype)
    for l in (ref, new):
        l.write("top", 0, dtop, diff=True)
        for i in range(l.num_blobs()):
            l.write("blob", i, np.ones(l.shape("blob", i)), diff=True)     # param diffs ACCUMULATE
        l.backward([False])
    for i in range(ref.num_blobs()):
        assert scaled_err(new.read("blob", i, diff=True), ref.read("blob", i, diff=True)) <= 1e-5
    with pytest.raises(refbind.RefError, match="backpropagate to EmbedLayer input"):
        new.backward([True])


@needs_dropin
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_embed_weight_source(dtype, tmp_path):
    """embed_param.weight_source through LayerSetUp of the reference layer and of the drop-in layer: the same table,
    bit for bit (embed_layer.cpp:46-113), and the same gather from it."""
    rng = np.random.default_rng(19)
    V, D = 31, 10
    path = tmp_path / "vectors.bin"
    vecs = rng.uniform(-1, 1, (V - 2, D)).astype("<f4")
    with open(path, "wb") as f:
        f.write(b"%d %d\n" % vecs.shape)
        for i, v in enumerate(vecs):
            f.write(b"w%d " % i + v.tobytes() + b"\n")
    idx = rng.integers(0, V, size=(4, 9)).astype(dtype)
    params = {"num_output": D, "input_dim": V, "embed.bias_term": 0, "weight_filler.type": "constant",
              "weight_filler.value": 0.125, "weight_source": str(path)}
    ref = refbind.RefLayer("Embed", [idx], params, dtype=dtype)
    refbind.dropin_lib().mmsref_set_mode(1)
    new = refbind.DropinLayer("Embed", [idx], params, dtype=dtype)
    assert new.read("blob", 0).tobytes() == ref.read("blob", 0).tobytes()
    ref.forward(); new.forward()
    assert np.array_equal(ref.read("top", 0), new.read("top", 0))


@needs_dropin
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_simcross(dtype, mode):
    rng = np.random.default_rng(5 + mode)
    N, Lq, La, D, mc = 5, 7, 9, 50, 3
    q = rng.uniform(-0.5, 0.5, size=(N, Lq, D)).astype(dtype)
    a = rng.uniform(-0.5, 0.5, size=(N, La, D)).astype(dtype)
    params = {"dist_mode": mode, "mesure_count": mc, "sim_cross.bias_term": 1, "weight_filler.type": "uniform",
              "weight_filler.min": -0.1, "weight_filler.max": 0.1, "bias_filler.type": "uniform",
              "bias_filler.min": -0.1, "bias_filler.max": 0.1}
    ref, new = pair("SimCross", [q, a], params, dtype)
    assert ref.shape("top", 0) == (N, mc if mode == 2 else 1, Lq, La)
    ref.forward(); new.forward()
    tol = 1e-3 if (dtype == np.float32 and mode == 2) else (1e-5 if dtype == np.float32 else 1e-11)
    assert scaled_err(new.read("top", 0), ref.read("top", 0)) <= tol
    dS = rng.uniform(-1, 1, size=ref.shape("top", 0)).astype(dtype)
    for l in (ref, new):
        l.write("top", 0, dS, diff=True)
        for i in range(l.num_blobs()):
            l.write("blob", i, np.full(l.shape("blob", i), 0.25), diff=True)
        l.backward([True, True])
    gtol = 1e-3 if dtype == np.float32 else 1e-9
    for i in range(2):
        assert scaled_err(new.read("bottom", i, diff=True), ref.read("bottom", i, diff=True)) <= gtol
    for i in range(ref.num_blobs()):
        # dM is overwritten, dB accumulates onto the 0.25 already there (sim_cross_layer.cpp:256, :301-304)
        assert scaled_err(new.read("blob", i, diff=True), ref.read("blob", i, diff=True)) <= gtol


@needs_dropin
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_simmatrix(dtype):
    rng = np.random.default_rng(7)
    N, K1, K2 = 33, 70, 45
    q = np.tanh(rng.standard_normal((N, K1))).astype(dtype)
    a = np.tanh(rng.standard_normal((N, K2))).astype(dtype)
    params = {"weight_filler.type": "xavier"}
    ref, new = pair("SimMatrix", [q, a], params, dtype)
    ref.forward(); new.forward()
    tol = 1e-3 if dtype == np.float32 else 1e-11
    assert scaled_err(new.read("top", 0), ref.read("top", 0)) <= tol
    # forward parks T = q W in bottom[1]'s diff buffer (sim_matrix_layer.cpp:58)
    assert scaled_err(new.read("bottom", 1, diff=True), ref.read("bottom", 1, diff=True)) <= tol
    ds = rng.standard_normal((N, 1)).astype(dtype)
    for l in (ref, new):
        l.write("top", 0, ds, diff=True)
        l.write("blob", 0, np.full(l.shape("blob", 0), 0.5), diff=True)
        l.backward([True, True])
    for kind, i in (("bottom", 0), ("bottom", 1), ("blob", 0)):
        assert scaled_err(new.read(kind, i, diff=True), ref.read(kind, i, diff=True)) <= tol


@needs_dropin
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pairrankloss(dtype):
    rng = np.random.default_rng(11)
    N = 257
    sa = rng.standard_normal((N, 1)).astype(dtype)
    sb = rng.standard_normal((N, 1)).astype(dtype)
    y = rng.choice([1.0, 0.0, -1.0], size=(N, 1)).astype(dtype)
    ref, new = pair("PairRankLoss", [sa, sb, y], {"margin": 0.7}, dtype)
    lr, ln = ref.forward(), new.forward()
    assert abs(lr - ln) <= 1e-5 * max(abs(lr), 1.0)
    ref.backward([True, True, False]); new.backward([True, True, False])
    for i in range(2):
        assert scaled_err(new.read("bottom", i, diff=True), ref.read("bottom", i, diff=True)) <= 1e-6
    with pytest.raises(refbind.RefError, match="cannot backpropagate to label inputs"):
        new.backward([True, True, True])


@needs_dropin
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fm(dtype):
    rng = np.random.default_rng(13)
    x = rng.uniform(-1, 1, size=(19, 4, 33)).astype(dtype)
    ref, new = pair("FM", [x], {"fm.bias_term": 1}, dtype)
    for l in (ref, new):
        l.write("blob", 0, np.array([0.3]))
    ref.forward(); new.forward()
    assert scaled_err(new.read("top", 0), ref.read("top", 0)) <= 1e-5
    dy = rng.standard_normal((19, 1)).astype(dtype)
    for l in (ref, new):
        l.write("top", 0, dy, diff=True)
        l.backward([True])
    assert scaled_err(new.read("bottom", 0, diff=True), ref.read("bottom", 0, diff=True)) <= 1e-5
    assert scaled_err(new.read("blob", 0, diff=True), ref.read("blob", 0, diff=True)) <= 1e-5


@needs_dropin
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ranking_metric_layers(dtype):
    """MAP / MRR / AUC / RankAccuracy: the reference's layers on the CPU vs the drop-in classes in Caffe::GPU mode
    (distinct scores: equal scores have no specified order in the reference).  The device sums in double and rounds
    once, the reference accumulates in Dtype."""
    rng = np.random.default_rng(17)
    n = 1517
    prob = np.zeros((n, 2), dtype)
    prob[:, 1] = (rng.permutation(n) + 0.5) / n
    prob[:, 0] = 1 - prob[:, 1]
    label = (rng.uniform(0, 1, n) < 0.17).astype(dtype)
    group = rng.integers(-2, 66, n).astype(dtype)
    b = rng.uniform(0, 1, n).astype(dtype)
    y = np.where(rng.uniform(0, 1, n) < 0.5, 1.0, -1.0).astype(dtype)
    tol = 2e-6 if dtype == np.float32 else 1e-12
    for typ, params, bots in (("MAP", {"map.fixed_axis": 1}, [prob, label, group]),
                              ("MRR", {"mrr.fixed_axis": 1}, [prob, label, group]),
                              ("AUC", {"auc.fixed_axis": 1}, [prob, label]),
                              ("AUC", {"auc.fixed_axis": 1, "auc.ignore_label": 3}, [prob, label]),
                              ("RankAccuracy", {}, [np.ascontiguousarray(prob[:, 1]), b, y])):
        ref, new = pair(typ, bots, params, dtype)
        ref.forward(); new.forward()
        r, g = float(ref.read("top", 0).reshape(-1)[0]), float(new.read("top", 0).reshape(-1)[0])
        assert 0.0 < r <= 1.0 and abs(g - r) <= tol * abs(r), (typ, r, g)


@needs_dropin
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bn(dtype):
    """The fork's BN layer: two TRAIN forwards (running statistics blend twice), backward, against the reference's
    own layer; float 1e-4 (var = E[x^2] - E[x]^2 cancels), double 1e-10."""
    rng = np.random.default_rng(23)
    x0 = (rng.standard_normal((9, 6, 12, 1)) * 0.7 + 0.3).astype(dtype)
    x1 = (x0 * 1.5 - 0.2).astype(dtype)
    params = {"scale_filler.type": "uniform", "scale_filler.min": 0.5, "scale_filler.max": 1.5,
              "shift_filler.type": "uniform", "shift_filler.min": -0.2, "shift_filler.max": 0.2, "bn_memory": 0.9}
    ref, new = pair("BN", [x0], params, dtype)
    tol = 1e-4 if dtype == np.float32 else 1e-10
    dz = rng.uniform(-1, 1, x0.shape).astype(dtype)
    for l in (ref, new):
        l.forward()
        l.write("bottom", 0, x1)
        l.forward()
        l.write("top", 0, dz, diff=True)
        for i in range(2):
            l.write("blob", i, np.full(l.shape("blob", i), 3.0), diff=True)      # overwritten by Backward
        l.backward([True])
    assert scaled_err(new.read("top", 0), ref.read("top", 0)) <= tol
    for i in (2, 3):
        assert scaled_err(new.read("blob", i), ref.read("blob", i)) <= tol        # running mean / variance
    for i in (0, 1):
        assert scaled_err(new.read("blob", i, diff=True), ref.read("blob", i, diff=True)) <= tol
    assert scaled_err(new.read("bottom", 0, diff=True), ref.read("bottom", 0, diff=True)) <= 10 * tol


# ---- no GPU needed: the drop-in library registers the five types and has no CPU path -----------
@needs_dropin
def test_dropin_registers_the_reference_layer_types_and_refuses_cpu_mode():
    refbind.dropin_lib().mmsref_set_mode(0)                # Caffe::CPU
    try:
        x = np.zeros((2, 3, 4), dtype=np.float32)
        shapes = {"Embed": ([np.zeros((2, 3), np.float32)], {"num_output": 4, "input_dim": 5}),
                  "SimCross": ([x, x], {"dist_mode": 2, "mesure_count": 2}),
                  "SimMatrix": ([x, x], {}), "FM": ([x], {}),
                  "PairRankLoss": ([np.zeros((2, 1), np.float32)] * 3, {}),
                  "MAP": ([np.zeros((2, 2), np.float32)] + [np.zeros((2,), np.float32)] * 2, {}),
                  "MRR": ([np.zeros((2, 2), np.float32)] + [np.zeros((2,), np.float32)] * 2, {}),
                  "AUC": ([np.zeros((2, 2), np.float32), np.zeros((2,), np.float32)], {}),
                  "RankAccuracy": ([np.zeros((2,), np.float32)] * 3, {}),
                  "BN": ([np.zeros((2, 3, 4, 1), np.float32)], {})}
        for type_, (bottoms, params) in shapes.items():
            layer = refbind.DropinLayer(type_, bottoms, params)
            with pytest.raises(refbind.RefError, match="runs on the GPU only"):
                layer.forward()
    finally:
        refbind.dropin_lib().mmsref_set_mode(1)


@needs_dropin
@pytest.mark.gpu
@pytest.mark.parametrize("world,fused", [(2, 0), (4, 0), (4, 1)])
def test_cpp_grad_exchange_glue(world, fused):
    """caffe_layers/mms_grad_exchange.cpp -- the C++ body of P2PSync's Params re-binding (parallel.cpp:60-115), on_start
    (:287-322) and on_gradients_ready (:325-380) -- with `world` solver replicas of reference Blobs on one device:
    after on_start every replica holds rank 0's weights; after on_gradients_ready every replica's diff is the mean of
    the replicas' gradients (fixed rank order: bit-exact); the fused form leaves identical updated weights everywhere
    and zeroed gradients.  Everything is read back through the reference's own Blob::cpu_data / cpu_diff."""
    import ctypes
    L = refbind.dropin_lib()
    counts = np.array([60 * 5, 5, 2 * 5 * 5, 18], dtype=np.int32)          # W, b, M, B (B ragged: padded to 16 bytes inside)
    total = int(counts.sum())
    rng = np.random.default_rng(world + fused)
    data = rng.uniform(-1, 1, (world, total)).astype(np.float32)
    diff = rng.normal(0, 1e-2, (world, total)).astype(np.float32)
    out_d, out_g = np.empty_like(data), np.empty_like(diff)
    L.mmsref_grad_exchange_run.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int] + [ctypes.c_void_p] * 2
    rc = L.mmsref_grad_exchange_run(world, len(counts), counts.ctypes.data, data.ctypes.data, diff.ctypes.data, fused,
                                    out_d.ctypes.data, out_g.ctypes.data)
    assert rc == 0, L.mmsref_last_error()
    acc = diff[0].copy()
    for r in range(1, world):
        acc = acc + diff[r]                                              # float adds in rank order, as the kernel does
    for r in range(world):
        np.testing.assert_array_equal(out_d[r], out_d[0])               # replicas identical
        if not fused:
            np.testing.assert_array_equal(out_d[r], data[0])            # on_start: rank 0's weights everywhere
            np.testing.assert_array_equal(out_g[r], acc * np.float32(1.0 / world))
        else:
            assert not out_g[r].any()                                   # ClearParamDiffs folded in
    if fused:
        w, g = data[0].copy(), acc.copy()
        cport.adadelta_step(w, g, np.zeros_like(w), np.zeros_like(w), grad_scale=1.0 / world, local_decay=5e-4, momentum=0.95,
                            delta=5e-7, local_rate=1.0)
        assert scaled_err(out_d[0], w) <= 2e-6
