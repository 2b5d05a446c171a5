"""Data formats either side of the path -- CPU tests, no GPU.

* weight_source loaders (csrc/formats.cu behind mms_load_weight_source_*): known-answer files written here, and the
  reference's own loader (EmbedLayer::LayerSetUp, embed_layer.cpp:46-113, compiled in place in oracle/_ref) on the
  same files -- bit-exact for float and for double (whose odd `(float*)` store the loader keeps).
* .caffemodel codec (formats.py): checked against the protobuf runtime itself, fed with descriptors that restate the
  field numbers of the reference's caffe.proto (:6-22, :64-96, :310-330) -- what our writer emits the protobuf library
  parses to the same message, what the protobuf library emits (packed and unpacked) our reader parses to the same arrays.
"""
import ctypes

import numpy as np
import pytest

from mms_answer_selection_b200 import build, formats
from mms_answer_selection_b200.blob import Blob
from oracle import refbind

needs_ref = pytest.mark.skipif(not refbind.ref_available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def L():
    lib = ctypes.CDLL(build.build())
    lib.mms_last_error.restype = ctypes.c_char_p
    for sfx in ("_f32", "_f64"):
        fn = getattr(lib, "mms_load_weight_source" + sfx)
        fn.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong,
                       ctypes.POINTER(ctypes.c_longlong)]
    return lib


def load(L, path, table):
    fn = L.mms_load_weight_source_f32 if table.dtype == np.float32 else L.mms_load_weight_source_f64
    n = ctypes.c_longlong(-1)
    rc = fn(str(path).encode(), ctypes.c_void_p(table.ctypes.data), table.shape[0], table.shape[1], ctypes.byref(n))
    return rc, int(n.value)


def vectors(nwords, dim, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, (nwords, dim)).astype(np.float32), ["w%d" % i for i in range(nwords)]


def write_txt(path, vecs, words):
    with open(path, "w") as f:
        for w, v in zip(words, vecs):
            f.write(w + " " + " ".join("%.9g" % x for x in v) + "\n")


def write_all(path, vecs, words, V, D):
    with open(path, "w") as f:
        f.write("0.5 %d %d\n" % (V - 1, D - 1))
        for i, (w, v) in enumerate(zip(words, vecs)):
            f.write("%d " % i + " ".join("%.9g" % x for x in v) + " " + w + "\n")


def write_bin(path, vecs, words):
    with open(path, "wb") as f:
        f.write(b"%d %d\n" % vecs.shape)
        for w, v in zip(words, vecs):
            f.write(w.encode() + b" " + v.astype("<f4").tobytes() + b"\n")


WRITERS = {"glove.txt": lambda p, v, w, V, D: write_txt(p, v, w),
           "dump.all": write_all,
           "vectors.bin": lambda p, v, w, V, D: write_bin(p, v, w)}


@pytest.mark.parametrize("fname", sorted(WRITERS))
def test_weight_source_known_answer(L, tmp_path, fname):
    V, D, n = 9, 5, 7
    vecs, words = vectors(n, D, 3)
    path = tmp_path / fname
    WRITERS[fname](path, vecs, words, V, D)
    table = np.full((V, D), 0.25, np.float32)
    rc, loaded = load(L, path, table)
    assert rc == 0, L.mms_last_error()
    assert loaded == n
    assert np.array_equal(table[:n], vecs)                      # %.9g round-trips a float exactly
    assert np.all(table[n:] == 0.25)                            # rows the file does not reach keep the filler's values


@needs_ref
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("fname", sorted(WRITERS))
def test_weight_source_matches_reference_loader(L, tmp_path, fname, dtype):
    """The reference's EmbedLayer::LayerSetUp on the same file: identical tables, bit for bit -- including the double
    instantiation, where the reference stores each value as a float into the first bytes of the double."""
    V, D, n = 23, 8, 21
    vecs, words = vectors(n, D, 5)
    path = tmp_path / fname
    WRITERS[fname](path, vecs, words, V, D)
    ref = refbind.RefLayer("Embed", [np.zeros((2, 3))], {
        "num_output": D, "input_dim": V, "embed.bias_term": 0, "weight_filler.type": "constant",
        "weight_filler.value": 0.25, "weight_source": str(path)}, dtype=dtype)
    want = ref.read("blob", 0)
    table = np.full((V, D), 0.25, dtype)
    rc, loaded = load(L, path, table)
    assert rc == 0 and loaded == n
    assert table.tobytes() == want.tobytes()
    if dtype == np.float32:
        assert np.array_equal(table[:n], vecs)


def test_weight_source_errors_are_loud(L, tmp_path):
    V, D = 4, 3
    table = np.zeros((V, D), np.float32)
    rc, _ = load(L, tmp_path / "missing.txt", table)
    assert rc == -1 and b"missing.txt" in L.mms_last_error()
    vecs, words = vectors(6, D, 1)                                # more records than rows: the reference overruns
    write_txt(tmp_path / "long.txt", vecs, words)
    rc, _ = load(L, tmp_path / "long.txt", table)
    assert rc == -1 and b"more than input_dim" in L.mms_last_error()
    write_all(tmp_path / "bad.all", vecs[:2], words[:2], V + 1, D)  # header must say (V-1, D-1)   embed_layer.cpp:67-68
    rc, _ = load(L, tmp_path / "bad.all", table)
    assert rc == -1 and b"header says" in L.mms_last_error()
    write_bin(tmp_path / "dim.bin", np.zeros((2, D + 1), np.float32), words[:2])   # CHECK_EQ(dim_t, N_)  :85
    rc, _ = load(L, tmp_path / "dim.bin", table)
    assert rc == -1 and b"dimensions" in L.mms_last_error()
    with open(tmp_path / "short.txt", "w") as f:
        f.write("w0 1 2 3\nw1 4 5\n")
    rc, _ = load(L, tmp_path / "short.txt", table)
    assert rc == -1 and b"expected 3 values" in L.mms_last_error()


# ------------------------------------------------------------------------------------------- .caffemodel
def caffe_messages(packed=True):
    """NetParameter / LayerParameter / BlobProto / BlobShape message classes built at run time from the field numbers
    of the reference's caffe.proto (no protoc in this image)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="caffe_subset_%d.proto" % packed, package="caffe%d" % packed, syntax="proto2")
    pk = "caffe%d." % packed

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, ftype, label, tname, pack in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if tname:
                f.type_name = "." + pk + tname
            if pack is not None:
                f.options.packed = pack
    OPT, REP = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    msg("BlobShape", [("dim", 1, F.TYPE_INT64, REP, None, packed)])                                 # caffe.proto:6-8
    msg("BlobProto", [("shape", 7, F.TYPE_MESSAGE, OPT, "BlobShape", None),                          # :10-22
                      ("data", 5, F.TYPE_FLOAT, REP, None, packed), ("diff", 6, F.TYPE_FLOAT, REP, None, packed),
                      ("double_data", 8, F.TYPE_DOUBLE, REP, None, packed),
                      ("double_diff", 9, F.TYPE_DOUBLE, REP, None, packed),
                      ("num", 1, F.TYPE_INT32, OPT, None, None), ("channels", 2, F.TYPE_INT32, OPT, None, None),
                      ("height", 3, F.TYPE_INT32, OPT, None, None), ("width", 4, F.TYPE_INT32, OPT, None, None)])
    msg("LayerParameter", [("name", 1, F.TYPE_STRING, OPT, None, None), ("type", 2, F.TYPE_STRING, OPT, None, None),  # :310-330
                           ("bottom", 3, F.TYPE_STRING, REP, None, None), ("top", 4, F.TYPE_STRING, REP, None, None),
                           ("phase", 10, F.TYPE_INT32, OPT, None, None), ("loss_weight", 5, F.TYPE_FLOAT, REP, None, None),
                           ("blobs", 7, F.TYPE_MESSAGE, REP, "BlobProto", None)])
    msg("NetParameter", [("name", 1, F.TYPE_STRING, OPT, None, None), ("force_backward", 5, F.TYPE_BOOL, OPT, None, None),  # :64-96
                         ("layer", 100, F.TYPE_MESSAGE, REP, "LayerParameter", None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName(pk + n))
    return get("NetParameter"), get("BlobProto")


def reference_net(Net, rng):
    net = Net(name="trec_qa", force_backward=True)
    arrays = {}
    l = net.layer.add(name="embed_q", type="Embed", bottom=["question"], top=["q"], phase=0)
    W, b = rng.uniform(-0.08, 0.08, (11, 6)).astype(np.float32), rng.standard_normal(6).astype(np.float32)
    for arr in (W, b):
        bp = l.blobs.add()
        bp.shape.dim.extend(arr.shape)
        bp.data.extend(arr.reshape(-1).tolist())
    arrays["embed_q"] = [W, b]
    l = net.layer.add(name="sim_cross", type="SimCross", bottom=["q", "a"], top=["S"], loss_weight=[1.0])
    M = rng.uniform(-0.1, 0.1, (2, 6, 6))                      # a double snapshot: double_data
    bp = l.blobs.add()
    bp.shape.dim.extend(M.shape)
    bp.double_data.extend(M.reshape(-1).tolist())
    bp.double_diff.extend((M * 2).reshape(-1).tolist())
    arrays["sim_cross"] = [M]
    l = net.layer.add(name="fc", type="InnerProduct")           # a legacy 4-D blob (num, channels, height, width)
    w = rng.standard_normal((3, 4)).astype(np.float32)
    bp = l.blobs.add(num=1, channels=1, height=3, width=4)
    bp.data.extend(w.reshape(-1).tolist())
    arrays["fc"] = [w]
    net.layer.add(name="relu", type="ReLU")                     # a layer without blobs
    return net, arrays


@pytest.mark.parametrize("packed", [True, False])
def test_caffemodel_reader_against_protobuf(packed):
    Net, _ = caffe_messages(packed)
    net, arrays = reference_net(Net, np.random.default_rng(7))
    got = formats.NetProto.parse(net.SerializeToString())
    assert got.name == "trec_qa"
    assert [(l.name, l.type, len(l.blobs)) for l in got.layers] == \
        [("embed_q", "Embed", 2), ("sim_cross", "SimCross", 1), ("fc", "InnerProduct", 1), ("relu", "ReLU", 0)]
    eq, sc, fc = got.layers[0], got.layers[1], got.layers[2]
    assert eq.blobs[0].blob_shape() == (11, 6) and eq.blobs[1].blob_shape() == (6,)
    assert np.array_equal(eq.blobs[0].values(np.float32).reshape(11, 6), arrays["embed_q"][0])
    assert np.array_equal(eq.blobs[1].values(np.float32), arrays["embed_q"][1])
    assert np.array_equal(sc.blobs[0].values(np.float64).reshape(2, 6, 6), arrays["sim_cross"][0])      # double_data wins
    assert np.array_equal(sc.blobs[0].values(np.float64, diff=True).reshape(2, 6, 6), arrays["sim_cross"][0] * 2)
    assert fc.blobs[0].blob_shape() == (1, 1, 3, 4)
    assert fc.blobs[0].shape_equals((3, 4)) and fc.blobs[0].shape_equals((1, 3, 4))                       # blob.cpp:401-405
    assert not fc.blobs[0].shape_equals((4, 3)) and not fc.blobs[0].shape_equals((1, 1, 1, 3, 4))
    assert eq.blobs[0].shape_equals((11, 6)) and not eq.blobs[0].shape_equals((1, 11, 6))


def test_caffemodel_writer_parses_with_protobuf_and_keeps_unknown_fields():
    Net, _ = caffe_messages(True)
    net, _ = reference_net(Net, np.random.default_rng(9))
    wire = net.SerializeToString()
    ours = formats.NetProto.parse(wire)
    # read-modify-write: replace M, leave everything else (bottom/top/phase/loss_weight/force_backward) alone
    newM = np.arange(72, dtype=np.float64).reshape(2, 6, 6)
    ours.layers[1].blobs[0] = formats.BlobProto.from_array(newM)
    back = Net()
    back.ParseFromString(ours.serialize())
    assert back.name == "trec_qa" and back.force_backward is True
    assert list(back.layer[0].bottom) == ["question"] and list(back.layer[0].top) == ["q"] and back.layer[0].phase == 0
    assert list(back.layer[1].loss_weight) == [1.0]
    assert list(back.layer[1].blobs[0].shape.dim) == [2, 6, 6]
    assert np.array_equal(np.array(back.layer[1].blobs[0].double_data), newM.reshape(-1))
    assert len(back.layer[1].blobs[0].data) == 0 and len(back.layer[1].blobs[0].double_diff) == 0
    assert back.layer[0] == net.layer[0] and back.layer[2] == net.layer[2] and back.layer[3] == net.layer[3]
    # Blob<float>::ToProto writes `data`, Blob<double>::ToProto writes `double_data` (blob.cpp:494-534)
    f = formats.BlobProto.from_array(np.ones((2, 3), np.float32), diff=np.full((2, 3), 2.0, np.float32))
    _, Blob_ = caffe_messages(True)
    bp = Blob_()
    bp.ParseFromString(f.serialize())
    assert list(bp.shape.dim) == [2, 3] and list(bp.data) == [1.0] * 6 and list(bp.diff) == [2.0] * 6
    assert not bp.HasField("num") and len(bp.double_data) == 0
    # a 0-axis blob (the scalar tops) keeps an empty shape message
    s = formats.BlobProto.parse(formats.BlobProto.from_array(np.zeros((), np.float32)).serialize())
    assert s.shape == () and s.blob_shape() == ()


class _FakeParam(object):
    def __init__(self, name):
        self.name = name


class _FakeLayer(object):
    """Stands in for a Layer of the host mirror (those need a GPU): a name, a type and blobs."""

    def __init__(self, name, type_, shapes, dtype=np.float32):
        self.layer_param_ = _FakeParam(name)
        self._type = type_
        self.blobs = [Blob(s, dtype=dtype, device="cpu") for s in shapes]

    def type(self):
        return self._type


def test_copy_trained_layers_from_follows_the_reference_rules(tmp_path):
    from mms_answer_selection_b200.layers import CheckError
    Net, _ = caffe_messages(True)
    net, arrays = reference_net(Net, np.random.default_rng(11))
    path = str(tmp_path / "snap.caffemodel")
    with open(path, "wb") as f:
        f.write(net.SerializeToString())
    target = [_FakeLayer("embed_q", "Embed", [(11, 6), (6,)]), _FakeLayer("other", "FM", [(1,)]),
              _FakeLayer("sim_cross", "SimCross", [(2, 6, 6)])]
    copied = formats.copy_trained_layers_from(target, path)
    assert copied == ["embed_q", "sim_cross"]                   # "fc" and "relu" have no namesake: ignored (net.cpp:751-754)
    assert np.array_equal(target[0].blobs[0].cpu_data(), arrays["embed_q"][0])
    assert np.array_equal(target[0].blobs[1].cpu_data(), arrays["embed_q"][1])
    assert np.array_equal(target[2].blobs[0].cpu_data(), arrays["sim_cross"][0].astype(np.float32))   # double -> float
    assert np.array_equal(target[2].blobs[0].cpu_diff(), (arrays["sim_cross"][0] * 2).astype(np.float32))
    with pytest.raises(CheckError, match="shape mismatch"):
        formats.copy_trained_layers_from([_FakeLayer("sim_cross", "SimCross", [(2, 6, 5)])], path)
    with pytest.raises(CheckError, match="Incompatible number of blobs"):
        formats.copy_trained_layers_from([_FakeLayer("embed_q", "Embed", [(11, 6)])], path)
    # Net::ToProto -> file -> CopyTrainedLayersFrom round trip of the mirror itself
    snap = str(tmp_path / "mine.caffemodel")
    formats.write_caffemodel(snap, formats.net_to_proto(target, name="mms"))
    fresh = [_FakeLayer("embed_q", "Embed", [(11, 6), (6,)]), _FakeLayer("sim_cross", "SimCross", [(2, 6, 6)])]
    formats.copy_trained_layers_from(fresh, snap)
    assert np.array_equal(fresh[0].blobs[0].cpu_data(), target[0].blobs[0].cpu_data())
    assert np.array_equal(fresh[1].blobs[0].cpu_data(), target[2].blobs[0].cpu_data())
    back = Net()
    back.ParseFromString(open(snap, "rb").read())
    assert [l.name for l in back.layer] == ["embed_q", "other", "sim_cross"] and back.layer[2].type == "SimCross"


def test_v1_layers_are_refused():
    with pytest.raises(ValueError, match="V1LayerParameter"):
        formats.NetProto.parse(bytes([0x12, 0x00]))             # field 2 (`layers`), empty message


def test_caffemodel_random_round_trips():
    """Random nets (shapes incl. 0-axis and 5-axis blobs, float and double blobs, with and without diffs) through our
    writer -> protobuf parser and protobuf writer -> our parser."""
    Net, _ = caffe_messages(True)
    rng = np.random.default_rng(99)
    for _ in range(25):
        layers = []
        for li in range(int(rng.integers(1, 5))):
            blobs = []
            for _b in range(int(rng.integers(0, 4))):
                shape = tuple(int(d) for d in rng.integers(1, 5, int(rng.integers(0, 6))))
                dt = np.float64 if rng.uniform() < 0.3 else np.float32
                data = rng.standard_normal(shape).astype(dt)
                diff = rng.standard_normal(shape).astype(dt) if rng.uniform() < 0.5 else None
                blobs.append((data, diff))
            layers.append(("layer%d" % li, "Type%d" % li, blobs))
        ours = formats.NetProto("net", [formats.LayerProto(n, t, [formats.BlobProto.from_array(d, g) for d, g in bl])
                                        for n, t, bl in layers])
        back = Net()
        back.ParseFromString(ours.serialize())
        again = formats.NetProto.parse(back.SerializeToString())
        assert [l.name for l in back.layer] == [n for n, _, _ in layers]
        for (n, t, bl), lp, lo in zip(layers, back.layer, again.layers):
            assert lp.type == t and len(lp.blobs) == len(bl) == len(lo.blobs)
            for (data, diff), bp, bo in zip(bl, lp.blobs, lo.blobs):
                assert tuple(bp.shape.dim) == data.shape == bo.blob_shape()
                field = bp.double_data if data.dtype == np.float64 else bp.data
                assert np.array_equal(np.array(field, data.dtype).reshape(data.shape), data)
                assert np.array_equal(bo.values(data.dtype).reshape(data.shape), data)
                if diff is not None:
                    assert np.array_equal(bo.values(data.dtype, diff=True).reshape(data.shape), diff)
                else:
                    assert len(bo.diff) == 0 and len(bo.double_diff) == 0
