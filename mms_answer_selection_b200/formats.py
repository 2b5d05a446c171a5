"""Data formats either side of the path, host side.

* ``load_weight_source``: ``embed_param.weight_source`` -- the pre-trained word-vector files EmbedLayer::LayerSetUp
  reads (reference src/caffe/layers/embed_layer.cpp:46-113).  The parser is C code of libmms_b200.so
  (``mms_load_weight_source_*``, csrc/formats.cu), shared with the C++ drop-in layer.
* ``.caffemodel``: the binary NetParameter a reference snapshot is (``Net::ToProto`` net.cpp:847-856 ->
  ``WriteProtoToBinaryFile``; read back by ``Net::CopyTrainedLayersFrom`` net.cpp:741-776 through
  ``Blob::FromProto`` / ``ToProto`` blob.cpp:447-536).  In a Caffe build with the drop-in layers Caffe itself does this
  (the layers' ``blobs_`` are ordinary Blobs); this module is for the Python host mirror, so that ``M`` / ``B`` /
  the embedding table of a reference snapshot flow in and out.  It is a protobuf *wire-format* codec for exactly the
  fields involved (caffe.proto:6-22 BlobShape / BlobProto, :64-96 NetParameter, :310-330 LayerParameter); every other
  field is skipped on read (its bytes are kept and written back on a read-modify-write).

HDF5 batches (hdf5_data_layer.cpp:27-69) and ``.caffemodel.h5`` are not handled: neither libhdf5 nor h5py exists in
this image, so nothing could be checked against a real file.
"""
import ctypes

import numpy as np

from ._lib import check, lib


# --------------------------------------------------------------------------------------- weight_source
def load_weight_source(path, blob):
    """Overwrite the first rows of ``blob`` (the (input_dim, num_output) table, already filled by the weight filler)
    with the vectors in ``path``; returns the number of records read.  embed_layer.cpp:46-113."""
    rows, dim = blob.shape
    table = np.ascontiguousarray(blob.cpu_data())
    fn = lib().mms_load_weight_source_f32 if blob.dtype == np.float32 else lib().mms_load_weight_source_f64
    loaded = ctypes.c_longlong(0)
    check(fn(str(path).encode(), ctypes.c_void_p(table.ctypes.data), rows, dim, ctypes.byref(loaded)))
    blob.set_cpu_data(table)
    return int(loaded.value)


# --------------------------------------------------------------------------------------- protobuf wire format
_VARINT, _FIXED64, _BYTES, _FIXED32 = 0, 1, 2, 5


def _read_varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("malformed varint")


def _write_varint(out, v):
    if v < 0:
        v += 1 << 64                      # int32 / int64 fields: two's complement, 10 bytes
    while v > 0x7F:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _fields(buf):
    """Yield (field number, wire type, value) of one message; value is an int (varint), a memoryview (bytes,
    fixed32, fixed64)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == _VARINT:
            val, pos = _read_varint(buf, pos)
        elif wt == _BYTES:
            ln, pos = _read_varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
            if len(val) != ln:
                raise ValueError("truncated field %d" % num)
        elif wt == _FIXED32:
            val, pos = buf[pos:pos + 4], pos + 4
        elif wt == _FIXED64:
            val, pos = buf[pos:pos + 8], pos + 8
        else:
            raise ValueError("unsupported wire type %d (groups are not used by caffe.proto)" % wt)
        yield num, wt, val


def _emit(out, num, wt, payload):
    _write_varint(out, (num << 3) | wt)
    if wt == _VARINT:
        _write_varint(out, payload)
    elif wt == _BYTES:
        _write_varint(out, len(payload))
        out += payload
    else:
        out += payload


def _signed(v, bits):
    v &= (1 << 64) - 1
    if bits == 32:
        v &= 0xFFFFFFFF
    return v - (1 << bits) if v >> (bits - 1) else v


class _Repeated(object):
    """Accumulates one repeated numeric field given packed and/or unpacked occurrences (parsers must take both)."""

    def __init__(self, dtype):
        self.dtype, self.parts = np.dtype(dtype), []

    def add(self, wt, val):
        if wt == _BYTES:
            self.parts.append(np.frombuffer(val, dtype=self.dtype))
        else:
            self.parts.append(np.frombuffer(bytes(val), dtype=self.dtype))

    def array(self):
        if not self.parts:
            return np.zeros(0, self.dtype)
        return self.parts[0] if len(self.parts) == 1 else np.concatenate(self.parts)


class BlobProto(object):
    """caffe.proto:10-22.  ``shape`` None means "no shape message"; the legacy num/channels/height/width are None
    when absent (``has_num()`` etc. decide how Blob::FromProto reads the shape, blob.cpp:449-464)."""

    def __init__(self):
        self.shape = None
        self.legacy = [None, None, None, None]
        self.data = np.zeros(0, "<f4")
        self.diff = np.zeros(0, "<f4")
        self.double_data = np.zeros(0, "<f8")
        self.double_diff = np.zeros(0, "<f8")

    @classmethod
    def parse(cls, buf):
        self = cls()
        rep = {5: _Repeated("<f4"), 6: _Repeated("<f4"), 8: _Repeated("<f8"), 9: _Repeated("<f8")}
        for num, wt, val in _fields(buf):
            if num in rep:
                rep[num].add(wt, val)
            elif num == 7 and wt == _BYTES:                      # BlobShape { repeated int64 dim = 1 [packed] }
                dims = [] if self.shape is None else list(self.shape)     # a repeated occurrence merges
                for n2, wt2, v2 in _fields(val):
                    if n2 != 1:
                        continue
                    if wt2 == _BYTES:
                        p = 0
                        while p < len(v2):
                            d, p = _read_varint(v2, p)
                            dims.append(_signed(d, 64))
                    else:
                        dims.append(_signed(v2, 64))
                self.shape = tuple(dims)
            elif 1 <= num <= 4 and wt == _VARINT:
                self.legacy[num - 1] = _signed(val, 32)
        self.data, self.diff = rep[5].array(), rep[6].array()
        self.double_data, self.double_diff = rep[8].array(), rep[9].array()
        return self

    def serialize(self):
        out = bytearray()
        for i, v in enumerate(self.legacy):
            if v is not None:
                _emit(out, i + 1, _VARINT, int(v))
        for num, arr, dt in ((5, self.data, "<f4"), (6, self.diff, "<f4")):
            if len(arr):
                _emit(out, num, _BYTES, np.ascontiguousarray(arr, dtype=dt).tobytes())
        if self.shape is not None:
            dims = bytearray()
            for d in self.shape:
                _write_varint(dims, int(d))
            body = bytearray()
            if len(dims):
                _emit(body, 1, _BYTES, dims)
            _emit(out, 7, _BYTES, body)
        for num, arr, dt in ((8, self.double_data, "<f8"), (9, self.double_diff, "<f8")):
            if len(arr):
                _emit(out, num, _BYTES, np.ascontiguousarray(arr, dtype=dt).tobytes())
        return bytes(out)

    # -- Blob::FromProto / ShapeEquals / ToProto --------------------------------------------------
    def blob_shape(self):
        """The shape Blob::FromProto(reshape=true) gives the blob (blob.cpp:449-464)."""
        if any(v is not None for v in self.legacy):
            return tuple(int(v or 0) for v in self.legacy)
        return tuple(self.shape or ())

    def shape_equals(self, shape):
        """Blob::ShapeEquals (blob.cpp:392-412): a legacy 4-D proto matches a blob of <= 4 axes whose shape,
        left-padded with ones, equals (num, channels, height, width)."""
        shape = tuple(int(s) for s in shape)
        if any(v is not None for v in self.legacy):
            if len(shape) > 4:
                return False
            padded = (1,) * (4 - len(shape)) + shape
            return padded == tuple(int(v or 0) for v in self.legacy)
        return shape == tuple(self.shape or ())

    def values(self, dtype, diff=False):
        """The array Blob::FromProto copies (blob.cpp:468-492): double_data wins over data when present."""
        dbl, flt = (self.double_diff, self.diff) if diff else (self.double_data, self.data)
        return (dbl if len(dbl) else flt).astype(dtype)

    @classmethod
    def from_array(cls, data, diff=None):
        """Blob<Dtype>::ToProto (blob.cpp:494-534): shape message, float blobs -> data/diff, double blobs ->
        double_data/double_diff."""
        self = cls()
        data = np.asarray(data)
        self.shape = tuple(data.shape)
        if data.dtype == np.float64:
            self.double_data = data.reshape(-1)
            if diff is not None:
                self.double_diff = np.asarray(diff, np.float64).reshape(-1)
        else:
            self.data = data.astype(np.float32).reshape(-1)
            if diff is not None:
                self.diff = np.asarray(diff, np.float32).reshape(-1)
        return self


class LayerProto(object):
    """LayerParameter, caffe.proto:310-330: name = 1, type = 2, blobs = 7; everything else is carried as raw bytes."""

    def __init__(self, name="", type="", blobs=()):
        self.name, self.type, self.blobs, self.other = name, type, list(blobs), []

    @classmethod
    def parse(cls, buf):
        self = cls()
        for num, wt, val in _fields(buf):
            if num == 1 and wt == _BYTES:
                self.name = bytes(val).decode("utf-8", "replace")
            elif num == 2 and wt == _BYTES:
                self.type = bytes(val).decode("utf-8", "replace")
            elif num == 7 and wt == _BYTES:
                self.blobs.append(BlobProto.parse(val))
            else:
                self.other.append((num, wt, val if wt == _VARINT else bytes(val)))
        return self

    def serialize(self):
        out = bytearray()
        if self.name:
            _emit(out, 1, _BYTES, self.name.encode())
        if self.type:
            _emit(out, 2, _BYTES, self.type.encode())
        for num, wt, val in self.other:
            _emit(out, num, wt, val)
        for b in self.blobs:
            _emit(out, 7, _BYTES, b.serialize())
        return bytes(out)


class NetProto(object):
    """NetParameter, caffe.proto:64-96: name = 1, layer = 100 (V1 ``layers`` = 2 is refused: the reference upgrades
    those files with upgrade_proto.cpp before use)."""

    def __init__(self, name="", layers=()):
        self.name, self.layers, self.other = name, list(layers), []

    @classmethod
    def parse(cls, buf):
        self = cls()
        for num, wt, val in _fields(memoryview(buf)):
            if num == 1 and wt == _BYTES:
                self.name = bytes(val).decode("utf-8", "replace")
            elif num == 100 and wt == _BYTES:
                self.layers.append(LayerProto.parse(val))
            elif num == 2:
                raise ValueError("V1LayerParameter `layers` found: upgrade the file with the reference's "
                                 "upgrade_net_proto_binary first")
            else:
                self.other.append((num, wt, val if wt == _VARINT else bytes(val)))
        return self

    def serialize(self):
        out = bytearray()
        if self.name:
            _emit(out, 1, _BYTES, self.name.encode())
        for num, wt, val in self.other:
            _emit(out, num, wt, val)
        for l in self.layers:
            _emit(out, 100, _BYTES, l.serialize())
        return bytes(out)


def read_caffemodel(path):
    with open(path, "rb") as f:
        return NetProto.parse(f.read())


def write_caffemodel(path, net_proto):
    with open(path, "wb") as f:
        f.write(net_proto.serialize())


# --------------------------------------------------------------------------------------- Net-level helpers
def copy_trained_layers_from(layers, source):
    """Net::CopyTrainedLayersFrom (net.cpp:741-776).  ``layers``: the target net's layer objects in net order (each
    with ``layer_param_.name`` and ``blobs``); ``source``: a NetProto or a ``.caffemodel`` path.  Source layers
    with no namesake are ignored; a blob-count or shape mismatch is an error; returns the copied layer names."""
    from .layers import CheckError
    net = read_caffemodel(source) if isinstance(source, str) else source
    by_name = {}
    for l in layers:
        by_name.setdefault(l.layer_param_.name, l)               # first match wins (:747-750)
    copied = []
    for src in net.layers:
        tgt = by_name.get(src.name)
        if tgt is None:
            continue                                            # "Ignoring source layer"
        if len(tgt.blobs) != len(src.blobs):
            raise CheckError("Incompatible number of blobs for layer %s" % src.name)
        for j, (tb, sb) in enumerate(zip(tgt.blobs, src.blobs)):
            if not sb.shape_equals(tb.shape):
                raise CheckError("Cannot copy param %d weights from layer '%s'; shape mismatch.  Source param shape "
                                 "is %s; target param shape is %s." % (j, src.name, sb.blob_shape(), tuple(tb.shape)))
            vals = sb.values(tb.dtype)
            if vals.size != tb.count():
                raise CheckError("Check failed: count_ == proto.data_size() (%d vs. %d)" % (tb.count(), vals.size))
            tb.set_cpu_data(vals.reshape(tb.shape))
            if len(sb.double_diff) or len(sb.diff):             # blob.cpp:481-492
                tb.set_cpu_diff(sb.values(tb.dtype, diff=True).reshape(tb.shape))
        copied.append(src.name)
    return copied


def net_to_proto(layers, name="", write_diff=False):
    """Net::ToProto (net.cpp:847-856) for the layers of the host mirror: one LayerParameter per layer with its name,
    type and blobs (shared blobs are written by every layer that holds them, as in the reference)."""
    out = NetProto(name)
    for l in layers:
        blobs = [BlobProto.from_array(b.cpu_data(), b.cpu_diff() if write_diff else None) for b in l.blobs]
        out.layers.append(LayerProto(l.layer_param_.name, l.type(), blobs))
    return out
