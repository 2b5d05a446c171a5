"""Host-side mirror of the reference's Caffe layers for the MMS hot path.

Same names, argument meaning and error behaviour as the reference classes
(paths relative to the reference tree):

  EmbedLayer         src/caffe/layers/embed_layer.{cpp,cu}
  SimCrossLayer      src/caffe/layers/sim_cross_layer.{cpp,cu}
  SimMatrixLayer     src/caffe/layers/sim_matrix_layer.{cpp,cu}
  PairRankLossLayer  src/caffe/layers/pair_rank_loss_layer.{cpp,cu}
  FMLayer            src/caffe/layers/fm_layer.{cpp,cu}

driven through the Layer API of include/caffe/layer.hpp: ``SetUp`` (= CheckBlobCounts,
LayerSetUp, Reshape, SetLossWeights, :67-77), ``Forward`` (= Reshape, Forward_gpu, loss
reduction, :451-487), ``Backward`` (:490-500).  Forward_gpu / Backward_gpu only marshal raw
device pointers and sizes into the C-ABI of libmms_b200.so -- exactly what the C++ drop-in
classes in caffe_layers/ do.  There is no Forward_cpu: Caffe::mode() is always GPU here.

A failed CHECK aborts the process in the reference (glog LOG(FATAL)); here it raises
``CheckError`` carrying the reference's message.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import Handle, c_p, check, lib
from .blob import Blob


class CheckError(RuntimeError):
    """A Caffe CHECK / LOG(FATAL) of the reference layer would have fired."""


def _check(cond, msg):
    if not cond:
        raise CheckError("Check failed: " + msg)


# ------------------------------------------------------------------------- parameters
class FillerParameter(dict):
    """caffe.proto:41-62.  Default: type 'constant', value 0."""

    def __init__(self, **kw):
        dict.__init__(self, type="constant", value=0.0, min=0.0, max=1.0, mean=0.0, std=1.0)
        self.update(kw)


_DEFAULTS = {
    # caffe.proto:471-477 (the reference's spelling: mesure_count)
    "sim_cross_param": dict(dist_mode=1, mesure_count=1, weight_filler=None, bias_term=True, bias_filler=None),
    "sim_matrix_param": dict(weight_filler=None),                       # caffe.proto:430-432
    "pair_rank_loss_param": dict(margin=1.0),                           # caffe.proto:479-481
    "fm_param": dict(bias_term=True),                                   # caffe.proto:418-420
    "embed_param": dict(num_output=0, input_dim=0, bias_term=True, weight_filler=None,   # :790-803
                        bias_filler=None, weight_source=""),
    "convolution_param": dict(num_output=0, bias_term=True, kernel_h=0, kernel_w=0, kernel_size=0, stride=1, pad=0,
                              group=1, weight_filler=None, bias_filler=None),   # caffe.proto ConvolutionParameter
    "pooling_param": dict(pool="MAX", kernel_h=0, kernel_w=0, kernel_size=0, stride=1, stride_h=0, stride_w=0,
                          pad=0, pad_h=0, pad_w=0, global_pooling=False),       # caffe.proto PoolingParameter
    "dropout_param": dict(dropout_ratio=0.5, seed=1701),                 # caffe.proto DropoutParameter (+ the mask seed)
    "bn_param": dict(bn_memory=0.9, scale_filler=None, shift_filler=None),      # caffe.proto:484-488
    "map_param": dict(fixed_axis=1),                                    # caffe.proto:422-424
    "mrr_param": dict(fixed_axis=1),                                    # caffe.proto:426-428
    "auc_param": dict(fixed_axis=1, axis=1, ignore_label=None),         # caffe.proto:465-469
}


class LayerParameter(object):
    """The subset of caffe.proto's LayerParameter (:310-416) the MMS layers read."""

    def __init__(self, type, name="", phase="TRAIN", loss_weight=(), dtype=np.float32, **params):
        self.type = type
        self.name = name or type
        self.phase = phase
        self.loss_weight = list(loss_weight)
        self.dtype = np.dtype(dtype)
        for key, dflt in _DEFAULTS.items():
            vals = dict(dflt)
            given = params.pop(key, {})
            unknown = set(given) - set(vals)
            if unknown:
                raise KeyError("unknown field(s) %s in %s" % (sorted(unknown), key))
            vals.update(given)
            for f in ("weight_filler", "bias_filler", "scale_filler", "shift_filler"):
                if f in vals:
                    vals[f] = FillerParameter(**(vals[f] or {}))
            setattr(self, key, vals)
        if params:
            raise KeyError("unknown LayerParameter field(s): %s" % sorted(params))


def _fill(blob, filler, rng):
    """include/caffe/filler.hpp: constant / uniform / gaussian / xavier (fan_in)."""
    shape, t = blob.shape, filler["type"]
    if t == "constant":
        arr = np.full(shape, filler["value"], dtype=blob.dtype)
    elif t == "uniform":
        arr = rng.uniform(filler["min"], filler["max"], size=shape).astype(blob.dtype)
    elif t == "gaussian":
        arr = rng.normal(filler["mean"], filler["std"], size=shape).astype(blob.dtype)
    elif t == "xavier":
        fan_in = blob.count() // max(shape[0], 1)
        scale = np.sqrt(3.0 / fan_in)
        arr = rng.uniform(-scale, scale, size=shape).astype(blob.dtype)
    else:
        raise CheckError("Unknown filler name: %s" % t)
    blob.set_cpu_data(arr)


# ------------------------------------------------------------------------- base class
class Layer(object):
    exact_num_bottom = None
    exact_num_top = 1

    def __init__(self, param):
        self.layer_param_ = param
        self.blobs_ = []
        self.param_propagate_down_ = []
        self.loss_ = []
        self.dtype = param.dtype
        self.sfx = "_f32" if self.dtype == np.float32 else "_f64"
        self.real = ctypes.c_float if self.dtype == np.float32 else ctypes.c_double
        self.handle = Handle()          # raises without a B200 / without the built library
        self.rng = np.random.default_rng(1701)
        # When set, Forward leaves the loss in `loss_dev_` (a device scalar per loss top) instead of
        # reading it back -- needed to record a step into a CUDA graph (no host sync while capturing).
        self.defer_loss_ = False
        self.loss_dev_ = {}

    # -- Caffe public API -------------------------------------------------------
    def type(self):
        return self.layer_param_.type

    @property
    def blobs(self):
        return self.blobs_

    def SetUp(self, bottom, top):
        self.CheckBlobCounts(bottom, top)
        self.LayerSetUp(bottom, top)
        self.Reshape(bottom, top)
        self.SetLossWeights(top)

    def CheckBlobCounts(self, bottom, top):
        if self.exact_num_bottom is not None:
            _check(len(bottom) == self.exact_num_bottom, "%s Layer takes %d bottom blob(s) as input."
                   % (self.type(), self.exact_num_bottom))
        if self.exact_num_top is not None:
            _check(len(top) == self.exact_num_top, "%s Layer produces %d top blob(s) as output."
                   % (self.type(), self.exact_num_top))

    def SetLossWeights(self, top):
        lw = self.layer_param_.loss_weight
        self.loss_ = [0.0] * len(top)
        if lw:
            _check(len(lw) == len(top), "loss_weight must be unspecified or specified once per top blob.")
            for i, w in enumerate(lw):
                if w == 0:
                    continue
                self.loss_[i] = float(w)
                top[i].diff.fill_(float(w))          # layer.hpp:414-428

    def Forward(self, bottom, top):
        self.Reshape(bottom, top)                    # layer.hpp:456
        self._bind_stream()
        self.Forward_gpu(bottom, top)
        return self.ForwardLoss(top)

    def ForwardLoss(self, top):
        """The loss part of Layer::Forward: sum over loss tops of dot(top.data, top.diff)."""
        self._bind_stream()
        loss = 0.0
        for i, t in enumerate(top):                  # layer.hpp:471-479
            if i < len(self.loss_) and self.loss_[i]:
                out = self.loss_dev_.get(i)
                if out is None or out.dtype != t.data.dtype:
                    out = self.loss_dev_[i] = torch.zeros(1, dtype=t.data.dtype, device=t.data.device)
                self._call("mms_dot", c_p(t.gpu_data()), c_p(t.gpu_diff()), t.count(), c_p(out.data_ptr()))
                if not self.defer_loss_:
                    loss += float(out.item())
        return loss

    def Backward(self, top, propagate_down, bottom):
        self._bind_stream()
        self.Backward_gpu(top, list(propagate_down), bottom)

    # -- helpers ----------------------------------------------------------------
    def _bind_stream(self):
        self.handle.set_stream(torch.cuda.current_stream().cuda_stream)

    def _call(self, name, *args):
        check(getattr(lib(), name + self.sfx)(self.handle.ptr, *args))

    def _new_blob(self, shape, like):
        return Blob(shape, dtype=self.dtype, device=like.device)

    def set_math(self, mode):
        self.handle.set_option(_lib.MMS_OPT_MATH, mode)


def _p(blob_or_none, diff=False):
    if blob_or_none is None:
        return c_p(0)
    return c_p(blob_or_none.gpu_diff() if diff else blob_or_none.gpu_data())


# ------------------------------------------------------------------------- Embed
class EmbedLayer(Layer):
    exact_num_bottom = 1

    def LayerSetUp(self, bottom, top):
        ep = self.layer_param_.embed_param
        self.N_ = int(ep["num_output"])
        _check(self.N_ > 0, "EmbedLayer num_output must be positive.")
        self.K_ = int(ep["input_dim"])
        _check(self.K_ > 0, "EmbedLayer input_dim must be positive.")
        self.bias_term_ = bool(ep["bias_term"])
        if not self.blobs_:                                   # embed_layer.cpp:18-44
            self.blobs_.append(self._new_blob((self.K_, self.N_), bottom[0]))
            _fill(self.blobs_[0], ep["weight_filler"], self.rng)
            if self.bias_term_:
                self.blobs_.append(self._new_blob((self.N_,), bottom[0]))
                _fill(self.blobs_[1], ep["bias_filler"], self.rng)
            if ep["weight_source"]:
                from .formats import load_weight_source      # embed_layer.cpp:46-113
                load_weight_source(ep["weight_source"], self.blobs_[0])
        self.param_propagate_down_ = [True] * len(self.blobs_)

    def Reshape(self, bottom, top):
        self.M_ = bottom[0].count()
        top[0].Reshape(tuple(bottom[0].shape) + (self.N_,))   # embed_layer.cpp:121-125

    def Forward_gpu(self, bottom, top):
        bias = self.blobs_[1] if self.bias_term_ else None
        self._call("mms_embed_forward", _p(bottom[0]), _p(self.blobs_[0]), _p(bias), _p(top[0]),
                   self.M_, self.N_, self.K_)

    def Backward_gpu(self, top, propagate_down, bottom):
        _check(not propagate_down[0], "!propagate_down[0] Can't backpropagate to EmbedLayer input.")
        dW = self.blobs_[0] if self.param_propagate_down_[0] else None
        db = self.blobs_[1] if (self.bias_term_ and self.param_propagate_down_[1]) else None
        self._call("mms_embed_backward", _p(bottom[0]), _p(top[0], True), _p(dW, True), _p(db, True),
                   self.M_, self.N_, self.K_)


    # -- two Embed layers that share their blobs (param sharing by name): both scatter-adds as one grouped pass
    def _pair_args(self, other, bottom, bottom_o):
        _check(other.blobs_[0].shares_storage_with(self.blobs_[0]) and other.N_ == self.N_ and other.K_ == self.K_,
               "the paired Embed layers must share their table")
        _check(self.dtype == np.float32, "the paired Embed backward is float32 only")
        return (_p(bottom), self.M_, _p(bottom_o), other.M_)

    def PlanPair(self, other, bottom, bottom_o):
        """Groups the token rows of both bottoms by id (mms_embed_plan_pair): needs the ids only, so it can run on a side
        stream beside the forward pass; the next BackwardPair on this layer with the same bottoms uses it."""
        self._bind_stream()
        i0, m0, i1, m1 = self._pair_args(other, bottom, bottom_o)
        check(lib().mms_embed_plan_pair_f32(self.handle.ptr, i0, m0, i1, m1, self.K_))

    def BackwardPair(self, other, top, top_o, bottom, bottom_o):
        """Backward of this layer and of ``other`` in one call (mms_embed_backward_pair)."""
        self._bind_stream()
        i0, m0, i1, m1 = self._pair_args(other, bottom, bottom_o)
        dW = self.blobs_[0] if self.param_propagate_down_[0] else None
        db = self.blobs_[1] if (self.bias_term_ and self.param_propagate_down_[1]) else None
        check(lib().mms_embed_backward_pair_f32(self.handle.ptr, i0, _p(top, True), m0, i1, _p(top_o, True), m1,
                                                _p(dW, True), _p(db, True), self.N_, self.K_))


# ------------------------------------------------------------------------- SimCross
class SimCrossLayer(Layer):
    exact_num_bottom = 2

    def LayerSetUp(self, bottom, top):
        _check(len(bottom) == 2, "bottom.size() == 2")
        _check(bottom[0].num() == bottom[1].num(), "bottom[0]->num() == bottom[1]->num()")
        _check(bottom[0].height() == bottom[1].height(), "bottom[0]->height() == bottom[1]->height()")
        sp = self.layer_param_.sim_cross_param
        self.dist_mode_ = int(sp["dist_mode"])
        self.bias_term_ = bool(sp["bias_term"])
        self.mc_ = int(sp["mesure_count"])
        if self.dist_mode_ == 2:                              # sim_cross_layer.cpp:17-46
            D = bottom[0].height()
            self.blobs_ = [self._new_blob((self.mc_, D, bottom[1].height()), bottom[0])]
            _fill(self.blobs_[0], sp["weight_filler"], self.rng)
            if self.bias_term_:
                self.blobs_.append(self._new_blob((self.mc_, bottom[0].channels(), bottom[1].channels()), bottom[0]))
                _fill(self.blobs_[1], sp["bias_filler"], self.rng)
        self.data0_norm_ = self._new_blob((), bottom[0])
        self.data1_norm_ = self._new_blob((), bottom[0])

    def Reshape(self, bottom, top):
        N, Lq, La = bottom[0].num(), bottom[0].channels(), bottom[1].channels()
        top[0].Reshape((N, self.mc_ if self.dist_mode_ == 2 else 1, Lq, La))   # :52-62
        if self.dist_mode_ == 0:
            self.data0_norm_.Reshape((N, Lq))
            self.data1_norm_.Reshape((N, La))

    def _dims(self, bottom):
        return (bottom[0].num(), bottom[0].channels(), bottom[1].channels(), bottom[0].height(),
                self.mc_ if self.dist_mode_ == 2 else 1)

    def Forward_gpu(self, bottom, top):
        N, Lq, La, D, mc = self._dims(bottom)
        Mw = self.blobs_[0] if self.dist_mode_ == 2 else None
        B = self.blobs_[1] if (self.dist_mode_ == 2 and self.bias_term_) else None
        n0 = self.data0_norm_ if self.dist_mode_ == 0 else None
        n1 = self.data1_norm_ if self.dist_mode_ == 0 else None
        self._call("mms_simcross_forward", self.dist_mode_, _p(bottom[0]), _p(bottom[1]), _p(Mw), _p(B),
                   _p(top[0]), _p(n0), _p(n1), N, Lq, La, D, mc)

    def Backward_gpu(self, top, propagate_down, bottom):
        N, Lq, La, D, mc = self._dims(bottom)
        Mw = self.blobs_[0] if self.dist_mode_ == 2 else None
        B = self.blobs_[1] if (self.dist_mode_ == 2 and self.bias_term_) else None
        n0 = self.data0_norm_ if self.dist_mode_ == 0 else None
        n1 = self.data1_norm_ if self.dist_mode_ == 0 else None
        self._call("mms_simcross_backward", self.dist_mode_, _p(bottom[0]), _p(bottom[1]), _p(Mw),
                   _p(top[0]), _p(top[0], True), _p(n0), _p(n1), _p(bottom[0], True), _p(bottom[1], True),
                   _p(Mw, True), _p(B, True), N, Lq, La, D, mc, int(propagate_down[0]), int(propagate_down[1]))

    def Prepare(self, bottom):
        """Rounds M ahead of the next Forward (mms_simcross_prepare): callable on a side stream beside the producers of the
        bottoms.  No-op for the other modes / double."""
        if self.dist_mode_ == 2 and self.dtype == np.float32:
            self._bind_stream()
            N, Lq, La, D, mc = self._dims(bottom)
            check(lib().mms_simcross_prepare_f32(self.handle.ptr, _p(self.blobs_[0]), D, mc))

    # -- the same backward in two calls (include/mms_b200.h: mms_simcross_backward_bottoms / _params) -------------
    def BackwardBottoms(self, top, bottom):
        """dq, da only; the weight gradient follows in BackwardParams.  Falls back to the whole Backward when the
        library does not take the shape in split form (BackwardParams is then a no-op)."""
        self._bind_stream()
        self._split_pending = False
        if self.dist_mode_ == 2 and self.dtype == np.float32:
            N, Lq, La, D, mc = self._dims(bottom)
            rc = lib().mms_simcross_backward_bottoms_f32(
                self.handle.ptr, _p(bottom[0]), _p(bottom[1]), _p(self.blobs_[0]), _p(top[0], True),
                _p(bottom[0], True), _p(bottom[1], True), N, Lq, La, D, mc)
            if rc == 0:
                self._split_pending = True
                return
            if rc != _lib.MMS_E_UNSUPPORTED:
                check(rc)
        self.Backward_gpu(top, [True, True], bottom)

    def BackwardParams(self, top, bottom):
        if not getattr(self, "_split_pending", False):
            return
        self._bind_stream()
        N, Lq, La, D, mc = self._dims(bottom)
        B = self.blobs_[1] if self.bias_term_ else None
        check(lib().mms_simcross_backward_params_f32(self.handle.ptr, _p(top[0], True), _p(self.blobs_[0], True),
                                                     _p(B, True), N, Lq, La, D, mc))
        self._split_pending = False


# ------------------------------------------------------------------------- SimMatrix
class SimMatrixLayer(Layer):
    exact_num_bottom = 2

    def LayerSetUp(self, bottom, top):
        _check(bottom[0].num() == bottom[1].num(), "bottom[0]->num() == bottom[1]->num()")
        self.K1_ = bottom[0].count(1)
        self.K2_ = bottom[1].count(1)
        if not self.blobs_:                                    # sim_matrix_layer.cpp:17-32
            self.blobs_.append(self._new_blob((self.K1_, self.K2_), bottom[0]))
            _fill(self.blobs_[0], self.layer_param_.sim_matrix_param["weight_filler"], self.rng)
        self.param_propagate_down_ = [True] * len(self.blobs_)

    def Reshape(self, bottom, top):
        _check(self.K1_ == bottom[0].count(1), "Input size incompatible with inner product parameters.")
        _check(self.K2_ == bottom[1].count(1), "Input size incompatible with inner product parameters.")
        self.M_ = bottom[0].count(0, 1)
        top[0].Reshape((bottom[0].shape[0], 1))                # :46-49

    def Forward_gpu(self, bottom, top):
        # T = q W lands in bottom[1]'s diff buffer, as in the reference (:58)
        self._call("mms_simmatrix_forward", _p(bottom[0]), _p(bottom[1]), _p(self.blobs_[0]), _p(top[0]),
                   _p(bottom[1], True), self.M_, self.K1_, self.K2_)

    def Backward_gpu(self, top, propagate_down, bottom):
        self._call("mms_simmatrix_backward", _p(bottom[0]), _p(bottom[1]), _p(self.blobs_[0]),
                   _p(top[0], True), _p(self.blobs_[0], True), _p(bottom[0], True), _p(bottom[1], True),
                   self.M_, self.K1_, self.K2_, int(self.param_propagate_down_[0]),
                   int(propagate_down[0]), int(propagate_down[1]))


# ------------------------------------------------------------------------- PairRankLoss
class PairRankLossLayer(Layer):
    exact_num_bottom = 3

    def LayerSetUp(self, bottom, top):
        if not self.layer_param_.loss_weight:                  # LossLayer::LayerSetUp, loss_layer.cpp
            self.layer_param_.loss_weight = [1.0]
        _check(bottom[0].num() == bottom[1].num(), "bottom[0]->num() == bottom[1]->num()")
        _check(bottom[0].num() == bottom[2].num(), "bottom[0]->num() == bottom[2]->num()")
        _check(bottom[0].count(1) == bottom[2].count(1), "bottom[0]->count(1) == bottom[2]->count(1)")
        _check(bottom[0].count(1) == bottom[1].count(1), "bottom[0]->count(1) == bottom[1]->count(1)")
        self.margin_ = float(self.layer_param_.pair_rank_loss_param["margin"])
        shp = (bottom[0].num(), bottom[0].channels(), 1, 1)    # pair_rank_loss_layer.cpp:21-22
        self.ordered_diff_ = self._new_blob(shp, bottom[0])
        self.similar_diff_ = self._new_blob(shp, bottom[0])

    def Reshape(self, bottom, top):
        _check(bottom[0].num() == bottom[1].num(), "The data and label should have the same number.")
        top[0].Reshape((1,))                                   # LossLayer::Reshape: scalar top

    def Forward_gpu(self, bottom, top):
        self._call("mms_pairrankloss_forward", _p(bottom[0]), _p(bottom[1]), _p(bottom[2]),
                   self.real(self.margin_), bottom[0].count(), _p(top[0]), _p(self.ordered_diff_),
                   _p(self.similar_diff_))

    def Backward_gpu(self, top, propagate_down, bottom):
        if propagate_down[2]:
            raise CheckError(self.type() + " Layer cannot backpropagate to label inputs.")
        if self.defer_loss_:                                   # recording a graph: no host read; the loss weight
            top_diff = float(self.loss_[0])                    # SetLossWeights put into top.diff (layer.hpp:414-428)
        else:
            top_diff = float(top[0].diff.reshape(-1)[0].item())    # top[0]->cpu_diff()[0]
        self._call("mms_pairrankloss_backward", _p(bottom[2]), _p(self.ordered_diff_), _p(self.similar_diff_),
                   self.real(top_diff), bottom[0].count(),
                   _p(bottom[0] if propagate_down[0] else None, True),
                   _p(bottom[1] if propagate_down[1] else None, True))


# ------------------------------------------------------------------------- FM
class FMLayer(Layer):
    exact_num_bottom = 1

    def LayerSetUp(self, bottom, top):
        self.bias_term_ = bool(self.layer_param_.fm_param["bias_term"])
        if self.bias_term_:                                    # fm_layer.cpp:13-19
            self.blobs_ = [self._new_blob((1,), bottom[0])]
        self.param_propagate_down_ = [True] * len(self.blobs_)

    def Reshape(self, bottom, top):
        top[0].Reshape((bottom[0].num(), 1))

    def Forward_gpu(self, bottom, top):
        b = self.blobs_[0] if self.bias_term_ else None
        self._call("mms_fm_forward", _p(bottom[0]), _p(b), _p(top[0]), bottom[0].num(),
                   bottom[0].channels(), bottom[0].height())

    def Backward_gpu(self, top, propagate_down, bottom):
        db = self.blobs_[0] if (self.bias_term_ and self.param_propagate_down_[0]) else None
        self._call("mms_fm_backward", _p(bottom[0]), _p(top[0], True), _p(bottom[0], True), _p(db, True),
                   bottom[0].num(), bottom[0].channels(), bottom[0].height(), int(propagate_down[0]))


# ------------------------------------------------------------------------- sentence encoder
class ConvolutionLayer(Layer):
    """ConvolutionLayer (conv_layer.cpp, base_conv_layer.cpp), stride 1, pad 0, group 1 -- the two geometries of the
    reference's nets: the sentence convolution (do_trec_qa_clean.py:352-358, 412-416: one input channel, kernel as wide
    as the input; mms_sentconv_*) and the 2-D convolutions over the similarity tensor (:470-477; mms_conv2d_*).  Other
    geometries are the stock Caffe layer's business and are refused."""
    exact_num_bottom = 1

    def LayerSetUp(self, bottom, top):
        cp = self.layer_param_.convolution_param
        self.kernel_h_ = int(cp["kernel_h"] or cp["kernel_size"])
        self.kernel_w_ = int(cp["kernel_w"] or cp["kernel_size"])
        _check(self.kernel_h_ > 0 and self.kernel_w_ > 0, "Filter dimensions must be nonzero.")   # base_conv_layer.cpp:52-54
        self.num_output_ = int(cp["num_output"])
        _check(self.num_output_ > 0, "num_output must be positive")
        self.bias_term_ = bool(cp["bias_term"])
        _check(bottom[0].num_axes() == 4, "convolution takes (N, C, H, W) bottoms")
        _check(int(cp["group"]) == 1 and int(cp["stride"]) == 1 and int(cp["pad"]) == 0,
               "Convolution (mms_b200): only stride 1, pad 0, group 1 runs here")
        self.channels_ = bottom[0].channels()
        self.sentence_ = self.channels_ == 1 and self.kernel_w_ == bottom[0].width()
        if not self.blobs_:                                   # base_conv_layer.cpp:137-160
            self.blobs_.append(self._new_blob((self.num_output_, self.channels_, self.kernel_h_, self.kernel_w_), bottom[0]))
            _fill(self.blobs_[0], cp["weight_filler"], self.rng)
            if self.bias_term_:
                self.blobs_.append(self._new_blob((self.num_output_,), bottom[0]))
                _fill(self.blobs_[1], cp["bias_filler"], self.rng)
        self.param_propagate_down_ = [True] * len(self.blobs_)

    def Reshape(self, bottom, top):
        _check(bottom[0].height() >= self.kernel_h_ and bottom[0].width() >= self.kernel_w_, "kernel larger than the input")
        _check(bottom[0].channels() == self.channels_, "Input size incompatible with convolution kernel.")
        top[0].Reshape((bottom[0].num(), self.num_output_, bottom[0].height() - self.kernel_h_ + 1,
                        bottom[0].width() - self.kernel_w_ + 1))

    def Forward_gpu(self, bottom, top):
        bias = self.blobs_[1] if self.bias_term_ else None
        if self.sentence_:
            self._call("mms_sentconv_forward", _p(bottom[0]), _p(self.blobs_[0]), _p(bias), _p(top[0]), bottom[0].num(),
                       bottom[0].height(), bottom[0].width(), self.num_output_, self.kernel_h_)
        else:
            self._call("mms_conv2d_forward", _p(bottom[0]), _p(self.blobs_[0]), _p(bias), _p(top[0]), bottom[0].num(),
                       self.channels_, bottom[0].height(), bottom[0].width(), self.num_output_, self.kernel_h_,
                       self.kernel_w_)

    def Backward_gpu(self, top, propagate_down, bottom):
        dW = self.blobs_[0] if self.param_propagate_down_[0] else None
        db = self.blobs_[1] if (self.bias_term_ and self.param_propagate_down_[1]) else None
        dx = bottom[0] if propagate_down[0] else None
        if self.sentence_:
            self._call("mms_sentconv_backward", _p(bottom[0]), _p(self.blobs_[0]), _p(top[0], True), _p(dW, True),
                       _p(db, True), _p(dx, True), bottom[0].num(), bottom[0].height(), bottom[0].width(),
                       self.num_output_, self.kernel_h_)
        else:
            self._call("mms_conv2d_backward", _p(bottom[0]), _p(self.blobs_[0]), _p(top[0], True), _p(dW, True),
                       _p(db, True), _p(dx, True), bottom[0].num(), self.channels_, bottom[0].height(), bottom[0].width(),
                       self.num_output_, self.kernel_h_, self.kernel_w_)


class DropoutLayer(Layer):
    """DropoutLayer (dropout_layer.cpp:13-75, dropout_layer.cu:10-60): TRAIN: top = bottom * (mask > threshold) * scale with
    one random 32-bit word per element (drawn by mms_dropout_mask, a counter-based generator seeded per forward, or set
    by the caller in ``rand_vec_``); TEST: a copy.  In place is allowed."""
    exact_num_bottom = 1

    def LayerSetUp(self, bottom, top):
        self.threshold_ = float(self.layer_param_.dropout_param["dropout_ratio"])
        _check(0.0 < self.threshold_ < 1.0, "threshold_ > 0. && threshold_ < 1.")
        self.scale_ = 1.0 / (1.0 - self.threshold_)
        self.uint_thres_ = int(4294967295 * self.threshold_)                   # UINT_MAX * threshold_ (dropout_layer.cpp:21)
        self.rand_vec_ = None
        self.seed_ = int(self.layer_param_.dropout_param["seed"])
        self.fixed_mask_ = False
        self.calls_ = 0

    def Reshape(self, bottom, top):
        top[0].Reshape(bottom[0].shape)
        n = bottom[0].count()
        if self.rand_vec_ is None or self.rand_vec_.numel() != n:
            self.rand_vec_ = torch.zeros(n, dtype=torch.int32, device=bottom[0].device)
            self.fixed_mask_ = False

    def set_mask(self, words):
        """Test hook: use these 32-bit words instead of drawing new ones every forward."""
        self.rand_vec_.copy_(torch.as_tensor(np.asarray(words, dtype=np.uint32).view(np.int32)).reshape(-1))
        self.fixed_mask_ = True

    def Forward_gpu(self, bottom, top):
        n = bottom[0].count()
        if self.layer_param_.phase == "TRAIN":
            if not self.fixed_mask_:
                check(lib().mms_dropout_mask(self.handle.ptr, c_p(self.rand_vec_.data_ptr()), n,
                                             ctypes.c_ulonglong(self.seed_ + 0x51ED27 * self.calls_)))
                self.calls_ += 1
            self._call("mms_dropout", _p(bottom[0]), c_p(self.rand_vec_.data_ptr()), _p(top[0]), n,
                       ctypes.c_uint(self.uint_thres_), self.real(self.scale_))
        elif top[0].gpu_data() != bottom[0].gpu_data():
            top[0].data.copy_(bottom[0].data)

    def Backward_gpu(self, top, propagate_down, bottom):
        if not propagate_down[0]:
            return
        n = bottom[0].count()
        if self.layer_param_.phase == "TRAIN":
            self._call("mms_dropout", _p(top[0], True), c_p(self.rand_vec_.data_ptr()), _p(bottom[0], True), n,
                       ctypes.c_uint(self.uint_thres_), self.real(self.scale_))
        elif top[0].gpu_diff() != bottom[0].gpu_diff():
            bottom[0].diff.copy_(top[0].diff)


class PoolingLayer(Layer):
    """PoolingLayer (pooling_layer.cpp:17-227), MAX and AVE, one top."""
    exact_num_bottom = 1

    def LayerSetUp(self, bottom, top):
        pp = self.layer_param_.pooling_param
        self.method_ = {"MAX": 0, "AVE": 1}.get(pp["pool"])
        _check(self.method_ is not None, "Unknown pooling method.")           # STOCHASTIC has no CPU path in the reference
        self.global_pooling_ = bool(pp["global_pooling"])
        if self.global_pooling_:
            self.kernel_h_, self.kernel_w_ = bottom[0].height(), bottom[0].width()
        else:
            self.kernel_h_ = int(pp["kernel_h"] or pp["kernel_size"])
            self.kernel_w_ = int(pp["kernel_w"] or pp["kernel_size"])
        _check(self.kernel_h_ > 0 and self.kernel_w_ > 0, "Filter dimensions cannot be zero.")
        self.pad_h_ = int(pp["pad_h"] or pp["pad"])
        self.pad_w_ = int(pp["pad_w"] or pp["pad"])
        self.stride_h_ = int(pp["stride_h"] or pp["stride"])
        self.stride_w_ = int(pp["stride_w"] or pp["stride"])
        _check(self.pad_h_ < self.kernel_h_ and self.pad_w_ < self.kernel_w_, "pad must be smaller than the kernel")
        self.max_idx_ = None

    def Reshape(self, bottom, top):
        _check(bottom[0].num_axes() == 4, "Input must have 4 axes, corresponding to (num, channels, height, width)")
        H, W = bottom[0].height(), bottom[0].width()
        if self.global_pooling_:
            self.kernel_h_, self.kernel_w_ = H, W
        ceil_div = lambda a, b: -(-a // b)
        self.pooled_height_ = ceil_div(H + 2 * self.pad_h_ - self.kernel_h_, self.stride_h_) + 1   # :88-91
        self.pooled_width_ = ceil_div(W + 2 * self.pad_w_ - self.kernel_w_, self.stride_w_) + 1
        if self.pad_h_ or self.pad_w_:                                        # :92-103
            if (self.pooled_height_ - 1) * self.stride_h_ >= H + self.pad_h_:
                self.pooled_height_ -= 1
            if (self.pooled_width_ - 1) * self.stride_w_ >= W + self.pad_w_:
                self.pooled_width_ -= 1
        shape = (bottom[0].num(), bottom[0].channels(), self.pooled_height_, self.pooled_width_)
        top[0].Reshape(shape)
        if self.method_ == 0 and (self.max_idx_ is None or tuple(self.max_idx_.shape) != shape):
            self.max_idx_ = torch.zeros(shape, dtype=torch.int32, device=bottom[0].device)       # Blob<int> max_idx_

    def _geom(self, bottom):
        return (bottom[0].num() * bottom[0].channels(), bottom[0].height(), bottom[0].width(), self.pooled_height_,
                self.pooled_width_, self.kernel_h_, self.kernel_w_, self.stride_h_, self.stride_w_, self.pad_h_,
                self.pad_w_, self.method_)

    def Forward_gpu(self, bottom, top):
        mask = c_p(self.max_idx_.data_ptr()) if self.method_ == 0 else c_p(0)
        self._call("mms_pool_forward", _p(bottom[0]), _p(top[0]), mask, *self._geom(bottom))

    def Backward_gpu(self, top, propagate_down, bottom):
        if not propagate_down[0]:
            return
        mask = c_p(self.max_idx_.data_ptr()) if self.method_ == 0 else c_p(0)
        self._call("mms_pool_backward", _p(top[0], True), mask, _p(bottom[0], True), *self._geom(bottom))


class TanHLayer(Layer):
    """TanHLayer (tanh_layer.cpp:11-37); in place (top[0] is bottom[0]) is allowed, as in the reference."""
    exact_num_bottom = 1

    def LayerSetUp(self, bottom, top):
        pass

    def Reshape(self, bottom, top):
        if top[0] is not bottom[0]:
            top[0].Reshape(bottom[0].shape)                                    # NeuronLayer::Reshape

    def Forward_gpu(self, bottom, top):
        self._call("mms_tanh_forward", _p(bottom[0]), _p(top[0]), bottom[0].count())

    def Backward_gpu(self, top, propagate_down, bottom):
        if propagate_down[0]:
            self._call("mms_tanh_backward", _p(top[0]), _p(top[0], True), _p(bottom[0], True), bottom[0].count())


class BNLayer(Layer):
    """The fork's batch normalisation (type "BN", bn_layer.cpp): blobs scale, shift, running mean, running variance,
    each (1, C, 1, 1)."""
    exact_num_bottom = 1

    def LayerSetUp(self, bottom, top):
        _check(top[0] is not bottom[0], "BN Layer does not allow in-place computation.")           # bn_layer.cpp:52-53
        bp = self.layer_param_.bn_param
        self.bn_memory_ = float(np.float32(bp["bn_memory"]))          # a float field of BNParameter (caffe.proto:485)
        self.var_eps_ = 1e-9                                                                        # :63
        C = bottom[0].channels()
        if not self.blobs_:                                                                         # :95-117
            for filler in (bp["scale_filler"], bp["shift_filler"]):
                self.blobs_.append(self._new_blob((1, C, 1, 1), bottom[0]))
                _fill(self.blobs_[-1], filler, self.rng)
            self.blobs_.append(self._new_blob((1, C, 1, 1), bottom[0]))
            self.blobs_.append(self._new_blob((1, C, 1, 1), bottom[0]))
            self.blobs_[2].data.zero_(); self.blobs_[3].data.zero_()
        self.param_propagate_down_ = [True] * len(self.blobs_)

    def Reshape(self, bottom, top):
        shape = (bottom[0].num(), bottom[0].channels(), bottom[0].height(), bottom[0].width())
        top[0].Reshape(shape)
        if getattr(self, "x_norm_", None) is None or tuple(self.x_norm_.shape) != shape:
            self.x_norm_ = self._new_blob(shape, bottom[0])                    # buffer_blob_.diff in the reference
            self.batch_mean_ = self._new_blob((shape[1],), bottom[0])
            self.batch_std_ = self._new_blob((shape[1],), bottom[0])           # batch_variance_ after the sqrt

    def Forward_gpu(self, bottom, top):
        N, C, HW = bottom[0].num(), bottom[0].channels(), bottom[0].height() * bottom[0].width()
        self._call("mms_bn_forward", _p(bottom[0]), _p(self.blobs_[0]), _p(self.blobs_[1]), _p(self.blobs_[2]),
                   _p(self.blobs_[3]), _p(top[0]), _p(self.x_norm_), _p(self.batch_mean_), _p(self.batch_std_), N, C, HW,
                   int(self.layer_param_.phase == "TRAIN"), self.real(self.bn_memory_), self.real(self.var_eps_))

    def Backward_gpu(self, top, propagate_down, bottom):
        N, C, HW = bottom[0].num(), bottom[0].channels(), bottom[0].height() * bottom[0].width()
        self._call("mms_bn_backward", _p(top[0], True), _p(self.x_norm_), _p(self.blobs_[0]), _p(self.batch_std_),
                   _p(self.blobs_[0], True), _p(self.blobs_[1], True), _p(bottom[0], True), N, C, HW)


# ------------------------------------------------------------------------- ranking metrics
class _GroupedRankLayer(Layer):
    """MAPLayer / MRRLayer (include/caffe/layers/{map,mrr}_layer.hpp): bottoms (predictions (N, C), labels (N),
    group ids (N)), a scalar top.  The reference implements Forward_cpu only; here the scores never leave the
    device."""
    exact_num_bottom = 3
    _param = None
    _which = 0

    def LayerSetUp(self, bottom, top):
        self.fixed_axis_ = int(getattr(self.layer_param_, self._param)["fixed_axis"])

    def Reshape(self, bottom, top):
        _check(self.fixed_axis_ <= bottom[0].count() // max(bottom[1].count(), 1),
               "top_k must be less than or equal to the number of classes.")           # map_layer.cpp:19-20
        outer, inner = bottom[0].count(0, 1), bottom[0].count(2)
        _check(outer * inner == bottom[1].count(), "Number of labels must match number of predictions; ")
        _check(outer * inner == bottom[2].count(), "Number of group ids must match number of predictions")
        top[0].Reshape((1,))                                   # scalar top (0 axes in the reference)

    def Forward_gpu(self, bottom, top):
        out = c_p(top[0].gpu_data())
        null = c_p(0)
        # the reference reads bottom_data[i * (fixed_axis + 1) + fixed_axis] (map_layer.cpp:50, mrr_layer.cpp:49)
        self._call("mms_rank_map_mrr", _p(bottom[0]), self.fixed_axis_ + 1, self.fixed_axis_, _p(bottom[1]),
                   _p(bottom[2]), bottom[0].num(), out if self._which == 0 else null, out if self._which == 1 else null)

    def Backward_gpu(self, top, propagate_down, bottom):
        _check(not any(propagate_down), "%s Layer cannot backpropagate." % self.type())   # map_layer.hpp: NOT_IMPLEMENTED


class MAPLayer(_GroupedRankLayer):
    _param, _which = "map_param", 0


class MRRLayer(_GroupedRankLayer):
    _param, _which = "mrr_param", 1


class AUCLayer(Layer):
    """AUCLayer (auc_layer.cpp:11-136) for (N, C) predictions and N labels."""
    exact_num_bottom = 2

    def LayerSetUp(self, bottom, top):
        ap = self.layer_param_.auc_param
        self.fixed_axis_ = int(ap["fixed_axis"])
        self.has_ignore_label_ = ap["ignore_label"] is not None
        self.ignore_label_ = int(ap["ignore_label"] or 0)

    def Reshape(self, bottom, top):
        _check(self.fixed_axis_ <= bottom[0].count() // max(bottom[1].count(), 1),
               "top_k must be less than or equal to the number of classes.")
        axis = int(self.layer_param_.auc_param["axis"])
        _check(axis == 1 and bottom[0].count(2) == 1,
               "AUC on the device takes (N, C) predictions (label axis 1, no inner dimensions)")
        _check(bottom[0].count(0, 1) == bottom[1].count(), "Number of labels must match number of predictions; ")
        top[0].Reshape((1,))                                   # scalar top (0 axes in the reference)

    def Forward_gpu(self, bottom, top):
        self._call("mms_rank_auc", _p(bottom[0]), bottom[0].count(1), self.fixed_axis_, _p(bottom[1]),
                   bottom[0].num(), int(self.has_ignore_label_), self.ignore_label_, c_p(top[0].gpu_data()))

    def Backward_gpu(self, top, propagate_down, bottom):
        _check(not any(propagate_down), "AUC Layer cannot backpropagate.")


class RankAccuracyLayer(Layer):
    """RankAccuracyLayer (rank_accuracy_layer.cpp:17-50): fraction of pairs with label * (a - b) > 0."""
    exact_num_bottom = 3

    def LayerSetUp(self, bottom, top):
        pass

    def Reshape(self, bottom, top):
        _check(bottom[0].count() == bottom[1].count(), "two pairs have the same dimension!.")
        _check(bottom[0].count() == bottom[2].count(), "pair should have the same dimension with the label!.")
        top[0].Reshape((1,))                                   # scalar top (0 axes in the reference)

    def Forward_gpu(self, bottom, top):
        self._call("mms_rank_accuracy", _p(bottom[0]), _p(bottom[1]), _p(bottom[2]), bottom[0].count(),
                   c_p(top[0].gpu_data()))

    def Backward_gpu(self, top, propagate_down, bottom):
        _check(not any(propagate_down), "RankAccuracy Layer cannot backpropagate.")


_REGISTRY = {"Embed": EmbedLayer, "SimCross": SimCrossLayer, "SimMatrix": SimMatrixLayer,
             "PairRankLoss": PairRankLossLayer, "FM": FMLayer, "MAP": MAPLayer, "MRR": MRRLayer, "AUC": AUCLayer,
             "RankAccuracy": RankAccuracyLayer, "Convolution": ConvolutionLayer, "Pooling": PoolingLayer,
             "TanH": TanHLayer, "BN": BNLayer, "Dropout": DropoutLayer}


def create_layer(param):
    """LayerRegistry::CreateLayer (include/caffe/layer_factory.hpp:73-81)."""
    if param.type not in _REGISTRY:
        raise CheckError("Unknown layer type: %s (known types: %s)" % (param.type, ", ".join(sorted(_REGISTRY))))
    return _REGISTRY[param.type](param)
