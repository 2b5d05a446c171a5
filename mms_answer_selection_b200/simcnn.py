"""SimCNNNet: the reference's ``network_v4`` up to the Flatten layer (examples/trec_qa_w2v_mms/do_trec_qa_clean.py:452-478)
-- the net it actually trains:

    question ids -> Embed --\\
                             +--> SimCross(dist_mode 2, mesure_count 4) -> S (N, 4, L, L)
    answer ids   -> Embed --/        -> Dropout(0.1)
                                     -> Convolution(5x5, 32) -> BN -> Pooling(AVE 4x4, stride 4) -> TanH (in place)
                                     -> Convolution(5x5, 64) -> BN -> Pooling(AVE 5x5)           -> TanH (in place)
                                     -> feat (N, 64, 1, 1)

executed layer by layer in net order the way caffe::Net does (net.cpp:535-591); Embed weights shared by name.  The
classifier head behind ``feat`` (Concat with the overlap features, two InnerProduct layers, SoftmaxWithLoss) is stock
Caffe and out of scope; ``feat`` is treated as a loss top whose per-element loss weights stand for the upstream
gradient, as MMSNet does for S.
"""
import numpy as np
import torch

from . import _lib
from .blob import Blob
from .layers import (BNLayer, ConvolutionLayer, DropoutLayer, EmbedLayer, LayerParameter, PoolingLayer, SimCrossLayer,
                     TanHLayer)


class SimCNNNet(object):
    def __init__(self, N, L=40, D=300, mc=4, V=60002, dtype=np.float32, device="cuda", dropout_ratio=0.1, phase="TRAIN"):
        self.N, self.L, self.D, self.mc, self.V = N, L, D, mc, V
        self.dtype = np.dtype(dtype)
        mk = lambda shape=(): Blob(shape, dtype=dtype, device=device)
        LP = lambda t, name, **kw: LayerParameter(t, name=name, dtype=dtype, phase=phase, **kw)
        self.idx_q, self.idx_a = mk((N, L)), mk((N, L))
        self.q, self.a, self.S, self.Sd = mk(), mk(), mk(), mk()
        ep = dict(num_output=D, input_dim=V, bias_term=True, weight_filler=dict(type="uniform", min=-0.08, max=0.08))
        self.embed_q = EmbedLayer(LP("Embed", "w2v_q", embed_param=ep))
        self.embed_a = EmbedLayer(LP("Embed", "w2v_a", embed_param=ep))
        self.sim = SimCrossLayer(LP("SimCross", "sim_cross", sim_cross_param=dict(
            dist_mode=2, mesure_count=mc, bias_term=True, weight_filler=dict(type="uniform", min=-0.1, max=0.1))))
        self.drop = DropoutLayer(LP("Dropout", "sim_drop", dropout_param=dict(dropout_ratio=dropout_ratio)))
        conv = lambda name, nout: ConvolutionLayer(LP("Convolution", name, convolution_param=dict(
            num_output=nout, kernel_h=5, kernel_w=5, weight_filler=dict(type="xavier"), bias_filler=dict(type="constant"))))
        bn = lambda name: BNLayer(LP("BN", name, bn_param=dict(scale_filler=dict(type="constant", value=1.0),
                                                                shift_filler=dict(type="constant", value=1e-3))))
        pool = lambda name, k, s: PoolingLayer(LP("Pooling", name, pooling_param=dict(
            pool="AVE", kernel_h=k, kernel_w=k, stride_h=s, stride_w=s)))
        self.conv0, self.bn0, self.pool0, self.tanh0 = conv("conv0", 32), bn("bn0"), pool("pool0", 4, 4), TanHLayer(LP("TanH", "relu0"))
        self.conv1, self.bn1, self.pool1, self.tanh1 = conv("conv1", 64), bn("bn1"), pool("pool1", 5, 1), TanHLayer(LP("TanH", "relu1"))
        self.c0, self.b0, self.p0 = mk(), mk(), mk()
        self.c1, self.b1, self.feat = mk(), mk(), mk()
        self.embed_q.SetUp([self.idx_q], [self.q])
        self.embed_a.SetUp([self.idx_a], [self.a])
        for j, b in enumerate(self.embed_q.blobs):
            self.embed_a.blobs[j].ShareData(b); self.embed_a.blobs[j].ShareDiff(b)
        self.sim.SetUp([self.q, self.a], [self.S])
        self.sim.handle.set_option(_lib.MMS_OPT_REUSE_FORWARD, 1)
        if self.dtype == np.float32:
            self.embed_q.handle.set_option(_lib.MMS_OPT_STAGE_TF32, 1)
            self.embed_a.handle.set_option(_lib.MMS_OPT_STAGE_TF32, 1)
        self.drop.SetUp([self.S], [self.Sd])
        self.conv0.SetUp([self.Sd], [self.c0]); self.bn0.SetUp([self.c0], [self.b0])
        self.pool0.SetUp([self.b0], [self.p0]); self.tanh0.SetUp([self.p0], [self.p0])
        self.conv1.SetUp([self.p0], [self.c1]); self.bn1.SetUp([self.c1], [self.b1])
        self.pool1.SetUp([self.b1], [self.feat]); self.tanh1.SetUp([self.feat], [self.feat])
        self._graph = None

    def layers(self):
        return [self.embed_q, self.embed_a, self.sim, self.drop, self.conv0, self.bn0, self.pool0, self.tanh0,
                self.conv1, self.bn1, self.pool1, self.tanh1]

    def params(self):
        """Learnable blobs in net order, shared blobs once (BN's running mean / variance included, lr_mult 0)."""
        out = list(self.embed_q.blobs) + list(self.sim.blobs)
        for l in (self.conv0, self.bn0, self.conv1, self.bn1):
            out += list(l.blobs)
        return out

    def set_inputs(self, idx_q, idx_a):
        self.idx_q.set_cpu_data(idx_q); self.idx_a.set_cpu_data(idx_a)

    def set_upstream_gradient(self, dfeat):
        self.feat.set_cpu_diff(np.asarray(dfeat).reshape(self.feat.shape))

    def ClearParamDiffs(self):
        for p in self.params():
            p.diff.zero_()

    def Forward(self):
        self.embed_q.Forward([self.idx_q], [self.q]); self.embed_a.Forward([self.idx_a], [self.a])
        self.sim.Forward([self.q, self.a], [self.S])
        self.drop.Forward([self.S], [self.Sd])
        self.conv0.Forward([self.Sd], [self.c0]); self.bn0.Forward([self.c0], [self.b0])
        self.pool0.Forward([self.b0], [self.p0]); self.tanh0.Forward([self.p0], [self.p0])
        self.conv1.Forward([self.p0], [self.c1]); self.bn1.Forward([self.c1], [self.b1])
        self.pool1.Forward([self.b1], [self.feat]); self.tanh1.Forward([self.feat], [self.feat])

    def Backward(self):
        self.tanh1.Backward([self.feat], [True], [self.feat]); self.pool1.Backward([self.feat], [True], [self.b1])
        self.bn1.Backward([self.b1], [True], [self.c1]); self.conv1.Backward([self.c1], [True], [self.p0])
        self.tanh0.Backward([self.p0], [True], [self.p0]); self.pool0.Backward([self.p0], [True], [self.b0])
        self.bn0.Backward([self.b0], [True], [self.c0]); self.conv0.Backward([self.c0], [True], [self.Sd])
        self.drop.Backward([self.Sd], [True], [self.S])
        self.sim.Backward([self.S], [True, True], [self.q, self.a])
        self.embed_q.Backward([self.q], [False], [self.idx_q]); self.embed_a.Backward([self.a], [False], [self.idx_a])

    def ForwardBackward(self):
        self.Forward()
        self.Backward()

    def capture(self, clear_diffs=True):
        def step():
            if clear_diffs:
                self.ClearParamDiffs()
            self.ForwardBackward()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=side):
            step()
        return self._graph

    def replay(self):
        self._graph.replay()
