"""MultiModalNet: BASELINE.json's configs[4] -- "multi-modal MMS with 4 embedding modalities, 300-d, M 1024x1024,
batch 16384" -- as one net over the reference's layers:

    q^_m, a^_m (N, K)  -- per modality m -->  SimMatrix_m   s_m = q^_m^T W_m a^_m          (sim_matrix_layer.cpp:53-95)
    Concat(s_0 .. s_{C-1}) -> x (N, C, 1) --> FM            y = sum_m x_m0 + b             (fm_layer.cpp:33-99; with Dm = 1
                                                                                            only the linear column exists)
    Slice(y) -> y+ (first half), y- (second half) + labels --> PairRankLoss, margin 1       (pair_rank_loss_layer.cpp:26-84)

forward and backward into every W_m, the FM bias and the bottoms.  Concat and Slice are stock Caffe plumbing (views /
strided copies of N*C floats, done with torch here); every arithmetic layer goes through the C-ABI.  The pairs are
built by splitting the batch, as SURVEY.md 8(d) prescribes.  Data-parallel: the batch (pairs of rows n and n + N/2) is
sharded over the ranks, the W_m / bias gradients are exchanged with GradientExchange.
"""
import numpy as np
import torch

from . import _lib
from .blob import Blob
from .layers import FMLayer, LayerParameter, PairRankLossLayer, SimMatrixLayer


class MultiModalNet(object):
    def __init__(self, N, K1=1024, K2=1024, modalities=4, margin=1.0, dtype=np.float32, device="cuda"):
        if N % 2:
            raise ValueError("N must be even: PairRankLoss pairs row n with row n + N/2")
        self.N, self.K1, self.K2, self.C = N, K1, K2, modalities
        self.dtype = np.dtype(dtype)
        mk = lambda shape=(): Blob(shape, dtype=dtype, device=device)
        self.q = [mk((N, K1)) for _ in range(modalities)]
        self.a = [mk((N, K2)) for _ in range(modalities)]
        self.s = [mk() for _ in range(modalities)]
        self.sim = []
        for m in range(modalities):
            lay = SimMatrixLayer(LayerParameter("SimMatrix", name="sim%d" % m, dtype=dtype,
                                                sim_matrix_param=dict(weight_filler=dict(type="xavier"))))
            lay.SetUp([self.q[m], self.a[m]], [self.s[m]])
            lay.handle.set_option(_lib.MMS_OPT_REUSE_FORWARD, 1)     # Backward right after Forward on unchanged bottoms
            self.sim.append(lay)
        self.x = mk((N, modalities, 1))
        self.y = mk()
        self.fm = FMLayer(LayerParameter("FM", name="fm", dtype=dtype, fm_param=dict(bias_term=True)))
        self.fm.SetUp([self.x], [self.y])
        h = N // 2
        self.y_pos, self.y_neg, self.label = mk((h, 1)), mk((h, 1)), mk((h, 1))
        self.y_pos.set_data(self.y.data[:h]); self.y_pos.set_diff(self.y.diff[:h])      # Slice: views of the FM top
        self.y_neg.set_data(self.y.data[h:]); self.y_neg.set_diff(self.y.diff[h:])
        self.loss_top = mk()
        self.loss = PairRankLossLayer(LayerParameter("PairRankLoss", name="loss", dtype=dtype,
                                                     pair_rank_loss_param=dict(margin=margin)))
        self.loss.SetUp([self.y_pos, self.y_neg, self.label], [self.loss_top])
        self._side = None
        self._graph = None

    def layers(self):
        return list(self.sim) + [self.fm, self.loss]

    def params(self):
        return [l.blobs[0] for l in self.sim] + list(self.fm.blobs)

    def set_inputs(self, qs, as_, label):
        for m in range(self.C):
            self.q[m].set_cpu_data(qs[m]); self.a[m].set_cpu_data(as_[m])
        self.label.set_cpu_data(np.asarray(label).reshape(-1, 1))

    def ClearParamDiffs(self):
        for p in self.params():
            p.diff.zero_()

    # -- the step ---------------------------------------------------------------------------------------------------
    def Forward(self):
        for m in range(self.C):
            self.sim[m].Forward([self.q[m], self.a[m]], [self.s[m]])
            self.x.data[:, m, :].copy_(self.s[m].data)                      # Concat (concat_layer.cpp)
        self.fm.Forward([self.x], [self.y])
        return self.loss.Forward([self.y_pos, self.y_neg, self.label], [self.loss_top])

    def Backward(self):
        self.loss.Backward([self.loss_top], [True, True, False], [self.y_pos, self.y_neg, self.label])
        self.fm.Backward([self.y], [True], [self.x])
        for m in range(self.C):
            self.s[m].diff.copy_(self.x.diff[:, m, :])                      # Concat backward
            self.sim[m].Backward([self.s[m]], [True, True], [self.q[m], self.a[m]])

    def ForwardBackward(self):
        loss = self.Forward()
        self.Backward()
        return loss

    def step(self, exch=None, clear_diffs=True):
        """ClearParamDiffs + Forward + Backward (+ the exchange of the parameter gradients)."""
        if clear_diffs:
            self.ClearParamDiffs()
        self.ForwardBackward()
        if exch is not None:
            exch.allreduce()

    def capture(self, exch=None, clear_diffs=True):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        self.loss.defer_loss_ = True                               # the loss stays on the device while capturing
        with torch.cuda.stream(side):
            for _ in range(2):
                self.step(exch, clear_diffs)
            if exch is not None:
                exch.check()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=side):
            self.step(exch, clear_diffs)
        return self._graph

    def replay(self):
        self._graph.replay()

    def loss_value(self):
        return float(self.loss_top.data.reshape(-1)[0].item())
