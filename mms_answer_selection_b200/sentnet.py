"""SentenceVectorNet: the sentence-vector variant of the reference's TREC-QA net
(examples/trec_qa_w2v_mms/do_trec_qa_clean.py:405-429, the block that feeds SimMatrix):

    question ids (N,L) -> Embed -> Reshape (N,1,L,D) -> Convolution(kh x D, C) -> BN -> Pooling(MAX, L-kh+1) -> TanH --\
                                                                                                                      SimMatrix
    answer   ids (N,L) -> Embed -> Reshape (N,1,L,D) -> Convolution(kh x D, C) -> BN -> Pooling(MAX, L-kh+1) -> TanH --/
        s (N,1) = q^T W a   ->   PairRankLoss(s[:N/2], s[N/2:], y)      (pairs built by splitting the batch)

i.e. the whole hot path BASELINE.json's north_star describes: embed the two sentences, encode them, score the pair with the
learned bilinear form, backpropagate a pairwise ranking loss into W, the filters and the embedding table.  The two
branches share every parameter by name, as the script declares them (`w2v-weights`, `conv_1_w`, `bn_1_shape` ...;
sharing = ShareData + ShareDiff, net.cpp:944-950).  Layers run in the reference's order (forward as declared, backward
reversed, net.cpp:535-591); consequence kept from the reference: BN OVERWRITES its scale / shift diffs (bn_layer.cpp:
271-292), so with shared BN parameters the branch whose backward runs last -- the question branch -- is the one whose
gradient the solver sees.
"""
import numpy as np
import torch

from . import _lib
from .blob import Blob
from .layers import (BNLayer, ConvolutionLayer, EmbedLayer, LayerParameter, PairRankLossLayer, PoolingLayer,
                     SimMatrixLayer, TanHLayer)


class SentenceVectorNet(object):
    def __init__(self, N, L=40, D=300, C=100, kh=5, V=60002, dtype=np.float32, device="cuda", margin=1.0):
        assert N % 2 == 0, "the batch is split into two halves of pairs for PairRankLoss"
        self.N, self.L, self.D, self.C, self.kh, self.V = N, L, D, C, kh, V
        self.dtype = np.dtype(dtype)
        mk = lambda shape=(): Blob(shape, dtype=dtype, device=device)
        h = N // 2
        ep = dict(num_output=D, input_dim=V, bias_term=True, weight_filler=dict(type="uniform", min=-0.08, max=0.08))
        cp = dict(num_output=C, kernel_h=kh, kernel_w=D, weight_filler=dict(type="xavier"),
                  bias_filler=dict(type="constant"))
        bp = dict(scale_filler=dict(type="constant", value=1.0), shift_filler=dict(type="constant", value=1e-3))
        pp = dict(pool="MAX", kernel_h=L - kh + 1, kernel_w=1)
        self.branches = []
        for side in ("q", "a"):
            b = dict(side=side, idx=mk((N, L)), emb=mk(), x=mk((N, 1, L, D)), y=mk(), z=mk(), v=mk())
            b["embed"] = EmbedLayer(LayerParameter("Embed", name="embed_" + side, dtype=dtype, embed_param=ep))
            b["conv"] = ConvolutionLayer(LayerParameter("Convolution", name="conv0_" + side, dtype=dtype, convolution_param=cp))
            b["bn"] = BNLayer(LayerParameter("BN", name="bn0_" + side, dtype=dtype, bn_param=bp))
            b["pool"] = PoolingLayer(LayerParameter("Pooling", name="pool0_" + side, dtype=dtype, pooling_param=pp))
            b["tanh"] = TanHLayer(LayerParameter("TanH", name="tanh0_" + side, dtype=dtype))
            b["embed"].SetUp([b["idx"]], [b["emb"]])
            # Reshape layer (N,L,D) -> (N,1,L,D): the same memory under another shape (reshape_layer.cpp shares data and diff)
            b["x"].set_data(b["emb"].data.view(N, 1, L, D))
            b["x"].set_diff(b["emb"].diff.view(N, 1, L, D))
            b["conv"].SetUp([b["x"]], [b["y"]])
            b["conv"].handle.set_option(_lib.MMS_OPT_REUSE_FORWARD, 1)     # Net::ForwardBackward: bottoms unchanged
            b["bn"].SetUp([b["y"]], [b["z"]])
            b["pool"].SetUp([b["z"]], [b["v"]])
            b["tanh"].SetUp([b["v"]], [b["v"]])                            # in place, as in the script
            self.branches.append(b)
        bq, ba = self.branches
        for name in ("embed", "conv", "bn"):                               # shared by param name
            for j, blob in enumerate(bq[name].blobs):
                ba[name].blobs[j].ShareData(blob)
                ba[name].blobs[j].ShareDiff(blob)
        self.s = mk()
        self.sim = SimMatrixLayer(LayerParameter("SimMatrix", name="sim", dtype=dtype,
                                                 sim_matrix_param=dict(weight_filler=dict(type="xavier"))))
        self.sim.SetUp([bq["v"], ba["v"]], [self.s])
        self.sim.handle.set_option(_lib.MMS_OPT_REUSE_FORWARD, 1)
        # PairRankLoss bottoms: the two halves of the score column (views), and the labels
        self.s_pos, self.s_neg, self.label = mk((h, 1)), mk((h, 1)), mk((h, 1))
        self.s_pos.set_data(self.s.data[:h]); self.s_pos.set_diff(self.s.diff[:h])
        self.s_neg.set_data(self.s.data[h:]); self.s_neg.set_diff(self.s.diff[h:])
        self.loss_top = mk()
        self.loss = PairRankLossLayer(LayerParameter("PairRankLoss", name="loss", dtype=dtype,
                                                     pair_rank_loss_param=dict(margin=margin)))
        self.loss.SetUp([self.s_pos, self.s_neg, self.label], [self.loss_top])
        self._graph = None

    # -- parameters (shared blobs once, net order) --------------------------------------------
    def layers(self):
        bq, ba = self.branches
        return [bq["embed"], ba["embed"], bq["conv"], bq["bn"], ba["conv"], ba["bn"], self.sim]

    def params(self):
        bq = self.branches[0]
        return list(bq["embed"].blobs) + list(bq["conv"].blobs) + list(bq["bn"].blobs) + list(self.sim.blobs)

    def learnable_params(self):
        """Without BN's running mean / variance (lr_mult 0 in the script)."""
        bq = self.branches[0]
        return list(bq["embed"].blobs) + list(bq["conv"].blobs) + list(bq["bn"].blobs[:2]) + list(self.sim.blobs)

    def set_inputs(self, idx_q, idx_a, label):
        self.branches[0]["idx"].set_cpu_data(idx_q)
        self.branches[1]["idx"].set_cpu_data(idx_a)
        self.label.set_cpu_data(np.asarray(label).reshape(-1, 1))

    def ClearParamDiffs(self):
        for p in self.params():
            p.diff.zero_()

    # -- the step --------------------------------------------------------------------------------
    def Forward(self):
        bq, ba = self.branches
        for b in (bq, ba):                                     # declared order: embeds, then conv/bn per branch ...
            b["embed"].Forward([b["idx"]], [b["emb"]])
        for b in (bq, ba):
            b["conv"].Forward([b["x"]], [b["y"]])
            b["bn"].Forward([b["y"]], [b["z"]])
        for b in (bq, ba):
            b["pool"].Forward([b["z"]], [b["v"]])
        for b in (bq, ba):
            b["tanh"].Forward([b["v"]], [b["v"]])
        self.sim.Forward([bq["v"], ba["v"]], [self.s])
        return self.loss.Forward([self.s_pos, self.s_neg, self.label], [self.loss_top])

    def Backward(self):
        bq, ba = self.branches
        self.loss.Backward([self.loss_top], [True, True, False], [self.s_pos, self.s_neg, self.label])
        self.sim.Backward([self.s], [True, True], [bq["v"], ba["v"]])
        for b in (ba, bq):
            b["tanh"].Backward([b["v"]], [True], [b["v"]])
        for b in (ba, bq):
            b["pool"].Backward([b["v"]], [True], [b["z"]])
        for b in (ba, bq):                                     # ... so the question branch's BN backward runs last
            b["bn"].Backward([b["z"]], [True], [b["y"]])
            b["conv"].Backward([b["y"]], [True], [b["x"]])
        for b in (ba, bq):
            b["embed"].Backward([b["emb"]], [False], [b["idx"]])

    def ForwardBackward(self):
        loss = self.Forward()
        self.Backward()
        return loss

    # -- CUDA-graph replay (see MMSNet.capture) ----------------------------------------------------
    def capture(self, clear_diffs=True):
        def step():
            if clear_diffs:
                self.ClearParamDiffs()
            self.ForwardBackward()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        for l in (self.loss,):
            l.defer_loss_ = True                               # loss stays on the device: no host sync while capturing
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

    def loss_value(self):
        return float(self.loss_top.data.reshape(-1)[0].item())
