"""MultiModalNet (BASELINE config C5: modalities x SimMatrix -> Concat -> FM -> Slice -> PairRankLoss) against the
chained CPU oracle of the same layers (sim_matrix_layer.cpp:53-95, fm_layer.cpp:33-99, pair_rank_loss_layer.cpp:26-84)."""
import numpy as np
import pytest
import torch

from conftest import scaled_err

pytestmark = pytest.mark.gpu

import mms_answer_selection_b200 as mms  # noqa: E402
from mms_answer_selection_b200 import synth  # noqa: E402
from mms_answer_selection_b200.multimodal import MultiModalNet  # noqa: E402
from oracle import cport  # noqa: E402


def _inputs(N, K1, K2, C, seed=3):
    rng = np.random.default_rng(seed)
    qs, as_, Ws = [], [], []
    for m in range(C):
        q, a, W = synth.make_sentence_vectors(N, K1, K2, seed=seed + m)
        qs.append(q); as_.append(a); Ws.append(W)
    label = (rng.random(N // 2) < 0.5).astype(np.float32)
    return qs, as_, Ws, label


@pytest.mark.parametrize("graph", [False, True])
def test_multimodal_net_vs_oracle(graph):
    N, K1, K2, C = 192, 96, 128, 4
    qs, as_, Ws, label = _inputs(N, K1, K2, C)
    net = MultiModalNet(N, K1, K2, C)
    for m in range(C):
        net.sim[m].blobs[0].set_cpu_data(Ws[m])
    net.fm.blobs[0].set_cpu_data(np.array([0.25], dtype=np.float32))
    net.set_inputs(qs, as_, label)
    if graph:
        net.capture()
        net.replay()
        torch.cuda.synchronize()
        loss = net.loss_value()
    else:
        net.ClearParamDiffs()
        loss = net.ForwardBackward()
    # oracle chain
    s = [cport.simmatrix_forward(qs[m], as_[m], Ws[m])[0] for m in range(C)]
    x = np.stack([v.reshape(-1) for v in s], axis=1).reshape(N, C, 1).astype(np.float32)
    y = cport.fm_forward(x, np.array([0.25], dtype=np.float32))
    h = N // 2
    ref_loss, ordered, similar = cport.pairrankloss_forward(y[:h].copy(), y[h:].copy(), label.reshape(-1, 1), 1.0)
    dyp, dyn = cport.pairrankloss_backward(label.reshape(-1, 1), ordered, similar, 1.0)
    dy = np.concatenate([dyp, dyn]).reshape(-1)
    dx, db = cport.fm_backward(x, dy)
    assert abs(loss - float(ref_loss)) <= 2e-3 * max(abs(float(ref_loss)), 1e-6)
    assert scaled_err(net.y.cpu_data(), y) <= 1e-3
    # db = sum(dy) = sum(dy+ + dy-) is zero up to rounding (PairRankLoss sends opposite gradients to its two bottoms)
    assert abs(float(net.fm.blobs[0].cpu_diff()[0]) - float(db[0])) <= 1e-6
    for m in range(C):
        dW = np.zeros_like(Ws[m])
        dW, dq, da = cport.simmatrix_backward(qs[m], as_[m], Ws[m], dx[:, m, 0].copy(), dW)
        # the hinge is a step function of the TF32-rounded scores: compare where the forward agrees on its side
        assert scaled_err(net.sim[m].blobs[0].cpu_diff(), dW) <= 2e-3, m
        assert scaled_err(net.q[m].cpu_diff(), dq) <= 2e-3, m
        assert scaled_err(net.a[m].cpu_diff(), da) <= 2e-3, m
