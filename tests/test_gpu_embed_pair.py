"""mms_embed_backward_pair / mms_embed_plan_pair (csrc/embed_sorted.cu): the scatter-add of two Embed layers that share a
table, token rows grouped by id.  Oracle: numpy's np.add.at in float64 -- the reference's atomicAdd loop
(src/caffe/layers/embed_layer.cu:29-39) summed exactly; float atomics land in arrival order, hence the 1e-5 tolerances."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mms_answer_selection_b200 import _lib  # noqa: E402

p = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def oracle(idx0, d0, idx1, d1, W0, b0, V):
    dW = W0.astype(np.float64).copy()
    db = b0.astype(np.float64).copy()
    for idx, d in ((idx0, d0), (idx1, d1)):
        if idx is None:
            continue
        ids = idx.astype(np.int64)
        ok = (ids >= 0) & (ids < V)
        np.add.at(dW, ids[ok], d[ok].astype(np.float64))
        db += d.astype(np.float64).sum(0)            # the reference's bias gradient sums every row (embed_layer.cu:72-76)
    return dW, db


def make_ids(rng, M, V, kind):
    if kind == "uniform":
        return rng.integers(0, V, size=M)
    if kind == "padded":                             # centre-padded sentences: long runs of the pad id
        ids = rng.integers(0, V - 1, size=M)
        ids[rng.random(M) < 0.55] = V - 1
        return ids
    if kind == "hot":                                # a few ids carry almost everything: runs of thousands of rows
        hot = rng.integers(0, V, size=5)
        ids = hot[rng.integers(0, 5, size=M)]
        cold = rng.random(M) < 0.1
        ids[cold] = rng.integers(0, V, size=int(cold.sum()))
        return ids
    raise ValueError(kind)


def run_pair(h, idx0, d0, idx1, d1, W0, b0, D, V, plan_first=False):
    dev = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
    i0, g0, i1, g1 = dev(idx0), dev(d0), dev(idx1), dev(d1)
    dW, db = dev(W0), dev(b0)
    M0 = 0 if idx0 is None else idx0.size
    M1 = 0 if idx1 is None else idx1.size
    L = _lib.lib()
    if plan_first:
        _lib.check(L.mms_embed_plan_pair_f32(h.ptr, p(i0), M0, p(i1), M1, V))
    _lib.check(L.mms_embed_backward_pair_f32(h.ptr, p(i0), p(g0), M0, p(i1), p(g1), M1, p(dW), p(db), D, V))
    torch.cuda.synchronize()
    return dW.cpu().numpy(), db.cpu().numpy()


@pytest.mark.parametrize("M0,M1,D,V,kind", [
    (20000, 20000, 300, 5000, "padded"),    # the QA net's shape in small
    (17777, 21233, 300, 50, "uniform"),     # every id hundreds of times, ragged sizes
    (30000, 9000, 300, 3000, "hot"),        # runs far beyond one chunk
    (33000, 0, 300, 2000, "padded"),        # one blob only
    (31500, 11700, 52, 400, "uniform"),     # one 16-byte group per lane, partly filled
    (19000, 21000, 512, 300, "padded"),     # the widest row the grouped kernels take
    (16385, 16383, 300, 100000, "uniform"), # just at the row threshold, almost every id once
    (5, 3, 300, 1000, "uniform"),           # below it: the per-layer kernels
])
@pytest.mark.parametrize("plan_first", [False, True])
def test_pair_backward_matches_exact_scatter_add(M0, M1, D, V, kind, plan_first):
    rng = np.random.default_rng(M0 + 3 * M1 + D)
    idx0 = make_ids(rng, M0, V, kind).astype(np.float32)
    idx1 = make_ids(rng, M1, V, kind).astype(np.float32) if M1 else None
    d0 = rng.standard_normal((M0, D)).astype(np.float32)
    d1 = rng.standard_normal((M1, D)).astype(np.float32) if M1 else None
    W0 = rng.standard_normal((V, D)).astype(np.float32)          # accumulate semantics: dW and db start non-zero
    b0 = rng.standard_normal(D).astype(np.float32)
    h = _lib.Handle()
    dW, db = run_pair(h, idx0, d0, idx1, d1, W0, b0, D, V, plan_first)
    rW, rb = oracle(idx0, d0, idx1, d1, W0, b0, V)
    scale = max(1.0, np.abs(rW).max())
    assert np.abs(dW - rW).max() <= 2e-6 * scale * max(1.0, np.sqrt(M0 + M1) / 30), "dW"
    assert np.abs(db - rb).max() <= 1e-5 * max(1.0, np.abs(rb).max()), "db"
    untouched = np.setdiff1d(np.arange(V), np.concatenate([idx0, idx1 if idx1 is not None else []]).astype(np.int64))
    np.testing.assert_array_equal(dW[untouched], W0[untouched])                 # rows nobody refers to: bit-identical


def test_pair_backward_agrees_with_the_two_per_layer_calls_and_flags_bad_ids():
    M, D, V = 24000, 300, 4000
    rng = np.random.default_rng(3)
    idx0 = make_ids(rng, M, V, "padded").astype(np.float32)
    idx1 = make_ids(rng, M, V, "padded").astype(np.float32)
    d0 = rng.standard_normal((M, D)).astype(np.float32)
    d1 = rng.standard_normal((M, D)).astype(np.float32)
    zW, zb = np.zeros((V, D), np.float32), np.zeros(D, np.float32)
    h = _lib.Handle()
    dW, db = run_pair(h, idx0, d0, idx1, d1, zW, zb, D, V)
    dev = lambda x: torch.from_numpy(x).cuda()
    W2, b2 = dev(zW), dev(zb)
    g = [dev(x) for x in (idx0, d0, idx1, d1)]                   # (kept alive: the calls are asynchronous)
    for i in (0, 2):
        _lib.check(_lib.lib().mms_embed_backward_f32(h.ptr, p(g[i]), p(g[i + 1]), p(W2), p(b2), M, D, V))
    torch.cuda.synchronize()
    assert np.abs(dW - W2.cpu().numpy()).max() <= 1e-5 * np.abs(dW).max()
    assert np.abs(db - b2.cpu().numpy()).max() <= 1e-5 * np.abs(db).max()
    # dW == NULL (param_propagate_down false) still sums the bias gradient; dbias == NULL alone is fine too
    L = _lib.lib()
    b3 = dev(zb)
    _lib.check(L.mms_embed_backward_pair_f32(h.ptr, p(g[0]), p(g[1]), M, p(g[2]), p(g[3]), M, p(None), p(b3), D, V))
    torch.cuda.synchronize()
    assert np.abs(b3.cpu().numpy() - db).max() <= 1e-5 * np.abs(db).max()
    # an id outside [0, V): skipped, and reported by mms_check_faults like the per-layer kernels
    bad = idx0.copy(); bad[17] = V + 5; bad[99] = -1
    W4, b4, gbad = dev(zW), dev(zb), dev(bad)
    _lib.check(L.mms_embed_backward_pair_f32(h.ptr, p(gbad), p(g[1]), M, p(None), p(None), 0, p(W4), p(b4), D, V))
    with pytest.raises(_lib.MMSError):
        h.check_faults()
    rW, _ = oracle(bad, d0, None, None, zW, zb, V)
    assert np.abs(W4.cpu().numpy() - rW).max() <= 1e-5 * np.abs(rW).max()


def test_net_step_with_grouped_scatter_equals_per_layer_scatter():
    import mms_answer_selection_b200 as mms
    from mms_answer_selection_b200 import synth
    N, L, D, mc, V = 450, 40, 300, 4, 3000           # 36 000 token rows: above the grouped kernels' threshold
    d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
    outs = []
    for grouped in (False, True):
        net = mms.MMSNet(N, L, D, mc, V, grouped_scatter=grouped)
        net.set_params(d["W"], d["b"], d["M"], d["B"]); net.set_inputs(d["idx_q"], d["idx_a"])
        net.set_upstream_gradient(d["dS"])
        net.capture(with_loss=True, clear_diffs=True)          # graph with the plan on its side branch
        for _ in range(2):
            net.replay()
        torch.cuda.synchronize()
        outs.append([b.cpu_diff() for b in net.params()])
        net.embed_q.handle.profile_enable(True)                # (the profile sees eager launches only)
        net.ClearParamDiffs(); net.ForwardBackward()
        torch.cuda.synchronize()
        assert ("embed_backward_short_runs" in net.embed_q.handle.profile_report()) == grouped
    for a, b in zip(*outs):
        assert np.abs(a - b).max() <= 1e-5 * max(np.abs(a).max(), 1e-30)


def test_prefetched_inputs_arrive_one_replay_later():
    """capture(host_inputs, prefetch_inputs=True): replay k computes on the ids replay k-1 fetched and fetches the next."""
    import mms_answer_selection_b200 as mms
    from mms_answer_selection_b200 import synth
    N, L, D, mc, V = 64, 40, 300, 4, 2000
    d1 = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V, seed=1)
    d2 = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V, seed=2)
    losses = {}
    for name, d in (("one", d1), ("two", d2)):                     # plain steps on each batch
        net = mms.MMSNet(N, L, D, mc, V)
        net.set_params(d1["W"], d1["b"], d1["M"], d1["B"]); net.set_inputs(d["idx_q"], d["idx_a"])
        net.set_upstream_gradient(d1["dS"])
        net.ClearParamDiffs(); losses[name] = float(net.ForwardBackward())
    net = mms.MMSNet(N, L, D, mc, V)
    net.set_params(d1["W"], d1["b"], d1["M"], d1["B"]); net.set_inputs(d1["idx_q"], d1["idx_a"])
    net.set_upstream_gradient(d1["dS"])
    hq = torch.from_numpy(d2["idx_q"].copy()).pin_memory(); ha = torch.from_numpy(d2["idx_a"].copy()).pin_memory()
    net.capture(with_loss=True, clear_diffs=True, host_inputs=(hq, ha), prefetch_inputs=True)
    net.set_inputs(d1["idx_q"], d1["idx_a"])                       # prime: the first replay computes on batch one ...
    torch.cuda.synchronize()
    first = net.replay_from_host()                                 # ... and fetches batch two
    hq.copy_(torch.from_numpy(d1["idx_q"])); ha.copy_(torch.from_numpy(d1["idx_a"]))
    second = net.replay_from_host()                                # computes on batch two, fetches batch one
    third = net.replay_from_host()
    assert abs(first - losses["one"]) <= 1e-5 * abs(losses["one"])
    assert abs(second - losses["two"]) <= 1e-5 * abs(losses["two"])
    assert abs(third - losses["one"]) <= 1e-5 * abs(losses["one"])
    assert abs(losses["one"] - losses["two"]) > 1e-3 * abs(losses["one"])       # the two batches do differ


def test_prepared_weights_match_and_follow_a_parameter_write():
    """mms_simcross_prepare: M rounded ahead of the forward gives the same S bit for bit; a parameter written after the
    prepare (Blob.set_cpu_data -> mms_invalidate_caches) is what the forward uses, not the stale rounded copy."""
    import mms_answer_selection_b200 as mms
    from mms_answer_selection_b200 import synth
    N, L, D, mc, V = 64, 40, 300, 4, 2000
    d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
    net = mms.MMSNet(N, L, D, mc, V)
    net.set_params(d["W"], d["b"], d["M"], d["B"]); net.set_inputs(d["idx_q"], d["idx_a"])
    net.Forward()                                                  # sizes the workspace, plain path
    S0 = net.S.cpu_data().copy()
    net.sim.handle.profile_enable(True)
    net.sim.Prepare([net.q, net.a]); net.Forward()
    torch.cuda.synchronize()
    rep = net.sim.handle.profile_report()
    assert rep["tf32_round_kernel"][0] == 1                        # the prepare's launch only: the forward skipped its own
    np.testing.assert_array_equal(net.S.cpu_data(), S0)
    net.sim.Prepare([net.q, net.a])
    net.sim.blobs[0].set_cpu_data(2.0 * d["M"])                    # after the prepare
    net.Forward()
    torch.cuda.synchronize()
    ref = mms.MMSNet(N, L, D, mc, V)
    ref.set_params(d["W"], d["b"], 2.0 * d["M"], d["B"]); ref.set_inputs(d["idx_q"], d["idx_a"])
    ref.Forward()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(net.S.cpu_data(), ref.S.cpu_data())


def test_one_staging_buffer_per_handle_never_serves_a_stale_copy():
    """MMS_OPT_STAGE_TF32 on ONE handle used for two Embed forwards: the second gather overwrites the staging buffer, so the
    first top must no longer resolve to it -- SimCross then rounds that top itself and the scores stay right."""
    from mms_answer_selection_b200 import synth
    N, L, D, mc, V = 20, 40, 300, 4, 500
    d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    W, b, Mw, B = dev(d["W"]), dev(d["b"]), dev(d["M"]), dev(d["B"])
    iq, ia = dev(d["idx_q"]), dev(d["idx_a"])
    outs = []
    for stage in (0, 1):
        he, hs = _lib.Handle(), _lib.Handle()
        he.set_option(_lib.MMS_OPT_STAGE_TF32, stage)
        q = torch.empty((N, L, D), device="cuda"); a = torch.empty((N, L, D), device="cuda")
        S = torch.empty((N, mc, L, L), device="cuda")
        Lb = _lib.lib()
        _lib.check(Lb.mms_embed_forward_f32(he.ptr, p(iq), p(W), p(b), p(q), N * L, D, V))
        _lib.check(Lb.mms_embed_forward_f32(he.ptr, p(ia), p(W), p(b), p(a), N * L, D, V))      # same handle
        _lib.check(Lb.mms_simcross_forward_f32(hs.ptr, 2, p(q), p(a), p(Mw), p(B), p(S), None, None, N, L, L, D, mc))
        torch.cuda.synchronize()
        outs.append(S.cpu().numpy())
    np.testing.assert_array_equal(outs[0], outs[1])
