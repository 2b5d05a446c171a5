"""GPU tests of the data-parallel gradient exchange (csrc/exchange.cu, parallel.py, MMSNet.ForwardBackwardExchange).

Reference semantics: P2PSync::on_gradients_ready sums the solvers' flat gradient buffers and scales by 1/solver_count
(src/caffe/parallel.cpp:325-380, :377); on_start broadcasts the weights (:287-322); the root then runs
SGDSolver::ApplyUpdate (sgd_solver.cpp:102-116) with the AdaDelta rule (adadelta_solver.cu:5-26).

* ``virtual ranks``: `world` exchange objects on ONE device attached to each other by plain pointers (the same-process
  form of the C-ABI a P2PSync thread per GPU would use); every rank's kernel runs on its own stream.  Checked
  bit-exactly against the fixed-order float sum the kernel promises, and the fused solver tail against the
  single-GPU mms_adadelta_step on the averaged gradient.
* ``multigpu``: two real processes / GPUs (cudaIpc mapping, run when the box has >= 2 GPUs): the exchanged flat
  gradient of a sharded batch against one GPU computing the concatenated batch.
"""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import mms_answer_selection_b200 as mms  # noqa: E402
from mms_answer_selection_b200 import _lib, synth  # noqa: E402
from mms_answer_selection_b200.parallel import _RawCuda  # noqa: E402

c_p = ctypes.c_void_p


class VirtualWorld(object):
    """`world` ranks of the exchange on one device, attached by pointers; one stream per rank."""

    def __init__(self, world, count, dtype=np.float32, ctas=8, timeout_ms=4000):
        self.L = _lib.lib()
        self.world, self.count, self.dtype = world, count, np.dtype(dtype)
        self.elem = self.dtype.itemsize
        self.x = []
        for r in range(world):
            x = c_p()
            _lib.check(self.L.mms_exchange_create(ctypes.byref(x), r, world, count, self.elem, c_p(0)))
            _lib.check(self.L.mms_exchange_set_option(x, _lib.MMS_EXCHANGE_OPT_CTAS, ctas))
            _lib.check(self.L.mms_exchange_set_option(x, _lib.MMS_EXCHANGE_OPT_TIMEOUT_MS, timeout_ms))
            self.x.append(x)
        bases = []
        for x in self.x:
            b = c_p()
            _lib.check(self.L.mms_exchange_base(x, ctypes.byref(b)))
            bases.append(b.value)
        arr = (c_p * world)(*[c_p(b) for b in bases])
        for x in self.x:
            _lib.check(self.L.mms_exchange_attach_ptrs(x, arr, c_p(0)))
        self.streams = [torch.cuda.Stream() for _ in range(world)]
        self.data, self.diff = [], []
        for x in self.x:
            pd, pg = c_p(), c_p()
            _lib.check(self.L.mms_exchange_buffers(x, ctypes.byref(pd), ctypes.byref(pg)))
            self.data.append(torch.as_tensor(_RawCuda(pd.value, count, self.dtype), device="cuda"))
            self.diff.append(torch.as_tensor(_RawCuda(pg.value, count, self.dtype), device="cuda"))

    def each(self, fn):
        """fn(rank, exchange, stream pointer) issued for every rank on its own stream, then all are checked."""
        torch.cuda.synchronize()
        for r in range(self.world):
            fn(r, self.x[r], c_p(self.streams[r].cuda_stream))
        rcs = [self.L.mms_exchange_check(self.x[r], c_p(self.streams[r].cuda_stream)) for r in range(self.world)]
        torch.cuda.synchronize()
        return rcs

    def close(self):
        for x in self.x:
            self.L.mms_exchange_destroy(x)
        self.x = []


def _fixed_order_mean(parts, scale):
    acc = parts[0].copy()
    for p in parts[1:]:
        acc = acc + p                     # float32 adds in rank order, as the kernel does
    return acc * parts[0].dtype.type(scale)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_allreduce_virtual_ranks_bit_exact(world, dtype):
    count = 100_003                       # ragged: not a multiple of 16 bytes, not a multiple of world
    vw = VirtualWorld(world, count, dtype)
    try:
        rng = np.random.default_rng(world)
        parts = [rng.normal(0, 1, count).astype(dtype) for _ in range(world)]
        for r in range(world):
            vw.diff[r].copy_(torch.from_numpy(parts[r]))
        fn = vw.L.mms_exchange_allreduce_f32 if dtype == np.float32 else vw.L.mms_exchange_allreduce_f64
        real = ctypes.c_float if dtype == np.float32 else ctypes.c_double
        rcs = vw.each(lambda r, x, st: _lib.check(fn(x, st, 0, 0, count, real(1.0 / world))))
        assert rcs == [0] * world
        want = _fixed_order_mean(parts, 1.0 / world)
        for r in range(world):
            np.testing.assert_array_equal(vw.diff[r].cpu().numpy(), want)
        # a second call on the same channel (new epoch), a sub-range bucket on another channel
        for r in range(world):
            vw.diff[r].copy_(torch.from_numpy(parts[r]))
        b, e = 1024, 50_000
        rcs = vw.each(lambda r, x, st: _lib.check(fn(x, st, 1, b, e, real(0.5))))
        assert rcs == [0] * world
        want2 = _fixed_order_mean([p[b:e] for p in parts], 0.5)
        for r in range(world):
            got = vw.diff[r].cpu().numpy()
            np.testing.assert_array_equal(got[b:e], want2)
            np.testing.assert_array_equal(got[:b], parts[r][:b])       # outside the bucket: untouched
            np.testing.assert_array_equal(got[e:], parts[r][e:])
    finally:
        vw.close()


def test_exchange_is_replayable_in_a_cuda_graph():
    """The epoch lives in device memory: the same recorded launch is a new exchange on every replay."""
    world, count = 2, 4096
    vw = VirtualWorld(world, count)
    try:
        for r in range(world):
            vw.diff[r].fill_(float(r + 1))
        graphs = []
        for r in range(world):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=vw.streams[r]):
                _lib.check(vw.L.mms_exchange_allreduce_f32(vw.x[r], c_p(vw.streams[r].cuda_stream), 0, 0, count,
                                                           ctypes.c_float(1.0)))
            graphs.append(g)
        for it in range(3):
            for r in range(world):
                with torch.cuda.stream(vw.streams[r]):
                    graphs[r].replay()
            torch.cuda.synchronize()
            # sum doubles every replay: 3, 6, 12 ... on both ranks
            for r in range(world):
                assert torch.all(vw.diff[r] == 3.0 * 2 ** it).item()
        assert [vw.L.mms_exchange_check(x, c_p(0)) for x in vw.x] == [0, 0]
    finally:
        vw.close()


def test_missing_peer_is_a_fault_not_a_hang():
    vw = VirtualWorld(2, 1024, timeout_ms=200)
    try:
        _lib.check(vw.L.mms_exchange_allreduce_f32(vw.x[0], c_p(vw.streams[0].cuda_stream), 0, 0, 1024,
                                                   ctypes.c_float(0.5)))      # rank 1 never calls
        rc = vw.L.mms_exchange_check(vw.x[0], c_p(vw.streams[0].cuda_stream))
        assert rc == _lib.MMS_E_FAULT
        assert b"did not arrive" in vw.L.mms_last_error()
    finally:
        vw.close()


def test_broadcast_virtual_ranks():
    world, count = 4, 30_001
    vw = VirtualWorld(world, count)
    try:
        rng = np.random.default_rng(5)
        parts = [rng.normal(0, 1, count).astype(np.float32) for _ in range(world)]
        for r in range(world):
            vw.data[r].copy_(torch.from_numpy(parts[r]))
        rcs = vw.each(lambda r, x, st: _lib.check(vw.L.mms_exchange_broadcast(x, st, 2, 1)))
        assert rcs == [0] * world
        for r in range(world):
            np.testing.assert_array_equal(vw.data[r].cpu().numpy(), parts[1])
    finally:
        vw.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fused_adadelta_tail_matches_allreduce_then_solver_step(dtype):
    """mms_exchange_adadelta == all-reduce(avg) + mms_adadelta_step on every rank, bit for bit (the same adadelta_one),
    with per-blob multipliers, three iterations; gradients are left zeroed and the weights identical on every rank."""
    world = 4
    counts = [6000 * 50, 52, 2 * 50 * 50, 2 * 12 * 12]                 # W, b, M, B
    offs, total = mms.parallel.flat_layout(counts, np.dtype(dtype).itemsize)
    lr, dec = [1.0, 2.0, 1.0, 1.0], [0.0, 0.0, 1.0, 1.0]
    vw = VirtualWorld(world, total, dtype)
    h = _lib.Handle()
    try:
        tdt = torch.float32 if dtype == np.float32 else torch.float64
        rng = np.random.default_rng(11)
        w0 = rng.uniform(-0.08, 0.08, total).astype(dtype)
        for r in range(world):
            vw.data[r].copy_(torch.from_numpy(w0))
        ref_w = torch.from_numpy(w0.copy()).cuda()
        ref_hg, ref_hu = torch.zeros(total, dtype=tdt, device="cuda"), torch.zeros(total, dtype=tdt, device="cuda")
        ends = [offs[i + 1][0] if i + 1 < len(offs) else total for i in range(len(offs))]
        seg_end = (ctypes.c_longlong * 4)(*ends)
        seg_rate = (ctypes.c_double * 4)(*lr)
        seg_decay = (ctypes.c_double * 4)(*[5e-4 * d for d in dec])
        fn = vw.L.mms_exchange_adadelta_f32 if dtype == np.float32 else vw.L.mms_exchange_adadelta_f64
        step = vw.L.mms_adadelta_step_f32 if dtype == np.float32 else vw.L.mms_adadelta_step_f64
        real = ctypes.c_float if dtype == np.float32 else ctypes.c_double
        for it in range(3):
            parts = [rng.normal(0, 1e-3, total).astype(dtype) for _ in range(world)]
            for r in range(world):
                vw.diff[r].copy_(torch.from_numpy(parts[r]))
            rcs = vw.each(lambda r, x, st: _lib.check(fn(x, st, 0, 0, total, real(1.0 / world), seg_end, seg_rate,
                                                         seg_decay, 4, real(0.95), real(5e-7), 1)))
            assert rcs == [0] * world
            # the unfused composition on one GPU: summed gradient (rank order), then the solver step blob by blob
            g = torch.from_numpy(_fixed_order_mean(parts, 1.0)).cuda()
            for i, (off, n) in enumerate(offs):
                end = ends[i]
                _lib.check(step(h.ptr, c_p(ref_w[off:end].data_ptr()), c_p(g[off:end].data_ptr()),
                                c_p(ref_hg[off:end].data_ptr()), c_p(ref_hu[off:end].data_ptr()), end - off,
                                real(1.0 / world), real(5e-4 * dec[i]), real(0.95), real(5e-7), real(lr[i]), 1))
            torch.cuda.synchronize()
            for r in range(world):
                assert torch.equal(vw.data[r], ref_w), (it, r)
                assert not vw.diff[r].any().item()
        # the history is sharded: rank r holds the slice it owns
        n16 = total // (16 // np.dtype(dtype).itemsize)
        vn = 16 // np.dtype(dtype).itemsize
        for r in range(world):
            pg, pu = c_p(), c_p()
            _lib.check(vw.L.mms_exchange_history(vw.x[r], ctypes.byref(pg), ctypes.byref(pu)))
            hg = torch.as_tensor(_RawCuda(pg.value, total, dtype), device="cuda")
            lo, hi = n16 * r // world * vn, n16 * (r + 1) // world * vn
            assert torch.equal(hg[lo:hi], ref_hg[lo:hi])
    finally:
        vw.close()


def test_split_backward_equals_whole_backward():
    """mms_simcross_backward_bottoms + _params == mms_simcross_backward (same kernels, other order)."""
    N, L, D, mc, V = 512, 40, 300, 4, 3000
    d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
    outs = []
    for split in (False, True):
        net = mms.MMSNet(N, L, D, mc, V)
        net.set_params(d["W"], d["b"], d["M"], d["B"]); net.set_inputs(d["idx_q"], d["idx_a"])
        net.set_upstream_gradient(d["dS"])
        net.ClearParamDiffs()
        net.Forward()
        if split:
            net.sim.BackwardBottoms([net.S], [net.q, net.a])
            assert net.sim._split_pending
            net.embed_q.Backward([net.q], [False], [net.idx_q])      # other handles' work in between is fine
            net.sim.BackwardParams([net.S], [net.q, net.a])
        else:
            net.sim.Backward([net.S], [True, True], [net.q, net.a])
        torch.cuda.synchronize()
        outs.append([net.q.cpu_diff(), net.a.cpu_diff(), net.sim.blobs[0].cpu_diff(), net.sim.blobs[1].cpu_diff()])
    P = 128 // L
    groups = -(-N // P)
    whole = (groups - groups % 148) * P * L        # token rows of the pair groups that run as whole tiles on 148 SMs
    for a, b, name in zip(outs[0], outs[1], ("dq", "da", "dM", "dB")):
        # split-K / per-CTA partial sums land with red.global.add in arrival order; so do the dq / da rows of the pair
        # groups left over after the last whole wave, whose measures are spread over CTAs
        assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max(), name
        if name in ("dq", "da"):
            np.testing.assert_array_equal(a.reshape(-1, D)[:whole], b.reshape(-1, D)[:whole], err_msg=name)


def test_exchange_step_single_rank_equals_plain_step():
    """world 1: ForwardBackwardExchange (flat buffers, split backward, two buckets) against ForwardBackward."""
    N, L, D, mc, V = 96, 40, 300, 4, 2000
    d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
    res = []
    for use_exchange in (False, True):
        net = mms.MMSNet(N, L, D, mc, V)
        net.set_params(d["W"], d["b"], d["M"], d["B"]); net.set_inputs(d["idx_q"], d["idx_a"])
        net.set_upstream_gradient(d["dS"])
        if use_exchange:
            ex = mms.GradientExchange(net.params())
            assert ex.backend == "p2p" and net.embed_a.blobs[0].data.data_ptr() == ex.flat_data.data_ptr()
            g = net.capture_exchange_step(ex)
            g.replay()
            torch.cuda.synchronize()
        else:
            net.ClearParamDiffs(); net.ForwardBackward()
        res.append([p.cpu_diff().copy() for p in net.params()])
    for a, b in zip(res[0], res[1]):
        assert np.abs(a - b).max() <= 2e-6 * max(np.abs(a).max(), 1e-30)


# ---------------------------------------------------------------------------------------------- two real GPUs
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _dp_worker(rank, world, port, N, cfg, out, symmetric):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        L, D, mc, V = cfg
        full = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
        n = N // world
        sl = slice(rank * n, (rank + 1) * n)
        net = mms.MMSNet(n, L, D, mc, V)
        net.set_params(full["W"], full["b"], full["M"], full["B"])
        net.set_inputs(full["idx_q"][sl], full["idx_a"][sl])
        # each worker's loss is the mean over ITS pairs: dS of the global batch, times world
        net.set_upstream_gradient(full["dS"][sl] * world)
        ex = mms.GradientExchange(net.params(), symmetric=symmetric)
        ex.broadcast_params(0)
        g = net.capture_exchange_step(ex)
        g.replay()
        ex.check()
        out[rank] = dict(diff=[p.cpu_diff().copy() for p in net.params()], multicast=ex.multicast)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("symmetric", [False, True])
def test_multigpu_exchanged_gradient_equals_single_gpu_batch(symmetric):
    """2 ranks x N/2 pairs, exchanged (averaged) flat gradient vs ONE GPU on the N-pair batch: the 1/n of
    parallel.cpp:377 turns the sum of the per-rank means into the global-batch mean."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    N, cfg = 256, (40, 300, 4, 5000)
    L, D, mc, V = cfg
    full = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
    net = mms.MMSNet(N, L, D, mc, V)
    net.set_params(full["W"], full["b"], full["M"], full["B"]); net.set_inputs(full["idx_q"], full["idx_a"])
    net.set_upstream_gradient(full["dS"])
    net.ClearParamDiffs(); net.ForwardBackward()
    torch.cuda.synchronize()
    want = [p.cpu_diff().copy() for p in net.params()]
    del net
    mgr = mp.Manager(); out = mgr.dict()
    try:
        mp.spawn(_dp_worker, args=(2, _free_port(), N, cfg, out, symmetric), nprocs=2, join=True)
    except Exception as e:        # noqa: BLE001
        if symmetric:
            pytest.skip("symmetric memory / multicast not available here: %s" % str(e).splitlines()[-1][:200])
        raise
    for j, w in enumerate(want):
        np.testing.assert_array_equal(out[0]["diff"][j], out[1]["diff"][j])          # replicas identical
        err = np.abs(out[0]["diff"][j] - w).max() / max(np.abs(w).max(), 1e-30)
        assert err <= 1e-3, (j, err)                                                 # TF32 contractions, other split
    print("multigpu exchange ok; multicast=%s" % out[0]["multicast"])


def test_staged_tf32_operands_change_nothing():
    """MMS_OPT_STAGE_TF32: the gather writes the rounded operand copy SimCross reads; every result is bit-identical to
    the path that rounds q / a in a pass of its own, and the rounding kernel no longer sees q / a."""
    N, L, D, mc, V = 384, 40, 300, 4, 4000
    d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
    outs, rounds = [], []
    for stage in (False, True):
        net = mms.MMSNet(N, L, D, mc, V, stage_tf32=stage)
        net.set_params(d["W"], d["b"], d["M"], d["B"]); net.set_inputs(d["idx_q"], d["idx_a"])
        net.set_upstream_gradient(d["dS"])
        net.sim.handle.profile_enable(True)
        net.ClearParamDiffs(); net.ForwardBackward()
        torch.cuda.synchronize()
        rounds.append(net.sim.handle.profile_report().get("tf32_round_kernel", (0, 0.0))[1])
        outs.append([net.q.cpu_data(), net.S.cpu_data(), net.q.cpu_diff(), net.a.cpu_diff()])
        # a refilled bottom drops the staged copy: new ids, forward again, compare with a fresh net
        if stage:
            d2 = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V, seed=5)
            net.set_inputs(d2["idx_q"], d2["idx_a"])
            net.Forward()
            ref = mms.MMSNet(N, L, D, mc, V, stage_tf32=False)
            ref.set_params(d["W"], d["b"], d["M"], d["B"]); ref.set_inputs(d2["idx_q"], d2["idx_a"])
            ref.Forward()
            torch.cuda.synchronize()
            np.testing.assert_array_equal(net.S.cpu_data(), ref.S.cpu_data())
    for a, b, name in zip(outs[0], outs[1], ("q", "S", "dq", "da")):
        np.testing.assert_array_equal(a, b, err_msg=name)
    assert rounds[1] < 0.8 * rounds[0]          # only M is rounded when q / a arrive staged (device time: a loose bound)


def test_stage_only_embed_tops_same_results_and_loud_failures():
    """MMS_OPT_STAGE_ONLY (MMSNet(keep_embed_tops=False)): the gather writes nothing but the TF32 operand copy.  S, dq,
    da and dM are bit-identical to the net that also writes the fp32 tops (same operands, same kernels), the fp32 top
    buffers are provably untouched, and a consumer that cannot use the staged copy refuses instead of reading them."""
    from mms_answer_selection_b200 import _lib
    N, L, D, mc, V = 384, 40, 300, 4, 4000
    d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V)
    outs = []
    for keep in (True, False):
        net = mms.MMSNet(N, L, D, mc, V, keep_embed_tops=keep)
        net.set_params(d["W"], d["b"], d["M"], d["B"]); net.set_inputs(d["idx_q"], d["idx_a"])
        net.set_upstream_gradient(d["dS"])
        net.q.data.fill_(-7.0); net.a.data.fill_(-7.0)
        net.ClearParamDiffs(); net.ForwardBackward()
        torch.cuda.synchronize()
        outs.append([net.S.cpu_data(), net.q.cpu_diff(), net.a.cpu_diff(), net.sim.blobs[0].cpu_diff()])
        if keep:
            assert not (net.q.cpu_data() == -7.0).any()
            w_keep = net.embed_q.blobs[0].cpu_diff()
        else:
            assert (net.q.cpu_data() == -7.0).all() and (net.a.cpu_data() == -7.0).all()      # never written
            w_only = net.embed_q.blobs[0].cpu_diff()
            # fp32 math cannot use the TF32 copy: refuse, do not read the unwritten top
            net.sim.handle.set_option(_lib.MMS_OPT_REUSE_FORWARD, 0)
            net.sim.set_math(_lib.MMS_MATH_FP32)
            with pytest.raises(Exception, match="STAGE_ONLY"):
                net.sim.Forward([net.q, net.a], [net.S])
            with pytest.raises(Exception, match="STAGE_ONLY"):
                net.sim.Backward([net.S], [True, True], [net.q, net.a])
            # a blanket invalidation makes the staged copy stale but does not materialise the tops: still a refusal
            net.sim.set_math(_lib.MMS_MATH_TF32)
            _lib.lib().mms_invalidate_caches()
            with pytest.raises(Exception, match="STAGE_ONLY"):
                net.sim.Forward([net.q, net.a], [net.S])
            net.Forward()                                          # gathering again repairs it
            torch.cuda.synchronize()
            np.testing.assert_array_equal(net.S.cpu_data(), outs[0][0])
    for a, b, name in zip(outs[0], outs[1], ("S", "dq", "da")):
        np.testing.assert_array_equal(a, b, err_msg=name)
    for a, b, name in ((outs[0][3], outs[1][3], "dM"), (w_keep, w_only, "dW")):                # float atomics: order differs
        assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max(), name
