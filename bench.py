#!/usr/bin/env python
"""Benchmark of the MMS hot path on B200 (see the contract in DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3]

A "step" is one forward+backward pass of the hot path (Embed x2 -> SimCross mode 2 ->
loss plumbing -> SimCross backward -> Embed scatter-add) over one batch of synthetic
TREC-QA-shaped QA pairs.  Workload by N, as BASELINE.json's configs assign them:
  N = 1      configs[1] (C2): the reference's batch of 50 QA pairs per step, q/a length 40,
             300-d embeddings, mesure_count 4, V = 60002 (an epoch is 1069 such steps);
  N = 2,4,8  configs[2] (C3): a GLOBAL batch of 4096 QA pairs sharded over the ranks (strong
             scaling), the flat gradient buffer all-reduced over NCCL and scaled by 1/N inside
             the step.
--workload overrides.  The line also carries `extra`: candidate scores/s (configs[3], candidates
sharded over the ranks), and at N = 1 the C3 and C5 single-GPU figures.

Prints ONE JSON line (rank 0).  `value` = QA pairs/s with inputs resident in HBM;
`e2e` = the same through the public API with pinned-host inputs, H2D of the step's token
ids and D2H of the step's loss inside the timed region.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "qa_pairs_per_sec_fwd_bwd"
UNIT = "QA pairs/s"


def flops_per_pair(L, D, mc):
    """Algorithmic FLOPs of SimCross mode 2 per QA pair, fwd+bwd:
    (6 L D^2 + 8 L^2 D) per measure."""
    return mc * (6.0 * L * D * D + 8.0 * L * L * D)


def kernel_flop_shares(L, D, mc):
    """How the algorithmic FLOPs of one (pair, measure) -- SURVEY.md 8(d): fwd 2LD(D+L), bwd minimum
    2LD^2 (A M^T) + 3 * 2L^2D (dQ, dA, dS A) + 2LD^2 (dM) -- are attributed to the kernels that produce them
    (DESIGN.md 3.1).  The dA kernel re-derives its D x D product instead of reading T from the forward, so it is
    credited with the 2L^2D of dA only; the shares add up to 6LD^2 + 8L^2D."""
    return {
        "simcross2_fwd_fused_kernel": mc * 2.0 * L * D * (D + L),
        "simcross2_bwd_fused_kernel<dQ>": mc * (2.0 * L * D * D + 4.0 * L * L * D),   # A M^T, dQ and U = dS A
        "simcross2_bwd_fused_kernel<dA>": mc * 2.0 * L * L * D,
        "tc_gemm_tma_kernel": mc * 2.0 * L * D * D,                                   # dM = Q^T U (small batches)
        "simcross2_dm_kernel": mc * 2.0 * L * D * D,                                  # dM = Q^T U (tc/simcross_dm.cu)
    }


def workload_config(name, world=1):
    """Per-rank configuration: C3 fixes the GLOBAL batch (4096), so a rank gets 4096 / world pairs."""
    from mms_answer_selection_b200 import synth
    c = dict(synth.CONFIGS[name])
    c["global_N"] = c["N"]
    if name == "c3":
        c["N"] = max(1, c["N"] // world)
    else:
        c["global_N"] = c["N"] * world
    return c


def pick_workload(args, world):
    return args.workload or ("c2" if world == 1 else "c3")


# ---------------------------------------------------------------------------- reference arm
def ref_step_builder(cfg, pairs, threads):
    """One fwd+bwd step over `pairs` QA pairs through the reference's own layer code
    (oracle/_ref), or through the C port of it when /root/reference was never built."""
    from mms_answer_selection_b200 import synth
    d = synth.make_qa_batch(N=pairs, L=cfg["L"], D=cfg["D"], mc=cfg["mc"], V=cfg["V"])
    from oracle import refbind
    if refbind.ref_available():
        refbind.set_ref_blas_threads(threads)
        ep = {"num_output": cfg["D"], "input_dim": cfg["V"]}
        eq = refbind.RefLayer("Embed", [d["idx_q"]], ep)
        ea = refbind.RefLayer("Embed", [d["idx_a"]], ep)
        for e in (eq, ea):
            e.write("blob", 0, d["W"]); e.write("blob", 1, d["b"])
        eq.forward(); ea.forward()
        sim = refbind.RefLayer("SimCross", [eq.read("top", 0), ea.read("top", 0)],
                               {"dist_mode": 2, "mesure_count": cfg["mc"], "loss_weight": 1.0})
        sim.write("blob", 0, d["M"]); sim.write("blob", 1, d["B"])
        sim.forward()
        sim.write("top", 0, d["dS"], diff=True)

        def step():
            eq.forward(); ea.forward()
            sim.write("bottom", 0, eq.read("top", 0)); sim.write("bottom", 1, ea.read("top", 0))
            loss = sim.forward()
            sim.backward([True, True])
            eq.write("top", 0, sim.read("bottom", 0, diff=True), diff=True)
            ea.write("top", 0, sim.read("bottom", 1, diff=True), diff=True)
            eq.backward([False]); ea.backward([False])
            return loss
        return step, "reference"
    from oracle import cport

    def step():
        q = cport.embed_forward(d["idx_q"], d["W"], d["b"]); a = cport.embed_forward(d["idx_a"], d["W"], d["b"])
        S, _, _ = cport.simcross_forward(2, q, a, d["M"], d["B"])
        dq, da, dM, dB = cport.simcross_backward(2, q, a, d["M"], S, d["dS"])
        dW, db = np.zeros_like(d["W"]), np.zeros_like(d["b"])
        cport.embed_backward(d["idx_q"], dq, dW, db); cport.embed_backward(d["idx_a"], da, dW, db)
        return float((S * d["dS"]).sum())
    return step, "port"


def time_reference(cfg, steps, warmup, budget_s):
    """Times the CPU implementation on a bounded sample: calibrates seconds/pair, then
    sizes the per-step sample so that (steps+warmup) steps fit in `budget_s`."""
    cores = os.cpu_count() or 1
    probe_pairs = 2
    step, kind = ref_step_builder(cfg, probe_pairs, cores)
    step()
    t0 = time.perf_counter(); step(); t_pair = (time.perf_counter() - t0) / probe_pairs
    full = cfg["N"]
    pairs = int(max(1, min(full, budget_s / max(steps + warmup, 1) / max(t_pair, 1e-9))))
    step, kind = ref_step_builder(cfg, pairs, cores)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    used = cores if kind == "reference" else 1
    return dict(value=pairs * steps / dt, ms_per_step=dt / steps * 1e3, pairs=pairs, kind=kind, cores=used,
                sample="%d steps of %d QA pairs (of the %d-pair batch), Embed x2 + SimCross(mode 2) fwd+bwd, "
                       "%s, BLAS threads=%d" % (steps, pairs, full,
                                                "reference layer code compiled in place (oracle/_ref)"
                                                if kind == "reference" else "C port (oracle/mms_oracle.c)", used))


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = pick_workload(args, args.gpus)
    cfg = workload_config(wl, args.gpus)
    r = time_reference(cfg, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if wl == "c3" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(wl, cfg), "pairs_per_step": r["pairs"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_name(name, cfg):
    return ("%s: synthetic TREC-QA-shaped batch, %d QA pairs/step/GPU (%d global), q/a len %d, %d-d embeddings, "
            "mesure_count %d, V=%d; Embed x2 -> SimCross(mode 2) fwd+bwd"
            % (name.upper(), cfg["N"], cfg["global_N"], cfg["L"], cfg["D"], cfg["mc"], cfg["V"]))


# ---------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------- our arm
def main_ours(args):
    import torch
    import torch.distributed as dist

    import mms_answer_selection_b200 as mms
    from mms_answer_selection_b200 import synth
    from mms_answer_selection_b200.parallel import GradientExchange

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = pick_workload(args, world)
    cfg = workload_config(wl, world)
    N, L, D, mc, V = cfg["N"], cfg["L"], cfg["D"], cfg["mc"], cfg["V"]

    d = synth.make_qa_batch(N=N, L=L, D=D, mc=mc, V=V, seed=synth.SEED + rank)
    net = mms.MMSNet(N, L, D, mc, V)
    w = synth.make_qa_batch(N=1, L=L, D=D, mc=mc, V=V, seed=synth.SEED)      # same weights on every rank
    net.set_params(w["W"], w["b"], w["M"], w["B"])
    net.set_inputs(d["idx_q"], d["idx_a"])
    net.set_upstream_gradient(d["dS"])
    exch = GradientExchange(net.params()) if world > 1 else None
    if exch:
        exch.broadcast_params(0)
    host_q = torch.from_numpy(d["idx_q"]).pin_memory()
    host_a = torch.from_numpy(d["idx_a"]).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2
    handles = [net.embed_q.handle, net.embed_a.handle, net.sim.handle]

    # The whole step (ClearParamDiffs + Forward + loss + Backward) is recorded once as a CUDA graph and
    # replayed: issued launch by launch from the host the ~15 short kernels are launch-bound.
    l0 = sum(h.launch_count() for h in handles)
    net.capture(with_loss=True, clear_diffs=True)
    launches_per_step = (sum(h.launch_count() for h in handles) - l0) // 3      # 2 warm-up passes + the capture
    net.capture(with_loss=True, clear_diffs=True, host_inputs=(host_q, host_a))  # + H2D / D2H nodes for the e2e step

    def device_step():
        net.replay(read_loss=False)
        if exch:
            exch.allreduce()

    def e2e_step():
        if not exch:
            # one graph launch: H2D of this step's inputs (pinned), the step, D2H of the loss (4 bytes), then a sync
            return net.replay_from_host()
        net.set_inputs_from_pinned(host_q, host_a)      # H2D of this step's inputs
        net.replay(read_loss=False)
        exch.allreduce()
        return float(net.sim.loss_dev_[0].item())       # D2H of the step's loss (4 bytes) + sync

    def eager_step():
        net.ClearParamDiffs()          # Net::ClearParamDiffs (solver.cpp:203): zeroes the V x D diff too
        net.ForwardBackward(with_loss=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """Per-step CUDA-event timing with an L2 flush (untimed) between steps."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for e0, e1 in evs:
            flush.fill_(1)
            e0.record()
            fn()
            e1.record()
        barrier()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        device_step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    total_ms = timed(device_step, args.steps)
    launches = launches_per_step * args.steps

    for _ in range(3):
        e2e_step()
    e2e_ms = timed(e2e_step, args.steps)
    clocks = sampler.stop()          # sampled every 5 ms over both timed regions (device-resident and end-to-end)

    # roofline of the dominant kernel, measured live with CUDA events around each launch of an
    # eagerly issued step.  The GPU is first given ~0.5 ms of other work (L2 flushes) so that every
    # launch of the step is already queued when it runs: the events then bracket device time, not
    # host launch latency.
    sim_state = net.sim.defer_loss_
    net.sim.defer_loss_ = True
    from mms_answer_selection_b200 import _lib as _mmslib
    net.sim.handle.set_option(_mmslib.MMS_OPT_CONCURRENCY, 0)     # one kernel at a time: the events bracket it alone
    for h in handles:
        h.profile_enable(True)
    prof_steps = min(args.steps, 20)
    for _ in range(prof_steps):
        for _ in range(8):
            flush.fill_(1)
        eager_step()
        torch.cuda.synchronize()
    net.sim.defer_loss_ = sim_state
    net.sim.handle.set_option(_mmslib.MMS_OPT_CONCURRENCY, 1)
    prof = {}
    for h in handles:
        for k, (n, ms) in h.profile_report().items():
            pn, pms = prof.get(k, (0, 0.0))
            prof[k] = (pn + n, pms + ms)
        h.profile_enable(False)
    step_ms_prof = sum(ms for _, ms in prof.values()) / prof_steps
    dom = max(prof.items(), key=lambda kv: kv[1][1])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roof = roofline_for(dom[0], dom[1], prof, prof_steps, cfg, peaks, wl)

    value = N * world * args.steps / (total_ms / 1e3)
    e2e_value = N * world * args.steps / (e2e_ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if wl == "c3" else "weak", "vs_baseline": None,
        "dtype": "f32 (TF32 tensor-core contractions, fp32 accumulate)",
        "data": "synthetic",
        "config": {"workload": workload_name(wl, cfg), "parallelism": "dp%d" % world,
                   "l2": "256 MiB L2 flush between timed steps (inputs+table < L2)",
                   "launch": "step recorded as one CUDA graph (%d kernels of libmms_b200.so per step)" % launches_per_step,
                   "grad_exchange": "one NCCL all-reduce (ncclAvg) of the flat gradient buffer" if world > 1 else "none"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(host_q.numel() * 4 + host_a.numel() * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernels_ms_per_step": {k: round(ms / prof_steps, 5) for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])},
        "kernel_step_ms_sum": step_ms_prof,
        "hbm_kernels": hbm_kernel_table(prof, prof_steps, cfg, peaks),
    }
    if not args.no_extra:
        line["extra"] = extras(world, rank, flush)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = time_reference(cfg, steps=3, warmup=1, budget_s=20.0)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                "sample": r["sample"]}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _max_ms(ms, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _time_ms(fn, iters, flush, world):
    """CUDA-event time of `iters` calls (L2 flushed before each, untimed), max over ranks."""
    import torch
    import torch.distributed as dist
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = []
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return _max_ms(sum(a.elapsed_time(b) for a, b in evs), world) / iters


def extras(world, rank, flush):
    """The other headline figures of BASELINE.json, measured briefly (a few iterations each)."""
    import ctypes

    import torch

    import mms_answer_selection_b200 as mms
    from mms_answer_selection_b200 import _lib, layers, synth
    from mms_answer_selection_b200.blob import Blob
    out = {}
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    # ---- configs[3]: reranking, 1k queries x 1M candidates, K = 1024, candidates sharded over the ranks
    c4 = synth.CONFIGS["c4"]
    Nq, K = c4["Nq"], c4["K"]
    Nc_local = c4["Nc"] // world
    g = torch.Generator(device="cuda").manual_seed(synth.SEED + rank)
    Q = torch.randn((Nq, K), device="cuda", generator=g) / K ** 0.5
    C = torch.randn((Nc_local, K), device="cuda", generator=g) / K ** 0.5
    W = (torch.rand((K, K), device="cuda", generator=g) * 2 - 1) * (3.0 / K) ** 0.5
    QW = torch.empty((Nq, K), device="cuda")
    scores = torch.empty((Nq, Nc_local), device="cuda")
    h = _lib.Handle()
    h.set_stream(torch.cuda.current_stream().cuda_stream)

    def rerank():
        _lib.check(_lib.lib().mms_rerank_scores_f32(h.ptr, p(Q), p(C), p(W), p(QW), p(scores), Nq, Nc_local, K, K))
    ms = _time_ms(rerank, 3, flush, world)
    flops = 2.0 * Nq * K * K + 2.0 * Nq * Nc_local * K
    out["candidate_scoring"] = {
        "workload": "C4: %d queries x %d candidates (%d per GPU), K=%d, scores = (Q W) C^T, all scores written"
                    % (Nq, Nc_local * world, Nc_local, K),
        "candidate_scores_per_sec": Nq * Nc_local * world / (ms / 1e3), "ms": ms,
        "tflops_per_gpu": flops / (ms / 1e3) / 1e12}
    # the same against PREPARED candidates: a static candidate set is rounded to TF32 once (outside the timed call, as a
    # static index would be) and every query batch is scored against that copy -- reported beside, never instead of,
    # the figure above, which pays for the rounded copy inside every call
    Cr = torch.empty((Nc_local, (K + 3) // 4 * 4), device="cuda")
    _lib.check(_lib.lib().mms_rerank_prepare_f32(h.ptr, p(C), p(Cr), Nc_local, K))

    def rerank_prepared():
        _lib.check(_lib.lib().mms_rerank_scores_prepared_f32(h.ptr, p(Q), p(Cr), p(W), p(QW), p(scores), Nq, Nc_local, K, K))
    ms_p = _time_ms(rerank_prepared, 3, flush, world)
    out["candidate_scoring"]["prepared_candidates"] = {
        "note": "candidate set rounded to TF32 once before the timed calls (mms_rerank_prepare), scores identical",
        "candidate_scores_per_sec": Nq * Nc_local * world / (ms_p / 1e3), "ms": ms_p,
        "tflops_per_gpu": flops / (ms_p / 1e3) / 1e12}
    del C, Cr, scores, Q, QW, W
    torch.cuda.empty_cache()
    if world > 1:
        return out
    # ---- configs[2] on ONE GPU (the base of the strong-scaling series the N > 1 runs report)
    c3 = synth.CONFIGS["c3"]
    d = synth.make_qa_batch(N=c3["N"], L=c3["L"], D=c3["D"], mc=c3["mc"], V=c3["V"])
    net = mms.MMSNet(c3["N"], c3["L"], c3["D"], c3["mc"], c3["V"])
    net.set_params(d["W"], d["b"], d["M"], d["B"])
    net.set_inputs(d["idx_q"], d["idx_a"])
    net.set_upstream_gradient(d["dS"])
    net.capture(with_loss=True, clear_diffs=True)
    ms = _time_ms(lambda: net.replay(read_loss=False), 5, flush, 1)
    fl = c3["N"] * flops_per_pair(c3["L"], c3["D"], c3["mc"])
    out["c3_single_gpu"] = {"workload": "C3: 4096 QA pairs/step on one GPU, D=300, mc=4 (fwd+bwd, graph replay)",
                            "qa_pairs_per_sec": c3["N"] / (ms / 1e3), "ms_per_step": ms,
                            "algorithmic_tflops": fl / (ms / 1e3) / 1e12}
    # the same step as a whole solver iteration: + the fused AdaDelta update of all 18.4 M parameters (scale,
    # weight decay, update, Net::Update and the next iteration's ClearParamDiffs in one launch per blob)
    net.ClearParamDiffs()
    solver = mms.AdaDeltaSolver(net.params(), lr_mult=[1.0, 2.0, 1.0, 1.0][:len(net.params())],
                                decay_mult=[0.0, 0.0, 1.0, 1.0][:len(net.params())])
    net.capture_train_step(solver)
    ms_t = _time_ms(net.replay_train_step, 5, flush, 1)
    out["c3_train_step_adadelta"] = {
        "workload": "C3 step + AdaDelta update of every learnable blob (fused optimizer launch replaces ClearParamDiffs)",
        "qa_pairs_per_sec": c3["N"] / (ms_t / 1e3), "ms_per_step": ms_t}
    del net, d, solver
    torch.cuda.empty_cache()
    # ---- configs[4]: 4 modalities, SimMatrix 1024 x 1024, batch 16384, PairRankLoss on (s+, s-)
    c5 = synth.CONFIGS["c5"]
    N5, K1, K2, nm = c5["N"], c5["K1"], c5["K2"], c5["modalities"]
    mods = []
    for m in range(nm):
        qv, av, Wv = synth.make_sentence_vectors(N5, K1, K2, seed=synth.SEED + m)
        lay = layers.SimMatrixLayer(layers.LayerParameter("SimMatrix", sim_matrix_param=dict(
            weight_filler=dict(type="xavier"))))
        bq, ba, top = Blob((N5, K1)), Blob((N5, K2)), Blob(())
        bq.set_cpu_data(qv); ba.set_cpu_data(av)
        lay.SetUp([bq, ba], [top])
        lay.handle.set_option(_lib.MMS_OPT_REUSE_FORWARD, 1)      # Backward right after Forward on unchanged bottoms
        lay.blobs[0].set_cpu_data(Wv)
        top.diff.fill_(1.0 / N5)
        mods.append((lay, bq, ba, top))

    def c5_step():
        for lay, bq, ba, top in mods:
            lay.blobs[0].diff.zero_()
            lay.Forward([bq, ba], [top])
            lay.Backward([top], [True, True], [bq, ba])
    ms_eager = _time_ms(c5_step, 3, flush, 1)
    # the ~50 launches of the step issued from Python are host-bound; record them once and replay (as MMSNet.capture does)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        c5_step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    c5_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(c5_graph, stream=side):
        c5_step()
    ms = _time_ms(c5_graph.replay, 5, flush, 1)
    out["c5_multimodal"] = {"workload": "C5: %d modalities x SimMatrix %dx%d, batch %d, fwd+bwd (dW, dq, da), step replayed "
                                        "as one CUDA graph (eager launches from Python: %.3f ms)" % (nm, K1, K2, N5, ms_eager),
                            "qa_pairs_per_sec": N5 / (ms / 1e3), "ms_per_step": ms,
                            "algorithmic_tflops": nm * 6.0 * N5 * K1 * K2 / (ms / 1e3) / 1e12}
    del mods, c5_graph
    torch.cuda.empty_cache()
    # ---- ranking metrics on the device: MAP + MRR over a reranking score slab, grouped by query
    nq, nc = 1000, 32768
    n = nq * nc
    gm = torch.Generator(device="cuda").manual_seed(synth.SEED)
    prob = torch.rand((n, 2), device="cuda", generator=gm)
    label = (torch.rand((n,), device="cuda", generator=gm) < 0.01).float()
    group = torch.randint(0, nq, (n,), device="cuda", generator=gm).float()
    res = torch.empty((2,), device="cuda")

    def rank_metrics():
        _lib.check(_lib.lib().mms_rank_map_mrr_f32(h.ptr, p(prob), 2, 1, p(label), p(group), n, p(res),
                                                   ctypes.c_void_p(res.data_ptr() + 4)))
    ms = _time_ms(rank_metrics, 3, flush, 1)
    out["ranking_metrics"] = {"workload": "MAP + MRR of %d scores in %d query groups (segmented sort + scans on the device, "
                                          "one call)" % (n, nq),
                              "scores_per_sec": n / (ms / 1e3), "ms": ms, "map": float(res[0]), "mrr": float(res[1]),
                              "input_gbs": 16.0 * n / (ms / 1e3) / 1e9}
    del prob, label, group
    torch.cuda.empty_cache()
    # ---- sentence encoder (sentence-vector variant): Convolution(5 x D) -> BN -> MAX over time ->
    #      TanH, forward + backward, with per-kernel device times (tools/sentenc_bench.py)
    import tools.sentenc_bench as sentenc_bench
    r = sentenc_bench.run(N=8192, iters=3)
    out["sentence_encoder"] = {k: r[k] for k in ("workload", "ms_per_step", "sentences_per_sec", "conv_algorithmic_tflops",
                                                 "conv_gemm_only_tflops", "hbm_gbs")}
    out["sentence_encoder"]["kernels_ms_per_step"] = {k: v["ms_per_step"] for k, v in r["kernels"].items()}
    torch.cuda.empty_cache()
    # ---- the whole sentence-vector variant as one net (north_star's path end to end): Embed x2 -> sentence encoder x2
    #      (shared parameters) -> SimMatrix -> PairRankLoss, forward + backward into W, the filters, BN and the table
    Ns = c3["N"]
    ds_ = synth.make_qa_batch(N=Ns, L=c3["L"], D=c3["D"], mc=1, V=c3["V"])
    snet = mms.SentenceVectorNet(Ns, c3["L"], c3["D"], 100, 5, c3["V"])
    snet.branches[0]["embed"].blobs[0].set_cpu_data(ds_["W"])
    lab = (np.random.default_rng(synth.SEED).uniform(0, 1, Ns // 2) < 0.5).astype(np.float32)
    snet.set_inputs(ds_["idx_q"], ds_["idx_a"], lab)
    snet.capture(clear_diffs=True)
    ms = _time_ms(snet.replay, 5, flush, 1)
    out["sentence_variant_step"] = {
        "workload": "sentence-vector variant, %d QA pairs/step: Embed x2 -> Convolution(5 x %d, 100) -> BN -> MAX over time -> "
                    "TanH (shared parameters) -> SimMatrix -> PairRankLoss, fwd+bwd incl. ClearParamDiffs, one CUDA graph"
                    % (Ns, c3["D"]),
        "qa_pairs_per_sec": Ns / (ms / 1e3), "ms_per_step": ms, "loss": snet.loss_value()}
    return out


def _ncu_row(wl, name):
    """The row of kernel `name` in the committed `ncu --set full` summary of this workload
    (profiles/r01_ncu_full_<wl>_fused.json, written by tools/ncu_summary.py), or None."""
    alias = {"simcross2_bwd_fused_kernel<dQ>": "simcross2_bwd_fused_kernel<0, 0>",
             "simcross2_bwd_fused_kernel<dA>": "simcross2_bwd_fused_kernel<1, 0>"}
    try:
        rows = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_%s_fused.json" % wl)))
    except Exception:
        return None
    want = alias.get(name, name)
    for r in rows:
        if r.get("kernel") == want or r.get("kernel", "").startswith(want + "<"):
            return r
    return None


def ncu_traffic(wl, name):
    """DRAM bytes per launch of kernel `name` from the committed ncu capture of this workload, or None."""
    r = _ncu_row(wl, name)
    if not r:
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v, u = r[k].split()
        tot += float(v) * scale.get(u, 1.0)
    return tot


def ncu_tensor_pipe(wl, name):
    """sm__pipe_tensor_cycles_active (% of peak, ncu) of kernel `name` from the committed capture, or None."""
    r = _ncu_row(wl, name)
    try:
        return float(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"].split()[0])
    except Exception:
        return None


def hbm_kernel_table(prof, steps, cfg, peaks):
    """Achieved HBM GB/s of the gather / scatter / reduce kernels of the step against the measured copy bandwidth:
    algorithmic bytes per step / the kernel's summed device time per step."""
    N, L, D, mc, V = cfg["N"], cfg["L"], cfg["D"], cfg["mc"], cfg["V"]
    rows = N * 2 * L                                   # token rows of q and a
    Dp = (D + 31) // 32 * 32                           # rounded copies: rows padded to 128-byte lines
    alg = {
        "embed_forward_vec": rows * (4 + 8 * D),                       # id + table row in, row out
        "embed_backward_runs": rows * (4 + 4 * D) + rows * 8 * D,      # id + dtop in, <= one RMW of the dW row
        "tf32_round_kernel": rows * 4 * (D + Dp) + mc * D * 4 * (D + Dp),
        "sum_kernel": 2 * 4 * N * mc * L * L,                          # loss = dot(S, dS)
        "bias_grad_kernel": 4 * N * mc * L * L,                        # dB += sum_n dS[n]
    }
    peak = peaks.get("hbm_gbs", 6650.0)
    out = {}
    for k, b in alg.items():
        if k in prof and prof[k][1] > 0:
            gbs = b / (prof[k][1] / steps / 1e3) / 1e9
            out[k] = {"algorithmic_bytes_per_step": int(b), "achieved_gbs": round(gbs, 1), "frac_of_peak": round(gbs / peak, 3)}
    out["peak_gbs"] = peak
    return out


def roofline_for(name, rec, prof, steps, cfg, peaks, wl):
    """Roofline entry for the dominant kernel `name` (launch count, total ms over `steps`)."""
    n, ms = rec
    per_launch_s = ms / n / 1e3
    N, L, D, mc, V = cfg["N"], cfg["L"], cfg["D"], cfg["mc"], cfg["V"]
    launches_per_step = n / steps
    shares = kernel_flop_shares(L, D, mc)
    tensor_kernels = ("simcross", "tc_", "simt_gemm")
    if any(t in name for t in tensor_kernels):
        peak = peaks.get("bf16_tflops", 1590.0) / 2.0       # TF32 dense = half the bf16 rate
        flops_launch = N * shares.get(name, 0.0) / max(launches_per_step, 1.0)
        achieved = flops_launch / per_launch_s / 1e12
        # the whole step: all algorithmic FLOPs over the summed device time of every contraction kernel
        fam_ms = sum(m for k, (_, m) in prof.items() if any(t in k for t in tensor_kernels))
        step_achieved = N * flops_per_pair(L, D, mc) * steps / (fam_ms / 1e3) / 1e12
        return {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": ncu_traffic(wl, name),
                "tensor_pipe_active_pct_ncu": ncu_tensor_pipe(wl, name),
                "flops_per_launch": flops_launch, "ms_per_launch": per_launch_s * 1e3,
                "step_contractions": {"achieved": step_achieved, "frac": step_achieved / peak,
                                      "ms_per_step": fam_ms / steps},
                "note": "algorithmic FLOPs attributed to this kernel (DESIGN.md 3.1) per launch / its mean launch time "
                        "(CUDA events, one kernel at a time); step_contractions = all 6LD^2+8L^2D FLOPs / summed time of "
                        "the contraction kernels; peak = measured cuBLAS bf16 burst / 2 (TF32), %s; traffic = DRAM "
                        "bytes per launch from profiles/r01_ncu_full_%s_fused.json" % ("of measured" if peaks else "of fallback", wl)}
    rows = N * 2 * L
    bytes_launch = rows * (4 + 8 * D) / max(launches_per_step, 1) * 1.0
    achieved = bytes_launch / per_launch_s / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": None,
            "note": "algorithmic bytes per launch / mean launch time; %s" % ("of measured" if peaks else "of fallback")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["c1", "c2", "c3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result.  Libraries write there too (NCCL prints its version banner to
    # stdout when NCCL_DEBUG is set): route fd 1 to stderr for the run and keep the real stdout for the result line.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


_RESULT_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


if __name__ == "__main__":
    sys.exit(main())
