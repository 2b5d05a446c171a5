#!/usr/bin/env python
"""Benchmark of the MMS hot path on B200 (see the contract in DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c1]
                    [--exchange p2p|p2p-multicast|nccl]

A "step" is one forward+backward pass of the hot path (Embed x2 -> SimCross mode 2 -> loss plumbing -> SimCross backward
-> Embed scatter-add, and at N > 1 the exchange of the parameter gradients) over one batch of synthetic TREC-QA-shaped
QA pairs.  The workload is the SAME at every N -- BASELINE.json's configs[2] (C3): a GLOBAL batch of 4096 QA pairs,
q/a length 40, 300-d embeddings, mesure_count 4, V = 60002 -- on one GPU at N = 1 and sharded 4096/N pairs per rank at
N = 2, 4, 8 (strong scaling), so the N = 1 line is the base of the 1 -> 8 series.  --workload overrides (c2 = the
reference's own batch of 50 pairs, c1 = its 50-d configuration); both are also measured in `extra` at N = 1, with the
other configurations of BASELINE.json: candidate scoring (configs[3], candidates sharded over the ranks) and the
multi-modal net (configs[4]).

Prints ONE JSON line (rank 0).  `value` = QA pairs/s with inputs resident in HBM; `e2e` = the same through the public
API with pinned-host inputs, H2D of the step's token ids and D2H of the step's loss inside the timed region.  At N > 1
the line also carries `comm` (the exchange timed alone, beside a plain NCCL all-reduce of the same buffer) and `parity`
(the exchanged gradient against rank 0 recomputing the whole 4096-pair batch alone).
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
NVLINK5_GBS_PER_DIRECTION = 900.0


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
    """Per-rank configuration: the GLOBAL batch is fixed, a rank gets N / world pairs."""
    from mms_answer_selection_b200 import synth
    c = dict(synth.CONFIGS[name])
    c["global_N"] = c["N"]
    c["N"] = max(1, c["N"] // world)
    return c


def pick_workload(args, world):
    return args.workload or "c3"


def bench_config(name, cfg, world):
    """The `config` object of the JSON line -- built by BOTH arms from the same arguments, so the driver sees identical
    dictionaries; anything arm-specific goes under `run`."""
    return {"workload": workload_name(name, cfg), "name": name, "global_batch": cfg["global_N"],
            "pairs_per_gpu": cfg["N"], "q_a_len": cfg["L"], "embedding_dim": cfg["D"], "mesure_count": cfg["mc"],
            "vocab_rows": cfg["V"], "parallelism": "dp%d" % world,
            "l2": "256 MiB L2 flush between timed steps (inputs + table < the 126 MB L2)"}


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
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(wl, cfg, args.gpus),
        "run": {"pairs_per_step": r["pairs"], "note": "rank 0's share of the global batch on the host cores; each "
                "step is a bounded sample of it"},
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
def build_net(cfg, world, rank, keep_embed_tops=False, grouped_scatter=True):
    """MMSNet over this rank's slice of the GLOBAL synthetic batch (same seed on every rank: the global batch is
    identical everywhere, rank r takes pairs [r N/n, (r+1) N/n)).  Each worker normalises its loss by ITS pair count, as
    every Caffe worker does: dS of the global batch times `world`; the exchange's 1/n (parallel.cpp:377) turns the
    sum of the per-rank means into the global-batch mean."""
    import mms_answer_selection_b200 as mms
    from mms_answer_selection_b200 import synth
    N, L, D, mc, V = cfg["N"], cfg["L"], cfg["D"], cfg["mc"], cfg["V"]
    full = synth.make_qa_batch(N=cfg["global_N"], L=L, D=D, mc=mc, V=V, seed=synth.SEED)
    sl = slice(rank * N, (rank + 1) * N)
    # keep_embed_tops=False: q / a leave the gather as the TF32 operand copy the contractions read and nothing else
    # (MMS_OPT_STAGE_ONLY); the fp32 tops a Caffe net would expose are not written -- nothing on this path reads them
    net = mms.MMSNet(N, L, D, mc, V, keep_embed_tops=keep_embed_tops, grouped_scatter=grouped_scatter)
    net.set_params(full["W"], full["b"], full["M"], full["B"])
    net.set_inputs(full["idx_q"][sl], full["idx_a"][sl])
    net.set_upstream_gradient(full["dS"][sl] * world)
    return net, full, sl


def elementwise_error_stats(got, ref, tol=1e-3, floor_frac=0.01):
    """Normwise AND element-wise error of `got` against `ref` (both numpy): max |got-ref| / max|ref| (the
    GradientChecker-style figure the tests bound), and among the elements with |ref| >= floor_frac * max|ref| the
    largest relative error and the fraction above `tol`."""
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    m = float(np.abs(ref).max()) if ref.size else 0.0
    if m == 0.0:
        return {"max_scaled_err": float(np.abs(got).max()) if got.size else 0.0, "elements_checked": 0}
    big = np.abs(ref) >= floor_frac * m
    rel = np.abs(got[big] - ref[big]) / np.abs(ref[big])
    return {"max_scaled_err": float(np.abs(got - ref).max() / m),
            "max_rel_err_above_%g_of_max" % floor_frac: float(rel.max()) if rel.size else 0.0,
            "frac_rel_err_gt_%g" % tol: float((rel > tol).mean()) if rel.size else 0.0,
            "elements_checked": int(big.sum())}


def measure_tf32_peak(seconds=1.5):
    """cuBLAS TF32 8192^3 through torch.matmul (allow_tf32), measured the way MEASURED_PEAKS.json measures bf16: best
    of 10 single launches (burst) and back-to-back launches for `seconds` (sustained)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn((n, n), device="cuda"); b = torch.randn((n, n), device="cuda"); c = torch.empty((n, n), device="cuda")
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record(); torch.cuda.synchronize()
        fl = 2.0 * n ** 3
        return {"tf32_tflops_burst": fl / (best / 1e3) / 1e12, "tf32_tflops_sustained": fl * reps / (e0.elapsed_time(e1) / 1e3) / 1e12,
                "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS TF32): best of 10 (burst), back to back for "
                       "%.1f s (sustained), CUDA events" % seconds}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def main_ours(args):
    import torch
    import torch.distributed as dist

    import mms_answer_selection_b200 as mms
    from mms_answer_selection_b200 import _lib as _mmslib
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

    net, full, sl = build_net(cfg, world, rank, keep_embed_tops=args.embed_tops == "fp32",
                              grouped_scatter=args.scatter == "grouped")
    exch, exch_note = None, "none"
    if world > 1:
        exch, exch_note = make_exchange(net.params(), args.exchange)
        exch.broadcast_params(0)
        exch.check()
    host_q = torch.from_numpy(np.ascontiguousarray(full["idx_q"][sl])).pin_memory()
    host_a = torch.from_numpy(np.ascontiguousarray(full["idx_a"][sl])).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2
    handles = [net.embed_q.handle, net.embed_a.handle, net.sim.handle]
    count_launches = lambda: sum(h.launch_count() for h in handles) + (exch.launch_count() if exch else 0)

    # The whole step (ClearParamDiffs + Forward + loss + Backward [+ exchange]) is recorded once as a CUDA graph and
    # replayed: issued launch by launch from the host the ~15 short kernels are launch-bound.
    in_graph = exch is not None and exch.backend == "p2p"
    l0 = count_launches()
    if in_graph:
        graph = net.capture_exchange_step(exch, with_loss=True, clear_diffs=True)
        launches_per_step = (count_launches() - l0) // 3                      # 2 warm-up passes + the capture
        graph_host = net.capture_exchange_step(exch, with_loss=True, clear_diffs=True, host_inputs=(host_q, host_a),
                                               prefetch_inputs=not args.no_input_prefetch)
    else:
        net.capture(with_loss=True, clear_diffs=True)
        launches_per_step = (count_launches() - l0) // 3 + (1 if exch else 0)
        net.capture(with_loss=True, clear_diffs=True, host_inputs=(host_q, host_a),      # + H2D / D2H nodes for the e2e step
                    prefetch_inputs=not args.no_input_prefetch)

    def device_step():
        if in_graph:
            graph.replay()
            return
        net.replay(read_loss=False)
        if exch:
            exch.allreduce()

    def e2e_step():
        if in_graph:
            # one graph launch: H2D of this step's inputs (pinned), the step incl. the exchange, D2H of the loss, a sync
            graph_host.replay()
            torch.cuda.current_stream().synchronize()
            return float(net._host_loss[0])
        if not exch:
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
        return _max_ms(ms, world)

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
    if exch:
        exch.check()

    # ---- parity of the exchanged gradient (N > 1): one more step, then rank 0 recomputes the GLOBAL batch alone
    parity = None
    if exch:
        device_step()
        torch.cuda.synchronize()
        parity = exchange_parity(net, exch, cfg, full, world, rank)

    # ---- the exchange alone (N > 1)
    comm = comm_benchmark(exch, world, rank, flush, args) if exch else None

    # roofline of the dominant kernel, measured live with CUDA events around each launch of an
    # eagerly issued step.  The GPU is first given ~0.5 ms of other work (L2 flushes) so that every
    # launch of the step is already queued when it runs: the events then bracket device time, not
    # host launch latency.
    sim_state = net.sim.defer_loss_
    net.sim.defer_loss_ = True
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
    tf32 = measure_tf32_peak() if (rank == 0 and world == 1 and not args.no_extra) else None
    roof = roofline_for(dom[0], dom[1], prof, prof_steps, cfg, peaks, wl, tf32)

    value = N * world * args.steps / (total_ms / 1e3)
    e2e_value = N * world * args.steps / (e2e_ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (TF32 tensor-core contractions, fp32 accumulate)",
        "data": "synthetic",
        "config": bench_config(wl, cfg, world),
        "run": {"launch": "step recorded as one CUDA graph (%d kernels of libmms_b200.so per step)" % launches_per_step,
                "grad_exchange": exch_note,
                "embed_tops": ("fp32 tops + TF32 operand copy" if net.keep_embed_tops else
                               "TF32 operand copy only (MMS_OPT_STAGE_ONLY; --embed-tops fp32 also writes the fp32 tops)"),
                "algorithmic_tflops_per_gpu": N * flops_per_pair(L, D, mc) / (total_ms / args.steps / 1e3) / 1e12},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(host_q.numel() * 4 + host_a.numel() * 4), "d2h_bytes_per_step": 4,
                "note": ("one graph launch + one stream synchronisation per step; the pinned ids are copied H2D every step "
                         + ("in front of the step's kernels" if args.no_input_prefetch else
                            "beside the step's kernels into staging buffers and handed to the id blobs at the end of the step "
                            "(input double buffering: step k computes on the ids step k-1 fetched; --no-input-prefetch puts "
                            "the copy in front of the kernels)")
                         + "; the loss is read back D2H every step")},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernels_ms_per_step": {k: round(ms / prof_steps, 5) for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])},
        "kernel_step_ms_sum": step_ms_prof,
        "hbm_kernels": hbm_kernel_table(prof, prof_steps, cfg, peaks, net.keep_embed_tops),
    }
    if tf32:
        line["peaks"] = dict(tf32, bf16_tflops_burst_driver=peaks.get("bf16_tflops"), hbm_gbs_driver=peaks.get("hbm_gbs"))
    if comm is not None:
        line["comm"] = comm
    if parity is not None:
        line["parity"] = parity
    if not args.no_extra:
        line["extra"] = extras(world, rank, flush, args)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = time_reference(cfg, steps=3, warmup=1, budget_s=20.0)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                "sample": r["sample"]}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def make_exchange(params, mode):
    """GradientExchange over the net's learnable blobs.  "auto" (default): csrc/exchange.cu over torch symmetric memory
    with the NVSwitch multicast mapping (multimem.ld_reduce / multimem.st) when a small probe allocation shows that
    every rank can set it up, else the same kernel over cudaIpc-mapped peer memory (plain peer loads / stores);
    "p2p" / "p2p-multicast" force one of the two; "nccl": the library all-reduce (baseline).  A peer-memory set-up that
    fails on this box falls back to NCCL -- loudly, in the line's `run.grad_exchange` -- so that a scaling series is
    never lost to a mapping problem."""
    import torch
    import torch.distributed as dist
    from mms_answer_selection_b200.blob import Blob
    from mms_answer_selection_b200.parallel import GradientExchange

    def all_ok(flag):
        t = torch.tensor([1 if flag else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item()) == 1

    if mode == "nccl":
        return GradientExchange(params, backend="nccl"), "one NCCL all-reduce (ncclAvg) of the flat gradient buffer after the graph"
    symmetric = mode == "p2p-multicast"
    if mode == "auto":
        ok = True
        try:
            probe = GradientExchange([Blob((4096,))], backend="p2p", symmetric=True)
            ok = probe.multicast
            probe.allreduce(); probe.check(); probe.close()
        except Exception as e:          # noqa: BLE001
            ok = False
            sys.stderr.write("symmetric-memory / multicast probe failed on this rank: %r\n" % (e,))
        symmetric = all_ok(ok)
    ok, ex = True, None
    try:
        ex = GradientExchange(params, backend="p2p", symmetric=symmetric)
    except Exception as e:          # noqa: BLE001
        ok = False
        sys.stderr.write("peer-memory exchange unavailable on this rank: %r\n" % (e,))
    if all_ok(ok):
        if ex.multicast:
            ex.set_option(1, 32)        # the NVSwitch does the adding: 32 CTAs saturate it and leave the SMs to dM / dB
        how = "multimem.ld_reduce/st over the NVSwitch multicast mapping" if ex.multicast else "peer loads/stores"
        return ex, ("csrc/exchange.cu (%s): table bucket on a private stream overlapping dM/dB, SimCross bucket after; "
                    "both inside the step's CUDA graph" % how)
    if ex is not None:
        raise RuntimeError("peer-memory exchange came up on some ranks only; blobs are already re-bound -- aborting")
    return GradientExchange(params, backend="nccl"), "FALLBACK: peer-memory mapping failed here; one NCCL all-reduce (ncclAvg) after the graph"


def exchange_parity(net, exch, cfg, full, world, rank):
    """The exchanged (averaged) flat gradient of the sharded step against ONE GPU computing the global batch (rank 0
    recomputes it alone), plus: are the replicas bit-identical?"""
    import torch
    import torch.distributed as dist

    import mms_answer_selection_b200 as mms
    flat = exch.flat_diff
    h = flat.view(torch.int32).to(torch.int64).sum().reshape(1)
    lo, hi = h.clone(), h.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out = {"replicas_bit_identical": bool(lo.item() == hi.item())}
    if rank == 0:
        L, D, mc, V = cfg["L"], cfg["D"], cfg["mc"], cfg["V"]
        ref = mms.MMSNet(cfg["global_N"], L, D, mc, V)
        ref.set_params(full["W"], full["b"], full["M"], full["B"])
        ref.set_inputs(full["idx_q"], full["idx_a"])
        ref.set_upstream_gradient(full["dS"])
        ref.ClearParamDiffs(); ref.ForwardBackward()
        torch.cuda.synchronize()
        names = ["dW_embed", "db_embed", "dM", "dB"]
        worst = 0.0
        for name, p, q in zip(names, net.params(), ref.params()):
            st = elementwise_error_stats(p.cpu_diff(), q.cpu_diff())
            out[name] = st
            worst = max(worst, st["max_scaled_err"])
        out["max_scaled_err"] = worst
        out["tolerance"] = 1e-3
        out["ok"] = bool(worst <= 1e-3 and out["replicas_bit_identical"])
        out["what"] = ("flat gradient after the exchange (mean over %d ranks of per-rank means) vs one GPU on the whole "
                       "%d-pair batch, TF32 contractions on both sides" % (world, cfg["global_N"]))
        del ref
        torch.cuda.empty_cache()
    dist.barrier()
    return out


def comm_benchmark(exch, world, rank, flush, args):
    """The gradient exchange timed ALONE on a scratch buffer of the step's size: our kernel (whole buffer in one launch;
    fused with the AdaDelta step), and a plain NCCL all-reduce of the same bytes for comparison.  busbw is NCCL's
    convention 2 (n-1)/n x bytes / time, against NVLink 5's 900 GB/s per direction."""
    import torch
    import torch.distributed as dist

    from mms_answer_selection_b200.blob import Blob
    from mms_answer_selection_b200.parallel import GradientExchange
    count = exch.count
    nbytes = count * exch.elem
    out = {"bytes": int(nbytes), "nvlink5_gbs_per_direction": NVLINK5_GBS_PER_DIRECTION}
    iters = 20

    def busbw(ms):
        return 2.0 * (world - 1) / world * nbytes / (ms / 1e3) / 1e9

    def entry(ms):
        return {"ms": ms, "algbw_gbs": nbytes / (ms / 1e3) / 1e9, "busbw_gbs": busbw(ms),
                "busbw_frac_of_nvlink5": busbw(ms) / NVLINK5_GBS_PER_DIRECTION}

    buf = torch.ones(count, dtype=exch.flat_diff.dtype, device="cuda")
    out["nccl_allreduce_avg"] = entry(_time_ms(lambda: dist.all_reduce(buf, op=dist.ReduceOp.AVG), iters, flush, world))
    del buf
    if exch.backend != "p2p":
        return out
    variants = [("p2p", False)]
    if exch.multicast or args.try_multicast:     # the symmetric-memory rendezvous is known to work here when the step uses it
        variants.append(("p2p_multicast", True))
    for name, symmetric in variants:
        try:
            b = Blob((count,))
            ok = 1
            try:
                x2 = GradientExchange([b], backend="p2p", symmetric=symmetric)
            except Exception as e:      # noqa: BLE001
                ok = 0
                err = repr(e)[:300]
            t = torch.tensor([ok], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0:
                out[name] = {"unavailable": err if not ok else "failed on another rank"}
                continue
            if symmetric and not x2.multicast:
                out[name] = {"unavailable": "no multicast support reported for this allocation"}
                continue
            x2.flat_diff.fill_(1.0)
            key = "allreduce" + ("_multimem" if x2.multicast else "")
            r = {key: entry(_time_ms(lambda: x2.allreduce(), iters, flush, world))}
            x2.check()
            for ctas in ((16, 32) if x2.multicast else (74, 296)):
                x2.set_option(1, ctas)
                r["%s_%dctas" % (key, ctas)] = entry(_time_ms(lambda: x2.allreduce(), iters, flush, world))
            x2.set_option(1, 0)
            x2.adadelta_step()                           # allocates the history outside the timed loop
            x2.check()
            r["fused_adadelta_step"] = entry(_time_ms(lambda: x2.adadelta_step(), iters, flush, world))
            x2.check()
            out[name] = r
            x2.close()
        except Exception as e:          # noqa: BLE001
            out[name] = {"error": repr(e)[:300]}
    return out


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


def small_batch_line(name, flush, steps=30):
    """One of the reference's own small configurations (C1: 50 pairs, D = 50; C2: 50 pairs, D = 300) on one GPU:
    device-resident and end-to-end step times, graph replay, L2 flushed between steps."""
    import torch
    cfg = workload_config(name, 1)
    net, full, sl = build_net(cfg, 1, 0)
    host_q = torch.from_numpy(np.ascontiguousarray(full["idx_q"])).pin_memory()
    host_a = torch.from_numpy(np.ascontiguousarray(full["idx_a"])).pin_memory()
    handles = [net.embed_q.handle, net.embed_a.handle, net.sim.handle]
    l0 = sum(h.launch_count() for h in handles)
    net.capture(with_loss=True, clear_diffs=True)
    per_step = (sum(h.launch_count() for h in handles) - l0) // 3
    net.capture(with_loss=True, clear_diffs=True, host_inputs=(host_q, host_a))
    ms = _time_ms(lambda: net.replay(read_loss=False), steps, flush, 1)
    ms_e2e = _time_ms(net.replay_from_host, steps, flush, 1)
    N, L, D, mc = cfg["N"], cfg["L"], cfg["D"], cfg["mc"]
    return {"workload": workload_name(name, cfg), "steps": steps, "qa_pairs_per_sec": N / (ms / 1e3), "ms_per_step": ms,
            "e2e_qa_pairs_per_sec": N / (ms_e2e / 1e3), "e2e_ms_per_step": ms_e2e, "kernels_per_step": per_step,
            "algorithmic_tflops": N * flops_per_pair(L, D, mc) / (ms / 1e3) / 1e12}


def multimodal_line(world, rank, flush, exchange_mode):
    """configs[4] (C5): 4 modalities x SimMatrix 1024 x 1024 -> Concat -> FM -> Slice -> PairRankLoss, batch 16384 (global;
    16384 / N per rank), forward + backward into every W_m, the FM bias and the bottoms, ClearParamDiffs and -- at
    N > 1 -- the exchange of the 16.8 MB of parameter gradients, recorded as one CUDA graph."""
    import torch

    from mms_answer_selection_b200 import synth
    from mms_answer_selection_b200.multimodal import MultiModalNet
    c5 = synth.CONFIGS["c5"]
    Ng, K1, K2, nm = c5["N"], c5["K1"], c5["K2"], c5["modalities"]
    n = Ng // world
    net = MultiModalNet(n, K1, K2, nm)
    g = torch.Generator(device="cuda").manual_seed(synth.SEED + rank)
    for m in range(nm):
        # q^, a^ ~ tanh(N(0,1)), W ~ xavier; generated on the device (16384 x 1024 x 8 tensors)
        net.q[m].data.copy_(torch.tanh(torch.randn((n, K1), device="cuda", generator=g)))
        net.a[m].data.copy_(torch.tanh(torch.randn((n, K2), device="cuda", generator=g)))
        gw = torch.Generator(device="cuda").manual_seed(synth.SEED + 100 + m)          # same weights on every rank
        net.sim[m].blobs[0].data.copy_((torch.rand((K1, K2), device="cuda", generator=gw) * 2 - 1) * (3.0 / K1) ** 0.5)
    net.label.data.copy_((torch.rand((n // 2, 1), device="cuda", generator=g) < 0.5).float())
    exch, note = None, "none"
    if world > 1:
        exch, note = make_exchange(net.params(), exchange_mode)
        if exch.backend == "p2p":
            note = "csrc/exchange.cu (%s): one launch for the 16.8 MB of W_m / bias gradients after Backward, inside the graph" % (
                "multimem.ld_reduce/st over the NVSwitch multicast mapping" if exch.multicast else "peer loads/stores")
    net.capture(exch)
    ms = _time_ms(net.replay, 10, flush, world)
    if exch:
        exch.check()
    out = {"workload": "C5: %d modalities x SimMatrix %dx%d -> Concat -> FM -> Slice -> PairRankLoss(margin 1), global batch "
                       "%d (%d per GPU), fwd+bwd (dW_m, db, dq, da) + ClearParamDiffs%s, one CUDA graph"
                       % (nm, K1, K2, Ng, n, " + gradient exchange" if exch else ""),
           "qa_pairs_per_sec": Ng / (ms / 1e3), "ms_per_step": ms, "loss": net.loss_value(), "grad_exchange": note,
           "algorithmic_tflops_per_gpu": nm * 6.0 * n * K1 * K2 / (ms / 1e3) / 1e12}
    return out


def embed_backward_modes(flush):
    """One Embed backward of the C3 step (163 840 token rows x 300, V = 60002, centre-padded ids) in both modes."""
    import ctypes

    import torch

    from mms_answer_selection_b200 import _lib, synth
    c3 = synth.CONFIGS["c3"]
    M, D, V = c3["N"] * c3["L"], c3["D"], c3["V"]
    rng = np.random.default_rng(synth.SEED)
    idx = torch.from_numpy(synth.make_indices(rng, c3["N"], c3["L"], V, 5, 40).reshape(-1)).cuda()
    g = torch.randn((M, D), device="cuda") * 1e-6
    dW = torch.zeros((V, D), device="cuda"); db = torch.zeros(D, device="cuda")
    h = _lib.Handle()
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    out = {"workload": "Embed backward, %d token rows x %d floats into a %d-row table (one of the two calls of a C3 step)" % (M, D, V)}
    for name, det in (("atomic", 0), ("deterministic", 1)):
        h.set_option(_lib.MMS_OPT_EMBED_DETERMINISTIC, det)
        fn = lambda: _lib.check(_lib.lib().mms_embed_backward_f32(h.ptr, p(idx), p(g), p(dW), p(db), M, D, V))
        ms = _time_ms(fn, 10, flush, 1)
        out[name] = {"ms": ms, "algorithmic_gbs": (M * (4 + 4 * D) + M * 8 * D) / (ms / 1e3) / 1e9}
    # both layers of the step (q: sentences of 3-20 tokens, a: 5-40) as one scatter-add grouped by id, plan included
    h.set_option(_lib.MMS_OPT_EMBED_DETERMINISTIC, 0)
    idx_q = torch.from_numpy(synth.make_indices(rng, c3["N"], c3["L"], V, 3, 20).reshape(-1)).cuda()
    gq = torch.randn((M, D), device="cuda") * 1e-6
    L_ = _lib.lib()
    pair = lambda: _lib.check(L_.mms_embed_backward_pair_f32(h.ptr, p(idx_q), p(gq), M, p(idx), p(g), M, p(dW), p(db), D, V))
    plan = lambda: _lib.check(L_.mms_embed_plan_pair_f32(h.ptr, p(idx_q), M, p(idx), M, V))
    two = lambda: (_lib.check(L_.mms_embed_backward_f32(h.ptr, p(idx_q), p(gq), p(dW), p(db), M, D, V)),
                   _lib.check(L_.mms_embed_backward_f32(h.ptr, p(idx), p(g), p(dW), p(db), M, D, V)))
    ms_pair, ms_plan, ms_two = _time_ms(pair, 10, flush, 1), _time_ms(plan, 10, flush, 1), _time_ms(two, 10, flush, 1)
    out["both_layers"] = {"rows": 2 * M, "per_layer_atomic_ms": ms_two, "grouped_by_id_ms": ms_pair,
                          "of_which_plan_ms": ms_plan,
                          "note": "grouped_by_id_ms includes the id grouping; MMSNet runs that part on a side stream beside the forward"}
    return out


def extras(world, rank, flush, args):
    """The other headline figures of BASELINE.json, measured briefly (a few iterations each)."""
    import ctypes

    import torch

    import mms_answer_selection_b200 as mms
    from mms_answer_selection_b200 import _lib, synth
    out = {}
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    # ---- configs[3]: reranking, 1k queries x 1M candidates, candidates sharded over the ranks; K = 1024 (C5's M) and K = 300
    c4 = synth.CONFIGS["c4"]
    h = _lib.Handle()
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    for K, key in ((c4["K"], "candidate_scoring"), (300, "candidate_scoring_k300")):
        Nq = c4["Nq"]
        Nc_local = c4["Nc"] // world
        g = torch.Generator(device="cuda").manual_seed(synth.SEED + rank)
        Q = torch.randn((Nq, K), device="cuda", generator=g) / K ** 0.5
        C = torch.randn((Nc_local, K), device="cuda", generator=g) / K ** 0.5
        W = (torch.rand((K, K), device="cuda", generator=g) * 2 - 1) * (3.0 / K) ** 0.5
        QW = torch.empty((Nq, K), device="cuda")
        scores = torch.empty((Nq, Nc_local), device="cuda")

        def rerank():
            _lib.check(_lib.lib().mms_rerank_scores_f32(h.ptr, p(Q), p(C), p(W), p(QW), p(scores), Nq, Nc_local, K, K))
        ms = _time_ms(rerank, 3, flush, world)
        flops = 2.0 * Nq * K * K + 2.0 * Nq * Nc_local * K
        out[key] = {
            "workload": "C4: %d queries x %d candidates (%d per GPU), K=%d, scores = (Q W) C^T, all scores written"
                        % (Nq, Nc_local * world, Nc_local, K),
            "candidate_scores_per_sec": Nq * Nc_local * world / (ms / 1e3), "ms": ms,
            "tflops_per_gpu": flops / (ms / 1e3) / 1e12}
        # the same against PREPARED candidates: a static candidate set is rounded to TF32 once (outside the timed call, as
        # a static index would be) and every query batch is scored against that copy -- reported beside, never instead
        # of, the figure above, which pays for the rounded copy inside every call
        Cr = torch.empty((Nc_local, (K + 3) // 4 * 4), device="cuda")
        _lib.check(_lib.lib().mms_rerank_prepare_f32(h.ptr, p(C), p(Cr), Nc_local, K))

        def rerank_prepared():
            _lib.check(_lib.lib().mms_rerank_scores_prepared_f32(h.ptr, p(Q), p(Cr), p(W), p(QW), p(scores), Nq, Nc_local, K, K))
        ms_p = _time_ms(rerank_prepared, 3, flush, world)
        out[key]["prepared_candidates"] = {
            "note": "candidate set rounded to TF32 once before the timed calls (mms_rerank_prepare), scores identical",
            "candidate_scores_per_sec": Nq * Nc_local * world / (ms_p / 1e3), "ms": ms_p,
            "tflops_per_gpu": flops / (ms_p / 1e3) / 1e12}
        # per-query top-k instead of the full matrix: score slabs folded into the lists while in L2, then
        # the lists of all ranks all-gathered and merged (ties by candidate index): the Nq x Nc scores are never written
        from mms_answer_selection_b200.rerank import Reranker
        rr = Reranker(W, k=100)
        rr._prepared = (Cr, Nc_local, K)
        ms_k = _time_ms(lambda: rr.topk(Q, C), 3, flush, world)
        ms_kp = _time_ms(lambda: rr.topk(Q, None), 3, flush, world)
        out[key]["top100_per_query"] = {
            "note": "local top-100 per query fused slab by slab + all-gather + merge; global lists on every rank",
            "candidate_scores_per_sec": Nq * Nc_local * world / (ms_k / 1e3), "ms": ms_k,
            "prepared_candidates_ms": ms_kp,
            "prepared_candidate_scores_per_sec": Nq * Nc_local * world / (ms_kp / 1e3)}
        del C, Cr, scores, Q, QW, W, rr
        torch.cuda.empty_cache()
    # ---- configs[4]: the multi-modal net, at every N
    out["c5_multimodal"] = multimodal_line(world, rank, flush, args.exchange)
    torch.cuda.empty_cache()
    if world > 1:
        return out
    # ---- the reference's own small configurations on one GPU (C1: D = 50; C2: D = 300; 50 pairs per step)
    out["c1_reference_config"] = small_batch_line("c1", flush)
    out["c2_reference_batch"] = small_batch_line("c2", flush)
    torch.cuda.empty_cache()
    # ---- the C3 step as a whole solver iteration: + the fused AdaDelta update of all 18.4 M parameters (scale, weight
    #      decay, update, Net::Update and the next iteration's ClearParamDiffs in one launch per blob)
    c3 = synth.CONFIGS["c3"]
    cfg3 = workload_config("c3", 1)
    net, full, sl = build_net(cfg3, 1, 0)
    net.ClearParamDiffs()
    solver = mms.AdaDeltaSolver(net.params(), lr_mult=[1.0, 2.0, 1.0, 1.0][:len(net.params())],
                                decay_mult=[0.0, 0.0, 1.0, 1.0][:len(net.params())])
    net.capture_train_step(solver)
    ms_t = _time_ms(net.replay_train_step, 5, flush, 1)
    out["c3_train_step_adadelta"] = {
        "workload": "C3 step + AdaDelta update of every learnable blob (fused optimizer launch replaces ClearParamDiffs)",
        "qa_pairs_per_sec": c3["N"] / (ms_t / 1e3), "ms_per_step": ms_t}
    del net, full, solver
    torch.cuda.empty_cache()
    # ---- Embed backward: run-merged float atomics (default) beside the order-independent path (MMS_OPT_EMBED_DETERMINISTIC)
    out["embed_backward_modes"] = embed_backward_modes(flush)
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
    # ---- the net the reference actually trains, up to Flatten: Embed x2 -> SimCross -> Dropout ->
    #      (Conv5x5 + BN -> AvePool -> TanH) x 2, forward + backward (tools/simcnn_bench.py)
    import tools.simcnn_bench as simcnn_bench
    r = simcnn_bench.run(N=c3["N"], iters=3)
    top = dict(list(r["kernels_ms_per_step"].items())[:12])
    out["sim_cnn_step"] = {"workload": r["workload"], "ms_per_step": r["ms_per_step"], "qa_pairs_per_sec": r["qa_pairs_per_sec"],
                           "kernel_ms_sum": r["kernel_ms_sum"], "conv_algorithmic_gflop_per_step": r["conv_algorithmic_gflop_per_step"],
                           "largest_kernels_ms_per_step": top}
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
    rows = None
    for rnd in ("r02", "r01"):          # the newest committed capture of this workload
        try:
            rows = json.load(open(os.path.join(ROOT, "profiles", "%s_ncu_full_%s_fused.json" % (rnd, wl))))
            break
        except Exception:
            continue
    if rows is None:
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


def hbm_kernel_table(prof, steps, cfg, peaks, keep_embed_tops=True):
    """Achieved HBM GB/s of the gather / scatter / reduce kernels of the step against the measured copy bandwidth:
    algorithmic bytes per step / the kernel's summed device time per step."""
    N, L, D, mc, V = cfg["N"], cfg["L"], cfg["D"], cfg["mc"], cfg["V"]
    rows = N * 2 * L                                   # token rows of q and a
    Dp = (D + 31) // 32 * 32                           # rounded copies: rows padded to 128-byte lines
    alg = {
        # id + table row in, row out, and (MMS_OPT_STAGE_TF32, what MMSNet runs) the rounded operand copy out as well
        # (keep_embed_tops=False: the fp32 row is not written)
        "embed_forward_vec": rows * (4 + (8 if keep_embed_tops else 4) * D + 4 * Dp),
        "embed_backward_runs": rows * (4 + 4 * D) + rows * 8 * D,      # id + dtop in, <= one RMW of the dW row
        # grouped form: the gradient rows once; the touched table rows (<= one per token, not counted) on top
        "embed_backward_short_runs+embed_backward_long_chunks": rows * (4 + 4 * D),
        "tf32_round_kernel": mc * D * 4 * (D + Dp),                    # only M: q and a arrive staged by the gather
        "sum_kernel": 2 * 4 * N * mc * L * L,                          # loss = dot(S, dS)
        "bias_grad_kernel": 4 * N * mc * L * L,                        # dB += sum_n dS[n]
    }
    peak = peaks.get("hbm_gbs", 6650.0)
    out = {}
    for k, b in alg.items():
        parts = k.split("+")
        if all(p_ in prof and prof[p_][1] > 0 for p_ in parts):
            ms = sum(prof[p_][1] for p_ in parts)
            gbs = b / (ms / steps / 1e3) / 1e9
            out[k] = {"algorithmic_bytes_per_step": int(b), "achieved_gbs": round(gbs, 1), "frac_of_peak": round(gbs / peak, 3)}
    out["peak_gbs"] = peak
    return out


def roofline_for(name, rec, prof, steps, cfg, peaks, wl, tf32=None):
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
        extra_peak = {}
        if tf32:
            extra_peak = {"peak_cublas_tf32_burst": tf32["tf32_tflops_burst"],
                          "peak_cublas_tf32_sustained": tf32["tf32_tflops_sustained"],
                          "frac_of_cublas_tf32_burst": achieved / tf32["tf32_tflops_burst"]}
        return {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": ncu_traffic(wl, name), **extra_peak,
                "tensor_pipe_active_pct_ncu": ncu_tensor_pipe(wl, name),
                "flops_per_launch": flops_launch, "ms_per_launch": per_launch_s * 1e3,
                "step_contractions": {"achieved": step_achieved, "frac": step_achieved / peak,
                                      "ms_per_step": fam_ms / steps},
                "note": "algorithmic FLOPs attributed to this kernel (DESIGN.md 3.1) per launch / its mean launch time "
                        "(CUDA events, one kernel at a time); step_contractions = all 6LD^2+8L^2D FLOPs / summed time of "
                        "the contraction kernels; peak = driver-measured cuBLAS bf16 burst / 2 (TF32 runs at half the bf16 "
                        "rate), %s -- the cuBLAS TF32 8192^3 figure measured in this run is given beside it; traffic = DRAM "
                        "bytes per launch from the committed ncu capture profiles/r0x_ncu_full_%s_fused.json"
                        % ("of measured" if peaks else "of fallback", wl)}
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
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "p2p-multicast", "nccl"])
    ap.add_argument("--try-multicast", action="store_true", help="comm: also time the symmetric-memory / multimem variant")
    ap.add_argument("--embed-tops", default="staged", choices=["staged", "fp32"],
                    help="staged: the gather writes only the TF32 operand copy SimCross reads; fp32: also the fp32 tops")
    ap.add_argument("--scatter", default="grouped", choices=["grouped", "per-layer"],
                    help="grouped: both Embed backwards as one scatter-add with rows grouped by id; per-layer: two atomic kernels")
    ap.add_argument("--no-input-prefetch", action="store_true",
                    help="e2e: copy the step's ids H2D in front of its kernels instead of beside the previous step's")
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
