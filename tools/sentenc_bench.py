#!/usr/bin/env python
"""Sentence encoder (Convolution 5 x D -> BN -> max-over-time Pooling -> TanH) forward + backward at a C3-sized batch:
per-kernel device times (CUDA events around every launch, handle profiling), algorithmic TFLOP/s of the three
convolution contractions and GB/s of the HBM-bound layers.   python tools/sentenc_bench.py [N] [C]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(N=8192, L=40, D=300, C=100, kh=5, iters=5):
    import torch

    import mms_answer_selection_b200 as mms
    T = L - kh + 1
    g = torch.Generator(device="cuda").manual_seed(3)
    conv = mms.create_layer(mms.LayerParameter("Convolution", convolution_param=dict(
        num_output=C, kernel_h=kh, kernel_w=D, weight_filler=dict(type="xavier"))))
    bn = mms.create_layer(mms.LayerParameter("BN", bn_param=dict(scale_filler=dict(type="constant", value=1.0),
                                                                  shift_filler=dict(type="constant", value=1e-3))))
    pool = mms.create_layer(mms.LayerParameter("Pooling", pooling_param=dict(pool="MAX", kernel_h=T, kernel_w=1)))
    tanh = mms.create_layer(mms.LayerParameter("TanH"))
    x, y, z, p = mms.Blob((N, 1, L, D)), mms.Blob(()), mms.Blob(()), mms.Blob(())
    x.data.copy_(torch.rand((N, 1, L, D), device="cuda", generator=g) * 0.16 - 0.08)
    conv.SetUp([x], [y]); bn.SetUp([y], [z]); pool.SetUp([z], [p]); tanh.SetUp([p], [p])
    from mms_answer_selection_b200 import _lib
    conv.handle.set_option(_lib.MMS_OPT_REUSE_FORWARD, 1)     # Backward right after Forward on an unchanged bottom
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        conv.Forward([x], [y]); bn.Forward([y], [z]); pool.Forward([z], [p]); tanh.Forward([p], [p])
        p.diff.fill_(1.0 / N)
        tanh.Backward([p], [True], [p]); pool.Backward([p], [True], [z]); bn.Backward([z], [True], [y])
        conv.Backward([y], [True], [x])

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    layers = (conv, bn, pool, tanh)
    for l in layers:
        l.handle.profile_enable(True)
    evs = []
    for _ in range(iters):
        for _ in range(4):
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record()
        evs.append((e0, e1))
        torch.cuda.synchronize()
    prof = {}
    for name, l in zip(("conv", "bn", "pool", "tanh"), layers):
        for k, (n, ms) in l.handle.profile_report().items():
            prof["%s/%s" % (name, k)] = {"launches_per_step": n / iters, "ms_per_step": round(ms / iters, 5)}
        l.handle.profile_enable(False)
    step_ms = sum(a.elapsed_time(b) for a, b in evs) / iters
    conv_ms = sum(v["ms_per_step"] for k, v in prof.items() if k.startswith("conv/"))
    gemm_ms = sum(v["ms_per_step"] for k, v in prof.items() if k.startswith("conv/tc_gemm") or k in ("conv/sentconv_fwd_kernel", "conv/sentconv_dx_kernel", "conv/sentconv_dw_kernel"))
    flops = 6.0 * N * T * kh * D * C
    act_bytes = 4.0 * N * C * T
    out = {
        "workload": "sentence encoder fwd+bwd: %d sentences x %d tokens x %d-d, Convolution(%d x %d, %d filters) -> BN -> "
                    "MAX over time -> TanH (eager launches, L2 flushed between steps)" % (N, L, D, kh, D, C),
        "ms_per_step": step_ms, "sentences_per_sec": N / (step_ms / 1e3),
        "conv_algorithmic_tflops": flops / (conv_ms / 1e3) / 1e12 if conv_ms else None,
        "conv_gemm_only_tflops": flops / (gemm_ms / 1e3) / 1e12 if gemm_ms else None,
        "hbm_gbs": {name: round(nbytes / 1e9 / (prof[key]["ms_per_step"] / 1e3), 1) for name, key, nbytes in (
            ("bn_statistics (2 launches: x; dtop and x_norm)", "bn/bn_channel_sums_rows_kernel", 3 * act_bytes),
            ("bn_normalize (x in; x_norm, top out)", "bn/bn_normalize_vec_kernel", 3 * act_bytes),
            ("bn_backward (dtop, x_norm in; dx out)", "bn/bn_backward_vec_kernel", 3 * act_bytes),
            ("max_over_time forward", "pool/pool_plane_max_vec_kernel", act_bytes),
            ("max_over_time backward", "pool/pool_plane_max_backward_vec_kernel", act_bytes),
            ("conv top copy + bias", "conv/sentconv_unpack_t_kernel", 2 * act_bytes),
            ("conv gradient transpose + bias grad", "conv/sentconv_pack_warp_kernel", 4.0 * N * L * 104 + act_bytes),
            ("TF32 rounding of x", "conv/tf32_round_kernel", 8.0 * N * L * D)) if key in prof},
        "kernels": prof,
    }
    return out


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    print(json.dumps(run(N=N, C=C, iters=iters), indent=1))
