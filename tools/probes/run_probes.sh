#!/bin/bash
# The stand-alone probes of round 2, logs into gpurun_out/.  Build them here first (no GPU needed), e.g.
#   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/gather4_probe tools/probes/gather4_probe.cu
# (tools/bin/ is git-ignored but travels to the GPU box).
mkdir -p gpurun_out
timeout 120 tools/bin/embed_bwd_probe > gpurun_out/probe_embed_bwd.log 2>&1; echo "embed_bwd rc=$?"
timeout 120 tools/bin/bulk_reduce_probe > gpurun_out/probe_bulk_reduce.log 2>&1; echo "bulk_reduce rc=$?"
timeout 120 tools/bin/gather4_probe > gpurun_out/probe_gather4.log 2>&1; echo "gather4 rc=$?"
