#!/bin/bash
# The stand-alone probes of round 2 (binaries built into tools/bin/ by `make -C tools/probes` or the nvcc lines in each
# file's header), logs into gpurun_out/.
mkdir -p gpurun_out
timeout 120 tools/bin/embed_bwd_probe > gpurun_out/probe_embed_bwd.log 2>&1; echo "embed_bwd rc=$?"
timeout 120 tools/bin/bulk_reduce_probe > gpurun_out/probe_bulk_reduce.log 2>&1; echo "bulk_reduce rc=$?"
timeout 120 tools/bin/gather4_probe > gpurun_out/probe_gather4.log 2>&1; echo "gather4 rc=$?"
