mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 120 python tools/simcross_bench.py c3 > gpurun_out/scb_c3_fused.log 2>&1; tail -2 gpurun_out/scb_c3_fused.log
timeout 120 python tools/simcross_bench.py c2 > gpurun_out/scb_c2_fused.log 2>&1; tail -2 gpurun_out/scb_c2_fused.log
MMS_TC_TRACE=1 timeout 120 python tools/simcross_bench.py c3 3 > gpurun_out/scb_c3_trace.log 2>&1; grep "trace\]" gpurun_out/scb_c3_trace.log | tail -4 > gpurun_out/fused_trace_c3.log; cat gpurun_out/fused_trace_c3.log
