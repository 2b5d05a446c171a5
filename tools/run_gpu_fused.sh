mkdir -p gpurun_out
MMS_TC_TRACE=1 timeout 120 python tools/simcross_bench.py c3 3 > gpurun_out/scb_c3_trace.log 2>&1; grep "trace" gpurun_out/scb_c3_trace.log | tail -4
