mkdir -p gpurun_out
MMS_TC_TRACE=1 timeout 120 python tools/simcross_bench.py c3 3 > gpurun_out/scb_c3_trace.log 2>&1; grep "trace\] fwd" gpurun_out/scb_c3_trace.log | tail -1
for w in c2 c3; do
timeout 120 python tools/simcross_bench.py $w > gpurun_out/scb_${w}_fused.log 2>&1; tail -2 gpurun_out/scb_${w}_fused.log
done
