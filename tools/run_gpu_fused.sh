mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
MMS_TC_TRACE=1 timeout 120 python tools/simcross_bench.py c2 3 > gpurun_out/scb_c2_trace.log 2>&1; grep "trace" gpurun_out/scb_c2_trace.log | tail -8
for w in c2 c3; do
timeout 120 python tools/simcross_bench.py $w > gpurun_out/scb_${w}_fused.log 2>&1; tail -2 gpurun_out/scb_${w}_fused.log
MMS_NO_FUSED=1 timeout 120 python tools/simcross_bench.py $w > gpurun_out/scb_${w}_unfused.log 2>&1; tail -2 gpurun_out/scb_${w}_unfused.log
done
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; tail -12 gpurun_out/gemm_bench.log
