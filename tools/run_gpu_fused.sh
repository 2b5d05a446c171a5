set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "simcross" > gpurun_out/pytest_fused.log 2>&1; echo "pytest rc=$?" 
tail -15 gpurun_out/pytest_fused.log
timeout 120 python tools/simcross_bench.py c2 > gpurun_out/scb_c2_fused.log 2>&1; tail -3 gpurun_out/scb_c2_fused.log
MMS_NO_FUSED=1 timeout 120 python tools/simcross_bench.py c2 > gpurun_out/scb_c2_unfused.log 2>&1; tail -3 gpurun_out/scb_c2_unfused.log
timeout 120 python tools/simcross_bench.py c3 > gpurun_out/scb_c3_fused.log 2>&1; tail -3 gpurun_out/scb_c3_fused.log
MMS_NO_FUSED=1 timeout 120 python tools/simcross_bench.py c3 > gpurun_out/scb_c3_unfused.log 2>&1; tail -3 gpurun_out/scb_c3_unfused.log
