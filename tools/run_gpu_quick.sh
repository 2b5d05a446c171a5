mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "simcross or SimCross or fused" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log
timeout 120 python tools/simcross_bench.py c3 > gpurun_out/scb_c3_fused.log 2>&1; tail -2 gpurun_out/scb_c3_fused.log
timeout 120 python tools/simcross_bench.py c2 > gpurun_out/scb_c2_fused.log 2>&1; tail -2 gpurun_out/scb_c2_fused.log
