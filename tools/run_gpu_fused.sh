mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; tail -6 gpurun_out/gemm_bench.log
timeout 120 python tools/simcross_bench.py c3 > gpurun_out/scb_c3_fused.log 2>&1; tail -2 gpurun_out/scb_c3_fused.log
