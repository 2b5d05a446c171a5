mkdir -p gpurun_out
timeout 60 tools/bin/smem_row_shift_test > gpurun_out/smem_row_shift.log 2>&1; cat gpurun_out/smem_row_shift.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
bash tools/run_gpu_bench_only.sh
