mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/sentenc_bench.py 8192 100 > gpurun_out/sentenc_bench.json 2> gpurun_out/sentenc_bench.err; echo "bench rc=$?"; cat gpurun_out/sentenc_bench.json; tail -5 gpurun_out/sentenc_bench.err
