mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_c2.log
timeout 600 python bench.py --steps 10 --warmup 3 --workload c3 --no-cpu-baseline --no-extra > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_c3.log
