mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_c2.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print(json.dumps(d['extra'].get('ranking_metrics')))"
