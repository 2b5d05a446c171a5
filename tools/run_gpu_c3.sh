mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin_layers.py -q -m gpu -x > gpurun_out/pytest_exp.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_exp.log
timeout 600 python bench.py --steps 20 --warmup 3 --workload c3 --no-cpu-baseline --no-extra > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c3.log").read().strip().splitlines()[-1])
print(round(d["value"]), "pairs/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"]))
print(d["kernels_ms_per_step"]); print(d["hbm_kernels"]["bias_grad_kernel"])
PY
