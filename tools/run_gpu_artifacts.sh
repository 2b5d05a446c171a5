# regenerates the measured artefacts that profiles/ keeps (plain runs, no profiler)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --workload c3 --no-cpu-baseline --no-extra > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
MMS_TC_TRACE=1 timeout 120 python tools/simcross_bench.py c3 3 > gpurun_out/scb_c3_trace.log 2>&1; grep "trace\]" gpurun_out/scb_c3_trace.log | tail -4 > gpurun_out/fused_trace_c3.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1
python - <<'PY'
import json
for f in ("bench_c2", "bench_c3"):
    d = json.loads(open("gpurun_out/%s.log" % f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), "pairs/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"]), "; roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "step", round(d["roofline"]["step_contractions"]["frac"], 3))
PY
