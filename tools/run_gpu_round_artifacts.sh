# regenerates the measured artefacts that profiles/ keeps for the round (plain runs first, profiler runs after them)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --workload c3 --no-cpu-baseline --no-extra > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 300 python tools/sentenc_bench.py 8192 100 > gpurun_out/sentenc_bench.json 2> gpurun_out/sentenc_bench.err; echo "sentenc rc=$?"
python - <<'PY'
import json
for f in ("bench_c2", "bench_c3"):
    d = json.loads(open("gpurun_out/%s.log" % f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), "pairs/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"]), "; roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "step", round(d["roofline"]["step_contractions"]["frac"], 3), d["clocks"])
    print(json.dumps(d.get("hbm_kernels")))
    for k, v in (d.get("extra") or {}).items():
        print(" ", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a not in ("workload", "kernels_ms_per_step", "hbm_gbs")})
PY
# ---- profiler passes (never a source of bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sentenc.csv python tools/sentenc_bench.py 8192 100 1 > gpurun_out/ncu_sentenc.log 2>&1; echo "launch list sentenc rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_tma|bn_|pool_plane|sentconv|tf32_round" -s 30 -c 15 -o gpurun_out/prof_sentenc -f python tools/sentenc_bench.py 8192 100 1 > gpurun_out/ncu_sentenc_full.log 2>&1; echo "full sentenc rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
# ---- the main path: launch lists and full captures of the fused kernels (C2: 50 pairs, C3: 4096 pairs)
bash tools/run_gpu_ncu.sh
