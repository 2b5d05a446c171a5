mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; tail -3 gpurun_out/bench_c2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c2.log").read().strip().splitlines()[-1])
print(round(d["value"]), "pairs/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"]))
for k, v in (d.get("extra") or {}).items():
    print(" ", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a not in ("kernels_ms_per_step", "hbm_gbs")})
PY
