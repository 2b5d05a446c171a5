mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; tail -2 gpurun_out/bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c2.log").read().strip().splitlines()[-1])
print(round(d["value"]), "pairs/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"]), d.get("cpu_baseline"))
r = json.loads(open("gpurun_out/bench_ref.log").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
