mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/sentenc_bench.py 8192 100 > gpurun_out/sentenc_bench.json 2> gpurun_out/sentenc_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/sentenc_bench.json"))
print(d["ms_per_step"], d["conv_algorithmic_tflops"], d["conv_gemm_only_tflops"]); print(json.dumps(d["hbm_gbs"],indent=0))
for k,v in d["kernels"].items(): print("%-45s %.4f x%d" % (k, v["ms_per_step"], v["launches_per_step"]))
PY
tail -5 gpurun_out/sentenc_bench.err
