mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sentenc.py -q -m gpu -x > gpurun_out/pytest_sentenc.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_sentenc.log
timeout 200 python tools/sentenc_bench.py 8192 100 > gpurun_out/sentenc_bench.json 2> gpurun_out/sentenc_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/sentenc_bench.json"))
print(d["ms_per_step"], d["conv_algorithmic_tflops"], d["conv_gemm_only_tflops"])
for k,v in d["kernels"].items():
    if k.startswith("conv/"): print("%-45s %.4f x%d" % (k, v["ms_per_step"], v["launches_per_step"]))
PY
tail -3 gpurun_out/sentenc_bench.err
