mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_sentenc.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/pytest_exp.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_exp.log
timeout 300 python tools/sentenc_bench.py 8192 100 > gpurun_out/sentenc_bench.json 2> gpurun_out/sentenc_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/sentenc_bench.json"))
print(d["ms_per_step"], d["conv_gemm_only_tflops"], {k.split("/")[1]: v["ms_per_step"] for k, v in d["kernels"].items() if "tc_gemm" in k})
PY
timeout 300 python tools/gemm_bench.py 2>&1 | tail -12
