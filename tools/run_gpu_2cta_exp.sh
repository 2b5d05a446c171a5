mkdir -p gpurun_out
for v in 0 1; do
  if [ $v = 1 ]; then export MMS_NO_2CTA=1; fi
  timeout 300 python tools/sentenc_bench.py 8192 100 > gpurun_out/sentenc_2cta_$v.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/sentenc_2cta_$v.json"))
print("MMS_NO_2CTA=$v step %.3f ms" % d["ms_per_step"], {k.split("/")[1]: v["ms_per_step"] for k, v in d["kernels"].items() if "tc_gemm" in k})
PY
done
