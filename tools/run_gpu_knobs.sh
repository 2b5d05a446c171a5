mkdir -p gpurun_out
for k in ${KNOBS:-0 31 1 2 4 8}; do
echo "== MMS_BWD_DEBUG=$k $(MMS_BWD_DEBUG=$k timeout 120 python tools/simcross_bench.py c3 5 2>&1 | tail -1 | grep -o "'simcross2_bwd[^)]*)" | tr '\n' ' ')"
done > gpurun_out/knobs.log 2>&1
cat gpurun_out/knobs.log
