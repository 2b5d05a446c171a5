mkdir -p gpurun_out
set -o pipefail
for w in c2 c3; do
python tools/trace_step.py $w notrace > gpurun_out/plain_$w.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$w.csv python tools/trace_step.py $w notrace > gpurun_out/ncu_$w.log 2>&1
echo "launch list $w rc=$?"
done
python tools/trace_step.py c3 notrace > gpurun_out/plain_c3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:simcross2 -s 8 -c 4 -o gpurun_out/prof_c3_fused -f python tools/trace_step.py c3 notrace > gpurun_out/ncu_c3_full.log 2>&1
echo "full c3 rc=$?"
python tools/trace_step.py c2 notrace > gpurun_out/plain_c2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"simcross2|tc_gemm" -s 8 -c 4 -o gpurun_out/prof_c2_fused -f python tools/trace_step.py c2 notrace > gpurun_out/ncu_c2_full.log 2>&1
echo "full c2 rc=$?"
ls -la gpurun_out/*.ncu-rep
