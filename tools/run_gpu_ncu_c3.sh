mkdir -p gpurun_out
python tools/trace_step.py c3 notrace > gpurun_out/plain_c3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:simcross2 -s 6 -c 3 -o gpurun_out/prof_c3_fused -f python tools/trace_step.py c3 notrace > gpurun_out/ncu_c3_full.log 2>&1
echo "full c3 rc=$?"
ls -la gpurun_out/*.ncu-rep
