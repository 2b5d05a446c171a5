mkdir -p gpurun_out
bash tools/run_gpu_quick.sh
python tools/trace_step.py c3 notrace > gpurun_out/plain_c3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"embed_backward|embed_forward|tc_gemm_tma|tf32_round" -s 12 -c 6 -o gpurun_out/prof_c3_misc -f python tools/trace_step.py c3 notrace > gpurun_out/ncu_c3_misc.log 2>&1
echo "misc c3 rc=$?"
