#!/bin/bash
# One parameterised driver for everything that runs on the B200 box through gpurun:
#     gpurun --timeout 1500 -- 'bash tools/run_gpu.sh tests smoke bench ncu'
#     gpurun --gpus 2 --timeout 900 -- 'bash tools/run_gpu.sh multi:2'
# Stages (any order, run left to right); every artefact lands in gpurun_out/ (copy what should be
# judged into profiles/).  Plain runs come first; numbers printed under ncu are never bench values.
#   tests[:expr]   pytest -m gpu (optionally -k expr)
#   smoke          __graft_entry__.smoke()
#   bench[:wl]     bench.py on one GPU (default workload; wl = c1|c2|c3), both arms
#   quick[:wl]     bench.py --no-extra --no-cpu-baseline --steps 20 (kernel iteration loop)
#   multi:N        bench.py on N GPUs under torchrun (both arms) + the 2-rank exchange parity test
#   ncu[:wl]       launch list + `--set full` capture of the SimCross kernels of an eager step
#   ncu-sentenc    the same for the sentence encoder
#   tool:<file>    python tools/<file>.py (remaining words after ':' are its arguments, '+'-separated)
mkdir -p gpurun_out
set -o pipefail
summ() { python tools/bench_summary.py "$1"; }
for stage in "$@"; do
  name=${stage%%:*}; arg=""; [[ "$stage" == *:* ]] && arg=${stage#*:}
  case $name in
    tests)
      timeout 1500 python -m pytest tests -q -m gpu -rs ${arg:+-k "$arg"} > gpurun_out/pytest_gpu.log 2>&1
      echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log ;;
    smoke)
      timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
      echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log ;;
    bench)
      wl=${arg:-default}
      timeout 1200 python bench.py ${arg:+--workload $arg} > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err
      echo "bench $wl rc=$?"; tail -3 gpurun_out/bench_$wl.err; summ gpurun_out/bench_$wl.log
      timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ${arg:+--workload $arg} > gpurun_out/bench_ref_$wl.log 2> gpurun_out/bench_ref_$wl.err
      echo "bench ref $wl rc=$?"; tail -1 gpurun_out/bench_ref_$wl.log | cut -c1-400 ;;
    quick)
      wl=${arg:-c3}
      timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/quick_$wl.log 2> gpurun_out/quick_$wl.err
      echo "quick $wl rc=$?"; tail -3 gpurun_out/quick_$wl.err; summ gpurun_out/quick_$wl.log ;;
    multi)
      N=${arg:-2}
      timeout 600 python -m pytest tests -q -m gpu -rs -k "multigpu" > gpurun_out/pytest_multigpu_n$N.log 2>&1
      echo "pytest multigpu rc=$?"; tail -4 gpurun_out/pytest_multigpu_n$N.log
      NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
      echo "bench N=$N rc=$?"; grep -v NCCL gpurun_out/bench_n$N.err | tail -3; summ gpurun_out/bench_n$N.log
      grep -E "NCCL INFO (Connected|Channel 00/|.*NVLS|.*Algo|comm .* rank 0 )" gpurun_out/bench_n$N.err | head -40 > gpurun_out/nccl_info_n$N.log
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2> gpurun_out/bench_ref_n$N.err
      echo "bench ref N=$N rc=$?"; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-400 ;;
    ncu)
      wl=${arg:-c3}
      python tools/trace_step.py $wl notrace > gpurun_out/plain_$wl.log 2>&1 &&
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$wl.csv \
        python tools/trace_step.py $wl notrace > gpurun_out/ncu_$wl.log 2>&1
      echo "launch list $wl rc=$?"
      ncu --set full --clock-control none --import-source on -k regex:"simcross2|tc_gemm|embed_|tf32_round|short_runs|long_chunks|plan_" -s 13 -c 13 \
        -o gpurun_out/prof_${wl}_fused -f python tools/trace_step.py $wl notrace > gpurun_out/ncu_${wl}_full.log 2>&1
      echo "full $wl rc=$?"; ls -la gpurun_out/prof_${wl}_fused.ncu-rep ;;
    ncu-sentenc)
      python tools/sentenc_bench.py 8192 100 1 > gpurun_out/plain_sentenc.log 2>&1 &&
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sentenc.csv \
        python tools/sentenc_bench.py 8192 100 1 > gpurun_out/ncu_sentenc.log 2>&1
      echo "launch list sentenc rc=$?"
      ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_tma|bn_|pool_plane|sentconv|tf32_round" -s 30 -c 15 \
        -o gpurun_out/prof_sentenc -f python tools/sentenc_bench.py 8192 100 1 > gpurun_out/ncu_sentenc_full.log 2>&1
      echo "full sentenc rc=$?" ;;
    tool)
      f=${arg%%+*}; rest=""; [[ "$arg" == *+* ]] && rest=${arg#*+}
      timeout 900 python tools/$f.py ${rest//+/ } > gpurun_out/tool_$f.log 2>&1
      echo "tool $f rc=$?"; tail -25 gpurun_out/tool_$f.log ;;
    *) echo "unknown stage $stage" ;;
  esac
done
