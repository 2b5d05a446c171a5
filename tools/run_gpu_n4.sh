mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-300; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.log").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), "pairs/s", round(d["ms_per_step"],4), "ms; e2e", round(d["e2e"]["value"]), d["config"]["grad_exchange"], d["clocks"])
print(json.dumps(d.get("extra")))
print(json.dumps(d.get("kernels_ms_per_step")))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2> gpurun_out/bench_ref_n$N.err; echo "rc=$?"; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-400
