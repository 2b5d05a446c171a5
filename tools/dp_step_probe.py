#!/usr/bin/env python
"""Where does the data-parallel C3 step spend its time?  Run under torchrun (N ranks):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_step_probe.py
Times, as CUDA graphs with an L2 flush between replays (max over ranks):
  compute   the sharded step without any exchange
  serial    Backward complete, then ONE exchange launch of the whole flat buffer (the reference's order)
  overlap   table bucket on a private stream beside dM / dB, SimCross bucket after (the product's step)
  exchange  the exchange alone (whole buffer), for several CTA counts
  fused     overlap form with the fused AdaDelta tail instead of the plain all-reduce (a whole solver iteration)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
import mms_answer_selection_b200 as mms

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = bench.workload_config("c3", world)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {"world": world, "pairs_per_gpu": cfg["N"]}
net, full, sl = bench.build_net(cfg, world, rank)
net.capture(with_loss=True, clear_diffs=True)
out["compute_ms"] = bench._time_ms(lambda: net.replay(read_loss=False), 20, flush, world)
symmetric = '--symmetric' in sys.argv
ex = mms.GradientExchange(net.params(), symmetric=symmetric)
out['multicast'] = ex.multicast
ex.broadcast_params(0); ex.check()
for name, kw in (("serial", dict(overlap=False)), ("overlap", dict(overlap=True))):
    g = net.capture_exchange_step(ex, **kw)
    out[name + "_ms"] = bench._time_ms(g.replay, 20, flush, world)
    ex.check()
for ctas in ((0, 64, 32, 16) if symmetric else (0, 296, 112, 74)):
    ex.set_option(1, ctas)
    out["exchange_alone_ms_ctas%d" % ctas] = bench._time_ms(lambda: ex.allreduce(), 20, flush, world)
    g = net.capture_exchange_step(ex, overlap=True)
    out["overlap_ms_ctas%d" % ctas] = bench._time_ms(g.replay, 20, flush, world)
ex.set_option(1, 0)
ex.check()
solver = mms.AdaDeltaSolver(net.params(), lr_mult=[1.0, 2.0, 1.0, 1.0], decay_mult=[0.0, 0.0, 1.0, 1.0])
g = net.capture_exchange_step(ex, solver=solver)
out["fused_solver_step_ms"] = bench._time_ms(g.replay, 20, flush, world)
ex.check()
if rank == 0:
    print(json.dumps(out, indent=1))
dist.barrier()
dist.destroy_process_group()
