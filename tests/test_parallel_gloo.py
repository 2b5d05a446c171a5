"""CPU test (gloo, world_size 2) of the data-parallel gradient exchange host logic:
flat-buffer layout, small-bucket-first ordering, sum over ranks and the 1/n scaling of
the reference (src/caffe/parallel.cpp:325-380).  The device scaling kernel is replaced by
a host function injected through ``scaler`` -- a test double that lives here, not in the
product (the product's default refuses non-CUDA tensors)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mms_answer_selection_b200.blob import Blob
from mms_answer_selection_b200.parallel import GradientExchange, _scale_on_device


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shapes = [(60, 5), (5,), (2, 5, 5), (2, 4, 4)]          # W, b, M, B
        rng = np.random.default_rng(100 + rank)
        blobs = []
        for shp in shapes:
            b = Blob(shp, device="cpu")
            b.set_cpu_data(rng.uniform(-1, 1, shp))
            b.set_cpu_diff(rng.uniform(-1, 1, shp))
            blobs.append(b)
        local = [b.cpu_diff().copy() for b in blobs]
        ex = GradientExchange(blobs, scaler=lambda flat, alpha: flat.mul_(alpha))
        # layout: smallest first, the V x D table last and alone in the second bucket
        assert [b.count() for b in ex.blobs] == sorted(b.count() for b in blobs)
        assert ex.split == ex.flat_diff.numel() - 300
        # blobs are views of the flat buffers
        blobs[0].diff[0, 0] = 42.0
        assert ex.flat_diff[ex.offsets[-1][0]].item() == 42.0
        blobs[0].diff[0, 0] = float(local[0][0, 0])
        ex.broadcast_params(0)
        ex.allreduce()
        out[rank] = dict(data=[b.cpu_data().copy() for b in blobs], diff=[b.cpu_diff().copy() for b in blobs],
                         local=local)
    finally:
        dist.destroy_process_group()


def test_gradient_exchange_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    for j in range(4):
        want = (r0["local"][j] + r1["local"][j]) / world          # sum, then * 1/n
        np.testing.assert_allclose(r0["diff"][j], want, rtol=1e-6, atol=1e-7)
        np.testing.assert_array_equal(r0["diff"][j], r1["diff"][j])   # replicas stay identical
        np.testing.assert_array_equal(r0["data"][j], r1["data"][j])   # params broadcast from rank 0


def test_device_scaler_refuses_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _scale_on_device(None, torch.zeros(4), 0.5)
