"""CPU tests (gloo, world_size 2) of the host logic of the data-parallel gradient exchange: the flat-buffer layout
(net order, 16-byte aligned blob offsets, buckets as contiguous ranges), parameter sharing that survives the re-binding
of the blobs (net.cpp:944-950 + parallel.cpp:110-115), the sum over ranks and the 1/n scaling of the reference
(src/caffe/parallel.cpp:325-380).  The device kernels are replaced by a host function injected through ``scaler`` -- a
test double that lives here, not in the product (the product's defaults refuse non-CUDA tensors)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mms_answer_selection_b200.blob import Blob
from mms_answer_selection_b200.parallel import GradientExchange, _scale_on_device, flat_layout


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_flat_layout_aligns_every_blob_to_16_bytes():
    offs, total = flat_layout([18000600, 300, 360000, 6400], 4)
    assert offs == [(0, 18000600), (18000600, 300), (18000900, 360000), (18360900, 6400)] and total == 18367300
    offs, total = flat_layout([7, 5, 3], 4)                  # ragged blobs are padded to 4 floats
    assert offs == [(0, 7), (8, 5), (16, 3)] and total == 20
    offs, total = flat_layout([7, 5, 3], 8)                  # doubles: 2 elements per 16 bytes
    assert offs == [(0, 7), (8, 5), (14, 3)] and total == 18


def test_sharing_survives_rebinding():
    """The advisor's repro: b.ShareData(a); GradientExchange([a, ...]) must leave a and b on the same storage."""
    a, b, c = Blob((6, 5), device="cpu"), Blob((6, 5), device="cpu"), Blob((3,), device="cpu")
    a.set_cpu_data(np.arange(30).reshape(6, 5))
    b.ShareData(a); b.ShareDiff(a)
    assert b.data.data_ptr() == a.data.data_ptr()
    ex = GradientExchange([a, c], backend="host", scaler=lambda f, al: f.mul_(al))
    assert a.data.data_ptr() == ex.flat_data.data_ptr()
    assert b.data.data_ptr() == a.data.data_ptr() and b.diff.data_ptr() == a.diff.data_ptr()
    assert b.shares_storage_with(a)
    b.diff[2, 3] = 7.0                                       # a sharer's gradient lands in the exchanged buffer
    assert ex.flat_diff[2 * 5 + 3].item() == 7.0
    np.testing.assert_array_equal(b.cpu_data(), np.arange(30).reshape(6, 5))
    with pytest.raises(ValueError):
        Blob((2, 2), device="cpu").ShareData(a)              # blob.cpp:148-151: counts must match


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shapes = [(60, 5), (5,), (2, 5, 5), (2, 3, 3)]          # W, b, M, B  (net order; B = 18 floats: ragged)
        rng = np.random.default_rng(100 + rank)
        blobs = []
        for shp in shapes:
            b = Blob(shp, device="cpu")
            b.set_cpu_data(rng.uniform(-1, 1, shp))
            b.set_cpu_diff(rng.uniform(-1, 1, shp))
            blobs.append(b)
        # a second Embed layer sharing W and b, as MMSNet's answer branch does
        sharers = [Blob(shapes[0], device="cpu"), Blob(shapes[1], device="cpu")]
        for s, o in zip(sharers, blobs[:2]):
            s.ShareData(o); s.ShareDiff(o)
        local = [b.cpu_diff().copy() for b in blobs]
        ex = GradientExchange(blobs, scaler=lambda flat, alpha: flat.mul_(alpha))
        assert ex.backend == "host"
        # layout: net order, 16-byte aligned offsets
        assert [o for o, _ in ex.offsets] == [0, 300, 308, 360] and ex.count == 380
        assert ex.bucket(0, 2) == (0, 308) and ex.bucket(2, None) == (308, 380) and ex.bucket() == (0, 380)
        # blobs and their sharers are views of the flat buffers
        sharers[0].diff[0, 0] += 1.0
        local[0][0, 0] += 1.0
        assert ex.flat_diff[0].item() == pytest.approx(float(local[0][0, 0]))
        assert sharers[0].data.data_ptr() == ex.flat_data.data_ptr()
        ex.broadcast_params(0)
        ex.allreduce(bucket=ex.bucket(0, 2))                    # the table bucket first ...
        ex.allreduce(bucket=ex.bucket(2, None))                 # ... then the SimCross bucket
        out[rank] = dict(data=[b.cpu_data().copy() for b in blobs], diff=[b.cpu_diff().copy() for b in blobs],
                         local=local, shared_diff=sharers[0].cpu_diff().copy())
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
    np.testing.assert_array_equal(r0["shared_diff"], r0["diff"][0])   # the sharer sees the exchanged gradient


def test_device_paths_refuse_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _scale_on_device(None, torch.zeros(4), 0.5)
    b = Blob((4,), device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GradientExchange([b], backend="p2p")
