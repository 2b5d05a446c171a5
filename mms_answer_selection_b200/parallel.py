"""Data-parallel gradient exchange for the MMS hot path.

Replaces the reference's P2PSync (src/caffe/parallel.cpp): one process per GPU instead of
one thread per GPU; the binary-tree peer-memcpy reduce + root scale + weight broadcast
(parallel.cpp:287-380) becomes one NCCL all-reduce(sum) of the flat gradient buffer over
NVLink followed by the same 1/n scaling (:377) on every rank -- every rank then applies
the identical update, so no weight broadcast is needed.

Like ``Params`` in the reference (parallel.cpp:60-115) the learnable blobs are re-bound to
views of two contiguous device buffers (data, diff) so that the exchange is a single
message.  The buffer is ordered SMALL PARAMS FIRST (M, B, b ... then the V x D embedding
table) and exchanged in two buckets, so the latency-bound small bucket can be issued as
soon as SimCross backward has produced it, overlapping the Embed scatter.
"""
import ctypes

import torch
import torch.distributed as dist

from ._lib import Handle, c_p, check, lib


def _scale_on_device(handle, flat, alpha):
    """x *= alpha through the C-ABI (mms_scale_f32/f64).  Device tensors only."""
    if not flat.is_cuda:
        raise RuntimeError("gradient scaling runs on the GPU only (no CPU fallback)")
    if flat.dtype == torch.float32:
        check(lib().mms_scale_f32(handle.ptr, c_p(flat.data_ptr()), flat.numel(), ctypes.c_float(alpha)))
    else:
        check(lib().mms_scale_f64(handle.ptr, c_p(flat.data_ptr()), flat.numel(), ctypes.c_double(alpha)))


class GradientExchange(object):
    def __init__(self, blobs, group=None, scaler=None, small_first=True):
        """``blobs``: learnable Blob list (shared blobs once).  ``scaler(flat, alpha)``
        overrides the device scaling (tests inject a host function for gloo/CPU runs)."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        order = sorted(range(len(blobs)), key=lambda i: blobs[i].count()) if small_first else list(range(len(blobs)))
        self.blobs = [blobs[i] for i in order]
        total = sum(b.count() for b in self.blobs)
        ref = self.blobs[0].data
        self.flat_data = torch.empty(total, dtype=ref.dtype, device=ref.device)
        self.flat_diff = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self.offsets = []
        off = 0
        for b in self.blobs:
            n = b.count()
            self.flat_data[off:off + n].copy_(b.data.reshape(-1))
            self.flat_diff[off:off + n].copy_(b.diff.reshape(-1))
            b.set_data(self.flat_data[off:off + n])
            b.set_diff(self.flat_diff[off:off + n])
            self.offsets.append((off, n))
            off += n
        # bucket boundary: everything but the largest blob goes first
        self.split = self.offsets[-1][0] if len(self.blobs) > 1 else 0
        self._handle = None
        self._scaler = scaler

    def _scale(self, flat, alpha):
        if self._scaler is not None:
            self._scaler(flat, alpha)
            return
        if self._handle is None:
            self._handle = Handle()
        self._handle.set_stream(torch.cuda.current_stream().cuda_stream)
        _scale_on_device(self._handle, flat, alpha)

    def broadcast_params(self, src=0):
        """Initial weight sync (the reference broadcasts every iteration, parallel.cpp:287-322;
        with replicated updates once is enough)."""
        if self.world > 1:
            dist.broadcast(self.flat_data, src=src, group=self.group)

    def zero_grads(self):
        self.flat_diff.zero_()

    def allreduce_small(self, async_op=False):
        if self.world > 1 and self.split > 0:
            return dist.all_reduce(self.flat_diff[:self.split], op=dist.ReduceOp.SUM, group=self.group,
                                   async_op=async_op)
        return None

    def allreduce_large(self, async_op=False):
        if self.world > 1:
            return dist.all_reduce(self.flat_diff[self.split:], op=dist.ReduceOp.SUM, group=self.group,
                                   async_op=async_op)
        return None

    def finish(self):
        """grad = sum over ranks / n   (parallel.cpp:377)."""
        if self.world > 1:
            self._scale(self.flat_diff, 1.0 / self.world)

    def allreduce(self):
        """Sum over ranks and the 1/n scale (parallel.cpp:287-380) for the whole flat buffer.  Over NCCL this is ONE
        collective with ncclAvg (the division happens inside the reduction: no second pass over the 73.5 MB
        buffer, no second launch); other backends (the gloo tests) sum and scale."""
        if self.world <= 1:
            return
        if self.flat_diff.is_cuda and dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat_diff, op=dist.ReduceOp.AVG, group=self.group)
            return
        self.allreduce_small()
        self.allreduce_large()
        self.finish()
