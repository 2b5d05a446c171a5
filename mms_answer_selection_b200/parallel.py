"""Data-parallel gradient exchange for the MMS hot path: the host side of ``mms_exchange_*`` (include/mms_b200.h).

Replaces the reference's P2PSync (src/caffe/parallel.cpp): one process per GPU instead of one thread per GPU; the
binary-tree peer-memcpy reduce, root-only 1/n scale and solver step, and tree broadcast of the weights
(parallel.cpp:287-380) become ONE kernel per rank and bucket that reads the rank's slice of the flat gradient buffer
from every peer over NVLink, averages it (and, in the fused form, applies the AdaDelta solver step to the rank's slice of
the weights) and stores the result into every peer (csrc/exchange.cu).

Like ``Params`` / ``GPUParams`` in the reference (parallel.cpp:60-115) the learnable blobs are re-bound to views of two
flat device buffers (data, diff).  Blobs keep the caller's order (net order) and every blob starts on a 16-byte
boundary, so that a bucket -- a run of consecutive blobs -- is a contiguous, vector-aligned range.  Parameter sharing
(net.cpp:944-950) survives the re-binding: a sharer resolves its storage through the owning blob (blob.py).

Backends
  "p2p"   csrc/exchange.cu over peer memory; buffers are allocated by the library and mapped into every rank through
          cudaIpc handles (gathered with torch.distributed, which is plumbing only).  ``symmetric=True`` allocates the
          buffers from torch's symmetric memory instead and hands the NVSwitch multicast address to the kernel
          (multimem.ld_reduce / multimem.st path).
  "nccl"  torch.distributed all_reduce(AVG) on torch-allocated buffers: the library-collective baseline the bench
          compares against.
  "host"  any other torch.distributed backend on CPU tensors (the gloo tests of the host logic): sum, then the scaler.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import Handle, c_p, check, lib


def _scale_on_device(handle, flat, alpha):
    """x *= alpha through the C-ABI (mms_scale_f32/f64).  Device tensors only."""
    if not flat.is_cuda:
        raise RuntimeError("gradient scaling runs on the GPU only (no CPU fallback)")
    if flat.dtype == torch.float32:
        check(lib().mms_scale_f32(handle.ptr, c_p(flat.data_ptr()), flat.numel(), ctypes.c_float(alpha)))
    else:
        check(lib().mms_scale_f64(handle.ptr, c_p(flat.data_ptr()), flat.numel(), ctypes.c_double(alpha)))


class _RawCuda(object):
    """A device range owned by libmms_b200 (or symmetric memory), presented to torch without a copy."""

    def __init__(self, ptr, count, np_dtype):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": np.dtype(np_dtype).str,
                                         "data": (int(ptr), False), "version": 2}


def flat_layout(counts, elem_bytes):
    """Element offsets of the blobs in the flat buffers: caller's order, every blob on a 16-byte boundary.
    Returns ([(offset, count)], total elements incl. padding)."""
    vn = 16 // elem_bytes
    offs, off = [], 0
    for n in counts:
        offs.append((off, int(n)))
        off += (int(n) + vn - 1) // vn * vn
    return offs, off


class GradientExchange(object):
    def __init__(self, blobs, group=None, backend=None, scaler=None, symmetric=False):
        """``blobs``: learnable Blob list, net order, shared blobs once (Net::learnable_params).  ``scaler(flat, alpha)``
        overrides the device scaling of the "host" backend (tests inject a host function for gloo/CPU runs)."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.blobs = list(blobs)
        ref = self.blobs[0].data
        self.np_dtype = np.float32 if ref.dtype == torch.float32 else np.float64
        self.elem = 4 if ref.dtype == torch.float32 else 8
        if backend is None:
            backend = "p2p" if ref.is_cuda else "host"
        if backend not in ("p2p", "nccl", "host"):
            raise ValueError("backend must be p2p, nccl or host")
        if backend == "p2p" and not ref.is_cuda:
            raise RuntimeError("the peer-memory exchange runs on the GPU only (no CPU fallback)")
        self.backend = backend
        self.offsets, total = flat_layout([b.count() for b in self.blobs], self.elem)
        self.count = total
        self._x = None
        self._symm = None
        self.multicast = False
        if backend == "p2p":
            self._create_p2p(total, symmetric)
        else:
            self.flat_data = torch.zeros(total, dtype=ref.dtype, device=ref.device)
            self.flat_diff = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        for b, (off, n) in zip(self.blobs, self.offsets):
            self.flat_data[off:off + n].copy_(b.data.reshape(-1))
            self.flat_diff[off:off + n].copy_(b.diff.reshape(-1))
            b.set_data(self.flat_data[off:off + n])
            b.set_diff(self.flat_diff[off:off + n])
            b._exchange = self                      # the flat buffers live as long as a blob is bound to them
        self._handle = None
        self._scaler = scaler

    # ------------------------------------------------------------------ construction of the peer-memory backend
    def _create_p2p(self, total, symmetric):
        L = lib()
        x = c_p()
        nbytes = int(L.mms_exchange_bytes(total, self.elem))
        ext = None
        if symmetric and self.world > 1:
            ext = self._symmetric_alloc(nbytes)
        check(L.mms_exchange_create(ctypes.byref(x), self.rank, self.world, total, self.elem,
                                    c_p(ext[0]) if ext else c_p(0)))
        self._x = x
        if self.world > 1:
            if ext:
                arr = (c_p * self.world)(*[c_p(int(p)) for p in ext[1]])
                check(L.mms_exchange_attach_ptrs(x, arr, c_p(int(ext[2]) if ext[2] else 0)))
                self.multicast = bool(ext[2])
            else:
                mine = ctypes.create_string_buffer(_lib.MMS_EXCHANGE_IPC_BYTES)
                check(L.mms_exchange_export_ipc(x, mine))
                dev = torch.device("cuda", torch.cuda.current_device())
                t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).to(dev)
                allh = [torch.empty_like(t) for _ in range(self.world)]
                dist.all_gather(allh, t, group=self.group)
                blob = b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh)
                check(L.mms_exchange_attach_ipc(x, ctypes.create_string_buffer(blob, len(blob))))
        pd, pg = c_p(), c_p()
        check(L.mms_exchange_buffers(x, ctypes.byref(pd), ctypes.byref(pg)))
        dev = torch.device("cuda", torch.cuda.current_device())
        self.flat_data = torch.as_tensor(_RawCuda(pd.value, total, self.np_dtype), device=dev)
        self.flat_diff = torch.as_tensor(_RawCuda(pg.value, total, self.np_dtype), device=dev)
        if self.world > 1:
            dist.barrier(group=self.group)          # every rank has mapped every peer before the first kernel

    def _symmetric_alloc(self, nbytes):
        """(local base, [peer bases], multicast base or 0) from torch's symmetric memory (CUDA VMM + NVSwitch
        multicast objects); the tensor is kept alive by this object."""
        import torch.distributed._symmetric_memory as symm_mem
        dev = torch.device("cuda", torch.cuda.current_device())
        t = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        name = (self.group or dist.group.WORLD).group_name
        hdl = symm_mem.rendezvous(t, group=name)
        self._symm = (t, hdl)
        mc = int(hdl.multicast_ptr or 0)                  # 0 when the allocation has no NVSwitch multicast object
        return int(t.data_ptr()), [int(p) for p in hdl.buffer_ptrs], mc

    def close(self):
        if self._x is not None:
            lib().mms_exchange_destroy(self._x)
            self._x = None

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001
            pass

    # ------------------------------------------------------------------ layout
    def bucket(self, first=0, last=None):
        """Element range [begin, end) of blobs first..last-1 (a contiguous run of the flat buffers)."""
        last = len(self.blobs) if last is None else last
        begin = self.offsets[first][0]
        end = self.count if last >= len(self.blobs) else self.offsets[last][0]
        return begin, end

    def launch_count(self):
        return int(lib().mms_exchange_launch_count(self._x)) if self._x is not None else 0

    def set_option(self, opt, value):
        check(lib().mms_exchange_set_option(self._x, opt, int(value)))

    # ------------------------------------------------------------------ the exchange
    def _stream(self):
        return c_p(torch.cuda.current_stream().cuda_stream)

    def _scale(self, flat, alpha):
        if self._scaler is not None:
            self._scaler(flat, alpha)
            return
        if self._handle is None:
            self._handle = Handle()
        self._handle.set_stream(torch.cuda.current_stream().cuda_stream)
        _scale_on_device(self._handle, flat, alpha)

    def broadcast_params(self, src=0, channel=0):
        """Weight sync from ``src`` (P2PSync::on_start, parallel.cpp:287-322; the reference does it every iteration,
        with replicated or fused updates once at start-up is enough)."""
        if self.world <= 1:
            return
        if self.backend == "p2p":
            check(lib().mms_exchange_broadcast(self._x, self._stream(), channel, src))
        else:
            dist.broadcast(self.flat_data, src=src, group=self.group)

    def zero_grads(self):
        self.flat_diff.zero_()

    def allreduce(self, bucket=None, channel=0):
        """grad <- sum over ranks / n on every rank (parallel.cpp:325-380 and the 1/n of :377), for the whole flat
        buffer or one bucket, on the current stream."""
        begin, end = bucket if bucket is not None else (0, self.count)
        if self.backend == "p2p":
            fn = lib().mms_exchange_allreduce_f32 if self.elem == 4 else lib().mms_exchange_allreduce_f64
            real = ctypes.c_float if self.elem == 4 else ctypes.c_double
            check(fn(self._x, self._stream(), channel, begin, end, real(1.0 / self.world)))
            return
        if self.world <= 1:
            return
        part = self.flat_diff[begin:end]
        if self.backend == "nccl":
            # ONE collective with ncclAvg: the division happens inside the reduction
            dist.all_reduce(part, op=dist.ReduceOp.AVG, group=self.group)
            return
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        self._scale(part, 1.0 / self.world)

    def adadelta_step(self, lr_mult=None, decay_mult=None, base_lr=1.0, momentum=0.95, delta=5e-7, weight_decay=5e-4,
                      iter_size=1, bucket_blobs=None, channel=0, clear_diffs=True):
        """The fused tail: sum over ranks, 1/(n iter_size) scale, L2 decay, AdaDelta update on the owner's slice and
        the new weights stored into every rank (mms_exchange_adadelta).  ``bucket_blobs=(first, last)`` restricts it
        to a run of blobs.  Peer-memory backend only."""
        if self.backend != "p2p":
            raise RuntimeError("the fused exchange + solver step needs the peer-memory backend")
        nb = len(self.blobs)
        lr_mult = list(lr_mult) if lr_mult is not None else [1.0] * nb
        decay_mult = list(decay_mult) if decay_mult is not None else [1.0] * nb
        first, last = bucket_blobs if bucket_blobs is not None else (0, nb)
        begin, end = self.bucket(first, last)
        ends = [self.offsets[i + 1][0] if i + 1 < nb else self.count for i in range(first, last)]
        n = len(ends)
        seg_end = (ctypes.c_longlong * n)(*ends)
        seg_rate = (ctypes.c_double * n)(*[base_lr * lr_mult[i] for i in range(first, last)])
        seg_decay = (ctypes.c_double * n)(*[weight_decay * decay_mult[i] for i in range(first, last)])
        fn = lib().mms_exchange_adadelta_f32 if self.elem == 4 else lib().mms_exchange_adadelta_f64
        real = ctypes.c_float if self.elem == 4 else ctypes.c_double
        check(fn(self._x, self._stream(), channel, begin, end, real(1.0 / (self.world * iter_size)), seg_end, seg_rate,
                 seg_decay, n, real(momentum), real(delta), 1 if clear_diffs else 0))

    def history(self):
        """This rank's (hist_g, hist_u) flat tensors of the fused solver step (valid inside the slices it owns)."""
        pg, pu = c_p(), c_p()
        check(lib().mms_exchange_history(self._x, ctypes.byref(pg), ctypes.byref(pu)))
        if not pg.value:
            return None, None
        dev = self.flat_data.device
        return (torch.as_tensor(_RawCuda(pg.value, self.count, self.np_dtype), device=dev),
                torch.as_tensor(_RawCuda(pu.value, self.count, self.np_dtype), device=dev))

    def check(self):
        """Synchronises the current stream; raises if a peer failed to arrive within the time-out."""
        if self.backend == "p2p":
            check(lib().mms_exchange_check(self._x, self._stream()))
        else:
            torch.cuda.current_stream().synchronize() if self.flat_diff.is_cuda else None
