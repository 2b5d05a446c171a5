"""Blob: the host-side mirror of caffe::Blob (reference include/caffe/blob.hpp,
src/caffe/blob.cpp) restricted to what the MMS layers use: an N-d shape, a ``data`` and a
``diff`` array of the same shape living in device memory, legacy num/channels/height/
width accessors, and explicit host<->device transfer (the lazy SyncedMemory state
machine of the reference, syncedmem.cpp:25-77, becomes explicit ``set_cpu_*`` /
``cpu_*`` calls).  PyTorch tensors provide the device allocation only."""
import numpy as np
import torch

_TORCH = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


class Blob(object):
    def __init__(self, shape=(), dtype=np.float32, device="cuda"):
        self.dtype = np.dtype(dtype)
        if self.dtype not in _TORCH:
            raise TypeError("Blob dtype must be float32 or float64 (Caffe's Dtype)")
        self.device = torch.device(device)
        self._shape = ()
        self._data = None
        self._diff = None
        # parameter sharing (net.cpp:944-950): caffe::Blob::ShareData makes two blobs hold the SAME SyncedMemory
        # object, so re-pointing that memory (P2PSync's Params rebinding the learnable blobs to a flat buffer,
        # parallel.cpp:110-115) is seen by every sharer.  Here a sharer keeps a reference to the owning Blob and
        # resolves data/diff through it on every access instead of snapshotting a tensor.
        self._data_owner = None
        self._diff_owner = None
        self.Reshape(shape)

    # -- shape ------------------------------------------------------------------
    def Reshape(self, shape):
        shape = tuple(int(s) for s in shape)
        if any(s < 0 for s in shape):
            raise ValueError("negative blob dimension")
        n = int(np.prod(shape)) if shape else 0
        if n != self.count():
            self._data = None   # reallocated lazily, like blob.cpp:40-44
            self._diff = None
            self._data_owner = None
            self._diff_owner = None
        self._shape = shape
        view = shape if shape else (0,)
        if self._data is not None:
            self._data = self._data.reshape(view)
        if self._diff is not None:
            self._diff = self._diff.reshape(view)

    @property
    def shape(self):
        return self._shape

    def num_axes(self):
        return len(self._shape)

    def count(self, start=0, end=None):
        dims = self._shape[start:end]
        if not self._shape:
            return 0
        return int(np.prod(dims)) if dims else 1

    def _legacy(self, i):
        if len(self._shape) > 4:
            raise ValueError("legacy accessors need <= 4 axes")   # blob.hpp:137-141
        return self._shape[i] if i < len(self._shape) else 1

    def num(self):
        return self._legacy(0)

    def channels(self):
        return self._legacy(1)

    def height(self):
        return self._legacy(2)

    def width(self):
        return self._legacy(3)

    # -- storage ----------------------------------------------------------------
    def _alloc(self):
        return torch.zeros(self._shape if self._shape else (0,), dtype=_TORCH[self.dtype], device=self.device)

    @property
    def data(self):
        if self._data_owner is not None:
            return self._data_owner.data.view(self._shape if self._shape else (0,))
        if self._data is None:
            self._data = self._alloc()
        return self._data

    @property
    def diff(self):
        if self._diff_owner is not None:
            return self._diff_owner.diff.view(self._shape if self._shape else (0,))
        if self._diff is None:
            self._diff = self._alloc()
        return self._diff

    def gpu_data(self):
        return self.data.data_ptr()

    def gpu_diff(self):
        return self.diff.data_ptr()

    def set_data(self, t):
        """Adopt an existing device tensor as data (no copy); used to build flat
        parameter buffers for the data-parallel exchange (cf. parallel.cpp:110-115)."""
        assert t.numel() == max(self.count(), 0) and t.dtype == _TORCH[self.dtype]
        if self._data_owner is not None:          # the shared storage is re-pointed, for every sharer
            self._data_owner.set_data(t)
            return
        self._data = t.view(self._shape)

    def set_diff(self, t):
        assert t.numel() == max(self.count(), 0) and t.dtype == _TORCH[self.dtype]
        if self._diff_owner is not None:
            self._diff_owner.set_diff(t)
            return
        self._diff = t.view(self._shape)

    def set_cpu_data(self, arr, non_blocking=False):
        src = torch.from_numpy(np.ascontiguousarray(arr, dtype=self.dtype).reshape(self._shape))
        self.data.copy_(src, non_blocking=non_blocking)
        if self.data.is_cuda:
            # a write the library cannot see: drop whatever it derived from older contents (rounded operand copies kept
            # by MMS_OPT_REUSE_FORWARD / MMS_OPT_STAGE_TF32 / mms_simcross_prepare)
            from . import _lib
            _lib.lib().mms_invalidate_caches()

    def set_cpu_diff(self, arr, non_blocking=False):
        src = torch.from_numpy(np.ascontiguousarray(arr, dtype=self.dtype).reshape(self._shape))
        self.diff.copy_(src, non_blocking=non_blocking)

    def cpu_data(self):
        return self.data.detach().cpu().numpy()

    def cpu_diff(self):
        return self.diff.detach().cpu().numpy()

    def _root(self, attr):
        b = self
        while getattr(b, attr) is not None:
            b = getattr(b, attr)
        return b

    def ShareData(self, other):
        """blob.cpp:148-151 (CHECK_EQ(count_, other.count())); net.cpp:944-950 parameter sharing."""
        if other.count() != self.count():
            raise ValueError("ShareData: count mismatch (%d vs %d)" % (self.count(), other.count()))
        root = other._root("_data_owner")
        self._data_owner = None if root is self else root
        self._data = None

    def ShareDiff(self, other):
        if other.count() != self.count():
            raise ValueError("ShareDiff: count mismatch (%d vs %d)" % (self.count(), other.count()))
        root = other._root("_diff_owner")
        self._diff_owner = None if root is self else root
        self._diff = None

    def shares_storage_with(self, other):
        return (self._root("_data_owner") is other._root("_data_owner") and
                self._root("_diff_owner") is other._root("_diff_owner"))
