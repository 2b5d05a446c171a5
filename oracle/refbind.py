"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the two CPU checkers.

* ``RefLayer``  drives ``oracle/_ref/libmms_ref.so``: the reference's own layer
  sources (``/root/reference/src/caffe/layers/{embed,sim_cross,sim_matrix,
  pair_rank_loss,fm}_layer.cpp``) compiled in place by ``oracle/Makefile`` and
  driven through ``LayerRegistry::CreateLayer`` / ``Layer::SetUp/Forward/Backward``.
* ``oracle_lib()`` loads ``oracle/_build/libmms_oracle.so``: the plain-C
  restatement in ``oracle/mms_oracle.c``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``mms_answer_selection_b200`` never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libmms_ref.so")
DROPIN_SO = os.path.join(_HERE, "_ref", "libmms_dropin.so")
REFCUDA_SO = os.path.join(_HERE, "_ref", "libmms_refcuda.so")
PRODUCT_SO = os.path.join(os.path.dirname(_HERE), "mms_answer_selection_b200", "libmms_b200.so")
ORACLE_SO = os.path.join(_HERE, "_build", "libmms_oracle.so")

_ref = None
_dropin = None
_oracle = None

_NP = {0: np.float32, 1: np.float64}
_KIND = {"bottom": 0, "top": 1, "blob": 2}


def build(ref=True, oracle=True, dropin=None):
    """Build the checkers (a no-op for ``_ref`` when /root/reference is absent).

    ``dropin`` (default: same as ``ref``, provided the product library has been built) also builds
    ``_ref/libmms_dropin.so``: the reference's Layer API / Blob / SyncedMemory in GPU mode linked with
    the product's C++ drop-in layers (mms_answer_selection_b200/caffe_layers/)."""
    targets = []
    if oracle:
        targets.append("oracle")
    if ref:
        targets.append("ref")
    if dropin is None:
        dropin = ref
    if dropin and os.path.exists(PRODUCT_SO):
        targets.append("dropin")
    if ref:
        targets.append("refcuda")
    subprocess.run(["make", "-C", _HERE, "-s"] + targets, check=True)


def ref_available():
    return os.path.exists(REF_SO)


def dropin_available():
    return os.path.exists(DROPIN_SO)


def _bind(path):
        L = ctypes.CDLL(path)
        L.mmsref_last_error.restype = ctypes.c_char_p
        L.mmsref_create.restype = ctypes.c_void_p
        L.mmsref_create.argtypes = [ctypes.c_char_p, ctypes.c_int]
        L.mmsref_destroy.argtypes = [ctypes.c_void_p]
        L.mmsref_set_i.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_longlong]
        L.mmsref_set_f.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
        L.mmsref_set_s.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        L.mmsref_seed.argtypes = [ctypes.c_uint]
        L.mmsref_add_bottom.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.mmsref_reshape_bottom.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.mmsref_setup.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.mmsref_num_blobs.argtypes = [ctypes.c_void_p]
        L.mmsref_shape.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.mmsref_write.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.mmsref_read.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.mmsref_forward.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.mmsref_backward.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.mmsref_time.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.mmsref_set_blas_threads.argtypes = [ctypes.c_int]
        return L


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(
                REF_SO + " missing: run `make -C oracle ref` where /root/reference exists")
        _ref = _bind(REF_SO)
    return _ref


def dropin_lib():
    """The reference's Layer API + Blob + SyncedMemory (GPU build) hosting the PRODUCT's layers."""
    global _dropin
    if _dropin is None:
        if not os.path.exists(DROPIN_SO):
            raise FileNotFoundError(
                DROPIN_SO + " missing: run `make -C oracle dropin` where /root/reference exists")
        _dropin = _bind(DROPIN_SO)
        _dropin.mmsref_set_mode.argtypes = [ctypes.c_int]
    return _dropin


_refcuda = None


def refcuda_available():
    return os.path.exists(REFCUDA_SO)


def refcuda_lib():
    """The reference's own adadelta_solver.cu + math_functions.cu (nvcc, sm_100a): the GPU-side solver step."""
    global _refcuda
    if _refcuda is None:
        if not os.path.exists(REFCUDA_SO):
            raise FileNotFoundError(REFCUDA_SO + " missing: run `make -C oracle refcuda` where /root/reference exists")
        L = ctypes.CDLL(REFCUDA_SO)
        L.mmsrefcu_apply_update.argtypes = [ctypes.c_int, ctypes.c_longlong] + [ctypes.c_void_p] * 4 + [ctypes.c_double] * 5
        _refcuda = L
    return _refcuda


def ref_apply_update(data, diff, h, h2, accum_normalization=1.0, local_decay=0.0, momentum=0.95, delta=5e-7,
                     local_rate=1.0):
    """One blob's SGDSolver::ApplyUpdate (GPU mode) through the reference's own CUDA routines, in place on torch CUDA
    tensors (float32 or float64); ``data=None`` runs the bare adadelta_update_gpu."""
    import torch
    dt = 0 if diff.dtype == torch.float32 else 1
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
    torch.cuda.synchronize()
    rc = refcuda_lib().mmsrefcu_apply_update(dt, diff.numel(), p(data), p(diff), p(h), p(h2), accum_normalization,
                                             local_decay, momentum, delta, local_rate)
    if rc != 0:
        raise RuntimeError("mmsrefcu_apply_update failed: %d" % rc)


class RefError(RuntimeError):
    """A CHECK / LOG(FATAL) fired inside the reference code."""


class RefLayer(object):
    """One reference layer instance plus its bottom/top blobs.

    ``params`` keys are the prototxt field names (see oracle/ref_harness.cpp):
    ints ``dist_mode, mesure_count, sim_cross.bias_term, num_output, input_dim,
    embed.bias_term, fm.bias_term``; floats ``margin, loss_weight,
    weight_filler.{value,min,max,mean,std}``; strings ``weight_filler.type,
    bias_filler.type, weight_source``.
    """

    _lib = staticmethod(lambda: ref_lib())

    def __init__(self, type_, bottoms, params=None, dtype=np.float32, num_top=1, seed=1701):
        self.L = self._lib()
        self.np = np.dtype(dtype)
        self.dt = 0 if self.np == np.float32 else 1
        self.h = self.L.mmsref_create(type_.encode(), self.dt)
        self.type = type_
        for k, v in (params or {}).items():
            kb = k.encode()
            if isinstance(v, str):
                rc = self.L.mmsref_set_s(self.h, kb, v.encode())
            elif isinstance(v, (bool, int, np.integer)):
                rc = self.L.mmsref_set_i(self.h, kb, int(v))
            else:
                rc = self.L.mmsref_set_f(self.h, kb, float(v))
            if rc != 0:
                raise KeyError("unknown reference param %r for %s" % (k, type_))
        self.nbottom = len(bottoms)
        for b in bottoms:
            b = np.ascontiguousarray(b, dtype=self.np)
            shp = (ctypes.c_int * b.ndim)(*b.shape)
            if self.L.mmsref_add_bottom(self.h, b.ndim, shp) < 0:
                raise RefError(self.L.mmsref_last_error().decode())
        for i, b in enumerate(bottoms):
            self.write("bottom", i, b)
        self.L.mmsref_seed(seed)
        self._ck(self.L.mmsref_setup(self.h, num_top))
        self.num_top = num_top

    def __del__(self):
        try:
            if self.h:
                self.L.mmsref_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RefError(self.L.mmsref_last_error().decode())

    def shape(self, kind, i):
        buf = (ctypes.c_int * 8)()
        nd = self.L.mmsref_shape(self.h, _KIND[kind], i, buf)
        return tuple(buf[a] for a in range(nd))

    def num_blobs(self):
        return self.L.mmsref_num_blobs(self.h)

    def write(self, kind, i, arr, diff=False):
        arr = np.ascontiguousarray(arr, dtype=self.np)
        assert arr.size == int(np.prod(self.shape(kind, i))), (arr.shape, self.shape(kind, i))
        self._ck(self.L.mmsref_write(self.h, _KIND[kind], i, int(diff), arr.ctypes.data))

    def read(self, kind, i, diff=False):
        out = np.empty(self.shape(kind, i), dtype=self.np)
        self._ck(self.L.mmsref_read(self.h, _KIND[kind], i, int(diff), out.ctypes.data))
        return out

    def forward(self):
        loss = ctypes.c_double(0)
        self._ck(self.L.mmsref_forward(self.h, ctypes.byref(loss)))
        return loss.value

    def backward(self, propagate_down=None):
        pd = propagate_down if propagate_down is not None else [True] * self.nbottom
        arr = (ctypes.c_int * self.nbottom)(*[int(bool(x)) for x in pd])
        self._ck(self.L.mmsref_backward(self.h, arr))

    def time(self, iters, backward=True, propagate_down=None):
        pd = propagate_down if propagate_down is not None else [True] * self.nbottom
        arr = (ctypes.c_int * self.nbottom)(*[int(bool(x)) for x in pd])
        ms = ctypes.c_double(0)
        self._ck(self.L.mmsref_time(self.h, iters, int(backward), arr, ctypes.byref(ms)))
        return ms.value


class DropinLayer(RefLayer):
    """Same driver, but the layer behind ``LayerRegistry::CreateLayer`` is the product's C++ drop-in
    class running through libmms_b200.so on the GPU (Caffe::GPU mode)."""
    _lib = staticmethod(lambda: dropin_lib())


def set_ref_blas_threads(n):
    ref_lib().mmsref_set_blas_threads(int(n))


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False, oracle=True)
        _oracle = ctypes.CDLL(ORACLE_SO)
    return _oracle
