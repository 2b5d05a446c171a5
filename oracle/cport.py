"""TEST INFRASTRUCTURE ONLY -- numpy-level wrappers over the plain-C oracle
(oracle/mms_oracle.c, built to oracle/_build/libmms_oracle.so).

Each function mirrors one reference routine; see mms_oracle_impl.h for the
file:line citations.  dtype float32 -> *_f32, float64 -> *_f64.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.
"""
import ctypes

import numpy as np

from .refbind import oracle_lib

_P = ctypes.c_void_p


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "_f32", ctypes.c_float
    if dtype == np.float64:
        return "_f64", ctypes.c_double
    raise TypeError("oracle supports float32/float64 only, got %s" % dtype)


def _ptr(x):
    return None if x is None else ctypes.c_void_p(x.ctypes.data)


def _c(x, dtype):
    return None if x is None else np.ascontiguousarray(x, dtype=dtype)


def _call(name, dtype, *args):
    suf, _ = _suffix(dtype)
    fn = getattr(oracle_lib(), name + suf)
    fn.restype = ctypes.c_int
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError("%s%s failed with code %d" % (name, suf, rc))


def embed_forward(idx, W, bias=None):
    dt = W.dtype
    idx_c, W_c, b_c = _c(idx, dt), _c(W, dt), _c(bias, dt)
    V, D = W_c.shape
    M = idx_c.size
    top = np.empty(idx_c.shape + (D,), dtype=dt)
    _call("mmso_embed_forward", dt, _ptr(idx_c), _ptr(W_c), _ptr(b_c), _ptr(top), M, D, V)
    return top


def embed_backward(idx, dtop, dW, dbias=None):
    """Accumulates into dW / dbias in place (Caffe param-diff semantics)."""
    dt = dW.dtype
    idx_c, g = _c(idx, dt), _c(dtop, dt)
    V, D = dW.shape
    assert dW.flags.c_contiguous and (dbias is None or dbias.flags.c_contiguous)
    _call("mmso_embed_backward", dt, _ptr(idx_c), _ptr(g), _ptr(dW), _ptr(dbias), idx_c.size, D, V)
    return dW, dbias


def simcross_forward(mode, q, a, M=None, B=None):
    dt = q.dtype
    q_c, a_c, M_c, B_c = _c(q, dt), _c(a, dt), _c(M, dt), _c(B, dt)
    N, Lq, D = q_c.shape
    La = a_c.shape[1]
    mc = M_c.shape[0] if mode == 2 else 1
    S = np.empty((N, mc, Lq, La), dtype=dt)
    n0 = np.zeros((N, Lq), dtype=dt)
    n1 = np.zeros((N, La), dtype=dt)
    _call("mmso_simcross_forward", dt, mode, _ptr(q_c), _ptr(a_c), _ptr(M_c), _ptr(B_c),
          _ptr(S), _ptr(n0), _ptr(n1), N, Lq, La, D, mc)
    return S, n0, n1


def simcross_backward(mode, q, a, M, S, dS, n0=None, n1=None, dB=None, prop=(True, True)):
    """Returns (dq, da, dM, dB).  dB accumulates into the array passed in (or zeros)."""
    dt = q.dtype
    q_c, a_c, M_c, S_c, G = _c(q, dt), _c(a, dt), _c(M, dt), _c(S, dt), _c(dS, dt)
    N, Lq, D = q_c.shape
    La = a_c.shape[1]
    mc = M_c.shape[0] if mode == 2 else 1
    dq = np.empty_like(q_c)
    da = np.empty_like(a_c)
    dM = np.zeros((mc, D, D), dtype=dt) if mode == 2 else None
    if mode == 2 and dB is None:
        dB = np.zeros((mc, Lq, La), dtype=dt)
    _call("mmso_simcross_backward", dt, mode, _ptr(q_c), _ptr(a_c), _ptr(M_c), _ptr(S_c), _ptr(G),
          _ptr(_c(n0, dt)), _ptr(_c(n1, dt)), _ptr(dq), _ptr(da), _ptr(dM), _ptr(dB),
          N, Lq, La, D, mc, int(prop[0]), int(prop[1]))
    return dq, da, dM, dB


def simmatrix_forward(q, a, W):
    dt = q.dtype
    q_c, a_c, W_c = _c(q, dt).reshape(q.shape[0], -1), _c(a, dt).reshape(a.shape[0], -1), _c(W, dt)
    N, K1 = q_c.shape
    K2 = a_c.shape[1]
    s = np.empty((N, 1), dtype=dt)
    T = np.empty((N, K2), dtype=dt)
    _call("mmso_simmatrix_forward", dt, _ptr(q_c), _ptr(a_c), _ptr(W_c), _ptr(s), _ptr(T), N, K1, K2)
    return s, T


def simmatrix_backward(q, a, W, ds, dW, prop=(True, True), prop_w=True):
    """dW accumulates in place; returns (dW, dq, da)."""
    dt = q.dtype
    q_c, a_c, W_c = _c(q, dt).reshape(q.shape[0], -1), _c(a, dt).reshape(a.shape[0], -1), _c(W, dt)
    g = _c(ds, dt).reshape(-1)
    N, K1 = q_c.shape
    K2 = a_c.shape[1]
    dq = np.zeros_like(q_c)
    da = np.zeros_like(a_c)
    _call("mmso_simmatrix_backward", dt, _ptr(q_c), _ptr(a_c), _ptr(W_c), _ptr(g), _ptr(dW), _ptr(dq),
          _ptr(da), N, K1, K2, int(prop_w), int(prop[0]), int(prop[1]))
    return dW, dq, da


def pairrankloss_forward(a, b, y, margin=1.0):
    dt = a.dtype
    suf, cty = _suffix(dt)
    a_c, b_c, y_c = _c(a, dt), _c(b, dt), _c(y, dt)
    loss = np.zeros(1, dtype=dt)
    ordered = np.empty_like(a_c)
    similar = np.empty_like(a_c)
    _call("mmso_pairrankloss_forward", dt, _ptr(a_c), _ptr(b_c), _ptr(y_c), cty(margin), a_c.size,
          _ptr(loss), _ptr(ordered), _ptr(similar))
    return loss[0], ordered, similar


def pairrankloss_backward(y, ordered, similar, top_diff=1.0, ge=False, prop=(True, True)):
    dt = ordered.dtype
    suf, cty = _suffix(dt)
    y_c = _c(y, dt)
    da = np.empty_like(ordered) if prop[0] else None
    db = np.empty_like(ordered) if prop[1] else None
    _call("mmso_pairrankloss_backward", dt, _ptr(y_c), _ptr(ordered), _ptr(similar), cty(top_diff),
          ordered.size, int(ge), _ptr(da), _ptr(db))
    return da, db


def fm_forward(x, bias=None):
    dt = x.dtype
    x_c, b_c = _c(x, dt), _c(bias, dt)
    N, C, Dm = x_c.shape[:3]
    y = np.empty((N, 1), dtype=dt)
    _call("mmso_fm_forward", dt, _ptr(x_c), _ptr(b_c), _ptr(y), N, C, Dm)
    return y


def fm_backward(x, dy, bias_term=True, prop0=True):
    dt = x.dtype
    x_c, g = _c(x, dt), _c(dy, dt).reshape(-1)
    N, C, Dm = x_c.shape[:3]
    dx = np.zeros_like(x_c)
    db = np.zeros(1, dtype=dt) if bias_term else None
    _call("mmso_fm_backward", dt, _ptr(x_c), _ptr(g), _ptr(dx), _ptr(db), N, C, Dm, int(prop0))
    return dx, db


def adadelta_step(data, diff, hist_g, hist_u, grad_scale=1.0, local_decay=0.0, momentum=0.95, delta=5e-7,
                  local_rate=1.0):
    """In place on contiguous arrays of one dtype (data may be None: the bare adadelta_update)."""
    dtype = diff.dtype
    _, real = _suffix(dtype)
    for x in (data, diff, hist_g, hist_u):
        assert x is None or (x.dtype == dtype and x.flags["C_CONTIGUOUS"])
    _call("mmso_adadelta_step", dtype, _ptr(data), _ptr(diff), _ptr(hist_g), _ptr(hist_u),
          ctypes.c_longlong(diff.size), real(grad_scale), real(local_decay), real(momentum), real(delta),
          real(local_rate))


def map_mrr(data, label, group, fixed_axis=1):
    """(MAP, MRR) of MAPLayer / MRRLayer; data (N, fixed_axis + 1)."""
    dtype = data.dtype
    d, l, g = _c(data, dtype), _c(label, dtype), _c(group, dtype)
    out = np.zeros(2, dtype)
    _call("mmso_map_mrr", dtype, _ptr(d), _ptr(l), _ptr(g), int(l.size), int(fixed_axis),
          ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(out.ctypes.data + out.itemsize))
    return out[0], out[1]


def auc(data, label, fixed_axis=1, ignore_label=None):
    dtype = data.dtype
    d, l = _c(data, dtype), _c(label, dtype)
    out = np.zeros(1, dtype)
    _call("mmso_auc", dtype, _ptr(d), _ptr(l), int(l.size), int(d.size // l.size), int(fixed_axis),
          int(ignore_label is not None), int(ignore_label or 0), _ptr(out))
    return out[0]


def rank_accuracy(a, b, label):
    dtype = a.dtype
    out = np.zeros(1, dtype)
    _call("mmso_rank_accuracy", dtype, _ptr(_c(a, dtype)), _ptr(_c(b, dtype)), _ptr(_c(label, dtype)), int(a.size),
          _ptr(out))
    return out[0]
