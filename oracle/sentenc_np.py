"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the sentence encoder's reference layers (float64 arithmetic unless
the inputs say otherwise), pinned against fixtures produced by the compiled reference (tests/golden/sentenc_golden.npz,
tests/test_oracle.py).  Never imported by the product.

  conv_*   ConvolutionLayer for a kernel as wide as the input   conv_layer.cpp:25-73, base_conv_layer.cpp:257-321
  bn_*     the fork's BNLayer                                     bn_layer.cpp:121-257 (forward), :261-384 (backward)
  pool_*   PoolingLayer MAX / AVE                                 pooling_layer.cpp:80-227
  tanh_*   TanHLayer                                              tanh_layer.cpp:11-37
"""
import numpy as np


def conv_forward(x, W, b):
    N, _, L, D = x.shape
    C, _, kh, _ = W.shape
    T = L - kh + 1
    win = np.stack([x[:, 0, t:t + kh, :] for t in range(T)], axis=1)            # (N, T, kh, D): the im2col rows
    y = np.einsum("ntid,cid->nct", win, W[:, 0])
    if b is not None:
        y = y + b[None, :, None]
    return y[..., None].astype(x.dtype)


def conv_backward(x, W, dy):
    """(dW, db, dx) of one backward call; the layer ADDS dW and db to the param diffs, dx overwrites."""
    N, _, L, D = x.shape
    C, _, kh, _ = W.shape
    T = L - kh + 1
    g = dy[..., 0]                                                               # (N, C, T)
    win = np.stack([x[:, 0, t:t + kh, :] for t in range(T)], axis=1)
    dW = np.einsum("nct,ntid->cid", g, win)[:, None]
    db = g.sum(axis=(0, 2))
    dx = np.zeros_like(x)
    for t in range(T):                                                           # col2im: overlapping windows add up
        dx[:, 0, t:t + kh, :] += np.einsum("nc,cid->nid", g[:, :, t], W[:, 0])
    return dW.astype(x.dtype), db.astype(x.dtype), dx


def bn_forward(x, scale, shift, run_mean, run_var, train=True, memory=0.9, eps=1e-9):
    """Returns top, x_norm, std, new running mean, new running variance."""
    if train:
        mean = x.mean(axis=(0, 2, 3))
        var = (x * x).mean(axis=(0, 2, 3)) - mean * mean                         # var = E[x^2] - E[x]^2   :131-165
        run_mean = (1 - memory) * mean + memory * run_mean.reshape(-1)           # :168-172
        run_var = (1 - memory) * var + memory * run_var.reshape(-1)
    else:
        mean, var = run_mean.reshape(-1), run_var.reshape(-1)                    # :177-182
    std = np.sqrt(var + eps)
    xn = (x - mean[None, :, None, None]) / std[None, :, None, None]
    top = xn * scale.reshape(1, -1, 1, 1) + shift.reshape(1, -1, 1, 1)
    return top, xn, std, run_mean, run_var


def bn_backward(dtop, xn, scale, std):
    m = dtop.shape[0] * dtop.shape[2] * dtop.shape[3]
    dscale = (dtop * xn).sum(axis=(0, 2, 3))                                     # overwritten, not accumulated  :271-292
    dshift = dtop.sum(axis=(0, 2, 3))
    t = dtop * scale.reshape(1, -1, 1, 1)
    s_xt = (xn * t).sum(axis=(0, 2, 3))[None, :, None, None]
    s_t = t.sum(axis=(0, 2, 3))[None, :, None, None]
    dx = (t - (xn * s_xt + s_t) / m) / std[None, :, None, None]                  # :296-384
    return dscale, dshift, dx


def pooled_size(n, k, s, pad):
    p = -(-(n + 2 * pad - k) // s) + 1
    if pad and (p - 1) * s >= n + pad:
        p -= 1
    return p


def pool_forward(x, kh, kw, sh=1, sw=1, ph=0, pw=0, method="MAX"):
    N, C, H, W = x.shape
    PH, PW = pooled_size(H, kh, sh, ph), pooled_size(W, kw, sw, pw)
    top = np.zeros((N, C, PH, PW), x.dtype)
    mask = np.full((N, C, PH, PW), -1, np.int64)
    for i in range(PH):
        for j in range(PW):
            hs, ws = i * sh - ph, j * sw - pw
            if method == "MAX":
                he, we = min(hs + kh, H), min(ws + kw, W)
                hs, ws = max(hs, 0), max(ws, 0)
                win = x[:, :, hs:he, ws:we].reshape(N, C, -1)
                arg = win.argmax(axis=2)                                         # first maximum in scan order  :150-163
                top[:, :, i, j] = np.take_along_axis(win, arg[..., None], 2)[..., 0]
                mask[:, :, i, j] = (hs + arg // (we - ws)) * W + ws + arg % (we - ws)
            else:
                he, we = min(hs + kh, H + ph), min(ws + kw, W + pw)
                size = (he - hs) * (we - ws)                                     # the divisor counts the padding  :186-203
                hs, ws, he, we = max(hs, 0), max(ws, 0), min(he, H), min(we, W)
                top[:, :, i, j] = x[:, :, hs:he, ws:we].sum(axis=(2, 3)) / size
    return top, mask


def pool_backward(dtop, mask, xshape, kh, kw, sh=1, sw=1, ph=0, pw=0, method="MAX"):
    N, C, H, W = xshape
    dx = np.zeros(xshape, dtop.dtype)
    PH, PW = dtop.shape[2:]
    flat = dx.reshape(N, C, -1)
    for i in range(PH):
        for j in range(PW):
            if method == "MAX":
                np.add.at(flat, (np.arange(N)[:, None], np.arange(C)[None, :], mask[:, :, i, j]), dtop[:, :, i, j])
            else:
                hs, ws = i * sh - ph, j * sw - pw
                he, we = min(hs + kh, H + ph), min(ws + kw, W + pw)
                size = (he - hs) * (we - ws)
                hs, ws, he, we = max(hs, 0), max(ws, 0), min(he, H), min(we, W)
                dx[:, :, hs:he, ws:we] += (dtop[:, :, i, j] / size)[:, :, None, None]
    return dx


def tanh_forward(x):
    return np.tanh(x)


def tanh_backward(y, dy):
    return dy * (1 - y * y)
