"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's ConvolutionLayer for stride 1, pad 0, group 1
(conv_layer.cpp:25-73: forward_cpu_gemm / weight_cpu_gemm / backward_cpu_gemm per sample over the im2col matrix,
base_conv_layer.cpp:257-321; im2col.cpp:24-57) and DropoutLayer (dropout_layer.cpp:38-69).  Pinned against fixtures
produced by the reference's own layer compiled in place (tests/golden/make_conv2d_golden.py, tests/test_oracle.py)."""
import numpy as np


def _cols(x, kh, kw):
    """im2col view: (N, OH, OW, C, kh, kw) -- im2col_cpu (im2col.cpp:24-57) without the copy."""
    win = np.lib.stride_tricks.sliding_window_view(x, (kh, kw), axis=(2, 3))      # (N, C, OH, OW, kh, kw)
    return win.transpose(0, 2, 3, 1, 4, 5)


def conv2d_forward(x, W, b=None):
    """top[n,o] = W[o] . col(x[n]) + b[o]   (base_conv_layer.cpp:257-281, forward_cpu_bias :283-287)."""
    Co, C, kh, kw = W.shape
    y = np.einsum("nyxckl,ockl->noyx", _cols(x, kh, kw), W, optimize=True)
    if b is not None:
        y = y + b.reshape(1, Co, 1, 1)
    return y.astype(x.dtype)


def conv2d_backward(x, W, dtop, dW, db=None, want_dx=True):
    """dW += sum_n dtop[n] col(x[n])^T (:300-310), db += sum dtop (:312-316), dx = col2im(W^T dtop[n]) (:289-298).
    dW / db accumulate in place; returns (dW, db, dx)."""
    Co, C, kh, kw = W.shape
    N, _, H, Wd = x.shape
    dW += np.einsum("noyx,nyxckl->ockl", dtop, _cols(x, kh, kw), optimize=True).astype(dW.dtype)
    if db is not None:
        db += dtop.sum(axis=(0, 2, 3)).astype(db.dtype)
    dx = None
    if want_dx:
        dx = np.zeros_like(x)
        OH, OW = H - kh + 1, Wd - kw + 1
        for ky in range(kh):
            for kx in range(kw):
                dx[:, :, ky:ky + OH, kx:kx + OW] += np.einsum("noyx,oc->ncyx", dtop, W[:, :, ky, kx], optimize=True)
    return dW, db, dx


def dropout(x, mask_words, ratio):
    """top = bottom * (mask > UINT_MAX * ratio) * 1 / (1 - ratio)   (dropout_layer.cu:10-45; same call for the gradient)."""
    thres = np.uint32(int(4294967295 * ratio))
    keep = (np.asarray(mask_words, dtype=np.uint32).reshape(x.shape) > thres)
    return (x * keep * x.dtype.type(1.0 / (1.0 - ratio))).astype(x.dtype)
