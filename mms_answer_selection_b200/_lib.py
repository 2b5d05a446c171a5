"""ctypes binding of libmms_b200.so -- the reference-side stub for a Python host
(INTEGRATION.md shows the C++ one for Caffe).  Signatures follow include/mms_b200.h."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmms_b200.so")
_lib = None

c_int, c_ll, c_p = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
c_f, c_d = ctypes.c_float, ctypes.c_double

MMS_MATH_TF32, MMS_MATH_FP32 = 0, 1
MMS_OPT_MATH, MMS_OPT_PRL_GE, MMS_OPT_SCRATCH_BYTES, MMS_OPT_EMBED_DETERMINISTIC = 1, 2, 3, 4
MMS_OPT_REUSE_FORWARD, MMS_OPT_CONCURRENCY, MMS_OPT_STAGE_TF32, MMS_OPT_STAGE_ONLY = 5, 6, 7, 8
EMBED_GROUPED_MIN_ROWS = 32768      # csrc/embed_sorted.cu kMinRows: below it mms_embed_backward_pair runs the per-layer kernels
MMS_E_INVALID, MMS_E_UNSUPPORTED, MMS_E_NOMEM, MMS_E_FAULT = -1, -2, -3, -4
MMS_EXCHANGE_OPT_CTAS, MMS_EXCHANGE_OPT_TIMEOUT_MS, MMS_EXCHANGE_OPT_MULTICAST = 1, 2, 3
MMS_EXCHANGE_MAX_WORLD, MMS_EXCHANGE_CHANNELS, MMS_EXCHANGE_IPC_BYTES = 8, 4, 64


class MMSError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libmms_b200 error %d: %s" % (code, message))
        self.code = code


def lib_path():
    return _LIB_PATH


def _typed(real):
    p = c_p
    return {
        "mms_embed_forward": [p, p, p, p, p, c_ll, c_int, c_int],
        "mms_embed_backward": [p, p, p, p, p, c_ll, c_int, c_int],
        "mms_simcross_forward": [p, c_int, p, p, p, p, p, p, p, c_int, c_int, c_int, c_int, c_int],
        "mms_simcross_backward": [p, c_int, p, p, p, p, p, p, p, p, p, p, p, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_int],
        "mms_simmatrix_forward": [p, p, p, p, p, p, c_int, c_int, c_int],
        "mms_simmatrix_backward": [p, p, p, p, p, p, p, p, c_int, c_int, c_int, c_int, c_int, c_int],
        "mms_pairrankloss_forward": [p, p, p, p, real, c_ll, p, p, p],
        "mms_pairrankloss_backward": [p, p, p, p, real, c_ll, p, p],
        "mms_fm_forward": [p, p, p, p, c_int, c_int, c_int],
        "mms_fm_backward": [p, p, p, p, p, c_int, c_int, c_int, c_int],
        "mms_dot": [p, p, p, c_ll, p],
        "mms_scale": [p, p, c_ll, real],
        "mms_adadelta_update": [p, p, p, p, c_ll, real, real, real],
        "mms_rank_map_mrr": [p, p, c_ll, c_ll, p, p, c_ll, p, p],
        "mms_rank_auc": [p, p, c_ll, c_ll, p, c_ll, c_int, c_int, p],
        "mms_rank_accuracy": [p, p, p, p, c_ll, p],
        "mms_sentconv_forward": [p, p, p, p, p, c_int, c_int, c_int, c_int, c_int],
        "mms_sentconv_backward": [p, p, p, p, p, p, p, c_int, c_int, c_int, c_int, c_int],
        "mms_pool_forward": [p, p, p, p, c_ll] + [c_int] * 11,
        "mms_pool_backward": [p, p, p, p, c_ll] + [c_int] * 11,
        "mms_tanh_forward": [p, p, p, c_ll],
        "mms_tanh_backward": [p, p, p, p, c_ll],
        "mms_bn_forward": [p, p, p, p, p, p, p, p, p, p, c_int, c_int, c_int, c_int, real, real],
        "mms_bn_backward": [p, p, p, p, p, p, p, p, c_int, c_int, c_int],
        "mms_adadelta_step": [p, p, p, p, p, c_ll, real, real, real, real, real, c_int],
        "mms_conv2d_forward": [p, p, p, p, p] + [c_int] * 7,
        "mms_conv2d_backward": [p, p, p, p, p, p, p] + [c_int] * 7,
        "mms_dropout": [p, p, p, p, c_ll, ctypes.c_uint, real],
    }


def lib():
    """Load libmms_b200.so.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise MMSError(0, "%s is missing: run `python -m mms_answer_selection_b200.build` "
                              "(no CPU/PyTorch fallback exists)" % _LIB_PATH)
        L = ctypes.CDLL(_LIB_PATH)
        L.mms_last_error.restype = ctypes.c_char_p
        L.mms_create.argtypes = [ctypes.POINTER(c_p)]
        L.mms_destroy.argtypes = [c_p]
        L.mms_set_stream.argtypes = [c_p, c_p]
        L.mms_set_option.argtypes = [c_p, c_int, c_ll]
        L.mms_get_option.argtypes = [c_p, c_int, ctypes.POINTER(c_ll)]
        L.mms_reserve_scratch.argtypes = [c_p, c_ll]
        L.mms_invalidate_caches.argtypes = []
        L.mms_check_faults.argtypes = [c_p]
        L.mms_launch_count.argtypes = [c_p]
        L.mms_launch_count.restype = ctypes.c_ulonglong
        L.mms_profile_enable.argtypes = [c_p, c_int]
        L.mms_profile_report.argtypes = [c_p, ctypes.c_char_p, ctypes.c_size_t]
        for suffix, real in (("_f32", c_f), ("_f64", c_d)):
            for name, args in _typed(real).items():
                fn = getattr(L, name + suffix)
                fn.argtypes = args
                fn.restype = c_int
        for suffix in ("_f32", "_f64"):
            fn = getattr(L, "mms_load_weight_source" + suffix)
            fn.argtypes = [ctypes.c_char_p, c_p, c_ll, c_ll, ctypes.POINTER(c_ll)]
            fn.restype = c_int
        L.mms_rerank_scores_f32.argtypes = [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_ll, c_int, c_int]
        L.mms_rerank_prepare_f32.argtypes = [c_p, c_p, c_p, c_ll, c_int]
        L.mms_rerank_scores_prepared_f32.argtypes = [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_ll, c_int, c_int]
        c_pp = ctypes.POINTER(c_p)
        L.mms_exchange_bytes.argtypes = [c_ll, c_int]
        L.mms_exchange_bytes.restype = c_ll
        L.mms_exchange_create.argtypes = [c_pp, c_int, c_int, c_ll, c_int, c_p]
        L.mms_exchange_destroy.argtypes = [c_p]
        L.mms_exchange_buffers.argtypes = [c_p, c_pp, c_pp]
        L.mms_exchange_base.argtypes = [c_p, c_pp]
        L.mms_exchange_export_ipc.argtypes = [c_p, c_p]
        L.mms_exchange_attach_ipc.argtypes = [c_p, c_p]
        L.mms_exchange_attach_ptrs.argtypes = [c_p, c_pp, c_p]
        L.mms_exchange_set_option.argtypes = [c_p, c_int, c_ll]
        L.mms_exchange_allreduce_f32.argtypes = [c_p, c_p, c_int, c_ll, c_ll, c_f]
        L.mms_exchange_allreduce_f64.argtypes = [c_p, c_p, c_int, c_ll, c_ll, c_d]
        seg = [ctypes.POINTER(c_ll), ctypes.POINTER(c_d), ctypes.POINTER(c_d), c_int]
        L.mms_exchange_adadelta_f32.argtypes = [c_p, c_p, c_int, c_ll, c_ll, c_f] + seg + [c_f, c_f, c_int]
        L.mms_exchange_adadelta_f64.argtypes = [c_p, c_p, c_int, c_ll, c_ll, c_d] + seg + [c_d, c_d, c_int]
        L.mms_exchange_broadcast.argtypes = [c_p, c_p, c_int, c_int]
        L.mms_exchange_history.argtypes = [c_p, c_pp, c_pp]
        L.mms_exchange_check.argtypes = [c_p, c_p]
        L.mms_exchange_launch_count.argtypes = [c_p]
        L.mms_exchange_launch_count.restype = ctypes.c_ulonglong
        L.mms_simcross_backward_bottoms_f32.argtypes = [c_p] * 7 + [c_int] * 5
        L.mms_simcross_backward_params_f32.argtypes = [c_p] * 4 + [c_int] * 5
        L.mms_rerank_topk_f32.argtypes = [c_p] * 7 + [c_int, c_ll, c_int, c_int, c_int, c_ll]
        L.mms_rerank_topk_prepared_f32.argtypes = [c_p] * 7 + [c_int, c_ll, c_int, c_int, c_int, c_ll]
        L.mms_topk_merge_f32.argtypes = [c_p, c_p, c_p, c_ll, c_ll, c_p, c_p, c_int, c_int]
        L.mms_simcross_prepare_f32.argtypes = [c_p, c_p, c_int, c_int]
        L.mms_embed_plan_pair_f32.argtypes = [c_p, c_p, c_ll, c_p, c_ll, c_int]
        L.mms_embed_backward_pair_f32.argtypes = [c_p, c_p, c_p, c_ll, c_p, c_p, c_ll, c_p, c_p, c_int, c_int]
        L.mms_dropout_mask.argtypes = [c_p, c_p, c_ll, ctypes.c_ulonglong]
        L.mms_tc_gemm_f32.argtypes = [c_p, c_p, c_ll, c_int, c_p, c_ll, c_int, c_p, c_ll, c_int, c_int, c_int,
                                      c_int, c_int]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise MMSError(rc, lib().mms_last_error().decode(errors="replace"))


class Handle(object):
    """Owns one mms_handle_t (per layer, like the reference's per-layer state)."""

    def __init__(self):
        self._h = c_p()
        check(lib().mms_create(ctypes.byref(self._h)))

    @property
    def ptr(self):
        return self._h

    def set_stream(self, cuda_stream):
        check(lib().mms_set_stream(self._h, c_p(cuda_stream)))

    def set_option(self, opt, value):
        check(lib().mms_set_option(self._h, opt, int(value)))

    def reserve_scratch(self, nbytes):
        check(lib().mms_reserve_scratch(self._h, int(nbytes)))

    def launch_count(self):
        return int(lib().mms_launch_count(self._h))

    def profile_enable(self, on=True):
        check(lib().mms_profile_enable(self._h, int(on)))

    def profile_report(self):
        """{kernel name: (launches, total_ms)} since profiling was enabled / last report."""
        buf = ctypes.create_string_buffer(1 << 16)
        check(lib().mms_profile_report(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.rsplit(" ", 2)
            out[name] = (int(n), float(ms))
        return out

    def check_faults(self):
        check(lib().mms_check_faults(self._h))

    def close(self):
        if self._h:
            lib().mms_destroy(self._h)
            self._h = c_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
