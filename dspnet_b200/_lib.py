"""ctypes binding of libdspmb.so (the C ABI declared in include/dspmb.h).

There is no fallback of any kind: if the shared library has not been built (``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C dspnet_b200/csrc``) importing the operators raises, and every
compute entry returns DSPMB_ERR_CUDA when no CUDA device is present.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DSPMB_LIB: another build of the same library (A/B of compile-time settings, scripts/ab_build.sh); never a CPU path
LIB_PATH = os.environ.get("DSPMB_LIB") or os.path.join(_HERE, "libdspmb.so")

OK = 0
ERR_BAD_ARG = -1
ERR_LABEL_PADDING = -2
ERR_MINING_CANDIDATES = -3
ERR_MINING_THRESH = -4
ERR_WORKSPACE = -5
ERR_CUDA = -6
(TUNE_DET_STREAM_VARIANT, TUNE_NMS_MASK_ROWS, TUNE_NMS_SMEM_ROWS, TUNE_SORT_SMEM_KEYS, TUNE_PHASES, TUNE_GRAPH_CACHE,
 TUNE_DET_PIPELINE, TUNE_TARGET_PIPELINE, TUNE_NMS_PIPELINE, TUNE_DET_PREFETCH, TUNE_TARGET_PREFETCH,
 TUNE_DET_SPLIT, TUNE_DET_LEAN, TUNE_TARGET_SMALL, TUNE_TARGET_SHORTLIST,
 TUNE_TARGET_PDL, TUNE_DET_SORT_PDL, TUNE_NMS_PDL) = range(18)
PHASES_ALL = 31

# every symbol include/dspmb.h declares (tests check that the built library exports all of them)
EXPORTS = (
    "dspmb_version", "dspmb_last_error", "dspmb_set_libm_mode", "dspmb_prior_f32", "dspmb_prior_multi_f32",
    "dspmb_target_workspace_bytes", "dspmb_target_f32", "dspmb_detection_workspace_bytes", "dspmb_detection_f32",
    "dspmb_status", "dspmb_nms_workspace_bytes", "dspmb_nms_f32", "dspmb_nms_host", "dspmb_test_expf",
    "dspmb_test_logf", "dspmb_profile_enable", "dspmb_profile_read", "dspmb_profile_kernel_name", "dspmb_detection_compact_f32", "dspmb_set_tuning", "dspmb_gather_buffer_bytes", "dspmb_p2p_alloc",
    "dspmb_p2p_open", "dspmb_p2p_close", "dspmb_p2p_free", "dspmb_detection_gather_f32", "dspmb_detection_gather_wait", "dspmb_detection_gather_read",
    "dspmb_detection_gather_ack", "dspmb_gather_error", "dspmb_gather_ctx_create", "dspmb_gather_ctx_side_stream",
    "dspmb_gather_ctx_destroy", "dspmb_gather_submit",
    "dspmb_bbox_overlaps_f64", "dspmb_detection_postfilter_f32", "dspmb_map_match_f32", "dspmb_last_launch_count",
    "dspmb_debug_trace", "dspmb_detection_heads_workspace_bytes", "dspmb_detection_heads_f32",
    "dspmb_multibox_loss_workspace_bytes", "dspmb_multibox_loss_f32",
)


class DspmbError(RuntimeError):
    """Raised where the reference would abort through CHECK_* / dmlc::Error (-> MXNetError in Python)."""

    def __init__(self, code, message):
        super().__init__("dspmb error %d: %s" % (code, message))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "dspnet_b200: %s is missing -- build the CUDA extension first (make -C dspnet_b200/csrc); "
            "there is no CPU fallback" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    c_int, c_float, c_double, c_void_p, c_size_t, c_long = (ctypes.c_int, ctypes.c_float, ctypes.c_double,
                                                              ctypes.c_void_p, ctypes.c_size_t, ctypes.c_long)
    fp = ctypes.POINTER(c_float)
    ip = ctypes.POINTER(c_int)
    L.dspmb_version.restype = c_int
    L.dspmb_last_launch_count.restype = c_int
    L.dspmb_last_error.restype = ctypes.c_char_p
    L.dspmb_set_libm_mode.argtypes = [c_int]
    L.dspmb_set_libm_mode.restype = c_int
    L.dspmb_prior_f32.argtypes = [c_void_p, c_int, c_int, fp, c_int, fp, c_int, c_float, c_float, c_float, c_float,
                                  c_int, c_void_p]
    L.dspmb_prior_multi_f32.argtypes = [c_void_p, c_int, ip, ip, fp, ip, fp, ip, fp, fp, c_int, c_void_p]
    L.dspmb_target_workspace_bytes.argtypes = [c_int] * 4
    L.dspmb_target_workspace_bytes.restype = c_size_t
    L.dspmb_target_f32.argtypes = [c_void_p] * 6 + [c_int] * 5 + [c_float] * 4 + [c_int, fp, c_void_p, c_void_p,
                                                                               c_void_p, c_size_t, c_void_p]
    L.dspmb_detection_workspace_bytes.argtypes = [c_int] * 3
    L.dspmb_detection_workspace_bytes.restype = c_size_t
    L.dspmb_detection_f32.argtypes = [c_void_p] * 4 + [c_int] * 3 + [c_float, c_int, fp, c_float, c_int, c_int,
                                                                     c_void_p, c_void_p, c_size_t, c_void_p]
    L.dspmb_detection_heads_workspace_bytes.argtypes = [c_int, c_int, c_int, ip, ip, c_int]
    L.dspmb_detection_heads_workspace_bytes.restype = c_size_t
    L.dspmb_detection_heads_f32.argtypes = [ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ip, ip, c_int, c_void_p,
                                            c_void_p, c_int, c_int, c_int, c_float, c_int, fp, c_float, c_int, c_int,
                                            c_void_p, c_void_p, c_size_t, c_void_p]
    L.dspmb_multibox_loss_workspace_bytes.argtypes = [c_int, c_int]
    L.dspmb_multibox_loss_workspace_bytes.restype = c_size_t
    L.dspmb_multibox_loss_f32.argtypes = [c_void_p] * 8 + [c_int, c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]
    L.dspmb_detection_compact_f32.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    L.dspmb_set_tuning.argtypes = [c_int, c_int]
    c_ll = ctypes.c_longlong
    L.dspmb_gather_buffer_bytes.argtypes = [c_int, c_int, c_int, c_int]
    L.dspmb_gather_buffer_bytes.restype = c_size_t
    L.dspmb_p2p_alloc.argtypes = [c_size_t, ctypes.POINTER(c_void_p), ctypes.c_char_p]
    L.dspmb_p2p_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(c_void_p)]
    L.dspmb_p2p_close.argtypes = [c_void_p]
    L.dspmb_p2p_free.argtypes = [c_void_p]
    L.dspmb_detection_gather_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                             ctypes.POINTER(c_void_p), c_int, c_ll, c_void_p]
    L.dspmb_detection_gather_wait.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_ll, c_void_p]
    L.dspmb_detection_gather_ack.argtypes = [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_void_p), c_int, c_ll, c_void_p]
    L.dspmb_gather_ctx_create.restype = c_void_p
    L.dspmb_gather_ctx_side_stream.argtypes = [c_void_p]
    L.dspmb_gather_ctx_side_stream.restype = c_void_p
    L.dspmb_gather_ctx_destroy.argtypes = [c_void_p]
    L.dspmb_gather_submit.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      ctypes.POINTER(c_void_p), c_int, c_ll, c_ll, c_int, c_void_p]
    L.dspmb_detection_gather_read.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.dspmb_debug_trace.argtypes = [c_void_p]
    L.dspmb_debug_stamps.argtypes = [c_void_p]
    L.dspmb_gather_error.argtypes = [c_void_p, c_int, c_int, c_int, c_int]
    L.dspmb_status.argtypes = [c_void_p, c_void_p]
    L.dspmb_nms_workspace_bytes.argtypes = [c_int]
    L.dspmb_nms_workspace_bytes.restype = c_size_t
    L.dspmb_nms_f32.argtypes = [c_void_p, c_int, c_int, c_double, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_size_t, c_void_p]
    L.dspmb_nms_host.argtypes = [ip, ip, fp, c_int, c_int, c_float, c_int]
    L.dspmb_bbox_overlaps_f64.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]
    L.dspmb_detection_postfilter_f32.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]
    L.dspmb_map_match_f32.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p]
    L.dspmb_profile_enable.argtypes = [c_int]
    L.dspmb_profile_read.argtypes = [fp, ip, c_int]
    L.dspmb_profile_kernel_name.argtypes = [c_int]
    L.dspmb_profile_kernel_name.restype = ctypes.c_char_p
    L.dspmb_test_expf.argtypes = [c_void_p, c_void_p, c_long, c_void_p]
    L.dspmb_test_logf.argtypes = [c_void_p, c_void_p, c_long, c_void_p]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is c_int and name not in ("dspmb_version", "dspmb_set_libm_mode"):
            fn.restype = c_int
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise DspmbError(rc, lib().dspmb_last_error().decode())


def float_array(values):
    values = [float(v) for v in values]
    return (ctypes.c_float * len(values))(*values)


def int_array(values):
    values = [int(v) for v in values]
    return (ctypes.c_int * len(values))(*values)
