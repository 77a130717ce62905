"""ctypes binding of libvdet_b200.so (include/vdet_b200.h).

There is NO CPU fallback: if the shared library is missing or cannot be loaded, every
operator raises ``RuntimeError`` (build it with ``python -m vdetlib_b200.build``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VDET_B200_LIB: measurement hook -- load a kernel variant built by tools/build_variant.py instead
LIB_PATH = os.environ.get("VDET_B200_LIB") or os.path.join(_HERE, "libvdet_b200.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_UNSUPPORTED = -1, -2, -3, -4
STATUS_ZERO_DIVISION = 1
STATUS_ALL_MISSING = 2
DTYPE_F32, DTYPE_F64 = 0, 1
POOL_ARGMAX_SCORE, POOL_ARGMAX_IOU, POOL_MAX_IOU = 0, 1, 2
PAD_ZERO, PAD_EDGE = 0, 1
LAYOUT_CLASS_MAJOR, LAYOUT_FRAME_MAJOR = 0, 1
KEEP_U16_LOCAL, KEEP_I32_ROW = 0, 1

_c = ctypes
_vp, _i32, _i64, _f64, _f32, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_double, _c.c_float, _c.c_size_t

# name -> (restype, argtypes): exactly the entry points include/vdet_b200.h declares
SIGNATURES = {
    "vdet_abi_version": (_i32, []),
    "vdet_last_error": (_c.c_char_p, []),
    "vdet_sm_count": (_i32, [_i32]),
    "vdet_set_reserved_sms": (_i32, [_i32]),
    "vdet_host_copy_stream": (_i32, [_vp, _vp, _sz]),
    "vdet_host_copy_stream_mt": (_i32, [_vp, _vp, _sz, _i32]),
    "vdet_nms_frames_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "vdet_nms_frames_f32": (_i32, [_vp, _i32, _vp, _i64, _i64, _vp, _i32, _i32, _vp, _i32, _f64,
                                   _vp, _vp, _vp, _i64, _i32, _vp, _vp, _sz, _vp]),
    "vdet_compact_keep": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "vdet_nms_workspace_bytes": (_sz, [_i64, _i32]),
    "vdet_nms_f32": (_i64, [_vp, _i64, _i32, _f64, _vp, _vp, _vp, _sz, _vp]),
    "vdet_vid_nms_f32": (_i64, [_vp, _i64, _i32, _f64, _vp, _vp, _vp, _sz, _vp]),
    "vdet_track_det_nms_f32": (_i64, [_vp, _i64, _i32, _vp, _i64, _i32, _f64, _vp, _vp, _vp, _sz, _vp]),
    "vdet_track_nms_step_f32": (_i32, [_vp, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _f64, _vp, _vp, _vp]),
    "vdet_segment_workspace_bytes": (_sz, [_i64]),
    "vdet_segment_by_frame": (_i32, [_vp, _i32, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vdet_iou_matrix_f32": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "vdet_iou_matrix_f64": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "vdet_iou_bitmask_f32": (_i32, [_vp, _i32, _f64, _vp, _vp, _vp]),
    "vdet_link_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "vdet_link_frames_f32": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _i64, _vp, _sz, _vp]),
    "vdet_spatial_maxpool": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _i64, _i32, _vp, _i32, _f64, _i32,
                                    _vp, _vp, _vp]),
    "vdet_score_completion_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "vdet_score_completion": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _f64, _vp, _vp, _sz, _vp]),
    "vdet_score_completion_bounded": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _f64, _vp, _vp, _vp, _sz, _vp]),
    "vdet_temporal_maxpool": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _vp, _i32, _f64, _vp]),
    "vdet_temporal_conv1d": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _i32, _i32, _i32, _vp]),
    "vdet_tubelet_interpolate_f64": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _vp, _vp]),
    "vdet_follow_links": (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _f32, _vp, _vp]),
    "vdet_gather_chain_scores_f32": (_i32, [_vp, _i32, _vp, _i32, _i32, _f32, _vp, _vp]),
    "vdet_sort_workspace_bytes": (_sz, [_i64]),
    "vdet_sort_by_score_desc": (_i32, [_vp, _i32, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "vdet_threshold_topk_f32": (_i32, [_vp, _vp, _i32, _i32, _i32, _f32, _i32, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load the library once; raise RuntimeError (never fall back) when it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "vdetlib_b200: %s is missing -- build it with `python -m vdetlib_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise RuntimeError("vdetlib_b200: cannot load %s: %s" % (LIB_PATH, e))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library diverge
        fn.restype = res
        fn.argtypes = args
    if lib.vdet_abi_version() != 2:
        raise RuntimeError("vdetlib_b200: ABI version mismatch")
    _lib = lib
    return lib


def last_error():
    return load().vdet_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    """Map a negative return code to the Python exception the adapters document."""
    if rc >= 0:
        return rc
    msg = "%s: %s" % (what, last_error()) if what else last_error()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
