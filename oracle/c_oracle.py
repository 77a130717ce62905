"""ctypes front-end of oracle/liboracle_nms.so (the C restatement of utils/nms.pyx).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from vdetlib_b200/.

The functions mirror the reference signatures (utils/nms.pyx:17,71,128): float32
2-D arrays in, Python ``list`` of ``int`` out, ``ZeroDivisionError('float division')``
when a visited pair has union == 0.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_nms.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "nms_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, f64, vp = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
        L.oracle_nms.restype = i64
        L.oracle_nms.argtypes = [vp, i64, f64, vp]
        L.oracle_vid_nms.restype = i64
        L.oracle_vid_nms.argtypes = [vp, i64, f64, vp]
        L.oracle_track_det_nms.restype = i64
        L.oracle_track_det_nms.argtypes = [vp, i64, vp, i64, f64, vp]
        L.oracle_pair_iou_f32.restype = None
        L.oracle_pair_iou_f32.argtypes = [vp, i64, vp, i64, vp]
        L.oracle_iou_bitmask.restype = None
        L.oracle_iou_bitmask.argtypes = [vp, i64, f64, vp]
        L.oracle_link_f32.restype = None
        L.oracle_link_f32.argtypes = [vp, vp, i64, i64, vp, vp]
        L.oracle_nms_frames.restype = i64
        L.oracle_nms_frames.argtypes = [vp, vp, vp, i64, i64, i64, f64, vp, vp, vp]
        _lib = L
    return _lib


def _f32_2d(a, ncol, name):
    a = np.asarray(a)
    if a.dtype != np.float32:
        # utils/nms.pyx:17 typed buffer: wrong dtype raises ValueError
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' but got '%s'" % a.dtype)
    if a.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % a.ndim)
    if a.shape[1] < ncol:
        raise IndexError("%s needs at least %d columns" % (name, ncol))
    return np.ascontiguousarray(a[:, :ncol])


def _finish(n, keep):
    if n < 0:
        raise ZeroDivisionError("float division")
    return [int(i) for i in keep[:n]]


def nms(dets, thresh):
    d = _f32_2d(dets, 5, "dets")
    keep = np.empty(max(len(d), 1), np.int64)
    n = lib().oracle_nms(d.ctypes.data, len(d), float(thresh), keep.ctypes.data)
    return _finish(n, keep)


def vid_nms(dets, thresh):
    d = _f32_2d(dets, 6, "dets")
    keep = np.empty(max(len(d), 1), np.int64)
    n = lib().oracle_vid_nms(d.ctypes.data, len(d), float(thresh), keep.ctypes.data)
    return _finish(n, keep)


def track_det_nms(tracks, dets, thresh):
    t = _f32_2d(tracks, 5, "tracks")
    d = _f32_2d(dets, 6, "dets")
    keep = np.empty(max(len(d), 1), np.int64)
    n = lib().oracle_track_det_nms(t.ctypes.data, len(t), d.ctypes.data, len(d),
                                   float(thresh), keep.ctypes.data)
    return _finish(n, keep)


def pair_iou_f32(a, b):
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 4)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 4)
    out = np.empty((len(a), len(b)), np.float32)
    lib().oracle_pair_iou_f32(a.ctypes.data, len(a), b.ctypes.data, len(b), out.ctypes.data)
    return out


def iou_bitmask(boxes, thresh):
    b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4)
    n = len(b)
    out = np.zeros((n, (n + 31) // 32), np.uint32)
    lib().oracle_iou_bitmask(b.ctypes.data, n, float(thresh), out.ctypes.data)
    return out


def link_f32(boxes, counts=None):
    b = np.ascontiguousarray(boxes, np.float32)
    T, nmax = b.shape[0], b.shape[1]
    if counts is None:
        counts = np.full(T, nmax, np.int32)
    counts = np.ascontiguousarray(counts, np.int32)
    succ = np.empty((max(T - 1, 0), nmax), np.int32)
    best = np.empty((max(T - 1, 0), nmax), np.float32)
    lib().oracle_link_f32(b.ctypes.data, counts.ctypes.data, T, nmax,
                          succ.ctypes.data, best.ctypes.data)
    return succ, best


def nms_frames(boxes, scores, thresh, counts=None):
    """Per-(frame, class) NMS on class-shared boxes. boxes [T,N,4], scores [T,N,C]."""
    b = np.ascontiguousarray(boxes, np.float32)
    s = np.ascontiguousarray(scores, np.float32)
    T, nmax = b.shape[0], b.shape[1]
    C = s.shape[2]
    if counts is None:
        counts = np.full(T, nmax, np.int32)
    counts = np.ascontiguousarray(counts, np.int32)
    km = np.zeros((T, C, nmax), np.uint8)
    ki = np.full((T, C, nmax), -1, np.int32)
    kc = np.zeros((T, C), np.int32)
    rc = lib().oracle_nms_frames(b.ctypes.data, s.ctypes.data, counts.ctypes.data, T, nmax, C,
                                 float(thresh), km.ctypes.data, ki.ctypes.data, kc.ctypes.data)
    if rc < 0:
        raise ZeroDivisionError("float division")
    return km, ki, kc
