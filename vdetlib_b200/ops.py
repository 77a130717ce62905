"""Tensor-level operators: torch CUDA tensors in, torch CUDA tensors out.

PyTorch is used for device memory and streams only; every operator is one or a few calls
into libvdet_b200.so (include/vdet_b200.h) on ``torch.cuda.current_stream()``.  There is no
CPU path: CPU tensors are rejected and a missing library raises RuntimeError.
"""
import ctypes

import numpy as np
import torch

from . import _lib

MISSING = -1e5    # utils/protocol.py:459, vdet/tubelet_cls.py:344,402


def default_device():
    """The CUDA device the adapters allocate on.  There is no CPU path: without a GPU this raises."""
    if not torch.cuda.is_available():
        raise RuntimeError("vdetlib_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _need(t, name, dtype=None, ndim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor (vdetlib_b200 has no CPU path)" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must be %d-D" % (name, ndim))
    return t


_ws_cache = {}


def _workspace(nbytes, device):
    """Per-device scratch buffer, grown on demand (the C ABI never allocates)."""
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def new_status(device):
    return torch.zeros(1, dtype=torch.int32, device=device)


def raise_for_status_word(word):
    """Translate a status word already on the host into the reference's Python exceptions."""
    s = int(word) & 0xffffffff
    if s & 0x80000000:
        raise RuntimeError("vdetlib_b200: internal frame-length mismatch")
    if s & _lib.STATUS_ZERO_DIVISION:
        raise ZeroDivisionError("float division")          # utils/nms.pyx:64
    if s & _lib.STATUS_ALL_MISSING:
        raise IndexError("list index out of range")        # vdet/tubelet_cls.py:295
    return s


def raise_for_status(status):
    """Same for the device status word (synchronises)."""
    return raise_for_status_word(status.item())


def host_copy_stream(dst_pinned, src, threads=0):
    """Fill a (pinned) host tensor from a C-contiguous NumPy array with non-temporal stores, over
    ``threads`` host threads (0: up to 8 for copies of 4 MB and more)."""
    if dst_pinned.device.type != "cpu" or not dst_pinned.is_contiguous():
        raise ValueError("host_copy_stream: destination must be a contiguous host tensor")
    nbytes = dst_pinned.numel() * dst_pinned.element_size()
    if not src.flags["C_CONTIGUOUS"] or src.nbytes != nbytes:
        raise ValueError("host_copy_stream: source must be C-contiguous with %d bytes" % nbytes)
    _lib.check(_lib.load().vdet_host_copy_stream_mt(dst_pinned.data_ptr(), src.ctypes.data, nbytes, int(threads)),
               "host_copy_stream")


def seg_offsets_uniform(n_frames, n_per_frame, device):
    return torch.arange(0, (n_frames + 1) * n_per_frame, n_per_frame, dtype=torch.int32, device=device)


# ------------------------------------------------------------------------------------------
# NMS
# ------------------------------------------------------------------------------------------
def nms_frames(boxes, scores, seg_offsets, thresh, max_seg_len, row_ids=None, want_mask=False,
               status=None, class_major=False, frame_major_out=False, out=None):
    """Greedy NMS of every (frame, class) on class-shared boxes (utils/nms.pyx:43-66 per problem).

    boxes [n,4] f32; scores [n,C] f32 (or [C,n] with class_major=True, or [n]);
    seg_offsets [S+1] i32.  Returns (keep_idx i32, keep_cnt i32, keep_mask|None, status):
    class-major outputs [C,n] / [C,S] by default; with frame_major_out=True flat [n*C] buffers
    whose (frame s, class c) block sits at seg_offsets[s]*C + c*n_s, and keep_cnt [S,C].
    ``out=(keep_idx, keep_cnt, keep_mask)`` writes into caller-provided buffers.
    """
    lib = _lib.load()
    _need(boxes, "boxes", torch.float32, 2)
    _need(scores, "scores", torch.float32)
    _need(seg_offsets, "seg_offsets", torch.int32, 1)
    boxes = boxes.contiguous()
    scores = scores.contiguous()
    n = boxes.shape[0]
    if scores.dim() == 1:
        C, ldr, ldc = 1, 1, 0
    elif class_major:
        C, ldr, ldc = scores.shape[0], 1, scores.shape[1]
    else:
        C, ldr, ldc = scores.shape[1], scores.shape[1], 1
    S = seg_offsets.numel() - 1
    dev = boxes.device
    if out is not None:
        keep_idx, keep_cnt, keep_mask = out
    elif frame_major_out:
        keep_idx = torch.empty(C * n, dtype=torch.int32, device=dev)
        keep_cnt = torch.empty((S, C), dtype=torch.int32, device=dev)
        keep_mask = torch.empty(C * n, dtype=torch.uint8, device=dev) if want_mask else None
    else:
        keep_idx = torch.empty((C, n), dtype=torch.int32, device=dev)
        keep_cnt = torch.empty((C, S), dtype=torch.int32, device=dev)
        keep_mask = torch.empty((C, n), dtype=torch.uint8, device=dev) if want_mask else None
    if status is None:
        status = new_status(dev)
    if row_ids is not None:
        _need(row_ids, "row_ids", torch.int32, 1)
    ws, ws_bytes = None, 0
    if max_seg_len > 1024:          # big frames keep their bit matrix in a global scratch slot per CTA
        ws_bytes = lib.vdet_nms_frames_workspace_bytes(int(max_seg_len), C, dev.index or 0)
        ws = _workspace(ws_bytes, dev)
    rc = lib.vdet_nms_frames_f32(_ptr(boxes), 4, _ptr(scores), ldr, ldc, _ptr(seg_offsets), S,
                                 int(max_seg_len), _ptr(row_ids), C, float(thresh),
                                 _ptr(keep_idx), _ptr(keep_cnt), _ptr(keep_mask), n,
                                 _lib.LAYOUT_FRAME_MAJOR if frame_major_out else _lib.LAYOUT_CLASS_MAJOR,
                                 _ptr(status), _ptr(ws), ws_bytes, _stream())
    _lib.check(rc, "nms_frames")
    return keep_idx, keep_cnt, keep_mask, status


def _nms_entry(fn_name, dets, ncol, thresh, tracks=None):
    lib = _lib.load()
    _need(dets, "dets", torch.float32, 2)
    if dets.shape[1] < ncol:
        raise IndexError("dets needs at least %d columns" % ncol)
    dets = dets.contiguous()
    n, ld = dets.shape
    dev = dets.device
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    ws = _workspace(lib.vdet_nms_workspace_bytes(n, dev.index or 0) + 1024, dev)
    st = ctypes.c_uint32(0)
    if tracks is None:
        cnt = getattr(lib, fn_name)(_ptr(dets), n, ld, float(thresh), _ptr(keep), ctypes.byref(st),
                                    _ptr(ws), ws.numel(), _stream())
    else:
        _need(tracks, "tracks", torch.float32, 2)
        if tracks.shape[1] < 5:
            raise IndexError("tracks needs at least 5 columns")
        tracks = tracks.contiguous()
        cnt = lib.vdet_track_det_nms_f32(_ptr(tracks), tracks.shape[0], tracks.shape[1], _ptr(dets), n, ld,
                                         float(thresh), _ptr(keep), ctypes.byref(st), _ptr(ws), ws.numel(),
                                         _stream())
    _lib.check(cnt, fn_name)
    if st.value & _lib.STATUS_ZERO_DIVISION:
        raise ZeroDivisionError("float division")
    return keep[:cnt]


def nms(dets, thresh):
    """utils.cython_nms.nms on a CUDA tensor [N,>=5]; returns int64 kept rows, descending score."""
    return _nms_entry("vdet_nms_f32", dets, 5, thresh)


def vid_nms(dets, thresh):
    """utils.cython_nms.vid_nms on a CUDA tensor [M,>=6] = (frame,x1,y1,x2,y2,score)."""
    return _nms_entry("vdet_vid_nms_f32", dets, 6, thresh)


def track_det_nms(tracks, dets, thresh):
    """utils.cython_nms.track_det_nms: tracks [Q,5], dets [K,6] CUDA tensors."""
    return _nms_entry("vdet_track_det_nms_f32", dets, 6, thresh, tracks=tracks)


def segment_by_frame(frames, row_valid=None, scores=None):
    """Stable grouping of rows by their float32 frame value (with ``scores``: rows of a frame in
    descending score, ties by ascending row).

    frames: 1-D float32 CUDA view (any stride).  Returns (row_ids i32 [n_packed], seg_offsets
    i32 [S+1], seg_frame f32 [S], max_seg_len).  Synchronises.
    """
    lib = _lib.load()
    _need(frames, "frames", torch.float32, 1)
    n = frames.numel()
    dev = frames.device
    ld = frames.stride(0) if n > 1 else 1
    row_ids = torch.empty(n + 1, dtype=torch.int32, device=dev)
    seg_off = torch.empty(n + 2, dtype=torch.int32, device=dev)
    seg_frame = torch.empty(n + 1, dtype=torch.float32, device=dev)
    ws = _workspace(lib.vdet_segment_workspace_bytes(n) + 1024, dev)
    n_segs, max_len, n_packed = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int64(0)
    if row_valid is not None:
        _need(row_valid, "row_valid", torch.uint8, 1)
    sld, sdt = 0, _lib.DTYPE_F32
    if scores is not None:
        _need(scores, "scores", None, 1)
        if scores.dtype not in (torch.float32, torch.float64):
            raise TypeError("segment_by_frame: float32 or float64 scores")
        sdt = _lib.DTYPE_F32 if scores.dtype == torch.float32 else _lib.DTYPE_F64
        sld = scores.stride(0) if n > 1 else 1
    rc = lib.vdet_segment_by_frame(_ptr(frames), ld, n, _ptr(row_valid), _ptr(scores), sld, sdt, _ptr(row_ids), _ptr(seg_off),
                                   _ptr(seg_frame), ctypes.byref(n_segs), ctypes.byref(max_len),
                                   ctypes.byref(n_packed), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "segment_by_frame")
    S = n_segs.value
    return row_ids[:n_packed.value], seg_off[:S + 1], seg_frame[:S], max_len.value


def track_nms_step(det_info, seg_offsets, row_ids, track_boxes, track_seg, thresh, keep, status):
    """One tracklet's suppression pass (vdet/track.py:172-183); updates ``keep`` in place."""
    lib = _lib.load()
    _need(det_info, "det_info", torch.float32, 2)
    _need(track_boxes, "track_boxes", torch.float32, 2)
    _need(track_seg, "track_seg", torch.int32, 1)
    _need(keep, "keep", torch.uint8, 1)
    rc = lib.vdet_track_nms_step_f32(_ptr(det_info), det_info.shape[0], _ptr(seg_offsets), _ptr(row_ids),
                                     seg_offsets.numel() - 1, _ptr(track_boxes), _ptr(track_seg),
                                     track_boxes.shape[0], float(thresh), _ptr(keep), _ptr(status), _stream())
    _lib.check(rc, "track_nms_step")


# ------------------------------------------------------------------------------------------
# IoU
# ------------------------------------------------------------------------------------------
def iou_matrix(a, b, out=None):
    """Dense IoU [A,B].  float64 inputs: utils/common.py:451-468 bit for bit; float32 inputs:
    the pair arithmetic of utils/nms.pyx:57-64."""
    lib = _lib.load()
    _need(a, "a", None, 2)
    _need(b, "b", a.dtype, 2)
    a = a.contiguous()
    b = b.contiguous()
    if out is None:
        out = torch.empty((a.shape[0], b.shape[0]), dtype=a.dtype, device=a.device)
    if a.dtype == torch.float32:
        rc = lib.vdet_iou_matrix_f32(_ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _stream())
    elif a.dtype == torch.float64:
        rc = lib.vdet_iou_matrix_f64(_ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _stream())
    else:
        raise TypeError("iou_matrix: float32 or float64 boxes")
    _lib.check(rc, "iou_matrix")
    return out


def iou_bitmask(boxes, thresh, status=None):
    lib = _lib.load()
    _need(boxes, "boxes", torch.float32, 2)
    boxes = boxes.contiguous()
    n = boxes.shape[0]
    mask = torch.zeros((n, (n + 31) // 32), dtype=torch.int32, device=boxes.device)
    if status is None:
        status = new_status(boxes.device)
    _lib.check(lib.vdet_iou_bitmask_f32(_ptr(boxes), n, float(thresh), _ptr(mask), _ptr(status), _stream()),
               "iou_bitmask")
    return mask, status


def link_frames(boxes, seg_offsets, max_seg_len, halo=None, halo_row_base=0, out=None, halo_count=None, ws=None):
    """Frame-to-frame link: for each box, the FIRST arg-max IoU box of the next frame.
    Returns (succ i32 [n] packed row of the successor / halo_row_base + halo index / -1,
    best_iou f32 [n]).  ``out=(succ, best_iou)`` writes into caller-provided buffers.
    ``halo_count``: optional int32 CUDA tensor [1] with the halo's box count (``halo`` is then a
    buffer of that capacity; ragged shards learn the count from the boundary exchange).
    A shard passes ``halo_row_base = n`` so that halo successors are >= n (see follow_links).
    ``ws``: uint8 CUDA workspace (``link_workspace_bytes``) for the x-sorted variant; by default one is taken from
    the per-stream cache -- callers that capture the call in a CUDA graph pass their own, fixed buffer."""
    lib = _lib.load()
    _need(boxes, "boxes", torch.float32, 2)
    _need(seg_offsets, "seg_offsets", torch.int32, 1)
    boxes = boxes.contiguous()
    n = boxes.shape[0]
    if out is not None:
        succ, best = out
    else:
        succ = torch.empty(n, dtype=torch.int32, device=boxes.device)
        best = torch.empty(n, dtype=torch.float32, device=boxes.device)
    n_halo = 0
    if halo is not None:
        _need(halo, "halo", torch.float32, 2)
        halo = halo.contiguous()
        n_halo = halo.shape[0]
    if halo_count is not None:
        _need(halo_count, "halo_count", torch.int32)
    S = seg_offsets.numel() - 1
    if ws is None:
        ws = _workspace(lib.vdet_link_workspace_bytes(n, S, n_halo), boxes.device)
    elif ws is False:                     # the full N x M scan (no x-sorting): what the sorted variant must equal
        ws = None
    rc = lib.vdet_link_frames_f32(_ptr(boxes), _ptr(seg_offsets), S, int(max_seg_len),
                                  _ptr(halo), n_halo, _ptr(halo_count), int(halo_row_base), _ptr(succ), _ptr(best),
                                  n, _ptr(ws), ws.numel() if ws is not None else 0, _stream())
    _lib.check(rc, "link_frames")
    return succ, best


def link_workspace_bytes(n_rows, n_segs, n_halo=0):
    return int(_lib.load().vdet_link_workspace_bytes(int(n_rows), int(n_segs), int(n_halo)))


def spatial_maxpool(tub_boxes, tub_seg, det_boxes, det_scores, det_seg_offsets, thresh=0.7,
                    mode=_lib.POOL_ARGMAX_SCORE):
    """vdet/tubelet_cls.py:330-347 for all tubelet boxes at once.  det_scores: 1-D (strided) view of
    the class column.  Returns (arg i32 [P] packed det row or -1, score f64 [P])."""
    lib = _lib.load()
    _need(tub_boxes, "tub_boxes", None, 2)
    _need(det_boxes, "det_boxes", tub_boxes.dtype, 2)
    _need(det_scores, "det_scores", None, 1)
    dt = {torch.float32: _lib.DTYPE_F32, torch.float64: _lib.DTYPE_F64}
    if tub_boxes.dtype not in dt or det_scores.dtype not in dt:
        raise TypeError("spatial_maxpool: float32/float64 only")
    tub_boxes = tub_boxes.contiguous()
    det_boxes = det_boxes.contiguous()
    P = tub_boxes.shape[0]
    arg = torch.empty(P, dtype=torch.int32, device=tub_boxes.device)
    score = torch.empty(P, dtype=torch.float64, device=tub_boxes.device)
    ld = det_scores.stride(0) if det_scores.numel() > 1 else 1
    rc = lib.vdet_spatial_maxpool(_ptr(tub_boxes), _ptr(tub_seg), P, _ptr(det_boxes), dt[tub_boxes.dtype],
                                  _ptr(det_scores), ld, dt[det_scores.dtype], _ptr(det_seg_offsets),
                                  det_seg_offsets.numel() - 1, float(thresh), int(mode), _ptr(arg),
                                  _ptr(score), _stream())
    _lib.check(rc, "spatial_maxpool")
    return arg, score


# ------------------------------------------------------------------------------------------
# temporal
# ------------------------------------------------------------------------------------------
def _rows(x, name):
    _need(x, name, None, 2)
    if x.dtype not in (torch.float32, torch.float64):
        raise TypeError("%s: float32 or float64 rows" % name)
    if x.stride(1) != 1:
        raise ValueError("%s: rows must be contiguous along the frame axis" % name)
    return _lib.DTYPE_F32 if x.dtype == torch.float32 else _lib.DTYPE_F64


def score_completion_(scores, lengths=None, miss_thr=-10.0, status=None, bounds=None):
    """do_score_completion (vdet/tubelet_cls.py:284-303) on [rows, L] score rows, IN PLACE.
    ``bounds`` [rows, 4] (same dtype): for rows that are one frame range of longer tubelets, the nearest valid score
    of the neighbouring shards as (left gap, left value, right gap, right value), gap < 0 = none (see dist.py)."""
    lib = _lib.load()
    dt = _rows(scores, "scores")
    if status is None:
        status = new_status(scores.device)
    ws_bytes = lib.vdet_score_completion_workspace_bytes(scores.shape[0], scores.shape[1], dt)
    ws = _workspace(ws_bytes, scores.device)
    if bounds is not None:
        _need(bounds, "bounds", scores.dtype, 2)
        if tuple(bounds.shape) != (scores.shape[0], 4):
            raise ValueError("score_completion_: bounds must be [rows, 4]")
        bounds = bounds.contiguous()
        rc = lib.vdet_score_completion_bounded(_ptr(scores), dt, scores.shape[0], scores.shape[1], scores.stride(0),
                                               _ptr(lengths), float(miss_thr), _ptr(bounds), _ptr(status), _ptr(ws),
                                               ws_bytes, _stream())
        _lib.check(rc, "score_completion")
        return status
    rc = lib.vdet_score_completion(_ptr(scores), dt, scores.shape[0], scores.shape[1], scores.stride(0),
                                   _ptr(lengths), float(miss_thr), _ptr(status), _ptr(ws), ws_bytes, _stream())
    _lib.check(rc, "score_completion")
    return status


def temporal_maxpool(scores, window, lengths=None, pad=MISSING, out=None):
    """score_proto_temporal_maxpool (vdet/tubelet_cls.py:399-409) on [rows, L] score rows."""
    lib = _lib.load()
    dt = _rows(scores, "scores")
    if window % 2 != 1:
        raise ValueError('Window size must be odd!')
    if out is None:
        out = torch.empty_like(scores)
    if out.stride(0) != scores.stride(0):
        raise ValueError("temporal_maxpool: out must have the input's row pitch")
    rc = lib.vdet_temporal_maxpool(_ptr(scores), _ptr(out), dt, scores.shape[0], scores.shape[1],
                                   scores.stride(0), _ptr(lengths), int(window), float(pad), _stream())
    _lib.check(rc, "temporal_maxpool")
    return out


def temporal_conv1d(x, taps, pad_mode="zero", lengths=None, out=None):
    """Depthwise temporal convolution: row r uses taps[r % n_channels]."""
    lib = _lib.load()
    dt = _rows(x, "x")
    _need(taps, "taps", x.dtype, 2)
    taps = taps.contiguous()
    if out is None:
        out = torch.empty_like(x)
    if out.stride(0) != x.stride(0):
        raise ValueError("temporal_conv1d: out must have the input's row pitch")
    mode = {"zero": _lib.PAD_ZERO, "edge": _lib.PAD_EDGE}[pad_mode]
    rc = lib.vdet_temporal_conv1d(_ptr(x), _ptr(out), dt, x.shape[0], x.shape[1], x.stride(0), _ptr(lengths),
                                  _ptr(taps), taps.shape[0], taps.shape[1], mode, _stream())
    _lib.check(rc, "temporal_conv1d")
    return out


def threshold_topk(scores, seg_offsets, max_seg_len, thresh=0.05, k=100):
    """vdet/video_det.py:88-100 for every (frame, class).  scores [n, C] f32.
    Returns (idx i32 [S, C, k] local row or -1, cnt i32 [S, C])."""
    lib = _lib.load()
    _need(scores, "scores", torch.float32, 2)
    scores = scores.contiguous()
    S = seg_offsets.numel() - 1
    C = scores.shape[1]
    idx = torch.empty((S, C, k), dtype=torch.int32, device=scores.device)
    cnt = torch.empty((S, C), dtype=torch.int32, device=scores.device)
    rc = lib.vdet_threshold_topk_f32(_ptr(scores), _ptr(seg_offsets), S, int(max_seg_len), C, float(thresh),
                                     int(k), _ptr(idx), _ptr(cnt), _stream())
    _lib.check(rc, "threshold_topk")
    return idx, cnt


def tubelet_interpolate(knot_x, knot_y, knot_off, dense_first, dense_off):
    """score_proto_interpolation's arithmetic for all tubelets at once (vdet/tubelet_cls.py:430-490).
    knot_x f64 [n], knot_y f64 [F, n], knot_off / dense_off i32 [K+1], dense_first i32 [K].
    Returns out f64 [F, n_dense]."""
    lib = _lib.load()
    _need(knot_x, "knot_x", torch.float64, 1)
    _need(knot_y, "knot_y", torch.float64, 2)
    knot_x, knot_y = knot_x.contiguous(), knot_y.contiguous()
    K = knot_off.numel() - 1
    n_dense = int(dense_off[-1].item()) if K > 0 else 0
    counts = (dense_off[1:] - dense_off[:-1]).long()
    dense_tub = torch.repeat_interleave(torch.arange(K, dtype=torch.int32, device=knot_x.device), counts)
    out = torch.empty((knot_y.shape[0], n_dense), dtype=torch.float64, device=knot_x.device)
    rc = lib.vdet_tubelet_interpolate_f64(_ptr(knot_x), _ptr(knot_y), knot_x.numel(), _ptr(knot_off), _ptr(dense_off),
                                          _ptr(dense_first), _ptr(dense_tub), K, knot_y.shape[0], n_dense, _ptr(out),
                                          _stream())
    _lib.check(rc, "tubelet_interpolate")
    return out


def follow_links(succ, link_iou, start_rows, n_frames, min_iou=0.0):
    """Follow the frame-to-frame links from ``start_rows`` for ``n_frames`` frames.
    Returns chain_rows i32 [n_frames, K] (packed row per frame, -1 after a chain ended).  A successor
    >= succ.numel() is a halo index (the chain leaves this shard) and ends the chain here."""
    lib = _lib.load()
    _need(succ, "succ", torch.int32, 1)
    _need(link_iou, "link_iou", torch.float32, 1)
    _need(start_rows, "start_rows", torch.int32, 1)
    K = start_rows.numel()
    out = torch.empty((int(n_frames), K), dtype=torch.int32, device=succ.device)
    _lib.check(lib.vdet_follow_links(_ptr(succ), _ptr(link_iou), succ.numel(), _ptr(start_rows), K, int(n_frames), float(min_iou),
                                     _ptr(out), _stream()), "follow_links")
    return out


def gather_chain_scores(scores, chain_rows, missing=MISSING):
    """Class scores along every chain: [K, C, n_frames] f32 rows for the temporal kernels."""
    lib = _lib.load()
    _need(scores, "scores", torch.float32, 2)
    _need(chain_rows, "chain_rows", torch.int32, 2)
    scores = scores.contiguous()
    T, K = chain_rows.shape
    C = scores.shape[1]
    out = torch.empty((K, C, T), dtype=torch.float32, device=scores.device)
    _lib.check(lib.vdet_gather_chain_scores_f32(_ptr(scores), C, _ptr(chain_rows), K, T, float(missing), _ptr(out),
                                                _stream()), "gather_chain_scores")
    return out


def sort_by_score_desc(scores, ids):
    """Stable sort of (score f32|f64, id i64) pairs by descending score (ties keep input order)."""
    lib = _lib.load()
    _need(scores, "scores", None, 1)
    if scores.dtype not in (torch.float32, torch.float64):
        raise TypeError("sort_by_score_desc: float32 or float64 scores")
    dt = _lib.DTYPE_F32 if scores.dtype == torch.float32 else _lib.DTYPE_F64
    _need(ids, "ids", torch.int64, 1)
    scores, ids = scores.contiguous(), ids.contiguous()
    n = scores.numel()
    so, io = torch.empty_like(scores), torch.empty_like(ids)
    ws = _workspace(lib.vdet_sort_workspace_bytes(n) + 1024, scores.device)
    _lib.check(lib.vdet_sort_by_score_desc(_ptr(scores), dt, _ptr(ids), n, _ptr(so), _ptr(io), _ptr(ws), ws.numel(),
                                           _stream()), "sort_by_score_desc")
    return so, io


def to_device(array, dtype=None, device=None):
    """numpy / array-like -> CUDA tensor on the current device."""
    a = np.ascontiguousarray(array, dtype=dtype)
    return torch.from_numpy(a).to(device or default_device())
