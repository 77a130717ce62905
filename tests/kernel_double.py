"""A CPU stand-in for the kernels, for tests only: every ``vdetlib_b200.ops`` operator the reference-named
adapters call, restated on CPU tensors with the oracle (oracle/nms_oracle.c, oracle/oracle_np.py).

With ``install(monkeypatch)`` the adapters of vdetlib_b200.utils / vdetlib_b200.vdet run end to end
without a GPU: ``ops.default_device()`` hands out the CPU device and each operator below takes the place of
its CUDA counterpart, following the contract documented in include/vdet_b200.h and vdetlib_b200/ops.py.
What such a run checks is the adapters' HOST logic -- proto walking, packing, index bookkeeping, in-place
rules, exceptions -- against the golden protos the reference produced.  It says nothing about the kernels:
those are checked by the ``-m gpu`` tests, where the same assertions run on the real library.  The product
never imports this module (nor oracle/): without a GPU ``ops.default_device()`` raises.
"""
import numpy as np
import torch

from oracle import c_oracle, oracle_np

MISSING = -1e5
POOL_ARGMAX_SCORE, POOL_ARGMAX_IOU, POOL_MAX_IOU = 0, 1, 2


def _i64(a):
    return torch.tensor(list(a), dtype=torch.int64)


def nms(dets, thresh):
    return _i64(c_oracle.nms(dets.numpy(), thresh))


def vid_nms(dets, thresh):
    return _i64(c_oracle.vid_nms(dets.numpy(), thresh))


def track_det_nms(tracks, dets, thresh):
    return _i64(c_oracle.track_det_nms(tracks.numpy(), dets.numpy(), thresh))


def iou_matrix(a, b, out=None):
    if a.dtype == torch.float64:
        if a.shape[0] == 0 or b.shape[0] == 0:
            res = np.zeros((a.shape[0], b.shape[0]))
        else:
            res = oracle_np.iou(a.numpy(), b.numpy())
        return torch.from_numpy(np.ascontiguousarray(res, dtype=np.float64))
    return torch.from_numpy(c_oracle.pair_iou_f32(a.numpy(), b.numpy()))


def nms_frames(boxes, scores, seg_offsets, thresh, max_seg_len, row_ids=None, want_mask=False,
               status=None, class_major=False, frame_major_out=False, out=None):
    """Class-major outputs only (what packed_vid_nms asks for)."""
    assert row_ids is None and not class_major and not frame_major_out and out is None
    b, s, off = boxes.numpy(), scores.numpy(), seg_offsets.numpy()
    if s.ndim == 1:
        s = s[:, None]
    n, C = s.shape
    S = len(off) - 1
    keep_idx = np.full((C, n), -1, np.int32)
    keep_cnt = np.zeros((C, S), np.int32)
    keep_mask = np.zeros((C, n), np.uint8)
    st = 0
    for f in range(S):
        a, e = off[f], off[f + 1]
        for c in range(C):
            d = np.concatenate([b[a:e], s[a:e, c:c + 1]], axis=1).astype(np.float32)
            try:
                k = np.asarray(c_oracle.nms(d, thresh), dtype=np.int64)
            except ZeroDivisionError:
                st |= 1
                continue
            keep_idx[c, a:a + len(k)] = a + k
            keep_cnt[c, f] = len(k)
            keep_mask[c, a + k] = 1
    status = torch.tensor([st], dtype=torch.int32)
    return (torch.from_numpy(keep_idx), torch.from_numpy(keep_cnt),
            torch.from_numpy(keep_mask) if want_mask else None, status)


def segment_by_frame(frames, row_valid=None, scores=None):
    f = frames.numpy().astype(np.float32) + np.float32(0.0)                 # -0.0 == +0.0
    rows = np.arange(len(f))
    if row_valid is not None:
        rows = rows[row_valid.numpy() != 0]
    if scores is not None:
        sc = scores.numpy()
        rows = rows[np.argsort(-sc[rows], kind="stable")]                    # descending score, ties: ascending row
    order = rows[np.argsort(f[rows], kind="stable")]
    seg_frame, counts = np.unique(f[order], return_counts=True)
    off = np.zeros(len(seg_frame) + 1, np.int32)
    np.cumsum(counts, out=off[1:])
    return (torch.from_numpy(order.astype(np.int32)), torch.from_numpy(off), torch.from_numpy(seg_frame.astype(np.float32)),
            int(counts.max()) if len(counts) else 0)


def track_nms_step(det_info, seg_offsets, row_ids, track_boxes, track_seg, thresh, keep, status):
    """vdet/track.py:172-183 per tracked box: track_det_nms of the box against the surviving detections of
    its frame (in det_info order = descending score); everything it does not return is dropped."""
    di, off, rid, k = det_info.numpy(), seg_offsets.numpy(), row_ids.numpy(), keep.numpy()
    for box, s in zip(track_boxes.numpy(), track_seg.numpy()):
        if s < 0:
            continue
        ids = rid[off[s]:off[s + 1]]
        ids = ids[k[ids] != 0]
        if not len(ids):
            continue
        t = np.concatenate([[di[ids[0], 0]], box]).astype(np.float32)[None, :]
        try:
            kept = c_oracle.track_det_nms(t, di[ids], thresh)
        except ZeroDivisionError:
            status |= 1
            continue
        k[ids] = 0
        k[ids[np.asarray(kept, dtype=np.int64)]] = 1


def spatial_maxpool(tub_boxes, tub_seg, det_boxes, det_scores, det_seg_offsets, thresh=0.7, mode=POOL_ARGMAX_SCORE):
    tb, seg, db, sc, off = (tub_boxes.numpy(), tub_seg.numpy(), det_boxes.numpy(), det_scores.numpy(),
                            det_seg_offsets.numpy())
    P = len(tb)
    arg = np.full(P, -1, np.int32)
    score = np.full(P, MISSING, np.float64)
    for p in range(P):
        s = seg[p]
        if s < 0 or s >= len(off) - 1 or off[s + 1] == off[s]:
            continue
        a, e = off[s], off[s + 1]
        ovr = oracle_np.iou([tb[p]], db[a:e]).ravel()
        if mode == POOL_ARGMAX_SCORE:
            idx = ovr > thresh
            if np.any(idx):
                cand = np.nonzero(idx)[0]
                m = cand[int(np.argmax(sc[a:e][idx]))]
                arg[p], score[p] = a + m, float(sc[a + m])
        else:
            j = int(np.argmax(ovr))
            arg[p] = a + j
            score[p] = float(sc[a + j]) if mode == POOL_ARGMAX_IOU else float(ovr[j])
    return torch.from_numpy(arg), torch.from_numpy(score)


def _completion_row_bounded(row, bnd, miss_thr):
    """A row that is one frame range of a longer tubelet: rebuild the part of the tubelet that matters (the nearest
    valid score on either side at its true distance, missing frames in between), complete it with the reference's
    loop and cut the range out again."""
    lg, lv, rg, rv = (float(x) for x in bnd)
    left = ([lv] + [MISSING] * int(lg)) if lg >= 0 else []
    right = ([MISSING] * int(rg) + [rv]) if rg >= 0 else []
    full = oracle_np.completion_row(np.concatenate([left, row, right]), miss_thr)
    return full[len(left):len(left) + len(row)]


def score_completion_(scores, lengths=None, miss_thr=-10.0, status=None, bounds=None):
    if bounds is not None:
        a, st = scores.numpy(), 0
        for i in range(a.shape[0]):
            try:
                a[i] = _completion_row_bounded(a[i].astype(np.float64), bounds[i].numpy(), miss_thr).astype(a.dtype)
            except IndexError:
                st |= 2
        return torch.tensor([st], dtype=torch.int32)
    a = scores.numpy()
    lens = lengths.numpy() if lengths is not None else np.full(a.shape[0], a.shape[1])
    st = 0
    for i in range(a.shape[0]):
        if lens[i] == 0:
            continue
        try:
            a[i, :lens[i]] = oracle_np.completion_row(a[i, :lens[i]], miss_thr)
        except IndexError:
            st |= 2
    return torch.tensor([st], dtype=torch.int32)


def temporal_maxpool(scores, window, lengths=None, pad=MISSING, out=None):
    if window % 2 != 1:
        raise ValueError('Window size must be odd!')
    a = scores.numpy()
    lens = lengths.numpy() if lengths is not None else np.full(a.shape[0], a.shape[1])
    o = np.array(a, copy=True)
    for i in range(a.shape[0]):
        o[i, :lens[i]] = oracle_np.temporal_maxpool_row(a[i, :lens[i]], window, pad)
    return torch.from_numpy(o)


def temporal_conv1d(x, taps, pad_mode="zero", lengths=None, out=None):
    a, t = x.numpy(), taps.numpy()
    lens = lengths.numpy() if lengths is not None else np.full(a.shape[0], a.shape[1])
    o = np.zeros_like(a)
    for i in range(a.shape[0]):
        o[i, :lens[i]] = oracle_np.temporal_conv1d(a[i:i + 1, :lens[i]], t[i % len(t)][None, :], pad_mode)[0]
    return torch.from_numpy(o)


def sort_by_score_desc(scores, ids):
    order = torch.from_numpy(np.argsort(-scores.numpy(), kind="stable"))
    return scores[order], ids[order]


def tubelet_interpolate(knot_x, knot_y, knot_off, dense_first, dense_off):
    xs, ys, ko, df, do = (knot_x.numpy(), knot_y.numpy(), knot_off.numpy(), dense_first.numpy(), dense_off.numpy())
    out = np.zeros((ys.shape[0], int(do[-1])), np.float64)
    for k in range(len(ko) - 1):
        kx = xs[ko[k]:ko[k + 1]]
        for q in range(do[k + 1] - do[k]):
            for f in range(ys.shape[0]):
                out[f, do[k] + q] = oracle_np.interp_value(kx, ys[f, ko[k]:ko[k + 1]], df[k] + q)
    return torch.from_numpy(out)


def threshold_topk(scores, seg_offsets, max_seg_len, thresh=0.05, k=100):
    s, off = scores.numpy(), seg_offsets.numpy()
    S, C = len(off) - 1, s.shape[1]
    idx = np.full((S, C, k), -1, np.int32)
    cnt = np.zeros((S, C), np.int32)
    for f in range(S):
        blk = s[off[f]:off[f + 1]]
        for j in range(C):
            inds = np.where(blk[:, j] > np.float32(thresh))[0]
            if len(inds) > k:
                inds = inds[np.argsort(-blk[inds, j], kind="stable")[:k]]
            idx[f, j, :len(inds)] = inds
            cnt[f, j] = len(inds)
    return torch.from_numpy(idx), torch.from_numpy(cnt)


DOUBLES = ("nms", "vid_nms", "track_det_nms", "iou_matrix", "nms_frames", "segment_by_frame", "track_nms_step",
           "spatial_maxpool", "score_completion_", "temporal_maxpool", "temporal_conv1d", "sort_by_score_desc",
           "tubelet_interpolate", "threshold_topk")


def install(monkeypatch):
    """Swap the operators of vdetlib_b200.ops for the CPU restatements (undone by monkeypatch)."""
    from vdetlib_b200 import ops
    monkeypatch.setattr(ops, "default_device", lambda: torch.device("cpu"))
    for name in DOUBLES:
        monkeypatch.setattr(ops, name, globals()[name])
    return torch.device("cpu")
