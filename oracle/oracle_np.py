"""NumPy / pure-Python CPU restatement of vdetlib's Python hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product (vdetlib_b200/)
never does and fails loudly when its CUDA library is missing.

Every function cites the reference lines it follows (paths relative to
/root/reference).  The array-level functions are the arithmetic; the proto-level
functions walk the protocol dicts the way the reference does and call the array-level
ones.  Parity pin: while /root/reference is mounted, tests/test_oracle_pin.py runs the
reference's OWN functions (oracle/ref_py2.py: in-memory Python-2 -> 3 patch, no
arithmetic change) against these on seeded inputs, and oracle/gen_golden.py stores the
reference's outputs under tests/golden/ for the GPU box, where the tree is absent.

The NMS family (utils/nms.pyx) lives in oracle/nms_oracle.c (C restatement) and
oracle/_ref (the real Cython, compiled from the reference tree).
"""
import copy
import os
from collections import defaultdict

import numpy as np

from . import c_oracle

MISSING = -1e5            # utils/protocol.py:459, vdet/tubelet_cls.py:344,402


# --------------------------------------------------------------------------------------
# array level
# --------------------------------------------------------------------------------------
def iou(boxes1, boxes2):
    """Dense IoU matrix, float64, +1 pixel convention.  utils/common.py:451-468."""
    b1 = np.asarray(boxes1).astype('float')
    b2 = np.asarray(boxes2).astype('float')
    ax1, ay1, ax2, ay2 = (b1[:, [k]] for k in range(4))                  # [A,1]  (IndexError on 1-D input, as the reference)
    bx1, by1, bx2, by2 = (b2[:, [k]].T for k in range(4))                # [1,B]
    iw = np.maximum(0, np.minimum(ax2, bx2) - np.maximum(ax1, bx1) + 1)  # :455-459
    ih = np.maximum(0, np.minimum(ay2, by2) - np.maximum(ay1, by1) + 1)  # :457-460
    area_a = (ax2 - ax1 + 1) * (ay2 - ay1 + 1)                            # :462-463
    area_b = (bx2 - bx1 + 1) * (by2 - by1 + 1)                            # :464-465
    inter = iw * ih
    return 1. * inter / (area_a + area_b - inter)                         # :467


def completion_row(scores, miss_thr=-10.0):
    """One tubelet's det_score row after do_score_completion.  vdet/tubelet_cls.py:284-303.

    Maximal runs of ``score <= -10`` are filled: a leading run with the first valid value
    to its right (:293-295), a trailing run with the last valid value to its left
    (:296-298), an interior run [i,j) with ``l + (r - l) * (k - i + 1) / (j - i + 1)``
    (:299-303).  An all-missing row raises IndexError like the reference (:295).
    """
    s = np.array(scores, dtype=np.float64)
    n = len(s)
    i = 0
    while i < n:
        if s[i] > miss_thr:
            i += 1
            continue
        j = i
        while j < n and s[j] <= miss_thr:
            j += 1
        if i == 0:
            if j >= n:
                raise IndexError("list index out of range")
            s[i:j] = s[j]
        elif j == n:
            s[i:j] = s[i - 1]
        else:
            l, r = float(s[i - 1]), float(s[j])
            for k in range(i, j):
                s[k] = l + (r - l) * (k - i + 1) / (j - i + 1)
        i = j
    return s


def temporal_maxpool_row(scores, window_size, pad=MISSING):
    """out[i] = max(scores[i-h .. i+h]), out-of-range = -1e5.  vdet/tubelet_cls.py:399-409
    (the pad / tile / roll / column-max / slice construction reduces to this window)."""
    s = np.asarray(scores, dtype=np.float64)
    h = window_size // 2
    n = len(s)
    ext = np.concatenate([np.full(h, pad), s, np.full(h, pad)])
    out = np.empty(n, np.float64)
    for i in range(n):
        out[i] = ext[i:i + 2 * h + 1].max()
    return out


def spatial_maxpool_frame(tub_boxes, det_boxes, det_scores, overlap_thres=0.7):
    """Rows of vdet/tubelet_cls.py:330-347 for all tubelet boxes of ONE frame.

    Returns (score[P] f64, arg[P] int64) where arg = index into det_boxes of the FIRST
    arg-max class score among dets with IoU > overlap_thres (strict), or -1 when none
    (score = -1e5, box unchanged).
    """
    tub_boxes = np.asarray(tub_boxes)
    det_scores = np.asarray(det_scores)
    P = len(tub_boxes)
    out_s = np.full(P, MISSING, np.float64)
    out_a = np.full(P, -1, np.int64)
    for p in range(P):
        overlaps = iou([tub_boxes[p]], det_boxes)
        idx = (overlaps > overlap_thres).ravel()
        if np.any(idx):
            cand = np.nonzero(idx)[0]
            m = int(np.argmax(det_scores[idx]))
            out_s[p] = float(det_scores[cand[m]])
            out_a[p] = cand[m]
    return out_s, out_a


def threshold_topk_frame(scores, boxes, thresh=0.05, max_per_image=100):
    """Post-CNN per-class score floor + cap for ONE frame.  vdet/video_det.py:88-100.

    scores [R, C] (column 0 = background), boxes [R, 4*C].  Returns a list of length C;
    entry j (j >= 1) is float32 [K,5] = (x1,y1,x2,y2,score): rows in ascending row order when
    K <= max_per_image, else the top max_per_image in descending score (:93-96).
    """
    scores = np.asarray(scores)
    boxes = np.asarray(boxes)
    out = [None] * scores.shape[1]
    for j in range(1, scores.shape[1]):
        inds = np.where(scores[:, j] > thresh)[0]
        cls_scores = scores[inds, j]
        cls_boxes = boxes[inds, j * 4:(j + 1) * 4]
        if len(cls_scores) > max_per_image:
            top = np.argsort(-cls_scores, kind='stable')[:max_per_image]
            cls_scores = cls_scores[top]
            cls_boxes = cls_boxes[top, :]
        out[j] = np.hstack((cls_boxes, cls_scores[:, np.newaxis])).astype(np.float32, copy=False)
    return out


def temporal_conv1d(x, taps, pad_mode="zero"):
    """Depthwise 1-D correlation along the frame axis (build-defined stand-in for
    score_conv_cls, vdet/tubelet_cls.py:15-51, whose Caffe net is not in the reference:
    SURVEY 8a row 12 / 8c "parity unpinned").

    x [..., C, L]; taps [C, w] (w odd).  out[c,i] = sum_k taps[c,k] * x[c, i+k-h],
    accumulated left to right in x's dtype with separate multiply and add (no FMA);
    out-of-range samples are 0 ("zero") or the nearest edge sample ("edge").
    """
    x = np.asarray(x)
    taps = np.asarray(taps, dtype=x.dtype)
    C, w = taps.shape
    h = w // 2
    L = x.shape[-1]
    if pad_mode == "zero":
        ext = np.concatenate([np.zeros(x.shape[:-1] + (h,), x.dtype), x,
                              np.zeros(x.shape[:-1] + (h,), x.dtype)], axis=-1)
    elif pad_mode == "edge":
        ext = np.concatenate([np.repeat(x[..., :1], h, -1), x, np.repeat(x[..., -1:], h, -1)], axis=-1)
    else:
        raise ValueError(pad_mode)
    acc = np.zeros_like(x)
    for k in range(w):
        term = (taps[:, k][:, None] * ext[..., k:k + L]).astype(x.dtype)
        acc = (acc + term).astype(x.dtype)
    return acc


# --------------------------------------------------------------------------------------
# protocol helpers (utils/protocol.py)
# --------------------------------------------------------------------------------------
def det_score(detection, class_index):
    """utils/protocol.py:323-327 (linear scan; -inf when the class is absent)."""
    for sc in detection['scores']:
        if sc['class_index'] == class_index:
            return sc['score']
    return float('-inf')


def tubelets_proto_from_tracks_proto(tracks_proto, class_index, class_names):
    """utils/protocol.py:448-464."""
    out = []
    for track in tracks_proto:
        boxes = []
        for box in track:
            b = copy.copy(box)
            b['track_score'] = b['score']
            b['det_score'] = MISSING
            del b['score']
            boxes.append(b)
        out.append({'gt': 0, 'class_index': class_index, 'class': class_names[class_index],
                    'boxes': boxes})
    return out


# --------------------------------------------------------------------------------------
# proto level
# --------------------------------------------------------------------------------------
def apply_image_nms(boxes, scores, thres=0.3):
    """vdet/image_det.py:117-123."""
    box_score = np.asarray(np.r_['-1', boxes, np.reshape(scores, (-1, 1))], dtype='float32')
    return c_oracle.nms(box_score, thres)


def apply_vid_nms(det_proto, class_index, thres=0.3):
    """vdet/video_det.py:51-61 -- note the hard-coded 0.3 (:57); ``thres`` is ignored."""
    rows = [[d['frame']] + list(d['bbox']) + [det_score(d, class_index)]
            for d in det_proto['detections']]
    dets = np.asarray(rows, dtype='float32').reshape(-1, 6)
    keep = c_oracle.vid_nms(dets, thresh=0.3)
    return {'video': det_proto['video'],
            'detections': [det_proto['detections'][i] for i in keep]}


def do_score_completion(score_proto):
    """vdet/tubelet_cls.py:284-303 (in place)."""
    for tubelet in score_proto['tubelets']:
        boxes = tubelet['boxes']
        if not boxes:
            continue
        row = completion_row([b['det_score'] for b in boxes])
        for b, v in zip(boxes, row):
            if not (b['det_score'] > -10):          # only missing entries are rewritten (:287)
                b['det_score'] = float(v)


def _frame_index(items, key='frame'):
    idx = defaultdict(list)
    for i, it in enumerate(items):
        idx[it[key]].append(i)
    return idx


def _pool_tubelets(vid_proto, tubelets, frame_dets, overlap_thres):
    """Common body of vdet/tubelet_cls.py:324-347 and :509-532.
    frame_dets(frame_id) -> (det_boxes ndarray [N,4], det_scores ndarray [N]) or None."""
    where = defaultdict(list)
    for i, tub in enumerate(tubelets):
        for j, box in enumerate(tub['boxes']):
            where[box['frame']].append((i, j))
    for frame in vid_proto['frames']:
        fid = frame['frame']
        got = frame_dets(fid)
        if got is None:
            continue
        det_boxes, det_scores = got
        for i, j in where[fid]:
            cur = tubelets[i]['boxes'][j]
            s, a = spatial_maxpool_frame([cur['bbox']], det_boxes, det_scores, overlap_thres)
            if a[0] >= 0:
                cur['det_score'] = float(s[0])
                cur['bbox'] = det_boxes[a[0]].tolist()
            else:
                cur['det_score'] = float(MISSING)


def dets_spatial_max_pooling(vid_proto, track_proto, det_proto, class_idx, class_names,
                             overlap_thres=0.7):
    """vdet/tubelet_cls.py:305-350."""
    assert vid_proto['video'] == track_proto['video']
    tubelets = tubelets_proto_from_tracks_proto(track_proto['tracks'], class_idx, class_names)
    by_frame = _frame_index(det_proto['detections'])
    dets = det_proto['detections']

    def frame_dets(fid):
        ids = by_frame.get(fid, [])
        if not ids:
            return None
        return (np.asarray([dets[i]['bbox'] for i in ids]),
                np.asarray([dets[i]['scores'][class_idx - 1]['score'] for i in ids]))   # :329 positional

    _pool_tubelets(vid_proto, tubelets, frame_dets, overlap_thres)
    sp = {'video': vid_proto['video'],
          'method': "spatial_max_pooling_IOU_{}".format(overlap_thres),
          'tubelets': tubelets}
    do_score_completion(sp)
    return sp


def raw_dets_spatial_max_pooling(vid_proto, track_proto, frame_to_det, class_idx, class_names,
                                 overlap_thres=0.7):
    """vdet/tubelet_cls.py:493-535."""
    assert vid_proto['video'] == track_proto['video']
    tubelets = tubelets_proto_from_tracks_proto(track_proto['tracks'], class_idx, class_names)

    def frame_dets(fid):
        if fid not in frame_to_det:
            return None
        det_boxes, det_scores = frame_to_det[fid]
        if det_boxes.size == 0:
            return None
        return det_boxes, det_scores[:, class_idx - 1].ravel()

    _pool_tubelets(vid_proto, tubelets, frame_dets, overlap_thres)
    sp = {'video': vid_proto['video'],
          'method': "spatial_max_pooling_IOU_{}".format(overlap_thres),
          'tubelets': tubelets}
    do_score_completion(sp)
    return sp


def anchor_propagate(vid_proto, track_proto, det_proto, class_idx, class_names):
    """vdet/tubelet_cls.py:353-383."""
    assert vid_proto['video'] == track_proto['video']
    tubelets = tubelets_proto_from_tracks_proto(track_proto['tracks'], class_idx, class_names)
    by_frame = _frame_index(det_proto['detections'])
    dets = det_proto['detections']
    for tub in tubelets:
        anchors = [b for b in tub['boxes'] if b['anchor'] == 0]
        assert len(anchors) == 1
        ids = by_frame.get(anchors[0]['frame'], [])
        det_boxes = np.asarray([dets[i]['bbox'] for i in ids])
        det_scores = np.asarray([dets[i]['scores'][class_idx - 1]['score'] for i in ids])
        overlaps = iou([anchors[0]['bbox']], det_boxes)[0]
        score = det_scores[np.argmax(overlaps)]
        for b in tub['boxes']:
            b['det_score'] = score
    return {'video': vid_proto['video'], 'method': "anchor_propagate", 'tubelets': tubelets}


def score_proto_temporal_maxpool(score_proto, window_size):
    """vdet/tubelet_cls.py:386-414 (shallow copy: the input tubelets ARE mutated, :393)."""
    if window_size == 1:
        return score_proto
    if window_size % 2 != 1:
        raise ValueError('Window size must be odd!')
    new = copy.copy(score_proto)
    new['method'] += '_temporal_maxpool_{}'.format(window_size)
    for tub in new['tubelets']:
        if tub['gt'] == 1:
            raise ValueError('Dangerous: Score file contains gt tracks!')
        row = temporal_maxpool_row([b['det_score'] for b in tub['boxes']], window_size)
        for b, v in zip(tub['boxes'], row):
            b['det_score'] = float(v)
    return new


def greedy_track_nms_step(det_info, keep, frame_to_det_ids, new_tracks, nms_thres):
    """The suppression step of greedily_track_from_det / _from_raw_dets:
    vdet/track.py:172-183 and :238-249.  ``keep`` (list of bool) is updated in place."""
    for tracklet in new_tracks:
        for box in tracklet:
            fid = box['frame']
            det_ids = [i for i in frame_to_det_ids[fid] if keep[i]]
            if len(det_ids) == 0:
                continue
            t = np.asarray([[fid] + list(box['bbox'])], dtype=np.float32)
            kp = set(c_oracle.track_det_nms(t, det_info[det_ids], nms_thres))
            for i, det_id in enumerate(det_ids):
                if i not in kp:
                    keep[det_id] = False


def greedily_track_from_raw_dets(vid_proto, det_info, track_method, class_idx, opts):
    """vdet/track.py:189-252 with the MATLAB retry (:227-234) left out (external engine)."""
    nms_thres = opts.nms_thres if getattr(opts, 'nms_thres', None) is not None else 0.3
    det_info = np.asarray(sorted(det_info[:, [0, 1, 2, 3, 4, 4 + class_idx]],
                                 key=lambda r: r[5], reverse=True), dtype=np.float32).reshape(-1, 6)
    frame_to_det_ids = defaultdict(list)
    for i, det in enumerate(det_info):
        frame_to_det_ids[det[0]].append(i)
    keep = [True] * len(det_info)
    cur = 0
    tracks = []
    while np.any(keep) and len(tracks) < opts.max_tracks:
        while cur < len(keep) and not keep[cur]:
            cur += 1
        if cur == len(keep):
            break
        top = det_info[cur]
        cur += 1
        if top[-1] < opts.thres:
            break
        new_tracks = track_method(vid_proto, int(top[0]), [int(v) for v in top[1:5]], opts)
        tracks.extend(new_tracks)
        greedy_track_nms_step(det_info, keep, frame_to_det_ids, new_tracks, nms_thres)
    return {'video': vid_proto['video'], 'method': track_method.__name__, 'tracks': tracks}, keep


def greedily_track_from_det(vid_proto, det_proto, track_method, score_fun, opts):
    """vdet/track.py:122-186."""
    nms_thres = opts.nms_thres if getattr(opts, 'nms_thres', None) is not None else 0.3
    assert vid_proto['video'] == det_proto['video']
    dets = sorted(det_proto['detections'], key=lambda x: score_fun(x), reverse=True)
    det_info = np.asarray([[d['frame']] + list(d['bbox']) + [score_fun(d)] for d in dets],
                          dtype=np.float32).reshape(-1, 6)
    frame_to_det_ids = defaultdict(list)
    for i, d in enumerate(dets):
        frame_to_det_ids[d['frame']].append(i)
    keep = [True] * len(dets)
    cur = 0
    tracks = []
    while np.any(keep) and len(tracks) < opts.max_tracks:
        while cur < len(keep) and not keep[cur]:
            cur += 1
        if cur == len(keep):
            break
        top = dets[cur]
        cur += 1
        if score_fun(top) < opts.thres:
            break
        new_tracks = track_method(vid_proto, top['frame'], [int(v) for v in top['bbox']], opts)
        tracks.extend(new_tracks)
        greedy_track_nms_step(det_info, keep, frame_to_det_ids, new_tracks, nms_thres)
    return {'video': vid_proto['video'], 'method': track_method.__name__, 'tracks': tracks}, keep


def interp_value(xs, ys, x):
    """One interpolated value as the reference computes it (vdet/tubelet_cls.py:416-428, :462-463):
    numpy.interp's formula inside the knots (scipy interp1d kind='linear' delegates to it), the
    extrap1d closed forms outside."""
    xs = np.asarray(xs, dtype=np.float64)
    ys = np.asarray(ys, dtype=np.float64)
    x = np.float64(x)
    if x < xs[0]:
        return ys[0] + (x - xs[0]) * (ys[1] - ys[0]) / (xs[1] - xs[0])
    if x > xs[-1]:
        return ys[-1] + (x - xs[-1]) * (ys[-1] - ys[-2]) / (xs[-1] - xs[-2])
    j = int(np.searchsorted(xs, x, side='right')) - 1
    if j == len(xs) - 1 or xs[j] == x:
        return ys[j]
    slope = (ys[j + 1] - ys[j]) / (xs[j + 1] - xs[j])
    return slope * (x - xs[j]) + ys[j]


def score_proto_interpolation(score_proto, vid_proto):
    """vdet/tubelet_cls.py:430-490."""
    new = {'video': score_proto['video'], 'method': score_proto['method'] + '_interpolation', 'tubelets': []}
    max_frames = len(vid_proto['frames'])
    for tubelet in score_proto['tubelets']:
        if tubelet['gt'] == 1:
            raise ValueError('Dangerous: Score file contains gt tracks!')
        if len(tubelet['boxes']) < 2:
            new['tubelets'].append(copy.copy(tubelet))
            continue
        boxes = tubelet['boxes']
        idx = np.asarray([b['frame'] for b in boxes])
        order = np.argsort(idx, kind='mergesort')
        xs = idx[order]
        fields = {
            'x1': [b['bbox'][0] for b in boxes], 'y1': [b['bbox'][1] for b in boxes],
            'x2': [b['bbox'][2] for b in boxes], 'y2': [b['bbox'][3] for b in boxes],
            'det_score': [b['det_score'] for b in boxes], 'anchor': [b['anchor'] for b in boxes]}
        fields = {k: np.asarray(v, dtype=np.float64)[order] for k, v in fields.items()}
        lo, hi = int(idx.min()), int(idx.max())
        if lo == 2:
            lo = 1
        if hi == max_frames - 1:
            hi = max_frames
        out = {key: tubelet[key] for key in ['gt', 'class', 'class_index']}
        out['boxes'] = []
        for d in range(lo, hi + 1):
            v = {k: float(interp_value(xs, fields[k], d)) for k in fields}
            out['boxes'].append({'frame': d, 'det_score': v['det_score'], 'anchor': v['anchor'],
                                 'bbox': [v['x1'], v['y1'], v['x2'], v['y2']]})
        new['tubelets'].append(out)
    return new


def tubelets_overlap(tubelets_proto, annot_proto, class_idx):
    """utils/protocol.py:467-489 (in place)."""
    for tubelet in tubelets_proto:
        c = tubelet['class_index']
        for tb in tubelet['boxes']:
            tb['gt_overlap'] = 0
            for annot_track in annot_proto['annotations']:
                for ab in annot_track['track']:
                    if ab['class_index'] != c:
                        break
                    if tb['frame'] == ab['frame']:
                        cur = float(iou([ab['bbox']], [tb['bbox']]).ravel()[0])
                        if cur > tb['gt_overlap']:
                            tb['gt_overlap'] = cur
        mean_iou = np.asarray([b['gt_overlap'] for b in tubelet['boxes']]).mean()
        if abs(mean_iou - 1) < np.finfo(float).eps:
            tubelet['gt'] = 1
    return tubelets_proto


def top_detections(det_proto, top_num, class_index):
    """utils/protocol.py:330-339."""
    if len(det_proto['detections']) < top_num:
        return copy.copy(det_proto)
    ranked = sorted(copy.copy(det_proto['detections']), key=lambda x: det_score(x, class_index), reverse=True)
    return {'video': det_proto['video'], 'detections': ranked[:top_num]}


def frame_top_detections(det_proto, top_num, class_index):
    """utils/protocol.py:341-351."""
    out = {'video': det_proto['video'], 'detections': []}
    for frame_id in list(set([d['frame'] for d in det_proto['detections']])):
        cur = sorted([d for d in det_proto['detections'] if d['frame'] == frame_id],
                     key=lambda x: det_score(x, class_index), reverse=True)
        out['detections'].extend(cur[:top_num])
    return out


def rcnn_sampling_dets_scoring(vid_proto, track_proto, det_proto, net, class_idx, rcnn_model, class_names,
                               overlap_thres=0.7, save_feat=False, save_all_sc=False, score_column=None):
    """vdet/tubelet_cls.py:196-260 with the CNN / SVM as callables (``net(path, boxes) -> features``,
    ``rcnn_model(features) -> scores``; ``score_column`` = ``index_vdet_to_det[class_idx] - 1`` of :217)."""
    tubelets_proto = tubelets_proto_from_tracks_proto(track_proto['tracks'], class_idx, class_names)
    for frame in vid_proto['frames']:
        frame_id = frame['frame']
        path = str(os.path.join(vid_proto['root_path'], frame['path']))
        boxes = [next((b['bbox'] for b in t['boxes'] if b['frame'] == frame_id), None) for t in tubelets_proto]   # :205
        valid_boxes = np.asarray([box for box in boxes if box is not None])
        valid_index = [i for i, box in enumerate(boxes) if box is not None]
        if len(valid_index) == 0:
            continue
        features = net(path, valid_boxes)                                   # :213
        scores = np.asarray(rcnn_model(features))                           # :214
        cls_scores = scores[:, score_column]                                # :215-218
        dets = [det for det in det_proto['detections'] if det['frame'] == frame_id]
        det_boxes = np.asarray([x['bbox'] for x in dets])
        det_scores = np.asarray([det_score(x, class_idx) for x in dets])
        for score, tubelet_id, feat, all_score in zip(cls_scores, valid_index, features, scores):
            cur_box = [box for box in tubelets_proto[tubelet_id]['boxes'] if box['frame'] == frame_id]
            assert len(cur_box) == 1
            if len(det_boxes) > 0:
                overlaps = iou([cur_box[0]['bbox']], det_boxes)
                conf_idx = (overlaps > overlap_thres).ravel()
            else:
                conf_idx = [False]
            if np.any(conf_idx):
                conf_boxes = det_boxes[conf_idx]
                conf_scores = det_scores[conf_idx]
                max_idx = np.argmax(conf_scores)
                max_score = conf_scores[max_idx]
                max_box = conf_boxes[max_idx].tolist()
            else:
                max_score = -np.inf
            if max_score > score:                                           # :239
                cur_box[0]['det_score'] = max_score
                cur_box[0]['bbox'] = max_box
                max_feat = net(path, [max_box])
                if save_feat:
                    cur_box[0]['feat'] = np.asarray(max_feat).ravel().tolist()
                if save_all_sc:
                    cur_box[0]['all_score'] = np.asarray(rcnn_model(max_feat)).ravel().tolist()
            else:
                cur_box[0]['det_score'] = score
                if save_feat:
                    cur_box[0]['feat'] = np.asarray(feat).ravel().tolist()
                if save_all_sc:
                    cur_box[0]['all_score'] = np.asarray(all_score).ravel().tolist()
    return tubelets_proto


def score_conv_cls(score_proto, net):
    """vdet/tubelet_cls.py:15-51: per tubelet, the 1-D channels go into ``(1, C, 1, L)`` blobs of ``net`` (anything with
    Caffe's ``blobs`` / ``forward`` surface) and ``probs[:, 1, :]`` comes back as ``conv_score``."""
    new_score_proto = copy.copy(score_proto)
    for tubelet in new_score_proto['tubelets']:
        track = {}
        track['length'] = len(tubelet['boxes'])
        track['gt'] = tubelet['gt']
        track['mean_iou'] = np.mean([[x['gt_overlap'] for x in tubelet['boxes']]])
        track['det_scores'] = [x['det_score'] for x in tubelet['boxes']]
        track['track_scores'] = [x['track_score'] for x in tubelet['boxes']]
        track['anchors'] = [x['anchor'] * 1. / track['length'] for x in tubelet['boxes']]
        track['abs_anchors'] = [abs(a) for a in track['anchors']]
        track['gt_overlaps'] = [x['gt_overlap'] for x in tubelet['boxes']]
        track['labels'] = [1 if v >= 0.5 else 0 for v in track['gt_overlaps']]
        if 'all_scores' in net.blobs.keys():
            track['all_scores'] = [x['all_score'] for x in tubelet['boxes']]
        if 'feats' in net.blobs.keys():
            track['feats'] = [x['feat'] for x in tubelet['boxes']]
        for blob_name in set(net.blobs.keys()).intersection(set(track.keys())):
            num_channels = net.blobs[blob_name].shape[1]
            net.blobs[blob_name].reshape(1, num_channels, 1, track['length'])
            net.blobs[blob_name].data[...] = np.asarray(track[blob_name], dtype='float32')
        blobs_out = net.forward()
        probs = blobs_out['probs'][:, 1, :]
        for box, prob in zip(tubelet['boxes'], probs.ravel()):
            box['conv_score'] = float(prob)
    return new_score_proto
