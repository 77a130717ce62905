"""Protocol-dict helpers the hot-path adapters need (reference utils/protocol.py).

The proto schemas (vid / box / det / track / score / annot, utils/protocol.py:7-192) are kept
verbatim: plain JSON-style dicts.  Only the helpers on the hot path are provided.  File I/O:
``proto_load`` / ``proto_dump`` as in the reference (JSON, transparent gzip) plus the packed
binary container of ``vdetlib_b200.utils.packed`` for paths ending in ``.vdetpk``.
"""
import copy
import gzip
import json
import os

import numpy as np

from ..vdet.dataset import imagenet_vdet_classes


def proto_load(file_path):
    """utils/protocol.py:209-220 (transparent .gz).  A ``.vdetpk`` path -- or a ``.vdetpk`` side-car
    next to the JSON file, preferred like the reference prefers ``.gz`` -- is read from the packed
    container instead of being parsed as JSON.  The side-car is a derived cache: it is used only while it is
    at least as new as the file it stands for (a regenerated JSON / .gz wins over a stale side-car)."""
    if os.path.splitext(file_path)[1] != '.vdetpk' and os.path.isfile(file_path + '.vdetpk'):
        sources = [q for q in (file_path, file_path + '.gz') if os.path.isfile(q)]
        car = file_path + '.vdetpk'
        if not sources or os.path.getmtime(car) >= max(os.path.getmtime(q) for q in sources):
            file_path = car
        else:
            import logging
            logging.warning("proto_load: %s is older than %s; ignoring the stale side-car", car, sources[0])
    if os.path.splitext(file_path)[1] == '.vdetpk':
        from .packed import proto_load_packed
        return proto_load_packed(file_path)
    if os.path.isfile(file_path + '.gz'):
        file_path += '.gz'
    if os.path.splitext(file_path)[1] == '.gz':
        with gzip.open(file_path, 'rt') as f:
            return json.load(f)
    with open(file_path, 'r') as f:
        return json.load(f)


def proto_dump(obj, file_path):
    """utils/protocol.py:223-236; ``.vdetpk`` paths are written as the packed container."""
    if os.path.splitext(file_path)[1] == '.vdetpk':
        from .packed import proto_dump_packed
        return proto_dump_packed(obj, file_path)
    if os.path.splitext(file_path)[1] == '.gz':
        with gzip.open(file_path, 'wt', compresslevel=1) as f:
            json.dump(obj, f, indent=2)
        return
    with open(file_path, 'w') as f:
        json.dump(obj, f, indent=2)


def det_score(detection, class_index):
    """Score of ``class_index`` in a det proto entry; -inf when absent.  utils/protocol.py:323-327."""
    for score in detection['scores']:
        if score['class_index'] == class_index:
            return score['score']
    return float('-inf')


def score_proto(class_names, scores):
    """utils/protocol.py:307-320."""
    if type(scores) is not list:
        scores = scores.tolist()
    return [{'class': name, 'class_index': idx, 'score': sc}
            for idx, (name, sc) in enumerate(zip(class_names, scores))]


def _class_scores(det_proto, class_index):
    return np.asarray([det_score(d, class_index) for d in det_proto['detections']], dtype=np.float64)


def top_detections(det_proto, top_num, class_index):
    """The ``top_num`` highest-scoring detections of the video.  utils/protocol.py:330-339
    (stable descending sort; fewer than top_num detections -> a shallow copy, unsorted, :331-332).
    The ranking is a stable radix sort of the float64 scores on the GPU."""
    import torch
    from .. import ops
    if len(det_proto['detections']) < top_num:
        return copy.copy(det_proto)
    scores = _class_scores(det_proto, class_index)                     # Python floats -> float64 keys
    dev = ops.default_device()
    ids = torch.arange(len(scores), dtype=torch.int64, device=dev)
    order = ops.sort_by_score_desc(torch.from_numpy(scores).to(dev), ids)[1].cpu().tolist()
    dets = det_proto['detections']
    return {'video': det_proto['video'], 'detections': [dets[i] for i in order[:top_num]]}


def frame_top_detections(det_proto, top_num, class_index):
    """The ``top_num`` best detections of every frame.  utils/protocol.py:341-351 (frames visited in
    the iteration order of ``set(frames)`` like the reference, stable descending sort inside a frame:
    one stable sort by score then by frame on the GPU)."""
    import torch
    from .. import ops
    new_det = {'video': det_proto['video'], 'detections': []}
    dets = det_proto['detections']
    if not dets:
        return new_det
    frame_idx = list(set([d['frame'] for d in dets]))
    scores = _class_scores(det_proto, class_index)
    frames = np.asarray([d['frame'] for d in dets], dtype=np.float32)
    if not np.array_equal(frames.astype(np.float64), np.asarray([d['frame'] for d in dets], dtype=np.float64)):
        raise NotImplementedError("frame ids must be exactly representable in float32 (|frame| < 2^24)")
    dev = ops.default_device()
    row_ids, seg_off, seg_frame, _ = ops.segment_by_frame(torch.from_numpy(frames).to(dev), None,
                                                           torch.from_numpy(scores).to(dev))
    row_ids, seg_off = row_ids.cpu().numpy(), seg_off.cpu().numpy()
    seg_of = {float(f): s for s, f in enumerate(seg_frame.cpu().tolist())}
    for frame_id in frame_idx:
        s = seg_of[float(np.float32(frame_id))]
        new_det['detections'].extend(dets[i] for i in row_ids[seg_off[s]:seg_off[s + 1]][:top_num])
    return new_det


def tubelets_proto_from_tracks_proto(tracks_proto, class_index):
    """Track proto -> tubelets with det_score = -1e5.  utils/protocol.py:448-464."""
    tubelets = []
    for track in tracks_proto:
        boxes = []
        for box in track:
            tb = copy.copy(box)
            tb['track_score'] = tb['score']
            tb['det_score'] = -1e5
            del tb['score']
            boxes.append(tb)
        tubelets.append({'gt': 0, 'class_index': class_index,
                         'class': imagenet_vdet_classes[class_index], 'boxes': boxes})
    return tubelets


def tubelets_overlap(tubelets_proto, annot_proto, class_idx):
    """Ground-truth overlap of every tubelet box (``gt_overlap``) and the ``gt`` flag of tubelets
    that coincide with an annotation.  utils/protocol.py:467-489: for a tubelet of class c, the best
    IoU against the same-frame boxes of annotation tracks whose boxes carry class c (a track is read
    up to its first box of another class, :476-478); ``gt = 1`` when the mean IoU is 1 (:486-488).
    IoU in float64 on the GPU; IN PLACE like the reference."""
    import torch
    from .. import _lib, ops
    classes = sorted(set(t['class_index'] for t in tubelets_proto))
    for c in classes:
        frame_boxes = {}
        for annot_track in annot_proto['annotations']:
            for annot_box in annot_track['track']:
                if annot_box['class_index'] != c:
                    break
                frame_boxes.setdefault(annot_box['frame'], []).append(annot_box['bbox'])
        tubs = [t for t in tubelets_proto if t['class_index'] == c]
        boxes = [b for t in tubs for b in t['boxes']]
        if not boxes:
            continue
        frames = sorted(frame_boxes)
        seg_of = {f: s for s, f in enumerate(frames)}
        seg_off = np.zeros(len(frames) + 1, dtype=np.int32)
        np.cumsum([len(frame_boxes[f]) for f in frames], out=seg_off[1:])
        ann = (np.concatenate([np.asarray(frame_boxes[f], dtype=np.float64).reshape(-1, 4) for f in frames])
               if frames else np.zeros((0, 4)))
        tb = np.asarray([b['bbox'] for b in boxes], dtype=np.float64).reshape(-1, 4)
        seg = np.asarray([seg_of.get(b['frame'], -1) for b in boxes], dtype=np.int32)
        dev = ops.default_device()
        dummy = torch.zeros(max(len(ann), 1), dtype=torch.float64, device=dev)
        arg, best = ops.spatial_maxpool(torch.from_numpy(tb).to(dev), torch.from_numpy(seg).to(dev),
                                        torch.from_numpy(ann).to(dev), dummy, torch.from_numpy(seg_off).to(dev),
                                        0.0, _lib.POOL_MAX_IOU)
        arg, best = arg.cpu().numpy(), best.cpu().numpy()
        for b, a, v in zip(boxes, arg, best):
            b['gt_overlap'] = float(v) if (a >= 0 and v > 0) else 0            # :472, :482-483
    for tubelet in tubelets_proto:
        ious = [box['gt_overlap'] for box in tubelet['boxes']]
        mean_iou = np.asarray(ious).mean()
        if abs(mean_iou - 1) < np.finfo(float).eps:
            tubelet['gt'] = 1
    return tubelets_proto


def merge_score_protos(proto_1, proto_2, scheme='combine'):
    """utils/protocol.py:504-525: 'combine' appends the tubelets of proto_2, 'max' keeps per box the
    entry with the larger det_score (host-side dict logic; no arithmetic beyond one comparison)."""
    assert scheme in ['combine', 'max']
    assert proto_1['video'] == proto_2['video']
    new_proto = copy.copy(proto_1)
    if proto_1['method'] != proto_2['method']:
        new_proto['method'] = '_'.join([proto_1['method'], proto_2['method']])
    if scheme == 'combine':
        new_proto['tubelets'].extend(copy.copy(proto_2['tubelets']))
    elif scheme == 'max':
        for tubelet1, tubelet2 in zip(new_proto['tubelets'], proto_2['tubelets']):
            assert tubelet1['gt'] == tubelet2['gt']
            assert tubelet1['class'] == tubelet2['class']
            assert tubelet1['class_index'] == tubelet2['class_index']
            for box1, box2 in zip(tubelet1['boxes'], tubelet2['boxes']):
                assert box1['frame'] == box2['frame']
                assert box1['anchor'] == box2['anchor']
                if box1['det_score'] < box2['det_score']:
                    for key in box1:
                        box1[key] = copy.copy(box2[key])
    return new_proto
