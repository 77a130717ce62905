"""Protocol-dict helpers the hot-path adapters need (reference utils/protocol.py).

The proto schemas (vid / box / det / track / score / annot, utils/protocol.py:7-192) are kept
verbatim: plain JSON-style dicts.  Only the helpers on the hot path are provided; JSON / .mat
file I/O is host-side and out of scope.
"""
import copy
import gzip
import json
import os

from ..vdet.dataset import imagenet_vdet_classes


def proto_load(file_path):
    """utils/protocol.py:209-220 (transparent .gz)."""
    if os.path.isfile(file_path + '.gz'):
        file_path += '.gz'
    if os.path.splitext(file_path)[1] == '.gz':
        with gzip.open(file_path, 'rt') as f:
            return json.load(f)
    with open(file_path, 'r') as f:
        return json.load(f)


def proto_dump(obj, file_path):
    """utils/protocol.py:223-236."""
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


def top_detections(det_proto, top_num, class_index):
    """utils/protocol.py:330-339."""
    if len(det_proto['detections']) < top_num:
        return copy.copy(det_proto)
    ranked = sorted(copy.copy(det_proto['detections']),
                    key=lambda x: det_score(x, class_index), reverse=True)
    return {'video': det_proto['video'], 'detections': ranked[:top_num]}


def frame_top_detections(det_proto, top_num, class_index):
    """utils/protocol.py:341-351."""
    out = {'video': det_proto['video'], 'detections': []}
    for frame_id in list(set(d['frame'] for d in det_proto['detections'])):
        cur = sorted([d for d in det_proto['detections'] if d['frame'] == frame_id],
                     key=lambda x: det_score(x, class_index), reverse=True)
        out['detections'].extend(cur[:top_num])
    return out


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
