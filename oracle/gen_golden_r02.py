#!/usr/bin/env python
"""Round-2 golden vectors from the REAL reference (run in this container; /root/reference mounted):

  * rcnn_sampling_dets_scoring (vdet/tubelet_cls.py:196-260) with the CNN / SVM replaced by oracle/fakes.py,
  * score_conv_cls (vdet/tubelet_cls.py:15-51): the blobs the reference hands to its net + the conv_scores it writes,
  * fast_rcnn_det_vid (vdet/video_det.py:64-106): the threshold + top-k epilogue behind a fake det_fun.

    python -m oracle.gen_golden_r02        ->  tests/golden/r02.json
TEST INFRASTRUCTURE ONLY."""
import copy
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import fakes, ref_py2          # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "r02.json")


def _plain(o):
    if isinstance(o, dict):
        return dict((k, _plain(v)) for k, v in o.items())
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    if isinstance(o, np.ndarray):
        return _plain(o.tolist())
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.integer,)):
        return int(o)
    return o


def conv_score_proto(P, R):
    """A score proto with every field score_conv_cls reads, from the golden track proto."""
    tubelets = R.tubelets_proto_from_tracks_proto(copy.deepcopy(P["track"]["tracks"]), 3)
    rng = np.random.default_rng(11)
    for t in tubelets:
        t['gt'] = 0
        for b in t['boxes']:
            b['det_score'] = float(rng.uniform(-1, 1))
            b['gt_overlap'] = float(rng.choice([0.0, 0.3, 0.5, 0.77]))
            b['all_score'] = [float(v) for v in rng.uniform(0, 1, 4)]
            b['feat'] = [float(v) for v in rng.uniform(0, 1, 3)]
    return {'video': P["track"]["video"], 'method': 'golden', 'tubelets': tubelets}


def main():
    R = ref_py2.RefFunctions()
    P = json.load(open(os.path.join(ROOT, "tests", "golden", "protos.json")))
    out = {}
    # ---- rcnn_sampling_dets_scoring -----------------------------------------------------------
    R.ns.update(googlenet_features=lambda img, boxes, net, layer: fakes.cnn_features(boxes),
                svm_scores=lambda feats, model: fakes.svm_scores_200(feats),
                svm_from_rcnn_model=lambda m: m,
                index_vdet_to_det=dict((c, c + 20) for c in range(31)))
    for cls, thr, sf in ((1, 0.7, False), (3, 0.5, True)):
        tub = R.rcnn_sampling_dets_scoring(copy.deepcopy(P["vid"]), copy.deepcopy(P["track"]), copy.deepcopy(P["det"]),
                                           None, cls, None, overlap_thres=thr, save_feat=sf, save_all_sc=sf)
        out["rcnn_%d" % cls] = {"class_idx": cls, "thr": thr, "save": sf, "score_column": cls + 20 - 1, "tubelets": _plain(tub)}
    # ---- score_conv_cls: marshalling + conv_score -----------------------------------------------
    for tag, channels in (("small", {"det_scores": 1, "track_scores": 1, "anchors": 1, "abs_anchors": 1,
                                     "gt_overlaps": 1, "labels": 1}),
                          ("two", {"det_scores": 1, "labels": 1})):
        sp = conv_score_proto(P, R)
        net = fakes.RecordingNet(channels)
        res = R.score_conv_cls(sp, net)
        out["conv_%s" % tag] = {"channels": channels, "score_proto_in": _plain(conv_score_proto(P, R)),
                                "blobs": [_plain(c) for c in net.calls],
                                "conv_scores": [[b['conv_score'] for b in t['boxes']] for t in res['tubelets']]}
    # ---- fast_rcnn_det_vid: threshold + top-k ------------------------------------------------------
    vid = copy.deepcopy(P["vid"])
    vid["frames"] = vid["frames"][:3]
    rng = np.random.default_rng(3)
    box_proto = {"video": vid["video"], "boxes": [
        {"frame": f["frame"], "bbox": [float(v) for v in rng.uniform(0, 500, 4)], "hash": "h"}
        for f in vid["frames"] for _ in range(140 + 10 * f["frame"])]}
    all_boxes = R.fast_rcnn_det_vid(7, vid, box_proto, fakes.det_fun, max_per_image=100, thresh=0.05)
    out["topk"] = {"net": 7, "vid": vid, "box_proto": box_proto,
                   "all_boxes": [[_plain(np.asarray(a, dtype=np.float32)) for a in per_cls] for per_cls in all_boxes]}
    json.dump(out, open(OUT, "w"))
    print("written", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
