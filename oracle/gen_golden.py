#!/usr/bin/env python
"""Generate tests/golden/*.npz|json from the REAL reference, run in this container.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (read-only mount); the committed outputs
are what the GPU box -- where that tree does not exist -- checks against.

Sources of truth:
  * utils/nms.pyx compiled as-is (oracle/build_ref.py: dtype-alias patch only),
  * the reference's own Python functions exec'd through oracle/ref_py2.py (mechanical
    Python-2 -> 3 patch only).

    python -m oracle.gen_golden
"""
import copy
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import build_ref, ref_py2          # noqa: E402
from vdetlib_b200 import synth                 # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def fake_tracker(vid_proto, frame_id, bbox, opts):
    """Deterministic stand-in for the external MATLAB trackers: the anchor box drifts by
    (+2,+1) px per frame over frames [frame_id-3, frame_id+3] clipped to the video."""
    n = len(vid_proto['frames'])
    track = []
    for f in range(max(1, frame_id - 3), min(n, frame_id + 3) + 1):
        d = f - frame_id
        track.append({'frame': f, 'bbox': [bbox[0] + 2 * d, bbox[1] + d, bbox[2] + 2 * d, bbox[3] + d],
                      'score': 1.0 / (1 + abs(d)), 'anchor': d, 'hash': 'x'})
    return [track]


class Opts(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


def main():
    os.makedirs(OUT, exist_ok=True)
    build_ref.build()
    cy = build_ref.load()
    R = ref_py2.RefFunctions(cy)
    classes = R.classes

    # ---- 1. nms: BASELINE config 1 (300 boxes, 1 class) and friends -------------------------
    g = {}
    for tag, (N, seed, integer) in {"c1": (300, 1000, False), "int": (300, 1001, True),
                                    "n33": (33, 1002, False), "n1000": (1000, 1003, False)}.items():
        b, s = synth.boxes_scores(1, N, 1, seed=seed, integer=integer)
        dets = np.concatenate([b[0], s[0]], axis=1).astype(np.float32)
        g["nms_%s_dets" % tag] = dets
        for thr in (0.3, 0.5, 0.7):
            g["nms_%s_keep_%02d" % (tag, int(thr * 10))] = np.asarray(cy.nms(dets, thr), np.int64)
    # hand-built edge cases: identical boxes, touching boxes where the "+1" matters, threshold met exactly
    edge = np.asarray([[0, 0, 9, 9, 0.9], [0, 0, 9, 9, 0.8], [10, 0, 19, 9, 0.7], [9, 0, 18, 9, 0.6],
                       [0, 0, 9, 4, 0.5], [0, 5, 9, 9, 0.4], [100, 100, 100, 100, 0.3],
                       [100, 100, 101, 100, 0.2]], np.float32)
    g["nms_edge_dets"] = edge
    for thr in (0.0, 0.05, 0.1, 0.3, 0.5, 1.0):
        g["nms_edge_keep_%03d" % int(thr * 100)] = np.asarray(cy.nms(edge, thr), np.int64)

    # ---- 2. vid_nms -------------------------------------------------------------------------
    b, s = synth.boxes_scores(20, 60, 1, seed=2000)
    rng = np.random.default_rng(2001)
    frames = np.repeat(np.arange(1, 21), 60).astype(np.float32)
    dets = np.concatenate([frames[:, None], b.reshape(-1, 4), s.reshape(-1, 1)], axis=1).astype(np.float32)
    # unique scores across the whole video, rows shuffled so frames are interleaved
    dets[:, 5] = rng.permutation(np.linspace(0.001, 0.999, len(dets))).astype(np.float32)
    dets = dets[rng.permutation(len(dets))]
    g["vid_dets"] = dets
    for thr in (0.3, 0.5):
        g["vid_keep_%02d" % int(thr * 10)] = np.asarray(cy.vid_nms(dets, thr), np.int64)

    # ---- 3. track_det_nms -------------------------------------------------------------------
    sub = dets[dets[:, 0] <= 3]
    tracks = np.asarray([[1] + list(sub[sub[:, 0] == 1][0, 1:5] + 3), [2] + list(sub[sub[:, 0] == 2][5, 1:5] - 2),
                         [2] + list(sub[sub[:, 0] == 2][9, 1:5]), [7, 0, 0, 50, 50]], np.float32)
    g["tdn_tracks"] = tracks
    g["tdn_dets"] = sub
    for thr in (0.3, 0.5):
        g["tdn_keep_%02d" % int(thr * 10)] = np.asarray(cy.track_det_nms(tracks, sub, thr), np.int64)
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **g)

    # ---- 4. iou (float64), completion, temporal max-pool ------------------------------------
    g = {}
    b, _ = synth.boxes_scores(2, 50, 1, seed=3000, integer=True)
    g["iou_a_int"], g["iou_b_int"] = b[0, :40].astype(np.int64), b[1].astype(np.int64)
    g["iou_int"] = R.iou(g["iou_a_int"], g["iou_b_int"])
    b, _ = synth.boxes_scores(2, 50, 1, seed=3001)
    g["iou_a_f"], g["iou_b_f"] = b[0, :40].astype(np.float64) * 1.0001, b[1].astype(np.float64) / 3.0
    g["iou_f"] = R.iou(g["iou_a_f"], g["iou_b_f"])
    rows = synth.score_rows(12, 97, seed=3002, missing_frac=0.3, dtype=np.float64, max_run=9)
    rows[0, :5] = -1e5
    rows[1, -7:] = -1e5
    rows[2, :] = -1e5
    rows[2, 40] = 0.25
    rows[3, 1:-1] = -1e5
    g["rows_in"] = rows
    sp = {'video': 'v', 'method': 'm',
          'tubelets': [{'gt': 0, 'boxes': [{'det_score': float(v)} for v in r]} for r in rows]}
    R.do_score_completion(sp)
    done = np.asarray([[bx['det_score'] for bx in t['boxes']] for t in sp['tubelets']])
    g["rows_completed"] = done
    for w in (3, 5, 9, 201):
        sp2 = copy.deepcopy(sp)
        out = R.score_proto_temporal_maxpool(sp2, w)
        g["rows_maxpool_%d" % w] = np.asarray([[bx['det_score'] for bx in t['boxes']] for t in out['tubelets']])
    np.savez_compressed(os.path.join(OUT, "arrays.npz"), **g)

    # ---- 5. proto-level: spatial max-pooling, anchor propagate, apply_vid_nms, greedy track --
    T, N, C = 8, 40, 5
    boxes, scores = synth.boxes_scores(T, N, C, seed=4000, integer=True, frame_offset=1e-4)
    vid = synth.vid_proto(T)
    det = synth.det_proto(boxes, scores, classes, integer=True)
    trk = synth.track_proto(boxes, 6, seed=4001)
    # a tubelet box on a frame without detections and one far from everything
    trk['tracks'][0].append({'frame': 99, 'bbox': [1, 1, 20, 20], 'hash': 'h', 'score': 0.1, 'anchor': 50})
    trk['tracks'][1][0]['bbox'] = [1200, 650, 1270, 710]
    protos = {"vid": vid, "det": det, "track": trk, "out": {}}
    for cls in (1, 3):
        protos["out"]["smp_%d" % cls] = R.dets_spatial_max_pooling(
            vid, copy.deepcopy(trk), copy.deepcopy(det), cls)
        protos["out"]["smp_%d_05" % cls] = R.dets_spatial_max_pooling(
            vid, copy.deepcopy(trk), copy.deepcopy(det), cls, overlap_thres=0.5)
        trk_ap = copy.deepcopy(trk)
        trk_ap['tracks'][0] = trk_ap['tracks'][0][:-1]      # anchor_propagate needs dets on the anchor frame
        for k, tr in enumerate(trk_ap['tracks']):            # exactly one anchor==0 box per track
            for q, bx in enumerate(tr):
                bx['anchor'] = q - len(tr) // 2
        protos["out"]["anchor_%d" % cls] = R.anchor_propagate(vid, trk_ap, copy.deepcopy(det), cls)
        protos["anchor_track"] = trk_ap
        kept = R.apply_vid_nms(copy.deepcopy(det), cls)
        protos["out"]["vid_nms_%d" % cls] = [d['hash'] for d in kept['detections']]
    frame_to_det = {t + 1: (boxes[t].astype(np.float64), scores[t].astype(np.float64)) for t in range(T)}
    protos["out"]["raw_smp_2"] = R.raw_dets_spatial_max_pooling(vid, copy.deepcopy(trk), frame_to_det, 2)
    # greedy tubelet proposal with the deterministic fake tracker
    opts = Opts(max_tracks=5, thres=0.5, nms_thres=0.3)
    tp = R.greedily_track_from_det(vid, copy.deepcopy(det), fake_tracker,
                                   lambda d: R.det_score(d, 2), opts)
    protos["out"]["greedy_det"] = tp
    det_info = np.concatenate([np.repeat(np.arange(1, T + 1), N)[:, None].astype(np.float64),
                               boxes.reshape(-1, 4).astype(np.float64),
                               scores.reshape(-1, C).astype(np.float64)], axis=1)
    opts = Opts(max_tracks=4, thres=0.6, nms_thres=None)
    protos["out"]["greedy_raw"] = R.greedily_track_from_raw_dets(vid, det_info, fake_tracker, 3, opts)
    # image nms
    protos["out"]["image_nms"] = R.apply_image_nms(boxes[0].astype(np.float64), scores[0, :, 0].astype(np.float64), 0.4)
    # ---- 6. score_proto_interpolation on strided tubelets ---------------------
    rng = np.random.default_rng(4100)
    vid40 = synth.vid_proto(40)
    tubs = []
    for k, frames in enumerate([list(range(2, 39, 3)), list(range(5, 40, 5)), [7], list(range(1, 41, 4)),
                                [3, 4, 9, 10, 30], list(range(2, 40, 2))]):
        bxs = []
        x1, y1 = float(rng.uniform(0, 500)), float(rng.uniform(0, 300))
        for q, fr in enumerate(frames):
            bb = [x1 + 3.5 * q, y1 + 1.25 * q, x1 + 3.5 * q + 80.0, y1 + 1.25 * q + 60.0]
            if k % 2 == 0:
                bb = [int(v) for v in bb]
            bxs.append({'frame': fr, 'bbox': bb, 'det_score': float(rng.uniform(-1, 1)), 'anchor': fr - frames[len(frames) // 2],
                        'track_score': 0.5, 'hash': 'h'})
        tubs.append({'gt': 0, 'class': classes[3], 'class_index': 3, 'boxes': bxs})
    sp_strided = {'video': vid40['video'], 'method': 'm', 'tubelets': tubs}
    protos["interp_in"] = sp_strided
    protos["interp_vid"] = vid40
    protos["out"]["interp"] = R.score_proto_interpolation(copy.deepcopy(sp_strided), vid40)
    # ---- 7. tubelets_overlap / merge_score_protos / top detections ----------
    annot = {'video': vid['video'], 'annotations': []}
    smp = protos["out"]["smp_3"]
    for k, tub in enumerate(smp['tubelets'][:4]):
        tr = []
        for q, bx in enumerate(tub['boxes']):
            if bx['frame'] > T:
                continue
            bb = list(bx['bbox']) if k == 0 else [bx['bbox'][0] + 3 * k, bx['bbox'][1] - 2 * k, bx['bbox'][2] + k, bx['bbox'][3] + 5]
            tr.append({'frame': bx['frame'], 'bbox': bb, 'class_index': 3 if (k < 3 or q < 2) else 7, 'class': 'c',
                       'name': 'n', 'generated': False, 'occluded': False})
        annot['annotations'].append({'id': str(k), 'track': tr})
    annot['annotations'].append({'id': 'other', 'track': [{'frame': 1, 'bbox': [0, 0, 50, 50], 'class_index': 9}]})
    protos["annot"] = annot
    tubs_in = copy.deepcopy(smp['tubelets'])
    tubs_in[0]['boxes'] = [b for b in tubs_in[0]['boxes'] if b['frame'] <= T]      # this one coincides with annotation 0
    protos["overlap_in"] = copy.deepcopy(tubs_in)
    protos["out"]["overlap"] = R.tubelets_overlap(tubs_in, annot, 3)
    p1, p2 = copy.deepcopy(protos["out"]["smp_1"]), copy.deepcopy(protos["out"]["smp_1_05"])
    protos["out"]["merge_max"] = R.merge_score_protos(p1, p2, scheme='max')
    p1, p2 = copy.deepcopy(protos["out"]["smp_1"]), copy.deepcopy(protos["out"]["smp_3"])
    protos["out"]["merge_combine"] = R.merge_score_protos(p1, p2, scheme='combine')
    protos["out"]["top_50"] = [d['hash'] for d in R.top_detections(copy.deepcopy(det), 50, 2)['detections']]
    protos["out"]["top_all"] = [d['hash'] for d in R.top_detections(copy.deepcopy(det), 100000, 2)['detections']]
    protos["out"]["frame_top_5"] = [d['hash'] for d in R.frame_top_detections(copy.deepcopy(det), 5, 4)['detections']]
    with open(os.path.join(OUT, "protos.json"), "w") as f:
        json.dump(protos, f, default=lambda o: o.tolist() if hasattr(o, "tolist") else float(o))
    print("golden vectors written to", OUT)


def det_mat_case():
    """Inputs of the raw-detection loader golden: a vid proto and, per frame, what its .mat file holds
    (None = no file).  Covers: ``<basename>.mat`` vs ``<path>.mat`` lookup, a frame without boxes, a
    missing file, float32 ('single') and float64 arrays, frames listed out of order."""
    rng = np.random.default_rng(5000)
    vid = {'video': 'golden_mat', 'root_path': '/nowhere',
           'frames': [{'frame': f, 'path': '%06d.JPEG' % f} for f in (1, 2, 3, 5, 4, 6)]}
    C = 4
    mats = {}
    for f, n, dt, by_path in ((1, 7, np.float64, False), (2, 0, np.float64, False), (3, 5, np.float32, True),
                              (4, 1, np.float64, False), (5, 9, np.float32, False)):
        boxes = np.round(rng.uniform(0, 600, (n, 4))).astype(dt) + (0.25 if dt is np.float32 else 0.0)
        zs = rng.normal(0, 2, (n, C)).astype(dt)
        mats[f] = (boxes, zs, by_path)
    return vid, mats                                    # frame 6: no file at all


def write_det_mats(vid, mats, det_dir):
    import scipy.io as sio
    for fr in vid['frames']:
        if fr['frame'] not in mats:
            continue
        boxes, zs, by_path = mats[fr['frame']]
        name = (fr['path'] if by_path else os.path.splitext(fr['path'])[0]) + '.mat'
        sio.savemat(os.path.join(det_dir, name), {'boxes': boxes, 'zs': zs})


def main_det_mat():
    """tests/golden/det_mat.npz: the reference's load_det_info / load_frame_to_det
    (utils/protocol.py:528-555) on .mat files written from det_mat_case()."""
    import tempfile
    build_ref.build()
    R = ref_py2.RefFunctions(build_ref.load())
    vid, mats = det_mat_case()
    g = {}
    with tempfile.TemporaryDirectory() as d:
        write_det_mats(vid, mats, d)
        info = R.load_det_info(vid, d)
        f2d = R.load_frame_to_det(vid, d)
    g["det_info"] = np.asarray(info)
    g["frames_with_file"] = np.asarray(sorted(f2d), np.int64)
    for f, (b, z) in f2d.items():
        g["f2d_boxes_%d" % f] = b
        g["f2d_zs_%d" % f] = z
    np.savez_compressed(os.path.join(OUT, "det_mat.npz"), **g)
    print("det_mat golden written:", g["det_info"].shape, g["frames_with_file"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "det_mat":
        main_det_mat()
    else:
        main()
        main_det_mat()
