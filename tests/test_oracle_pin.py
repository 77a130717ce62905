"""Pin the CPU oracle (oracle/) before anything trusts it.

* against the committed golden vectors generated from the REAL reference
  (oracle/gen_golden.py: compiled utils/nms.pyx + the reference's own Python functions);
* against the real reference live, when oracle/_ref (and /root/reference) are available.
No GPU needed.
"""
import copy

import numpy as np
import pytest

from oracle import c_oracle, oracle_np
from vdetlib_b200 import synth
from vdetlib_b200.vdet.dataset import imagenet_vdet_classes as CLASSES

import helpers


# ---------------------------------------------------------------------------------------
# C restatement of utils/nms.pyx  vs golden
# ---------------------------------------------------------------------------------------
def test_c_oracle_nms_golden():
    g = helpers.golden_npz("nms.npz")
    n = 0
    for key in g.files:
        if key.startswith("nms_") and "_keep_" in key:
            tag, thr = key[4:].split("_keep_")
            thr = int(thr) / (100.0 if len(thr) == 3 else 10.0)
            assert c_oracle.nms(g["nms_%s_dets" % tag], thr) == g[key].tolist(), key
            n += 1
    assert n >= 18


def test_c_oracle_vid_and_track_golden():
    g = helpers.golden_npz("nms.npz")
    for thr in (0.3, 0.5):
        assert c_oracle.vid_nms(g["vid_dets"], thr) == g["vid_keep_%02d" % int(thr * 10)].tolist()
        assert c_oracle.track_det_nms(g["tdn_tracks"], g["tdn_dets"], thr) == \
            g["tdn_keep_%02d" % int(thr * 10)].tolist()


def test_c_oracle_known_answers():
    # identical boxes: the lower score is suppressed at any threshold <= 1
    d = np.asarray([[0, 0, 9, 9, 0.9], [0, 0, 9, 9, 0.8]], np.float32)
    assert c_oracle.nms(d, 1.0) == [0]
    # "+1" convention: [0..9] and [9..18] share one pixel column: inter 1*10, union 190
    d = np.asarray([[0, 0, 9, 9, 0.9], [9, 0, 18, 9, 0.8]], np.float32)
    assert c_oracle.nms(d, 10.0 / 190.0) == [0]            # threshold met exactly -> suppressed (>=)
    assert c_oracle.nms(d, np.nextafter(np.float32(10.0 / 190.0), np.float32(1))) == [0, 1]
    # frame gating: same boxes on different frames never suppress
    v = np.asarray([[1, 0, 0, 9, 9, 0.9], [2, 0, 0, 9, 9, 0.8], [1, 0, 0, 9, 9, 0.7]], np.float32)
    assert c_oracle.vid_nms(v, 0.3) == [0, 1]
    # track_det_nms: round 1 kills det 0 (overlaps the track box), round 2 runs on the rest
    t = np.asarray([[1, 0, 0, 9, 9]], np.float32)
    dd = np.asarray([[1, 0, 0, 9, 9, 0.9], [1, 50, 50, 60, 60, 0.5], [1, 51, 51, 60, 60, 0.4],
                     [2, 0, 0, 9, 9, 0.3]], np.float32)
    assert c_oracle.track_det_nms(t, dd, 0.3) == [1, 3]
    # degenerate boxes with union == 0 raise like Cython's cdivision=False
    z = np.asarray([[5, 5, 4, 4, 0.9], [7, 7, 6, 6, 0.8]], np.float32)
    with pytest.raises(ZeroDivisionError):
        c_oracle.nms(z, 0.3)
    with pytest.raises(ValueError):
        c_oracle.nms(z.astype(np.float64), 0.3)
    assert c_oracle.nms(np.zeros((0, 5), np.float32), 0.3) == []


def test_c_oracle_vs_real_cython(ref_cython):
    rng = np.random.default_rng(7)
    for trial in range(40):
        n = int(rng.integers(1, 400))
        thr = float(rng.choice([0.0, 0.1, 0.3, 0.5, 0.7, 0.95]))
        d = helpers.unique_score_dets(rng, n)
        if trial % 3 == 0:
            d[:, :4] = np.round(d[:, :4])
        assert c_oracle.nms(d, thr) == ref_cython.nms(d, thr)
        v = helpers.unique_score_dets(rng, n, with_frame=5)
        assert c_oracle.vid_nms(v, thr) == ref_cython.vid_nms(v, thr)
        q = int(rng.integers(0, 6))
        t = helpers.unique_score_dets(rng, max(q, 1), with_frame=5)[:q, :5]
        assert c_oracle.track_det_nms(t, v, thr) == [int(i) for i in ref_cython.track_det_nms(t, v, thr)]


def test_c_oracle_strided_input(ref_cython):
    rng = np.random.default_rng(8)
    big = helpers.unique_score_dets(rng, 120)
    wide = np.zeros((120, 9), np.float32)
    wide[:, ::2] = big
    view = wide[:, ::2]                       # non-contiguous, accepted by the typed buffer
    assert c_oracle.nms(view, 0.4) == ref_cython.nms(view, 0.4)


def test_bitmask_and_frames_consistent_with_nms():
    b, s = synth.boxes_scores(3, 70, 4, seed=11)
    km, ki, kc = c_oracle.nms_frames(b, s, 0.3)
    for t in range(3):
        mask = c_oracle.iou_bitmask(b[t], 0.3)
        iou = c_oracle.pair_iou_f32(b[t], b[t])
        bits = ((mask[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(70, -1)[:, :70]
        assert np.array_equal(bits.astype(bool), iou.astype(np.float64) >= 0.3)
        for c in range(4):
            keep = c_oracle.nms(np.concatenate([b[t], s[t, :, c:c + 1]], 1), 0.3)
            assert ki[t, c, :kc[t, c]].tolist() == keep
            assert np.flatnonzero(km[t, c]).tolist() == sorted(keep)


# ---------------------------------------------------------------------------------------
# NumPy restatement vs golden
# ---------------------------------------------------------------------------------------
def test_np_oracle_arrays_golden():
    g = helpers.golden_npz("arrays.npz")
    assert np.array_equal(oracle_np.iou(g["iou_a_int"], g["iou_b_int"]), g["iou_int"])
    assert np.array_equal(oracle_np.iou(g["iou_a_f"], g["iou_b_f"]), g["iou_f"])
    rows = g["rows_in"]
    done = np.stack([oracle_np.completion_row(r) for r in rows])
    assert np.array_equal(done, g["rows_completed"])
    for w in (3, 5, 9, 201):
        mp = np.stack([oracle_np.temporal_maxpool_row(r, w) for r in done])
        assert np.array_equal(mp, g["rows_maxpool_%d" % w])
    with pytest.raises(IndexError):
        oracle_np.completion_row([-1e5, -1e5])


def _proto_inputs():
    p = helpers.golden_protos()
    T, N, C = 8, 40, 5
    boxes, scores = synth.boxes_scores(T, N, C, seed=4000, integer=True, frame_offset=1e-4)
    # the stored det proto must be what synth regenerates (guards the fixture against drift)
    det = synth.det_proto(boxes, scores, CLASSES, integer=True)
    assert det == p["det"]
    return p, boxes, scores


def test_np_oracle_protos_golden():
    p, boxes, scores = _proto_inputs()
    vid, det, trk, out = p["vid"], p["det"], p["track"], p["out"]
    for cls in (1, 3):
        for suffix, thr in (("", 0.7), ("_05", 0.5)):
            got = oracle_np.dets_spatial_max_pooling(vid, copy.deepcopy(trk), copy.deepcopy(det), cls, CLASSES, thr)
            assert got == out["smp_%d%s" % (cls, suffix)]
        got = oracle_np.anchor_propagate(vid, copy.deepcopy(p["anchor_track"]), copy.deepcopy(det), cls, CLASSES)
        assert got == out["anchor_%d" % cls]
        kept = oracle_np.apply_vid_nms(copy.deepcopy(det), cls)
        assert [d['hash'] for d in kept['detections']] == out["vid_nms_%d" % cls]
    f2d = {t + 1: (boxes[t].astype(np.float64), scores[t].astype(np.float64)) for t in range(8)}
    got = oracle_np.raw_dets_spatial_max_pooling(vid, copy.deepcopy(trk), f2d, 2, CLASSES)
    assert got == out["raw_smp_2"]
    assert oracle_np.apply_image_nms(boxes[0].astype(np.float64), scores[0, :, 0].astype(np.float64), 0.4) == out["image_nms"]


def test_np_oracle_greedy_tracking_golden():
    p, boxes, scores = _proto_inputs()
    vid, det, out = p["vid"], p["det"], p["out"]
    opts = helpers.Opts(max_tracks=5, thres=0.5, nms_thres=0.3)
    tp, _ = oracle_np.greedily_track_from_det(vid, copy.deepcopy(det), helpers.fake_tracker,
                                              lambda d: oracle_np.det_score(d, 2), opts)
    assert tp == out["greedy_det"]
    T, N, C = 8, 40, 5
    det_info = np.concatenate([np.repeat(np.arange(1, T + 1), N)[:, None].astype(np.float64),
                               boxes.reshape(-1, 4).astype(np.float64),
                               scores.reshape(-1, C).astype(np.float64)], axis=1)
    opts = helpers.Opts(max_tracks=4, thres=0.6, nms_thres=None)
    tp, _ = oracle_np.greedily_track_from_raw_dets(vid, det_info, helpers.fake_tracker, 3, opts)
    assert tp == out["greedy_raw"]


# ---------------------------------------------------------------------------------------
# NumPy restatement vs the reference's own functions, live (only with /root/reference)
# ---------------------------------------------------------------------------------------
def test_np_oracle_vs_reference_live(ref_py):
    rng = np.random.default_rng(21)
    for _ in range(10):
        a = rng.uniform(0, 200, (int(rng.integers(1, 30)), 4)); a[:, 2:] += a[:, :2]
        b = rng.uniform(0, 200, (int(rng.integers(1, 30)), 4)); b[:, 2:] += b[:, :2]
        if rng.integers(0, 2):
            a, b = np.round(a).astype(np.int64), np.round(b).astype(np.int64)
        assert np.array_equal(oracle_np.iou(a, b), ref_py.iou(a, b))
    for trial in range(20):
        L = int(rng.integers(1, 80))
        row = synth.score_rows(1, L, seed=100 + trial, missing_frac=0.4, dtype=np.float64, max_run=6)[0]
        sp = {'video': 'v', 'method': 'm', 'tubelets': [{'gt': 0, 'boxes': [{'det_score': float(v)} for v in row]}]}
        ref_py.do_score_completion(sp)
        ref_done = np.asarray(helpers.tubelet_scores(sp)[0])
        assert np.array_equal(oracle_np.completion_row(row), ref_done)
        for w in (3, 7):
            out = ref_py.score_proto_temporal_maxpool(copy.deepcopy(sp), w)
            assert np.array_equal(oracle_np.temporal_maxpool_row(ref_done, w),
                                  np.asarray(helpers.tubelet_scores(out)[0]))


def test_np_oracle_interpolation_golden():
    p = helpers.golden_protos()
    got = oracle_np.score_proto_interpolation(copy.deepcopy(p["interp_in"]), p["interp_vid"])
    assert got == p["out"]["interp"]
    lens = [len(t['boxes']) for t in got['tubelets']]
    assert lens[2] == 1 and lens[0] == 38 and lens[5] == 38        # single box kept; ends stretched (:472-475)


def test_np_oracle_overlap_and_top_golden():
    p = helpers.golden_protos()
    got = oracle_np.tubelets_overlap(copy.deepcopy(p["overlap_in"]), p["annot"], 3)
    assert got == p["out"]["overlap"]
    assert got[0]['gt'] == 1 and all(t['gt'] == 0 for t in got[1:])
    det = p["det"]
    assert [d['hash'] for d in oracle_np.top_detections(copy.deepcopy(det), 50, 2)['detections']] == p["out"]["top_50"]
    assert [d['hash'] for d in oracle_np.frame_top_detections(copy.deepcopy(det), 5, 4)['detections']] == p["out"]["frame_top_5"]
