"""Randomised check of the NumPy / C oracle against the reference's OWN functions, live.

Needs /root/reference (this container); on the GPU box, where that tree does not exist, every test here
skips.  The committed golden vectors pin a handful of cases per function; this file widens the pin to many
random protos per function, through oracle/ref_py2.py (the reference's source, mechanically made Python 3)
and oracle/_ref (its nms.pyx compiled as is).  The oracle is what the GPU parity tests compare with, so an
oracle == reference statement over random inputs carries over to every GPU test that uses the oracle."""
import copy

import numpy as np
import pytest

from oracle import c_oracle, oracle_np
from vdetlib_b200 import synth
from vdetlib_b200.vdet.dataset import imagenet_vdet_classes as CLASSES

import helpers

pytestmark = pytest.mark.filterwarnings("ignore::DeprecationWarning")      # raised inside the reference's own code


def _case(seed, T=6, N=30, C=4, integer=True):
    boxes, scores = synth.boxes_scores(T, N, C, seed=seed, integer=integer, frame_offset=1e-4)
    vid = synth.vid_proto(T)
    det = synth.det_proto(boxes, scores, CLASSES, integer=integer)
    trk = synth.track_proto(boxes, 5, seed=seed + 1)
    return boxes, scores, vid, det, trk


@pytest.mark.parametrize("seed", range(6))
def test_spatial_maxpool_anchor_vidnms_random(ref_py, seed):
    boxes, scores, vid, det, trk = _case(7000 + seed, integer=bool(seed % 2))
    for cls in (1, 3):
        for thr in (0.7, 0.5, 0.2):
            want = ref_py.dets_spatial_max_pooling(vid, copy.deepcopy(trk), copy.deepcopy(det), cls, overlap_thres=thr)
            got = oracle_np.dets_spatial_max_pooling(vid, copy.deepcopy(trk), copy.deepcopy(det), cls, CLASSES, thr)
            assert got == want
        f2d = {t + 1: (boxes[t].astype(np.float64), scores[t].astype(np.float64)) for t in range(len(boxes))}
        want = ref_py.raw_dets_spatial_max_pooling(vid, copy.deepcopy(trk), f2d, cls, overlap_thres=0.6)
        assert oracle_np.raw_dets_spatial_max_pooling(vid, copy.deepcopy(trk), f2d, cls, CLASSES, 0.6) == want
        # exactly one anchor == 0 box per track (anchor_propagate's assertion, tubelet_cls.py:369)
        trk_ap = copy.deepcopy(trk)
        for tr in trk_ap['tracks']:
            for q, bx in enumerate(tr):
                bx['anchor'] = q - len(tr) // 2
        want = ref_py.anchor_propagate(vid, copy.deepcopy(trk_ap), copy.deepcopy(det), cls)
        assert oracle_np.anchor_propagate(vid, copy.deepcopy(trk_ap), copy.deepcopy(det), cls, CLASSES) == want
        want = ref_py.apply_vid_nms(copy.deepcopy(det), cls)
        got = oracle_np.apply_vid_nms(copy.deepcopy(det), cls)
        assert [d['hash'] for d in got['detections']] == [d['hash'] for d in want['detections']]


@pytest.mark.parametrize("seed", range(4))
def test_greedy_tracking_random(ref_py, seed):
    T, N, C = 10, 40, 3
    boxes, scores, vid, det, _ = _case(7100 + seed, T, N, C)
    opts = helpers.Opts(max_tracks=6, thres=0.3, nms_thres=0.3 if seed % 2 else None)
    want = ref_py.greedily_track_from_det(vid, copy.deepcopy(det), helpers.fake_tracker,
                                          lambda d: ref_py.det_score(d, 2), opts)
    got, _ = oracle_np.greedily_track_from_det(vid, copy.deepcopy(det), helpers.fake_tracker,
                                               lambda d: oracle_np.det_score(d, 2), opts)
    assert got == want
    det_info = np.concatenate([np.repeat(np.arange(1, T + 1), N)[:, None].astype(np.float64),
                               boxes.reshape(-1, 4).astype(np.float64), scores.reshape(-1, C).astype(np.float64)], axis=1)
    want = ref_py.greedily_track_from_raw_dets(vid, det_info, helpers.fake_tracker, 1 + seed % C, opts)
    got, _ = oracle_np.greedily_track_from_raw_dets(vid, det_info, helpers.fake_tracker, 1 + seed % C, opts)
    assert got == want


@pytest.mark.parametrize("seed", range(4))
def test_interpolation_overlap_top_random(ref_py, seed):
    rng = np.random.default_rng(7200 + seed)
    n_frames = 60
    vid = synth.vid_proto(n_frames)
    tubs = []
    for k in range(12):
        frames = np.sort(rng.choice(np.arange(1, n_frames + 1), size=int(rng.integers(1, 25)), replace=False)).tolist()
        tubs.append({'gt': 0, 'class': CLASSES[2], 'class_index': 2, 'boxes': [
            {'frame': fr, 'bbox': (rng.uniform(0, 700, 4) if k % 2 else rng.integers(0, 700, 4)).tolist(),
             'det_score': float(rng.normal()), 'anchor': fr - frames[0], 'track_score': 0.5, 'hash': 'h'} for fr in frames]})
    sp = {'video': vid['video'], 'method': 'm', 'tubelets': tubs}
    want = ref_py.score_proto_interpolation(copy.deepcopy(sp), vid)
    assert oracle_np.score_proto_interpolation(copy.deepcopy(sp), vid) == want
    # ground-truth overlap against random annotation tracks
    annot = {'video': vid['video'], 'annotations': []}
    for k in range(5):
        tr = []
        for bx in tubs[k]['boxes']:
            jit = rng.integers(-8, 9, 4)
            tr.append({'frame': bx['frame'], 'bbox': [float(v) + float(j) * (k > 0) for v, j in zip(bx['bbox'], jit)],
                       'class_index': 2 if rng.random() < 0.9 else 5})
        annot['annotations'].append({'id': str(k), 'track': tr})
    want = ref_py.tubelets_overlap(copy.deepcopy(tubs), annot, 2)
    assert oracle_np.tubelets_overlap(copy.deepcopy(tubs), annot, 2) == want
    _, _, _, det, _ = _case(7300 + seed, 5, 25, 3)
    for top in (7, 40, 10000):
        want = ref_py.top_detections(copy.deepcopy(det), top, 2)
        assert oracle_np.top_detections(copy.deepcopy(det), top, 2) == want
        want = ref_py.frame_top_detections(copy.deepcopy(det), min(top, 9), 1)
        assert oracle_np.frame_top_detections(copy.deepcopy(det), min(top, 9), 1) == want


@pytest.mark.parametrize("seed", range(3))
def test_c_oracle_vs_cython_random_frames(ref_cython, seed):
    """nms_frames (what the GPU bench path is compared with) == the compiled reference per (frame, class)."""
    T, N, C = 4, 200 + 37 * seed, 5
    b, s = synth.boxes_scores(T, N, C, seed=7400 + seed, integer=bool(seed % 2))
    for thr in (0.3, 0.5):
        km, ki, kc = c_oracle.nms_frames(b, s, thr)
        for t in range(T):
            for c in range(C):
                dets = np.concatenate([b[t], s[t, :, c:c + 1]], axis=1).astype(np.float32)
                want = ref_cython.nms(dets, thr)
                assert ki[t, c, :kc[t, c]].tolist() == want and kc[t, c] == len(want)
                assert np.array_equal(np.nonzero(km[t, c])[0], np.sort(want))
    succ, best = c_oracle.link_f32(b)
    iou32 = c_oracle.pair_iou_f32(b[0], b[1])
    assert np.array_equal(succ[0], np.argmax(iou32, axis=1)) and np.array_equal(best[0], iou32.max(axis=1))
