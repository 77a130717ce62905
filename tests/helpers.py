"""Shared test helpers (oracle access + golden fixtures)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_protos():
    with open(os.path.join(GOLDEN, "protos.json")) as f:
        return json.load(f)


def fake_tracker(vid_proto, frame_id, bbox, opts):
    """Same deterministic tracker oracle/gen_golden.py used."""
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


def tubelet_scores(score_proto):
    return [[b['det_score'] for b in t['boxes']] for t in score_proto['tubelets']]


def tubelet_boxes(score_proto):
    return [[list(b['bbox']) for b in t['boxes']] for t in score_proto['tubelets']]


def unique_score_dets(rng, n, with_frame=None, scale=300.0):
    """Random [n,5] or [n,6] float32 dets with unique scores."""
    x1 = rng.uniform(0, scale, n); y1 = rng.uniform(0, scale, n)
    w = rng.uniform(5, scale / 2, n); h = rng.uniform(5, scale / 2, n)
    s = rng.permutation(np.linspace(0.01, 0.99, n))
    cols = [x1, y1, x1 + w, y1 + h, s]
    if with_frame is not None:
        cols = [rng.integers(0, with_frame, n).astype(np.float64)] + cols
    return np.stack(cols, axis=1).astype(np.float32)
