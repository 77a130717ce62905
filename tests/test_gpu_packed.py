"""GPU side of the packed protos: all classes of a packed det proto through one NMS
launch, compared with the oracle's per-class vid_nms and the golden keep lists of apply_vid_nms."""
import numpy as np
import pytest

from oracle import c_oracle
from vdetlib_b200 import synth
from vdetlib_b200.utils import packed, protocol
from vdetlib_b200.vdet.dataset import imagenet_vdet_classes

import helpers

pytestmark = pytest.mark.gpu


def _matrix(det, cls):
    """apply_vid_nms's float32 matrix (vdet/video_det.py:53-56)."""
    return np.asarray([[d['frame']] + list(d['bbox']) + [protocol.det_score(d, cls)] for d in det['detections']],
                      dtype='float32').reshape(-1, 6)


def test_packed_vid_nms_golden():
    p = helpers.golden_protos()
    det = p["det"]
    pd = packed.PackedDets.from_det_proto(det)
    keeps = packed.packed_vid_nms(pd, 0.3)
    assert len(keeps) == len(pd.class_index)
    for c, cls in enumerate(pd.class_index):
        assert keeps[c].tolist() == c_oracle.vid_nms(_matrix(det, cls), 0.3), cls
    for cls in (1, 3):            # the reference's own apply_vid_nms output (hashes in keep order)
        got = [det['detections'][i]['hash'] for i in keeps[pd.class_index.index(cls)]]
        assert got == p["out"]["vid_nms_%d" % cls]


@pytest.mark.parametrize("T,N,C,shuffle", [(6, 300, 30, False), (5, 77, 4, True), (3, 1100, 3, True)])
def test_packed_vid_nms_synthetic(T, N, C, shuffle):
    b, s = synth.boxes_scores(T, N, C, seed=T * N + C)
    rng = np.random.default_rng(N)
    # unique scores across the whole video per class (the global order is then pinned)
    for c in range(C):
        s[:, :, c] = rng.permutation(np.linspace(0.001, 0.999, T * N)).astype(np.float32).reshape(T, N)
    frames = np.repeat(np.arange(1, T + 1), N)
    boxes, scores = b.reshape(-1, 4).astype(np.float64), s.reshape(-1, C).astype(np.float64)
    if shuffle:                  # rows of different frames interleaved, ragged frames
        keep = rng.permutation(T * N)[: T * N - N // 3]
        frames, boxes, scores = frames[keep], boxes[keep], scores[keep]
    pd = packed.PackedDets("v", frames, boxes, scores, imagenet_vdet_classes[1:C + 1], list(range(1, C + 1)))
    keeps = packed.packed_vid_nms(pd, 0.3)
    for c in range(C):
        dets = np.concatenate([frames[:, None], boxes, scores[:, c:c + 1]], axis=1).astype(np.float32)
        assert keeps[c].tolist() == c_oracle.vid_nms(dets, 0.3)
    assert packed.packed_vid_nms(packed.PackedDets("v", [], np.zeros((0, 4)), np.zeros((0, 2))), 0.3)[0].size == 0
