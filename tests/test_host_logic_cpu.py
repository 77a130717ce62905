"""Host-side logic that needs no GPU: synthetic generators, threshold rounding rule, protocol helpers."""
import copy

import numpy as np

from vdetlib_b200 import synth
from vdetlib_b200.utils import protocol
from vdetlib_b200.vdet.dataset import imagenet_vdet_classes


def test_synth_is_deterministic_and_unique():
    b1, s1 = synth.boxes_scores(4, 300, 30, seed=3)
    b2, s2 = synth.boxes_scores(4, 300, 30, seed=3)
    assert np.array_equal(b1, b2) and np.array_equal(s1, s2)
    assert b1.dtype == np.float32 and s1.shape == (4, 300, 30)
    for t in range(4):
        for c in (0, 29):
            assert len(np.unique(s1[t, :, c])) == 300
    assert (b1[..., 2] >= b1[..., 0]).all() and (b1[..., 3] >= b1[..., 1]).all()
    assert b1[..., 2].max() <= 1279 and b1[..., 3].max() <= 719
    _, s3 = synth.boxes_scores(5, 50, 2, seed=1, frame_offset=1e-4)
    assert len(np.unique(s3[:, :, 0])) == 250


def test_threshold_round_up_rule():
    """(double)ovr >= thresh  <=>  ovr >= ceil_f32(thresh) for every float32 ovr (nms.pyx:65)."""
    def ceil_f32(t):
        f = np.float32(t)
        return f if float(f) >= t else np.nextafter(f, np.float32(np.inf))
    rng = np.random.default_rng(0)
    for t in [0.3, 0.5, 0.7, 0.05, 1.0, 0.0, 1e-9] + rng.uniform(0, 1, 50).tolist():
        T = ceil_f32(t)
        for ovr in [T, np.nextafter(T, np.float32(-1)), np.nextafter(T, np.float32(2)), np.float32(t)]:
            assert (float(ovr) >= t) == bool(ovr >= T)
    assert float(ceil_f32(0.7)) > 0.7 and float(ceil_f32(0.3)) > 0.3 and float(ceil_f32(0.5)) == 0.5


def test_protocol_helpers():
    assert len(imagenet_vdet_classes) == 31 and imagenet_vdet_classes[0] == "__background__"
    det = {"scores": [{"class_index": 2, "score": 0.5}]}
    assert protocol.det_score(det, 2) == 0.5 and protocol.det_score(det, 3) == float('-inf')
    tracks = [[{"frame": 1, "bbox": [0, 0, 1, 1], "score": 0.25, "anchor": 0}]]
    tub = protocol.tubelets_proto_from_tracks_proto(tracks, 5)
    assert tub[0]["boxes"][0] == {"frame": 1, "bbox": [0, 0, 1, 1], "anchor": 0, "track_score": 0.25, "det_score": -1e5}
    assert "score" in tracks[0][0]                       # the input track is untouched (shallow copy per box)
    assert tub[0]["class"] == imagenet_vdet_classes[5] and tub[0]["gt"] == 0
    # top_detections / frame_top_detections rank on the GPU: covered by tests/test_gpu_temporal_tubelet.py
    dp = {"video": "v", "detections": [{"frame": 1, "scores": [{"class_index": 1, "score": 0.1}]}]}
    assert protocol.top_detections(dp, 5, 1) == dp                 # fewer than top_num: shallow copy (:331-332)


def test_adapters_refuse_to_run_without_a_gpu():
    """No CPU fallback: the device every adapter allocates on is CUDA or nothing (the CPU runs of
    tests/test_adapters_cpu.py swap the operators for the oracle explicitly)."""
    import pytest
    import torch
    from vdetlib_b200 import ops
    from vdetlib_b200.utils import cython_nms, common
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        ops.default_device()
    with pytest.raises(RuntimeError):
        cython_nms.nms(np.zeros((3, 5), np.float32), 0.3)
    with pytest.raises(RuntimeError):
        common.iou(np.zeros((2, 4)), np.zeros((2, 4)))
