"""Deterministic stand-ins for what the post-CNN path receives from outside vdetlib (Caffe nets, SVMs, images).

TEST INFRASTRUCTURE ONLY: shared by oracle/gen_golden_r02.py (which runs the REFERENCE's functions with them) and by
the tests (which run the repo's adapters and the NumPy restatements with the same ones)."""
import numpy as np


def cnn_features(boxes):
    """'pool5 features' of boxes [P,4]: a fixed function of the coordinates, 6 values per box."""
    b = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
    return np.stack([b[:, 0], b[:, 1], b[:, 2], b[:, 3], b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], axis=1)


def svm_scores_200(features):
    """'SVM scores' [P,200]: column k = frac((f . w_k)) in [0,1) -- both branches of `max_score > score` occur."""
    f = np.asarray(features, dtype=np.float64).reshape(-1, 6)
    k = np.arange(200, dtype=np.float64)
    w = np.stack([(k * 7 + 3) % 11, (k * 5 + 1) % 13, (k * 3 + 2) % 7, (k + 5) % 17, (k * 2 + 1) % 5, (k * 11) % 19]) / 97.0
    return np.mod(f.dot(w), 1.0)


class RecordingNet(object):
    """Caffe's surface as score_conv_cls uses it (vdet/tubelet_cls.py:36-48): ``blobs`` with ``shape`` / ``reshape`` /
    ``data``, ``forward()`` -> {'probs': (1, 2, 1, L)}.  Records every forward's input blobs."""

    class Blob(object):
        def __init__(self, channels):
            self.data = np.zeros((1, channels, 1, 1), dtype=np.float32)

        @property
        def shape(self):
            return self.data.shape

        def reshape(self, *shape):
            self.data = np.zeros(shape, dtype=np.float32)

    def __init__(self, channels):
        self.blobs = dict((name, self.Blob(c)) for name, c in channels.items())
        self.calls = []

    def forward(self):
        snap = dict((name, b.data.copy()) for name, b in self.blobs.items())
        self.calls.append(snap)
        L = snap[sorted(snap)[0]].shape[3]
        z = np.zeros(L, dtype=np.float64)
        for name in sorted(snap):
            z += snap[name].astype(np.float64).sum(axis=(0, 1, 2)) * (1 + len(name) % 3)
        p1 = 1.0 / (1.0 + np.exp(-z))
        return {'probs': np.stack([1 - p1, p1]).reshape(1, 2, 1, L).astype(np.float32)}


def det_fun(net, im, orig_boxes):
    """Fast R-CNN's im_detect as fast_rcnn_det_vid calls it (vdet/video_det.py:83): scores [R, C], boxes [R, 4C]."""
    rng = np.random.default_rng(int(net) + len(orig_boxes))
    R = len(orig_boxes)
    C = 31
    scores = rng.uniform(0, 1, (R, C)).astype(np.float32) ** 3
    scores[:, 5] = rng.uniform(0.06, 1, R)                    # one class with more than max_per_image survivors
    boxes = np.tile(np.asarray(orig_boxes, dtype=np.float32).reshape(R, 4), (1, C)) + \
        np.repeat(np.arange(C, dtype=np.float32), 4)[None, :]
    return scores, boxes
