"""Seeded synthetic inputs of SURVEY 8(d) (NumPy only; shared by tests and bench.py).

Frame 1280x720; x1~U[0,1180), y1~U[0,620), w,h~U[16,256), clipped to the frame; half of the
boxes are jittered copies (+-10% of w/h) of N/20 cluster seeds so that NMS has realistic
suppression rates; class scores are a random permutation of linspace(0.001, 0.999, N) per
(frame, class) -- unique, which removes the reference's un-pinnable argsort tie order.
"""
import numpy as np

FRAME_W, FRAME_H = 1280, 720


def boxes_scores(n_frames, n_boxes, n_classes, seed=0, integer=False, frame_offset=0.0):
    """boxes [T,N,4] float32, scores [T,N,C] float32.

    ``frame_offset`` > 0 adds t*frame_offset to frame t's scores so that scores are unique
    across the WHOLE video (needed to pin vid_nms's global keep order)."""
    rng = np.random.default_rng(seed)
    T, N, C = n_frames, n_boxes, n_classes
    x1 = rng.uniform(0, FRAME_W - 100, (T, N))
    y1 = rng.uniform(0, FRAME_H - 100, (T, N))
    w = rng.uniform(16, 256, (T, N))
    h = rng.uniform(16, 256, (T, N))
    n_seed = max(N // 20, 1)
    n_jit = N // 2
    if n_jit > 0:
        src = rng.integers(0, n_seed, (T, n_jit))
        tt = np.arange(T)[:, None]
        jx = rng.uniform(-0.1, 0.1, (T, n_jit))
        jy = rng.uniform(-0.1, 0.1, (T, n_jit))
        jw = rng.uniform(-0.1, 0.1, (T, n_jit))
        jh = rng.uniform(-0.1, 0.1, (T, n_jit))
        sw, sh = w[tt, src], h[tt, src]
        x1[:, N - n_jit:] = x1[tt, src] + jx * sw
        y1[:, N - n_jit:] = y1[tt, src] + jy * sh
        w[:, N - n_jit:] = sw * (1 + jw)
        h[:, N - n_jit:] = sh * (1 + jh)
    x1 = np.clip(x1, 0, FRAME_W - 2)
    y1 = np.clip(y1, 0, FRAME_H - 2)
    x2 = np.minimum(x1 + w, FRAME_W - 1)
    y2 = np.minimum(y1 + h, FRAME_H - 1)
    boxes = np.stack([x1, y1, x2, y2], axis=-1)
    if integer:
        boxes = np.round(boxes)
    base = np.linspace(0.001, 0.999, N)
    scores = rng.permuted(np.broadcast_to(base, (T, C, N)).copy(), axis=-1)
    scores = np.ascontiguousarray(scores.transpose(0, 2, 1))
    if frame_offset:
        scores = scores + np.arange(T)[:, None, None] * frame_offset
    return boxes.astype(np.float32), scores.astype(np.float32)


def score_rows(n_rows, length, seed=0, missing_frac=0.05, dtype=np.float32, max_run=20):
    """[n_rows, length] tubelet score rows in (0,1) with runs (length U{1..max_run}) of -1e5."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 1.0, (n_rows, length)).astype(dtype)
    n_runs = max(int(missing_frac * length / ((1 + max_run) / 2.0)), 1) if missing_frac > 0 else 0
    for r in range(n_rows):
        starts = rng.integers(0, length, n_runs)
        lens = rng.integers(1, max_run + 1, n_runs)
        for s, l in zip(starts, lens):
            x[r, s:s + l] = -1e5
        if not (x[r] > -10).any():
            x[r, rng.integers(0, length)] = 0.5
    return x


def gaussian_taps(n_channels, window, dtype=np.float32):
    h = window // 2
    k = np.arange(-h, h + 1, dtype=np.float64)
    sig = max(window / 4.0, 0.5)
    g = np.exp(-0.5 * (k / sig) ** 2)
    g /= g.sum()
    return np.tile(g.astype(dtype), (n_channels, 1))


# ---- protocol-dict builders (for the proto-level adapters) --------------------------------
def vid_proto(n_frames, name="synthetic_vid"):
    return {"video": name, "root_path": "/nonexistent",
            "frames": [{"frame": t + 1, "path": "%06d.JPEG" % (t + 1)} for t in range(n_frames)]}


def det_proto(boxes, scores, class_names, name="synthetic_vid", integer=True):
    """boxes [T,N,4], scores [T,N,C-1] (classes 1..C-1) -> det proto (utils/protocol.py:77-110)."""
    dets = []
    T, N = boxes.shape[:2]
    for t in range(T):
        for i in range(N):
            bb = [int(v) for v in boxes[t, i]] if integer else [float(v) for v in boxes[t, i]]
            dets.append({"frame": t + 1, "bbox": bb, "hash": "%d_%d" % (t, i),
                         "scores": [{"class": class_names[c + 1], "class_index": c + 1,
                                     "score": float(scores[t, i, c])} for c in range(scores.shape[2])]})
    return {"video": name, "detections": dets}


def track_proto(boxes, n_tracks, seed=0, name="synthetic_vid", jitter=6):
    """Tracks that follow detection i of every frame with integer jitter (so that IoU > 0.7 mostly holds)."""
    rng = np.random.default_rng(seed)
    T, N = boxes.shape[:2]
    tracks = []
    for k in range(n_tracks):
        start = int(rng.integers(0, max(T // 3, 1)))
        stop = int(rng.integers(start + 1, T + 1))
        anchor = int(rng.integers(start, stop))
        tr = []
        for t in range(start, stop):
            i = int(rng.integers(0, N))
            b = boxes[t, i] + rng.integers(-jitter, jitter + 1, 4)
            tr.append({"frame": t + 1, "bbox": [int(v) for v in b], "hash": "t%d_%d" % (k, t),
                       "score": float(rng.uniform()), "anchor": t - anchor})
        tracks.append(tr)
    return {"video": name, "method": "synthetic", "tracks": tracks}
