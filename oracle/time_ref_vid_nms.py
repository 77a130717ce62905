#!/usr/bin/env python
"""The reference's TRUE whole-video path on the CPU: apply_vid_nms calls vid_nms (utils/nms.pyx:71-125) on ALL
rows of the video for one class -- it visits every pair of rows, also across frames, only to skip them
(:110-112), so its cost grows with (T*N)^2.  bench.py's cpu_baseline / --impl reference time the per-(frame,
class) `nms` loop instead, which is the favourable restatement for the reference; this script measures the real
thing on slices of config 2 and extrapolates quadratically.  CPU only.

    python -m oracle.time_ref_vid_nms > profiles/rNN_ref_vid_nms_cpu.json

TEST INFRASTRUCTURE (lives under oracle/ because only tests/, smoke() and bench.py's CPU legs may execute the
oracle): it times the reference, never the product.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, c_oracle          # noqa: E402
from vdetlib_b200 import synth                  # noqa: E402

N, C = 300, 30
mod = None
try:
    build_ref.build()
    mod = build_ref.load()
except Exception:
    pass
vid_nms = mod.vid_nms if mod is not None else c_oracle.vid_nms
nms = mod.nms if mod is not None else c_oracle.nms
out = {"kind": "reference" if mod is not None else "port", "boxes_per_frame": N, "slices": []}
for T in (5, 10, 20, 40):
    b, s = synth.boxes_scores(T, N, 1, seed=7)
    dets = np.concatenate([np.repeat(np.arange(1, T + 1), N)[:, None].astype(np.float32), b.reshape(-1, 4),
                           s.reshape(-1, 1)], axis=1).astype(np.float32)
    t0 = time.perf_counter()
    keep = vid_nms(dets, 0.3)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    per_frame = sum(len(nms(np.ascontiguousarray(dets[t * N:(t + 1) * N, 1:]), 0.3)) for t in range(T))
    dt_pf = time.perf_counter() - t0
    assert per_frame == len(keep)
    out["slices"].append({"frames": T, "rows": T * N, "vid_nms_s": round(dt, 4), "per_frame_nms_loop_s": round(dt_pf, 4),
                          "kept": len(keep)})
# quadratic fit t = a * rows^2 through the largest slice
last = out["slices"][-1]
a = last["vid_nms_s"] / last["rows"] ** 2
rows_c2 = 1000 * N
out["extrapolated_config2"] = {
    "rows": rows_c2, "vid_nms_s_per_class": round(a * rows_c2 ** 2, 1),
    "vid_nms_s_all_30_classes": round(a * rows_c2 ** 2 * C, 1),
    "boxes_per_s_true_reference_path": round(rows_c2 / (a * rows_c2 ** 2 * C), 2),
    "note": "quadratic extrapolation of the largest measured slice; labelled as such (SURVEY 8d ii)"}
print(json.dumps(out, indent=1))
