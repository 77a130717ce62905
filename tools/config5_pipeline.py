#!/usr/bin/env python
"""BASELINE config 5 on ONE GPU's share: one video of 2000 frames x 2000 boxes x 30 classes through the
whole post-CNN path -- NMS (all classes) + frame-to-frame link + tubelet rows + score completion +
temporal max-pool + temporal conv -- with per-stage CUDA-event timings.  On 8 GPUs the 8 videos are
independent replicas (no communication), so the 8-GPU figure is 8x this one.

    python tools/config5_pipeline.py [--frames 2000] [--boxes 2000] [--tubelets 256]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from vdetlib_b200 import ops, synth

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2000)
ap.add_argument("--boxes", type=int, default=2000)
ap.add_argument("--classes", type=int, default=30)
ap.add_argument("--tubelets", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
T, N, C, K = a.frames, a.boxes, a.classes, a.tubelets
dev = torch.device("cuda", 0)
db = torch.empty((T * N, 4), dtype=torch.float32, device=dev)
ds = torch.empty((T * N, C), dtype=torch.float32, device=dev)
for t0 in range(0, T, 250):
    n = min(250, T - t0)
    b, s = synth.boxes_scores(n, N, C, seed=700 + t0)
    db[t0 * N:(t0 + n) * N] = torch.from_numpy(b.reshape(-1, 4)).to(dev)
    ds[t0 * N:(t0 + n) * N] = torch.from_numpy(s.reshape(-1, C)).to(dev)
seg = ops.seg_offsets_uniform(T, N, dev)
status = ops.new_status(dev)
taps = torch.from_numpy(synth.gaussian_taps(C, 9)).to(dev)
start = torch.arange(0, K, dtype=torch.int32, device=dev)
stages = ["nms", "link", "chains", "rows", "completion", "maxpool", "conv"]
times = {k: [] for k in stages}


def run():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
    ev[0].record()
    keep_idx, keep_cnt, keep_mask, _ = ops.nms_frames(db, ds, seg, 0.3, N, want_mask=True, status=status, frame_major_out=True)
    ev[1].record()
    succ, best = ops.link_frames(db, seg, N)
    ev[2].record()
    rows = ops.follow_links(succ, best, start, T, 0.3)
    ev[3].record()
    x = ops.gather_chain_scores(ds, rows).view(K * C, T)
    ev[4].record()
    ops.score_completion_(x, status=status)
    ev[5].record()
    m = ops.temporal_maxpool(x, 5)
    ev[6].record()
    y = ops.temporal_conv1d(m, taps)
    ev[7].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(len(stages))], keep_cnt, rows, y


run()
for _ in range(a.reps):
    ms, keep_cnt, rows, y = run()
    for k, v in zip(stages, ms):
        times[k].append(v)
st = int(status.item()) & 0x7fffffff
med = {k: float(np.median(v)) for k, v in times.items()}
total = sum(med.values())
alive = float((rows >= 0).float().mean().item())
out = {"config": "config 5, one video: %d frames x %d boxes x %d classes, %d tubelets" % (T, N, C, K),
       "ms": med, "ms_total": total, "boxes_per_s": T * N / (total / 1e3),
       "boxes_per_s_nms_link": T * N / ((med["nms"] + med["link"]) / 1e3),
       "kept_fraction": float(keep_cnt.sum().item()) / (T * N * C), "chain_alive_fraction": alive,
       "status_word": st, "note": "8 videos on 8 GPUs are independent replicas: aggregate = 8x"}
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config5.json"), "w"), indent=1)
