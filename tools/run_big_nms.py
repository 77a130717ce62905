#!/usr/bin/env python
"""Runs the big-frame NMS kernel (frames of 2000 boxes, 30 classes) a few times -- the target of
    ncu --set full -k regex:nms_frames_big -s 1 -c 1 ... python tools/run_big_nms.py
and, without ncu, prints its CUDA-event time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vdetlib_b200 import ops, synth

T = int(sys.argv[1]) if len(sys.argv) > 1 else 296
N, C = 2000, 30
dev = torch.device("cuda", 0)
b, s = synth.boxes_scores(T, N, C, seed=5)
db = torch.from_numpy(b.reshape(-1, 4)).to(dev)
ds = torch.from_numpy(s.reshape(-1, C)).to(dev)
seg = ops.seg_offsets_uniform(T, N, dev)
status = ops.new_status(dev)
for _ in range(2):
    ops.nms_frames(db, ds, seg, 0.3, N, want_mask=True, status=status, frame_major_out=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ops.nms_frames(db, ds, seg, 0.3, N, want_mask=True, status=status, frame_major_out=True)
e1.record()
torch.cuda.synchronize()
print("big nms %d frames: %.3f ms" % (T, e0.elapsed_time(e1) / 3))
