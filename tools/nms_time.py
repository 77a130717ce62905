#!/usr/bin/env python
"""Time vdet_nms_frames_f32 alone (CUDA events, rotating input sets).  usage: python tools/nms_time.py [T N C reps]
Environment: VDET_NMS_PER_SM caps the resident CTAs per SM (measurement hook of nms_frames.cu)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vdetlib_b200 import ops, synth          # noqa: E402

T, N, C, reps = (int(x) for x in (sys.argv[1:5] + ["1000", "300", "30", "30"][len(sys.argv) - 1:]))
dev = torch.device("cuda", 0)
sets = []
for k in range(4):
    b, s = synth.boxes_scores(T, N, C, seed=2000 + k)
    sets.append((torch.from_numpy(b.reshape(-1, 4)).to(dev), torch.from_numpy(s.reshape(-1, C)).to(dev)))
seg = ops.seg_offsets_uniform(T, N, dev)
st = ops.new_status(dev)
out = None
for k in range(3):
    out = ops.nms_frames(sets[k % 4][0], sets[k % 4][1], seg, 0.3, N, want_mask=True, status=st, frame_major_out=True)
torch.cuda.synchronize()
a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for k in range(reps):
    ops.nms_frames(sets[k % 4][0], sets[k % 4][1], seg, 0.3, N, want_mask=True, status=st, frame_major_out=True,
                   out=(out[0], out[1], out[2]))
b_.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b_) / reps
print(json.dumps({"T": T, "N": N, "C": C, "per_sm_cap": os.environ.get("VDET_NMS_PER_SM"), "ms": round(ms, 4),
                  "boxes_per_s": round(T * N / ms * 1e3), "kept": int(out[1].sum().item())}))
