#!/usr/bin/env python
"""Time vdet_nms_frames_f32 over a list of shapes under the library's measurement hooks (CUDA events, 4 rotating
input sets).  usage: python tools/nms_shapes.py ["T,N,C" ...]
Per shape: the library's own plan, then VDET_NMS_THREADS=256 / 320 (the two CTA shapes); "same" = identical outputs."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vdetlib_b200 import ops, synth          # noqa: E402

shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [
    (1000, 300, 30), (1000, 300, 1), (1000, 300, 8), (1000, 300, 12), (1000, 150, 30), (1000, 64, 30), (1000, 256, 30),
    (40, 300, 30), (2000, 300, 30)]
dev = torch.device("cuda", 0)
hooks = [{}, {"VDET_NMS_THREADS": "256"}, {"VDET_NMS_THREADS": "320"}]
rows = []
for T, N, C in shapes:
    sets = []
    for k in range(4):
        b, s = synth.boxes_scores(T, N, C, seed=2000 + k)
        sets.append((torch.from_numpy(b.reshape(-1, 4)).to(dev), torch.from_numpy(s.reshape(-1, C)).to(dev)))
    seg = ops.seg_offsets_uniform(T, N, dev)
    st = ops.new_status(dev)
    ref = None
    rec = {"T": T, "N": N, "C": C}
    for h in hooks:
        os.environ.pop("VDET_NMS_THREADS", None)
        os.environ.update(h)
        out = None
        for k in range(3):
            out = ops.nms_frames(sets[k % 4][0], sets[k % 4][1], seg, 0.3, N, want_mask=True, status=st, frame_major_out=True)
        torch.cuda.synchronize()
        reps = 24
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k in range(reps):
            ops.nms_frames(sets[k % 4][0], sets[k % 4][1], seg, 0.3, N, want_mask=True, status=st, frame_major_out=True, out=out[:3])
        b_.record()
        torch.cuda.synchronize()
        ops.nms_frames(sets[0][0], sets[0][1], seg, 0.3, N, want_mask=True, status=st, frame_major_out=True, out=out[:3])
        torch.cuda.synchronize()
        sig = (out[0].clone(), out[1].clone())
        if ref is None:
            ref = sig
        same = bool(torch.equal(sig[1], ref[1])) and bool(torch.equal(sig[0], ref[0]))
        name = ",".join("%s=%s" % (k[9:].lower(), v) for k, v in sorted(h.items())) or "default"
        rec[name] = round(a.elapsed_time(b_) / reps, 4)
        rec["same"] = rec.get("same", True) and same
    rows.append(rec)
    print(json.dumps(rec), flush=True)
