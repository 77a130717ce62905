#!/usr/bin/env python
"""Where does the end-to-end step go?  Stage timestamps of VideoPostProcessor.run_staged's pipeline
(re-implemented here with timing events), raw copy rates in the same process, and variants.

    python tools/e2e_probe.py -> gpurun_out/e2e_probe.json
"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from vdetlib_b200 import ops, synth
from vdetlib_b200.vdet.video_det import VideoPostProcessor

T, N, C = 1000, 300, 30
dev = torch.device("cuda", 0)
b, s = synth.boxes_scores(T, N, C, seed=3)
out = {}


def ev():
    return torch.cuda.Event(enable_timing=True)


def staged_timeline(pp, reps=30):
    """One pipelined step with timestamps (ms after the step's start) of: boxes on device, all scores
    on device, link done, last NMS done, all results on host."""
    cur = torch.cuda.current_stream()
    rows = []
    for _ in range(reps):
        t0 = ev(); t0.record(cur)
        pp.s_in.wait_stream(cur); pp.s_out.wait_stream(cur)
        ev_in = []
        with torch.cuda.stream(pp.s_in):
            pp.d_boxes.copy_(pp.h_boxes, non_blocking=True)
            e_boxes = ev(); e_boxes.record(pp.s_in)
            for f0, f1 in pp.chunks:
                r0, r1 = f0 * N, f1 * N
                pp.d_scores[r0:r1].copy_(pp.h_scores[r0:r1], non_blocking=True)
                e = ev(); e.record(pp.s_in); ev_in.append(e)
        cur.wait_event(e_boxes)
        ops.link_frames(pp.d_boxes, pp.seg_offsets, N, None, out=(pp.d_succ, pp.d_iou))
        e_link = ev(); e_link.record(cur)
        pp.s_out.wait_event(e_link)
        with torch.cuda.stream(pp.s_out):
            pp.h_succ.copy_(pp.d_succ, non_blocking=True); pp.h_iou.copy_(pp.d_iou, non_blocking=True)
        for k, (f0, f1) in enumerate(pp.chunks):
            r0, r1 = f0 * N, f1 * N
            cur.wait_event(ev_in[k])
            ops.nms_frames(pp.d_boxes[r0:r1], pp.d_scores[r0:r1], pp.chunk_seg[f1 - f0], 0.3, N, want_mask=True,
                           status=pp.status, frame_major_out=True,
                           out=(pp.d_idx[r0 * C:r1 * C], pp.d_cnt[f0:f1], pp.d_mask[r0 * C:r1 * C]))
            e_n = ev(); e_n.record(cur)
            pp.s_out.wait_event(e_n)
            with torch.cuda.stream(pp.s_out):
                pp.h_mask[r0 * C:r1 * C].copy_(pp.d_mask[r0 * C:r1 * C], non_blocking=True)
                pp.h_cnt[f0:f1].copy_(pp.d_cnt[f0:f1], non_blocking=True)
        e_out = ev(); e_out.record(pp.s_out)
        t_issue = time.perf_counter()
        pp.s_out.synchronize()
        rows.append([t0.elapsed_time(e_boxes), t0.elapsed_time(ev_in[-1]), t0.elapsed_time(e_link),
                     t0.elapsed_time(e_n), t0.elapsed_time(e_out)])
    return [round(float(x), 3) for x in np.median(np.asarray(rows), axis=0)]


def loop_ms(fn, reps):
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for nch in (8, 1, 4, 16):
    pp = VideoPostProcessor(T, N, C, 0.3, dev, n_chunks=nch)
    pp.stage(b, s)
    for _ in range(100):
        pp.run_staged()
    key = "chunks_%d" % nch
    out[key] = {"step_ms": round(loop_ms(pp.run_staged, 50), 3)}
    out[key]["timeline_ms[boxes_in, scores_in, link, last_nms, all_out]"] = staged_timeline(pp)
    # host-side cost of issuing one step (no waiting): time until everything is enqueued
    t = time.perf_counter()
    for _ in range(20):
        pp.run_staged()
    out[key]["wall_ms"] = round((time.perf_counter() - t) / 20 * 1e3, 3)
    if nch == 8:
        pp8 = pp

pp = pp8
# raw copy rates in this process, right after the steps
out["raw_h2d_scores_ms"] = round(loop_ms(lambda: pp.d_scores.copy_(pp.h_scores, non_blocking=True), 20), 3)
out["raw_h2d_scores_GBs"] = round(pp.h_scores.numel() * 4 / out["raw_h2d_scores_ms"] / 1e6, 1)
out["raw_d2h_mask_ms"] = round(loop_ms(lambda: pp.h_mask.copy_(pp.d_mask, non_blocking=True), 20), 3)
out["raw_d2h_mask_GBs"] = round(pp.h_mask.numel() / out["raw_d2h_mask_ms"] / 1e6, 1)

# same step with a spinning host thread (does host-side power management matter?)
stop = False


def spin():
    x = 0
    while not stop:
        x += 1


th = [threading.Thread(target=spin) for _ in range(2)]
for t_ in th:
    t_.start()
for _ in range(50):
    pp.run_staged()
out["chunks_8_with_spinning_threads_step_ms"] = round(loop_ms(pp.run_staged, 50), 3)
stop = True
for t_ in th:
    t_.join()

# kernels only (inputs resident)
out["device_only_step_ms"] = round(loop_ms(lambda: pp.run_device(pp.d_boxes, pp.d_scores), 50), 3)
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "e2e_probe.json"), "w"), indent=1)
