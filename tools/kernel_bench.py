#!/usr/bin/env python
"""Per-kernel timings at the BASELINE config sizes -> gpurun_out/kernels.json (+ a markdown table).

Every row of SURVEY 8a gets a measured number against its roofline: CUDA events around `reps`
back-to-back launches after warm-up, on buffers larger than L2 or rotated so that reuse distance
exceeds the 126 MB L2.  Algorithmic bytes follow SURVEY 8d.  Not a bench.py replacement: this is
the evidence file for DESIGN.md section 5.

    python tools/kernel_bench.py [--quick]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from vdetlib_b200 import _lib, ops, synth

DEV = torch.device("cuda", 0)
QUICK = "--quick" in sys.argv
try:
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    HBM_SRC = "measured"
except Exception:
    HBM, HBM_SRC = 6650.0, "fallback"

rows = []


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(reps):
        fn(k)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def add(name, config, ms, alg_bytes, units=None, unit_name=None, note=""):
    gbs = alg_bytes / (ms / 1e3) / 1e9
    r = {"kernel": name, "config": config, "ms": ms, "algorithmic_bytes": alg_bytes, "achieved_gbs": gbs,
         "hbm_frac": gbs / HBM, "note": note}
    if units:
        r[unit_name + "_per_s"] = units / (ms / 1e3)
    rows.append(r)
    print("%-28s %-46s %9.3f ms %9.1f GB/s (%5.1f%% of HBM) %s" %
          (name, config, ms, gbs, 100 * gbs / HBM, ("%.3e %s/s" % (units / (ms / 1e3), unit_name)) if units else ""))


def sets_of(T, N, C, nsets, seed):
    out = []
    for k in range(nsets):
        b, s = synth.boxes_scores(T, N, C, seed=seed + k)
        out.append((torch.from_numpy(b.reshape(-1, 4)).to(DEV), torch.from_numpy(s.reshape(-1, C)).to(DEV)))
    return out


# ---- NMS: config 2 (1000 x 300 x 30) and config 5 per GPU (2000 frames x 2000 boxes x 30) ----
T, N, C = 1000, 300, 30
S = sets_of(T, N, C, 4, 100)
seg = ops.seg_offsets_uniform(T, N, DEV)
st = ops.new_status(DEV)
ms = timeit(lambda k: ops.nms_frames(S[k % 4][0], S[k % 4][1], seg, 0.3, N, want_mask=True, status=st))
add("nms_frames_kernel<16>", "C2 1000fr x 300 x 30cls thr0.3", ms, T * N * (16 + 4 * C) + T * N * C * 5 + 4 * T * C,
    T * N, "boxes", "issue/ALU bound")
ms = timeit(lambda k: ops.link_frames(S[k % 4][0], seg, N))
add("link_frames_kernel", "C2 1000fr x 300", ms, T * N * 24, T * N, "boxes", "FP32 issue bound")
ms1 = timeit(lambda k: ops.nms_frames(S[k % 4][0], S[k % 4][1][:, 0].contiguous(), seg, 0.3, N, status=st))
add("nms_frames_kernel<16>", "C2 boxes, 1 class (apply_vid_nms shape)", ms1, T * N * 20 + T * N * 4 + 4 * T, T * N, "boxes")
del S

T5, N5 = (200, 2000) if QUICK else (500, 2000)
S5 = sets_of(T5, N5, C, 2, 200)
seg5 = ops.seg_offsets_uniform(T5, N5, DEV)
ms = timeit(lambda k: ops.nms_frames(S5[k % 2][0], S5[k % 2][1], seg5, 0.3, N5, want_mask=True, status=st), reps=5, warm=1)
add("nms_frames_big_kernel", "C5 shard %dfr x 2000 x 30cls" % T5, ms, T5 * N5 * (16 + 4 * C) + T5 * N5 * C * 5, T5 * N5, "boxes",
    "N^2 pair evals dominate")
ms = timeit(lambda k: ops.link_frames(S5[k % 2][0], seg5, N5), reps=5, warm=1)
add("link_frames_kernel", "C5 shard %dfr x 2000" % T5, ms, T5 * N5 * 24, T5 * N5, "boxes")
del S5

# ---- link: config 3 per GPU (2500 frames x 1000 boxes) ----
T3, N3 = (500, 1000) if QUICK else (2500, 1000)
b3, _ = synth.boxes_scores(T3, N3, 1, seed=300)
d3 = torch.from_numpy(b3.reshape(-1, 4)).to(DEV)
seg3 = ops.seg_offsets_uniform(T3, N3, DEV)
ms = timeit(lambda k: ops.link_frames(d3, seg3, N3), reps=5, warm=1)
add("link_frames_kernel", "C3 shard %dfr x 1000" % T3, ms, T3 * N3 * 24, T3 * N3, "boxes", "%.2e pair IoU/s" % (T3 * N3 * N3 / (ms / 1e3)))
del d3

# ---- IoU matrix ----
A = 8192 if QUICK else 16384
bb, _ = synth.boxes_scores(1, A, 1, seed=77)
xa = torch.from_numpy(bb[0]).to(DEV)
mat = torch.empty((A, A), dtype=torch.float32, device=DEV)
ms = timeit(lambda k: ops.iou_matrix(xa, xa, out=mat), reps=10)
add("iou_matrix_f32_kernel", "%d x %d f32" % (A, A), ms, 4 * A * A + 32 * A, A * A, "pairs", "HBM-write bound (BASELINE's IoU kernel)")
A2 = A - 3
mat2 = torch.empty((A2, A2), dtype=torch.float32, device=DEV)
ms = timeit(lambda k: ops.iou_matrix(xa[:A2], xa[:A2], out=mat2), reps=10)
add("iou_matrix_f32_kernel", "%d x %d f32 (row pitch not /4: scalar stores)" % (A2, A2), ms, 4 * A2 * A2, A2 * A2, "pairs")
del mat, mat2
A64 = A // 2
xa64 = xa[:A64].double()
mat64 = torch.empty((A64, A64), dtype=torch.float64, device=DEV)
ms = timeit(lambda k: ops.iou_matrix(xa64, xa64, out=mat64), reps=5)
add("iou_matrix_f64_kernel", "%d x %d f64 (utils.common.iou)" % (A64, A64), ms, 8 * A64 * A64, A64 * A64, "pairs", "FP64 issue bound")
del mat64

# ---- temporal: config 4 per GPU (256 tubelets x 30 classes x 10000 frames) ----
K, L = (32 * 30, 10000) if QUICK else (256 * 30, 10000)
x = torch.from_numpy(synth.score_rows(64, L, seed=4, missing_frac=0.05)).to(DEV).repeat(K // 64, 1).contiguous()
work = x.clone()
def comp(k):
    work.copy_(x)
    ops.score_completion_(work, status=st)
ms_copy = timeit(lambda k: work.copy_(x), reps=10)
ms = timeit(comp, reps=10) - ms_copy
add("score_completion_kernel", "C4 shard %d rows x %d f32" % (K, L), ms, 8 * K * L, K * L, "scores", "HBM streaming; copy time subtracted")
ops.score_completion_(work, status=st)
out = torch.empty_like(work)
for w in (3, 9):
    ms = timeit(lambda k: ops.temporal_maxpool(work, w, out=out), reps=10)
    add("temporal_maxpool_kernel", "C4 shard %d x %d f32 w=%d" % (K, L, w), ms, 8 * K * L, K * L, "scores", "HBM streaming")
taps = torch.from_numpy(synth.gaussian_taps(30, 9)).to(DEV)
ms = timeit(lambda k: ops.temporal_conv1d(work, taps, "zero", out=out), reps=10)
add("temporal_conv1d_kernel", "C4 shard %d x %d f32 w=9" % (K, L), ms, 8 * K * L, K * L, "scores", "HBM streaming")
ms = timeit(lambda k: out.copy_(work), reps=10)
add("(torch copy_ reference)", "%d x %d f32" % (K, L), ms, 8 * K * L, K * L, "scores", "what a plain device copy reaches")
del x, work, out

# ---- spatial max-pool, top-k, drop-in entry points (latency dominated) ----
Tm, Nm, P = 1000, 300, 20000
bm, sm = synth.boxes_scores(Tm, Nm, 1, seed=9)
rng = np.random.default_rng(1)
tub = bm.reshape(-1, 4)[rng.integers(0, Tm * Nm, P)].astype(np.float64) + rng.integers(-4, 5, (P, 4))
tseg = torch.from_numpy((rng.integers(0, Tm, P)).astype(np.int32)).to(DEV)
dbm = torch.from_numpy(bm.reshape(-1, 4).astype(np.float64)).to(DEV)
dsm = torch.from_numpy(sm.reshape(-1).astype(np.float64)).to(DEV)
dt = torch.from_numpy(tub).to(DEV)
segm = ops.seg_offsets_uniform(Tm, Nm, DEV)
ms = timeit(lambda k: ops.spatial_maxpool(dt, tseg, dbm, dsm, segm, 0.7), reps=10)
add("spatial_maxpool_kernel", "%d tubelet boxes vs 300 dets/frame, f64" % P, ms, P * 44 + Tm * Nm * 40, P * Nm, "pairs", "FP64 issue bound")
R, Ck = 300, 31
sc = torch.rand((Tm * R, Ck), device=DEV) * 0.4
ms = timeit(lambda k: ops.threshold_topk(sc, ops.seg_offsets_uniform(Tm, R, DEV), R, 0.05, 100), reps=10)
add("threshold_topk_kernel", "1000fr x 300 rows x 31cls k=100", ms, Tm * R * Ck * 4 + Tm * Ck * 404, Tm * R, "rows")
d1 = torch.from_numpy(np.concatenate([bm[0], sm[0]], 1)).to(DEV)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    ops.nms(d1, 0.5)
torch.cuda.synchronize()
rows.append({"kernel": "vdet_nms_f32 (sync entry)", "config": "C1: 300 boxes, 1 class, thr 0.5", "ms": (time.perf_counter() - t0) / 50 * 1e3,
             "note": "host wall clock per call incl. launch + D2H of the count; reference Cython: ~0.8-1.3 ms"})
print("%-28s %-46s %9.3f ms (wall, per call)" % (rows[-1]["kernel"], rows[-1]["config"], rows[-1]["ms"]))

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"hbm_peak_gbs": HBM, "hbm_peak_source": HBM_SRC, "gpu": torch.cuda.get_device_name(0), "rows": rows},
          open(os.path.join(ROOT, "gpurun_out", "kernels.json"), "w"), indent=1)
