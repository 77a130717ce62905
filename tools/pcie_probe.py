#!/usr/bin/env python
"""Pinned host<->device copy bandwidth of this box (the bound of bench.py's e2e figure).

    python tools/pcie_probe.py  -> gpurun_out/pcie.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda", 0)
out = {}


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for mb in (1, 4, 16, 41, 128):
    n = mb * 1000 * 1000
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h2d = timed(lambda: d.copy_(h, non_blocking=True))
    d2h = timed(lambda: h.copy_(d, non_blocking=True))
    s2 = torch.cuda.Stream()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)

    def both():
        s2.wait_stream(torch.cuda.current_stream())
        d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s2)

    bi = timed(both)
    out["%dMB" % mb] = {"h2d_GBs": n / h2d / 1e6, "d2h_GBs": n / d2h / 1e6, "bidir_each_GBs": n / bi / 1e6,
                        "h2d_ms": h2d, "d2h_ms": d2h}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "pcie.json"), "w"), indent=1)
