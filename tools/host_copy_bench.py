#!/usr/bin/env python
"""Staging copy into the upload buffer (vdet_host_copy_stream_mt): streaming stores over 1..8 host threads
against numpy.copyto, for one config-2 shard (40.8 MB).  CPU only.
    python tools/host_copy_bench.py > profiles/rNN_host_copy.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vdetlib_b200 import _lib          # noqa: E402

lib = _lib.load()
n = 40_800_000
src = np.random.default_rng(0).integers(0, 256, n, dtype=np.uint8)
dst = np.empty(n, np.uint8)


def best_ms(fn, reps=12):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


out = {"bytes": n, "cpu_count": os.cpu_count(), "ms": {}}
for th in (1, 2, 4, 8, 0):
    ms = best_ms(lambda: lib.vdet_host_copy_stream_mt(dst.ctypes.data, src.ctypes.data, n, th))
    out["ms"]["auto" if th == 0 else "%d threads" % th] = {"ms": round(ms, 3), "GBs": round(n / ms / 1e6, 1)}
ms = best_ms(lambda: np.copyto(dst, src))
out["ms"]["numpy.copyto (ordinary stores)"] = {"ms": round(ms, 3), "GBs": round(n / ms / 1e6, 1)}
assert np.array_equal(dst, src)
print(json.dumps(out, indent=1))
