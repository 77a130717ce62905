#!/usr/bin/env python
"""Why is a lone 41 MB pinned H2D copy slow on this box?  Variants of the same copy, per-iteration times.

    python tools/pcie_probe2.py  -> gpurun_out/pcie2.json
"""
import ctypes
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda", 0)
torch.cuda.init()
torch.zeros(1, device=dev)
rt = ctypes.CDLL("libcudart.so.12")
N = 40_800_000
out = {}


def series(fn, reps=12):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    return ts


def host_alloc(n, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags))
    assert rc == 0, rc
    return p


def memcpy_async(dst, src, n, kind, stream):
    rc = rt.cudaMemcpyAsync(ctypes.c_void_p(dst), ctypes.c_void_p(src), ctypes.c_size_t(n), ctypes.c_int(kind),
                            ctypes.c_void_p(stream))
    assert rc == 0, rc


d = torch.empty(N, dtype=torch.uint8, device=dev)
d2 = torch.empty(N, dtype=torch.uint8, device=dev)
h = torch.empty(N, dtype=torch.uint8).pin_memory()
h.fill_(3)
h2 = torch.empty(N, dtype=torch.uint8).pin_memory()
cur = torch.cuda.current_stream()
s2, s3 = torch.cuda.Stream(), torch.cuda.Stream()

out["A_h2d_alone"] = series(lambda: d.copy_(h, non_blocking=True))


def chunks(k):
    step = (N + k - 1) // k
    for i in range(0, N, step):
        d[i:i + step].copy_(h[i:i + step], non_blocking=True)


out["B_h2d_10_chunks_one_stream"] = series(lambda: chunks(10))
out["B2_h2d_40_chunks_one_stream"] = series(lambda: chunks(40))


def two_streams():
    half = N // 2
    s2.wait_stream(cur)
    d[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2):
        d[half:].copy_(h[half:], non_blocking=True)
    cur.wait_stream(s2)


out["C_h2d_two_streams"] = series(two_streams)


def with_d2h():
    s2.wait_stream(cur)
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    cur.wait_stream(s2)


out["D_h2d_with_concurrent_d2h"] = series(with_d2h)
out["A2_h2d_alone_again"] = series(lambda: d.copy_(h, non_blocking=True))
out["E_d2h_alone"] = series(lambda: h2.copy_(d2, non_blocking=True))

# write-combined and plain cudaHostAlloc buffers through cudart directly
for name, flags in (("F_cudaHostAlloc_default", 0), ("G_cudaHostAlloc_writecombined", 4), ("H_cudaHostAlloc_portable_mapped", 3)):
    p = host_alloc(N, flags)
    ctypes.memset(p, 5, N)
    out[name] = series(lambda: memcpy_async(d.data_ptr(), p.value, N, 1, cur.cuda_stream))

# registered pageable memory
import numpy as np
arr = np.full(N, 7, dtype=np.uint8)
rc = rt.cudaHostRegister(ctypes.c_void_p(arr.ctypes.data), ctypes.c_size_t(N), ctypes.c_uint(0))
if rc == 0:
    out["I_cudaHostRegister"] = series(lambda: memcpy_async(d.data_ptr(), arr.ctypes.data, N, 1, cur.cuda_stream))

# a kernel pulling from mapped pinned memory (zero copy): torch can't wrap a raw host pointer as a CUDA
# tensor, so use the elementwise copy kernel on a tensor that aliases the mapped buffer via UVA
pm = host_alloc(N, 2)      # cudaHostAllocMapped
ctypes.memset(pm, 9, N)
dp = ctypes.c_void_p()
assert rt.cudaHostGetDevicePointer(ctypes.byref(dp), pm, ctypes.c_uint(0)) == 0
out["J_mapped_devptr_equals_hostptr"] = bool(dp.value == pm.value)
out["J_memcpy_d2d_kind_from_mapped"] = series(lambda: memcpy_async(d.data_ptr(), dp.value, N, 4, cur.cuda_stream))  # cudaMemcpyDefault

# host-side location of the buffers
try:
    out["numa_nodes"] = len([x for x in os.listdir("/sys/devices/system/node") if x.startswith("node")])
    out["cpu_count"] = os.cpu_count()
    out["affinity"] = len(os.sched_getaffinity(0))
except Exception as e:
    out["numa_err"] = str(e)
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "pcie2.json"), "w"), indent=1)
