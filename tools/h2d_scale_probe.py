#!/usr/bin/env python
"""Host<->device ceiling of the box at N = 1/2/4/8 ranks, without any kernel.

    python tools/h2d_scale_probe.py                                   # 1 rank
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/h2d_scale_probe.py [--bind]

Every rank drives its own GPU (LOCAL_RANK) with the byte counts of one configs[1] step
(40.8 MB up: boxes + scores; DOWN = the bytes the step's results take) and reports GB/s per rank;
rank 0 prints one JSON line with the per-rank numbers, the aggregate (total bytes / slowest rank)
and the efficiency against N x the 1-rank figure the caller passes with --ref (optional).

Legs (each: 5 warm-up rounds, then `--reps` rounds between barriers):
  h2d         cudaMemcpyAsync pinned -> device, one stream
  d2h         device -> pinned
  duplex      both directions at once on two streams
  stage       pageable -> pinned with the library's streaming multi-thread copy (no GPU at all)
  stage+h2d   what a real producer costs: stage shard k+1 on the host threads while shard k uploads
`--bind`: pin the rank to its own slice of the CPUs (os.sched_setaffinity) BEFORE the pinned
buffers are allocated and first touched, so that their pages land on that slice's memory node.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np            # noqa: E402
import torch                  # noqa: E402
import torch.distributed as dist      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", action="store_true")
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--up-mb", type=float, default=40.8)
    ap.add_argument("--down-mb", type=float, default=8.7)
    ap.add_argument("--threads", type=int, default=0, help="staging threads per rank (0: cpus/ranks, max 8)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpus = sorted(os.sched_getaffinity(0))
    my_cpus = cpus
    if args.bind and world > 1:
        per = max(len(cpus) // world, 1)
        my_cpus = cpus[local * per:(local + 1) * per] or cpus
        os.sched_setaffinity(0, my_cpus)
    threads = args.threads or max(1, min(8, len(cpus) // world))
    if world > 1:
        dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from vdetlib_b200 import _lib
    lib = _lib.load()

    up = int(args.up_mb * 1e6) // 64 * 64
    down = int(args.down_mb * 1e6) // 64 * 64
    h_up = [torch.empty(up, dtype=torch.uint8).pin_memory() for _ in range(2)]
    h_down = torch.empty(down, dtype=torch.uint8).pin_memory()
    for t in h_up:
        t.fill_(1)                       # first touch after the (optional) binding
    h_down.fill_(2)
    d_up = torch.empty(up, dtype=torch.uint8, device=dev)
    d_down = torch.zeros(down, dtype=torch.uint8, device=dev)
    src = [np.random.default_rng(k).integers(0, 255, up, dtype=np.uint8) for k in range(3)]   # pageable "producer" shards
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def stage(k, slot):
        lib.vdet_host_copy_stream_mt(h_up[slot].data_ptr(), src[k % 3].ctypes.data, up, threads)

    def leg_h2d(n):
        with torch.cuda.stream(s1):
            for _ in range(n):
                d_up.copy_(h_up[0], non_blocking=True)
        return up * n, 0

    def leg_d2h(n):
        with torch.cuda.stream(s2):
            for _ in range(n):
                h_down.copy_(d_down, non_blocking=True)
        return 0, down * n

    def leg_duplex(n):
        leg_h2d(n)
        leg_d2h(n)
        return up * n, down * n

    def leg_stage(n):
        for k in range(n):
            stage(k, k & 1)
        return up * n, 0

    def leg_stage_h2d(n):
        # slot k&1 is staged while the other slot uploads; an event per slot says when its upload is done
        evs = [torch.cuda.Event(), torch.cuda.Event()]
        for k in range(n):
            slot = k & 1
            if k >= 2:
                evs[slot].synchronize()
            stage(k, slot)
            with torch.cuda.stream(s1):
                d_up.copy_(h_up[slot], non_blocking=True)
                evs[slot].record(s1)
            with torch.cuda.stream(s2):
                h_down.copy_(d_down, non_blocking=True)
        return up * n, down * n

    def leg_stage_thread_h2d(n):
        # the same with the staging on a second Python thread (ctypes releases the GIL): the main thread only enqueues
        ready = [threading.Semaphore(0), threading.Semaphore(0)]
        free = [threading.Semaphore(1), threading.Semaphore(1)]

        def producer():
            for k in range(n):
                slot = k & 1
                free[slot].acquire()
                stage(k, slot)
                ready[slot].release()
        th = threading.Thread(target=producer)
        th.start()
        evs = [None, None]
        for k in range(n):
            slot = k & 1
            ready[slot].acquire()
            with torch.cuda.stream(s1):
                d_up.copy_(h_up[slot], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s1)
            with torch.cuda.stream(s2):
                h_down.copy_(d_down, non_blocking=True)
            if evs[slot ^ 1] is not None:
                evs[slot ^ 1].synchronize()
                free[slot ^ 1].release()
                evs[slot ^ 1] = None
            evs[slot] = ev
        for slot in (0, 1):
            if evs[slot] is not None:
                evs[slot].synchronize()
                free[slot].release()
        th.join()
        return up * n, down * n

    cudart = torch.cuda.cudart()

    def leg_register(n):
        # pin the producer's pageable shard in place (cudaHostRegister), upload from it, unpin: no staging copy at all
        for k in range(n):
            a = src[k % 3]
            rc = cudart.cudaHostRegister(a.ctypes.data, up, 0)
            assert int(rc) == 0, rc
            t = torch.from_numpy(a)
            with torch.cuda.stream(s1):
                d_up.copy_(t, non_blocking=True)
            s1.synchronize()
            cudart.cudaHostUnregister(a.ctypes.data)
        return up * n, 0

    def leg_register_only(n):
        for k in range(n):
            a = src[k % 3]
            cudart.cudaHostRegister(a.ctypes.data, up, 0)
            cudart.cudaHostUnregister(a.ctypes.data)
        return up * n, 0

    legs = [("register_only", leg_register_only), ("register+h2d", leg_register), ("h2d", leg_h2d), ("d2h", leg_d2h), ("duplex", leg_duplex), ("stage", leg_stage),
            ("stage+h2d", leg_stage_h2d), ("stage_thread+h2d", leg_stage_thread_h2d)]
    out = {}
    for name, fn in legs:
        fn(5)
        barrier()
        t0 = time.perf_counter()
        bu, bd = fn(args.reps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        rec = {"ms_per_round": 1e3 * dt / args.reps, "up_GBs": bu / dt / 1e9, "down_GBs": bd / dt / 1e9}
        if world > 1:
            allr = [None] * world
            dist.all_gather_object(allr, rec)
        else:
            allr = [rec]
        slow = max(r["ms_per_round"] for r in allr)
        out[name] = {"ms_per_round_max": round(slow, 4),
                     "ms_per_round_by_rank": [round(r["ms_per_round"], 4) for r in allr],
                     "aggregate_up_GBs": round(world * (bu / args.reps) / (slow / 1e3) / 1e9, 2),
                     "aggregate_down_GBs": round(world * (bd / args.reps) / (slow / 1e3) / 1e9, 2),
                     "per_rank_up_GBs": [round(r["up_GBs"], 2) for r in allr]}
    if rank == 0:
        print(json.dumps({"probe": "h2d_scale", "ranks": world, "bind": bool(args.bind), "cpus_visible": len(cpus),
                          "cpus_per_rank": len(my_cpus), "stage_threads": threads, "up_bytes": up, "down_bytes": down,
                          "reps": args.reps, "legs": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
