#!/usr/bin/env python
"""Where does the sharded device step go?  Under torchrun (N ranks): host enqueue time per step
(perf_counter around step_device, no synchronisation) against device time per step (CUDA events),
and the same for the pieces (exchange only, NMS only, link only).  Rank 0 prints one JSON object.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_probe.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vdetlib_b200 import ops, synth                                   # noqa: E402
from vdetlib_b200.dist import ShardedVideoPostProcessor               # noqa: E402

rank, local, world = (int(os.environ.get(k, "0")) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
world = max(world, 1)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
T, N, C, K = 1000, 300, 30, 60
sets = []
for k in range(4):
    b, s = synth.boxes_scores(T, N, C, seed=2000 + 100 * rank + k)
    sets.append((torch.from_numpy(b.reshape(-1, 4)).to(dev), torch.from_numpy(s.reshape(-1, C)).to(dev)))
pp = ShardedVideoPostProcessor(T, N, C, 0.3, dev)
p = pp.pp


def measure(fn):
    for k in range(5):
        fn(k)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for k in range(K):
        fn(k)
    b_.record()
    host = (time.perf_counter() - t0) * 1e3 / K
    torch.cuda.synchronize()
    return {"host_enqueue_ms": round(host, 4), "device_ms": round(a.elapsed_time(b_) / K, 4)}


out = {"world": world, "rank": rank}
out["step_device"] = measure(lambda k: pp.step_device(*sets[k % 4]))
out["nms_only"] = measure(lambda k: ops.nms_frames(sets[k % 4][0], sets[k % 4][1], p.seg_offsets, 0.3, N, want_mask=True,
                                                   status=p.status, frame_major_out=True, out=(p.d_idx, p.d_cnt, p.d_mask)))
out["link_only"] = measure(lambda k: ops.link_frames(sets[k % 4][0], p.seg_offsets, N, None, out=(p.d_succ, p.d_iou)))
if world > 1:
    def exch(k):
        pp._exchange(sets[k % 4][0][:N])
        torch.cuda.current_stream().wait_stream(pp.side)
    out["exchange_only"] = measure(exch)
else:
    out["step_device_eager"] = measure(lambda k: p.run_device(sets[k % 4][0], sets[k % 4][1], None, graph=False))
if rank == 0:
    print(json.dumps(out, indent=1))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
