"""Multi-process host logic on CPU: gloo backend, world_size 2 (and 3), 127.0.0.1 rendezvous."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_all_frames():
    from vdetlib_b200.dist import shard_range
    for n in (0, 1, 7, 1000, 5000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vdetlib_b200.dist import BoundaryExchange, shard_range
    # a 10-frame video with ragged frames; every rank knows the whole thing for checking only
    rng = np.random.default_rng(0)
    counts = [5, 3, 8, 1, 6, 2, 7, 4, 3, 8]
    frames = [rng.uniform(0, 100, (c, 4)).astype(np.float32) for c in counts]
    a, b = shard_range(len(frames), world, rank)
    ex = BoundaryExchange(max_boxes=8, device="cpu")
    h = ex.start(torch.from_numpy(frames[a]))
    halo = ex.finish(h)
    if rank == world - 1:
        ok = halo is None
    else:
        ok = halo is not None and np.array_equal(halo.numpy(), frames[b])
    # uniform-count fast path (no count read-back)
    h = ex.start(torch.from_numpy(frames[a]))
    halo2 = ex.finish(h, count_hint=None if rank == world - 1 else counts[b])
    ok = ok and (halo2 is None if rank == world - 1 else np.array_equal(halo2.numpy(), frames[b]))
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write("ok" if ok else "bad")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_boundary_exchange_gloo(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"
