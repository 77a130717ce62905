"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the frame-sharded pipeline with the NCCL
boundary all-gather equals the single-GPU result on the whole video.  Also a single-GPU
"virtual shard" check that exercises the same halo logic with one device."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from vdetlib_b200 import ops, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _full_reference(b, s, thr):
    T, N, C = s.shape
    dev = torch.device("cuda", 0)
    db = torch.from_numpy(b.reshape(-1, 4)).to(dev)
    ds = torch.from_numpy(s.reshape(-1, C)).to(dev)
    seg = ops.seg_offsets_uniform(T, N, dev)
    _, cnt, mask, st = ops.nms_frames(db, ds, seg, thr, N, want_mask=True)
    succ, iou = ops.link_frames(db, seg, N)
    ops.raise_for_status(st)
    return cnt.cpu().numpy(), mask.cpu().numpy(), succ.cpu().numpy(), iou.cpu().numpy()


def test_virtual_shards_single_gpu():
    """S logical shards on one device: link of every shard with the next shard's first frame as halo."""
    T, N, C, thr, S = 24, 100, 3, 0.3, 4
    b, s = synth.boxes_scores(T, N, C, seed=12)
    cnt, mask, succ, iou = _full_reference(b, s, thr)
    dev = torch.device("cuda", 0)
    per = T // S
    for k in range(S):
        sb = torch.from_numpy(b[k * per:(k + 1) * per].reshape(-1, 4)).to(dev)
        seg = ops.seg_offsets_uniform(per, N, dev)
        halo = torch.from_numpy(b[(k + 1) * per]).to(dev) if k < S - 1 else None
        su, io = ops.link_frames(sb, seg, N, halo, halo_row_base=per * N)
        want = succ[k * per * N:(k + 1) * per * N].astype(np.int64)
        # inside the shard successors are shard-local rows; across the boundary they are rows + halo index,
        # i.e. the global successor minus the shard's first row again -- one formula for both
        local = np.where(want >= 0, want - k * per * N, -1)
        assert np.array_equal(su.cpu().numpy().astype(np.int64), local)
        # chains followed on the shard end at the boundary instead of wrapping into the shard's first frame
        start = torch.arange((per - 2) * N, (per - 2) * N + 8, dtype=torch.int32, device=dev)
        rows = ops.follow_links(su, io, start, 3).cpu().numpy()
        assert np.array_equal(rows[1], local[(per - 2) * N:(per - 2) * N + 8]) and np.all(rows[2] == -1)
        assert np.array_equal(io.cpu().numpy(), iou[k * per * N:(k + 1) * per * N])


def _worker(rank, world, port, T, N, C, thr, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from vdetlib_b200.dist import ShardedVideoPostProcessor, shard_range
    b, s = synth.boxes_scores(T, N, C, seed=12)
    a, e = shard_range(T, world, rank)
    pp = ShardedVideoPostProcessor(e - a, N, C, thr, dev)
    res = pp.step_host(b[a:e], s[a:e])
    res = {"keep_mask": res.keep_mask(), "keep_cnt": np.array(res["keep_cnt"]), "succ": np.array(res["succ"]),
           "link_iou": np.array(res["link_iou"]), "keep_idx": np.array(res["keep_idx"]), "keep_off": np.array(res["keep_off"])}
    # the device-resident step (what bench.py's `value` times) on the same shard must give the same arrays
    dev_res = pp.step_device(pp.pp.d_boxes, pp.pp.d_scores)
    torch.cuda.synchronize()
    for key in ("keep_mask", "keep_cnt", "succ", "link_iou"):
        assert np.array_equal(dev_res[key].cpu().numpy().reshape(res[key].shape), res[key]), key
    # two steps in flight through the boundary exchange: new shards from pageable memory, eager and graph replay
    for graph in (False, True):
        t0 = pp.submit_host(b[a:e], s[a:e], graph=graph)
        t1 = pp.submit_host(b[a:e].copy(), s[a:e].copy(), graph=graph)
        r0 = pp.collect(t0)
        r0 = {k: np.array(r0[k], copy=True) for k in ("keep_idx", "keep_off", "succ", "link_iou")}
        r1 = pp.collect(t1)
        for key in ("keep_idx", "keep_off", "succ", "link_iou"):
            assert np.array_equal(r0[key], res[key]) and np.array_equal(np.asarray(r1[key]), res[key]), (key, graph)
    # frame-sharded vid_nms of one class with the global keep-order merge
    from vdetlib_b200.dist import sharded_vid_nms
    s_glob = s[:, :, 0] + np.arange(T)[:, None] * 1e-5                      # unique across the video
    dets = np.concatenate([np.repeat(np.arange(1, T + 1), N)[:, None], b.reshape(-1, 4), s_glob.reshape(-1, 1)],
                          axis=1).astype(np.float32)
    merged = sharded_vid_nms(torch.from_numpy(dets[a * N:e * N]).to(dev), 0.3, a * N)
    res = dict(res)
    res["vid_keep"] = merged.cpu().numpy()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{k: np.array(v) for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_pipeline_matches_single_gpu(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    T, N, C, thr = 8 * world, 300, 30, 0.3          # equal shards: every rank holds 8 frames
    b, s = synth.boxes_scores(T, N, C, seed=12)
    cnt, mask, succ, iou = _full_reference(b, s, thr)
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    mp.spawn(_worker, args=(world, port, T, N, C, thr, str(tmp_path)), nprocs=world, join=True)
    per = T // world
    from oracle import c_oracle
    s_glob = s[:, :, 0] + np.arange(T)[:, None] * 1e-5
    dets = np.concatenate([np.repeat(np.arange(1, T + 1), N)[:, None], b.reshape(-1, 4), s_glob.reshape(-1, 1)],
                          axis=1).astype(np.float32)
    want_keep = c_oracle.vid_nms(dets, 0.3)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert got["vid_keep"].tolist() == want_keep                            # same global order on every rank
        lo, hi = r * per * N, (r + 1) * per * N
        assert np.array_equal(got["keep_cnt"], cnt[:, r * per:(r + 1) * per].T)                  # [frames, C]
        assert np.array_equal(got["keep_mask"], mask[:, lo:hi].reshape(C, per, N).transpose(1, 0, 2))
        assert np.array_equal(got["link_iou"], iou[lo:hi])
        want = succ[lo:hi].astype(np.int64)
        local = np.where(want >= 0, want - lo, -1)          # boundary successors: rows + halo index == global - lo
        assert np.array_equal(got["succ"].astype(np.int64), local)
