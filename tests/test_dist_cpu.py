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
    ok = True
    for _ in range(2):                     # second round: the count word is only rewritten when it changes
        h = ex.start(torch.from_numpy(frames[a]))
        halo, halo_count = ex.finish(h)
        if rank == world - 1:
            ok = ok and halo is None and halo_count is None
        else:
            # a fixed-capacity buffer + the neighbour's count as an int32 device word (no host read-back needed)
            ok = ok and halo.shape == (8, 4) and halo_count.dtype == torch.int32 and int(halo_count[0]) == counts[b]
            ok = ok and np.array_equal(halo.numpy()[:counts[b]], frames[b])
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


def _temporal_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_np
    from vdetlib_b200.dist import ShardedTemporalRows, frame_sharded_window_op, shard_range
    rng = np.random.default_rng(1)
    rows = rng.uniform(-1, 1, (11, 53))                                  # every rank knows the whole block
    ok = True
    # by tubelet: local stage, then gather
    sh = ShardedTemporalRows(rows.shape[0])
    mine = np.stack([oracle_np.temporal_maxpool_row(r, 5) for r in sh.local(rows)]) if sh.stop > sh.start \
        else np.zeros((0, rows.shape[1]))
    full = sh.gather(torch.from_numpy(mine)).numpy()
    want = np.stack([oracle_np.temporal_maxpool_row(r, 5) for r in rows])
    ok = ok and np.array_equal(full, want)
    # by frame: halo exchange, then the window op on the haloed block
    a, b = shard_range(rows.shape[1], world, rank)
    local = torch.from_numpy(rows[:, a:b].copy())
    for w in (3, 9):
        h = w // 2

        def maxpool(ext):
            return torch.from_numpy(np.stack([oracle_np.temporal_maxpool_row(r, w) for r in ext.numpy()]))
        got = frame_sharded_window_op(local, h, maxpool, pad_value=-1e5).numpy()
        want = np.stack([oracle_np.temporal_maxpool_row(r, w) for r in rows])[:, a:b]
        ok = ok and np.array_equal(got, want)
        taps = rng.uniform(-1, 1, (1, w))

        def conv(ext, mode):
            return torch.from_numpy(oracle_np.temporal_conv1d(ext.numpy(), np.repeat(taps, ext.shape[0], 0), mode))
        for mode, pad in (("zero", 0.0), ("edge", None)):
            got = frame_sharded_window_op(local, h, lambda e: conv(e, mode), pad_value=pad).numpy()
            want = oracle_np.temporal_conv1d(rows, np.repeat(taps, rows.shape[0], 0), mode)[:, a:b]
            ok = ok and np.array_equal(got, want)
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write("ok" if ok else "bad")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_temporal_sharding_gloo(tmp_path, world):
    """BASELINE config 4's two splits: by tubelet (+ gather) and by frame with a halo of
    w//2 columns; the window op is the NumPy oracle, so the result must equal the unsharded one bit for bit."""
    port = _free_port()
    mp.spawn(_temporal_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"


def test_temporal_sharding_single_process():
    """Without a process group the helpers degenerate to the plain op / identity."""
    from oracle import oracle_np
    from vdetlib_b200.dist import ShardedTemporalRows, frame_sharded_window_op
    rows = np.random.default_rng(2).uniform(-1, 1, (4, 20))
    sh = ShardedTemporalRows(4)
    assert (sh.start, sh.stop) == (0, 4) and sh.gather(torch.from_numpy(rows)).numpy() is not None

    def maxpool(ext):
        return torch.from_numpy(np.stack([oracle_np.temporal_maxpool_row(r, 7) for r in ext.numpy()]))
    got = frame_sharded_window_op(torch.from_numpy(rows), 3, maxpool, pad_value=-1e5).numpy()
    assert np.array_equal(got, np.stack([oracle_np.temporal_maxpool_row(r, 7) for r in rows]))
    with pytest.raises(ValueError):
        frame_sharded_window_op(torch.from_numpy(rows[:, :2]), 3, maxpool, pad_value=-1e5)


def _completion_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kernel_double
    from oracle import oracle_np
    from vdetlib_b200 import ops
    from vdetlib_b200.dist import frame_sharded_completion_, shard_range
    ops.score_completion_ = kernel_double.score_completion_              # the kernel's stand-in on CPU tensors
    rng = np.random.default_rng(3)
    rows = rng.uniform(0, 1, (9, 47))                                    # every rank knows the whole block, for checking
    for r in range(9):
        for _ in range(6):
            a = rng.integers(0, 47)
            rows[r, a:a + rng.integers(1, 14)] = -1e5
    rows[1, :] = -1e5; rows[1, 20] = 0.25                                # one valid score: every shard but one is empty
    rows[2, :30] = -1e5                                                  # a leading run over several shards
    rows[3, 10:] = -1e5                                                  # a trailing run over several shards
    rows[4, :] = -1e5                                                    # no valid score at all -> IndexError / status 2
    want = []
    for r in rows:
        try:
            want.append(oracle_np.completion_row(r))
        except IndexError:
            want.append(None)
    a, b = shard_range(rows.shape[1], world, rank)
    local = torch.from_numpy(rows[:, a:b].copy())
    status = frame_sharded_completion_(local, -10.0)
    ok = bool(int(status[0]) & 2)                                        # row 4
    for r in range(9):
        if want[r] is not None:
            ok = ok and np.array_equal(local.numpy()[r], want[r][a:b])
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write("ok" if ok else "bad")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_frame_sharded_completion_gloo(tmp_path, world):
    """Score completion of rows sharded by frame: a 4-value summary per (row, rank) is all-gathered, every
    rank completes its columns with the positions of the whole tubelet; runs that span one or several shard
    boundaries, shards without any valid score, all-missing rows.  Equal to the unsharded rows bit for bit."""
    port = _free_port()
    mp.spawn(_completion_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"


def _vid_nms_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    from vdetlib_b200 import ops, synth
    from vdetlib_b200.dist import shard_range, sharded_vid_nms
    # the two kernels behind sharded_vid_nms, replaced by the oracle on CPU tensors: what runs here is the
    # cross-rank merge (count exchange, padded all-gathers, global stable order)
    ops.vid_nms = lambda dets, thresh: torch.tensor(c_oracle.vid_nms(dets.numpy(), thresh), dtype=torch.int64)

    def sort_desc(scores, ids):
        order = np.argsort(-scores.numpy(), kind="stable")
        return scores[torch.from_numpy(order)], ids[torch.from_numpy(order)]
    ops.sort_by_score_desc = sort_desc
    T, N = 7, 40                                                    # uneven shards at world 2 and 3
    b, s = synth.boxes_scores(T, N, 1, seed=21)
    sc = s[:, :, 0] + np.arange(T)[:, None] * 1e-5                  # unique across the video
    sc[5] = sc[2]                                                   # ... except two frames with equal scores
    dets = np.concatenate([np.repeat(np.arange(1, T + 1), N)[:, None], b.reshape(-1, 4), sc.reshape(-1, 1)],
                          axis=1).astype(np.float32)
    a, e = shard_range(T, world, rank)
    got = sharded_vid_nms(torch.from_numpy(dets[a * N:e * N]), 0.3, a * N).numpy().tolist()
    want = c_oracle.vid_nms(dets, 0.3)
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write("ok" if got == want else "bad %d %d" % (len(got), len(want)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_vid_nms_merge_gloo(tmp_path, world):
    """vid_nms of a video sharded by frame: suppression is local, the reference's GLOBAL
    descending-score keep order (ties: ascending row) comes from one merge across ranks."""
    port = _free_port()
    mp.spawn(_vid_nms_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"


def _sharded_pipeline_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cuda_fake
    from oracle import c_oracle
    from vdetlib_b200 import synth
    mpatch = pytest.MonkeyPatch()
    cuda_fake.install(mpatch)                                      # fake streams, oracle-backed launches, CPU tensors
    from vdetlib_b200.dist import ShardedVideoPostProcessor, shard_range
    T, N, C, thr = 4 * world, 40, 3, 0.3
    b, s = synth.boxes_scores(T, N, C, seed=12)
    km, _, kc = c_oracle.nms_frames(b, s, thr)
    ls, lb = c_oracle.link_f32(b)                                  # [T-1, N]: successor index inside frame t+1
    a, e = shard_range(T, world, rank)
    per = e - a
    pp = ShardedVideoPostProcessor(per, N, C, thr, torch.device("cpu"))
    ok = True
    # two NEW-shard steps in flight through the boundary exchange (per-slot staging), a re-submission of the staged
    # shard, then a synchronous step
    tickets = [pp.submit_host(b[a:e], s[a:e], graph=False), pp.submit_host(b[a:e], s[a:e], graph=False)]
    results = []
    for t in tickets:
        r = pp.collect(t)
        results.append({"keep_mask": r.keep_mask(), "keep_cnt": np.array(r["keep_cnt"]), "succ": np.array(r["succ"]),
                        "link_iou": np.array(r["link_iou"])})
    r = pp.collect(pp.submit_staged(graph=False))
    results.append({"keep_mask": r.keep_mask(), "keep_cnt": r["keep_cnt"], "succ": r["succ"], "link_iou": r["link_iou"]})
    r = pp.step_host(b[a:e], s[a:e])
    results.append({"keep_mask": r.keep_mask(), "keep_cnt": r["keep_cnt"], "succ": r["succ"], "link_iou": r["link_iou"]})
    # the device-resident step (exchange beside the NMS, then the link) on the staged shard
    dev_res = pp.step_device(pp.pp.d_boxes, pp.pp.d_scores)
    results.append({"keep_mask": dev_res["keep_mask"].numpy(), "keep_cnt": dev_res["keep_cnt"].numpy(),
                    "succ": dev_res["succ"].numpy(), "link_iou": dev_res["link_iou"].numpy()})
    for res in results:
        ok = ok and np.array_equal(res["keep_mask"], km[a:e]) and np.array_equal(res["keep_cnt"], kc[a:e])
        succ = res["succ"].reshape(per, N)
        iou = res["link_iou"].reshape(per, N)
        for t in range(per):
            g = a + t                                              # global frame
            if g < T - 1:
                # inside the shard: packed local row of frame t+1; across the boundary: rows + index into the
                # neighbour's first frame, i.e. BEYOND the local rows (never aliasing them)
                base = (t + 1) * N
                ok = ok and np.array_equal(succ[t] - base, ls[g]) and np.array_equal(iou[t], lb[g])
            else:
                ok = ok and np.all(succ[t] == -1)
    # ragged shards: every rank a different number of frames and boxes; the neighbour's first-frame count is
    # only known through the exchange (device word), chains end at the shard boundary
    rng = np.random.default_rng(5)
    counts_all = [rng.integers(0, N + 1, 3 + r).astype(np.int32) for r in range(world)]
    for c in counts_all:
        c[0] = max(int(c[0]), 1)
    frames_all = [[synth.boxes_scores(1, N, C, seed=100 * r + f) for f in range(len(counts_all[r]))] for r in range(world)]
    cnt = counts_all[rank]
    fb = [frames_all[rank][f][0][0, :cnt[f]] for f in range(len(cnt))]
    fs = [frames_all[rank][f][1][0, :cnt[f]] for f in range(len(cnt))]
    pp2 = ShardedVideoPostProcessor(3 + world, N, C, thr, torch.device("cpu"), n_chunks=2)
    r = pp2.collect(pp2.submit_host(np.concatenate(fb), np.concatenate(fs), counts=cnt, graph=False))
    rows = int(cnt.sum())
    last_a = rows - int(cnt[-1])
    if cnt[-1] > 0:
        if rank < world - 1:
            nxt = frames_all[rank + 1][0][0][0, :counts_all[rank + 1][0]]
            iou = c_oracle.pair_iou_f32(fb[-1], nxt)
            ok = ok and np.array_equal(r["succ"][last_a:], rows + np.argmax(iou, axis=1))
            ok = ok and np.array_equal(r["link_iou"][last_a:], iou.max(axis=1))
        else:
            ok = ok and np.all(r["succ"][last_a:] == -1)
    for f in range(len(cnt)):
        for c in range(C):
            d = np.concatenate([fb[f], fs[f][:, c:c + 1]], axis=1).astype(np.float32)
            ok = ok and np.array_equal(r.keep_list(f, c), np.asarray(c_oracle.nms(d, thr), dtype=np.int64))
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write("ok" if ok else "bad")
    mpatch.undo()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pipeline_host_logic_gloo(tmp_path, world):
    """The frame-sharded staged step (ShardedVideoPostProcessor.submit_host / collect / step_host) across ranks
    on the CPU harness: the boundary all-gather is real (gloo), the launches are the oracle.  Every rank's
    shard must equal its slice of the single-process result, the last frame of a shard linking into the next
    rank's first frame.  (The same with NCCL and the real kernels: tests/test_gpu_multi.py.)"""
    port = _free_port()
    mp.spawn(_sharded_pipeline_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"
