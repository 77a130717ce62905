"""GPU parity of the NMS family against the oracle and the golden vectors (bit-exact indices).

Every call goes through the C ABI of libvdet_b200.so (vdetlib_b200._lib / ops)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from vdetlib_b200 import ops, synth
from vdetlib_b200.utils import cython_nms as gpu

import helpers

pytestmark = pytest.mark.gpu


def test_nms_golden_all():
    g = helpers.golden_npz("nms.npz")
    n = 0
    for key in g.files:
        if key.startswith("nms_") and "_keep_" in key:
            tag, thr = key[4:].split("_keep_")
            thr = int(thr) / (100.0 if len(thr) == 3 else 10.0)
            got = gpu.nms(g["nms_%s_dets" % tag], thr)
            assert got == g[key].tolist(), key
            assert all(type(i) is int for i in got)
            n += 1
    assert n >= 18


def test_vid_and_track_golden():
    g = helpers.golden_npz("nms.npz")
    for thr in (0.3, 0.5):
        assert gpu.vid_nms(g["vid_dets"], thresh=thr) == g["vid_keep_%02d" % int(thr * 10)].tolist()
        assert gpu.track_det_nms(g["tdn_tracks"], g["tdn_dets"], thr) == g["tdn_keep_%02d" % int(thr * 10)].tolist()


def test_config1_300_boxes_thr05():
    """BASELINE config 1: single 300-box frame, 1 class, IoU 0.5, bit-exact vs the CPU path."""
    b, s = synth.boxes_scores(1, 300, 1, seed=1000)
    dets = np.concatenate([b[0], s[0]], axis=1).astype(np.float32)
    assert gpu.nms(dets, 0.5) == c_oracle.nms(dets, 0.5)


@pytest.mark.parametrize("seed", range(6))
def test_nms_random_vs_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(12):
        n = int(rng.choice([1, 2, 31, 32, 33, 64, 65, 100, 257, 300, 512, 513, 777, 1024]))
        thr = float(rng.choice([0.0, 0.1, 0.3, 0.5, 0.7, 0.95, 1.0]))
        d = helpers.unique_score_dets(rng, n)
        if rng.integers(0, 2):
            d[:, :4] = np.round(d[:, :4])
        assert gpu.nms(d, thr) == c_oracle.nms(d, thr), (n, thr)


@pytest.mark.parametrize("seed", range(4))
def test_vid_nms_random_vs_oracle(seed):
    rng = np.random.default_rng(200 + seed)
    for _ in range(6):
        n = int(rng.integers(1, 3000))
        nf = int(rng.choice([1, 2, 7, 40]))
        thr = float(rng.choice([0.1, 0.3, 0.5]))
        v = helpers.unique_score_dets(rng, n, with_frame=nf)
        if rng.integers(0, 2):
            v[:, 0] = v[:, 0] * 3 + 1           # sparse frame ids
        assert gpu.vid_nms(v, thr) == c_oracle.vid_nms(v, thr), (n, nf, thr)
        q = int(rng.integers(0, 8))
        t = helpers.unique_score_dets(rng, max(q, 1), with_frame=nf)[:q, :5]
        assert gpu.track_det_nms(t, v, thr) == c_oracle.track_det_nms(t, v, thr), (n, nf, thr, q)


def test_known_answers_and_errors():
    d = np.asarray([[0, 0, 9, 9, 0.9], [9, 0, 18, 9, 0.8]], np.float32)
    assert gpu.nms(d, 10.0 / 190.0) == [0]
    assert gpu.nms(d, float(np.nextafter(np.float32(10.0 / 190.0), np.float32(1)))) == [0, 1]
    assert gpu.nms(d, 0.7) == [0, 1]            # 0.7 is not a float32: the threshold rounds UP on the host
    v = np.asarray([[1, 0, 0, 9, 9, 0.9], [2, 0, 0, 9, 9, 0.8], [1, 0, 0, 9, 9, 0.7]], np.float32)
    assert gpu.vid_nms(v, 0.3) == [0, 1]
    t = np.asarray([[1, 0, 0, 9, 9]], np.float32)
    dd = np.asarray([[1, 0, 0, 9, 9, 0.9], [1, 50, 50, 60, 60, 0.5], [1, 51, 51, 60, 60, 0.4],
                     [2, 0, 0, 9, 9, 0.3]], np.float32)
    assert gpu.track_det_nms(t, dd, 0.3) == [1, 3]
    assert gpu.nms(np.zeros((0, 5), np.float32), 0.3) == []
    assert gpu.vid_nms(np.zeros((0, 6), np.float32), 0.3) == []
    assert gpu.track_det_nms(t, np.zeros((0, 6), np.float32), 0.3) == []
    assert gpu.track_det_nms(np.zeros((0, 5), np.float32), dd, 0.3) == c_oracle.track_det_nms(
        np.zeros((0, 5), np.float32), dd, 0.3)
    with pytest.raises(ValueError):
        gpu.nms(d.astype(np.float64), 0.3)
    with pytest.raises(ValueError):
        gpu.nms(d[0], 0.3)
    with pytest.raises(TypeError):
        gpu.nms(d.tolist(), 0.3)
    z = np.asarray([[5, 5, 4, 4, 0.9], [7, 7, 6, 6, 0.8]], np.float32)
    with pytest.raises(ZeroDivisionError):
        gpu.nms(z, 0.3)
    # a zero-union pair that the reference never visits (j is suppressed first) must NOT raise
    z2 = np.asarray([[0, 0, 20, 20, 0.9], [5, 5, 4, 4, 0.8], [7, 7, 6, 6, 0.7], [100, 100, 120, 120, 0.6]], np.float32)
    try:
        want = c_oracle.nms(z2, 0.3)
    except ZeroDivisionError:
        want = ZeroDivisionError
    if want is ZeroDivisionError:
        with pytest.raises(ZeroDivisionError):
            gpu.nms(z2, 0.3)
    else:
        assert gpu.nms(z2, 0.3) == want


def test_strided_and_wide_input():
    rng = np.random.default_rng(8)
    big = helpers.unique_score_dets(rng, 120)
    wide = np.zeros((120, 9), np.float32)
    wide[:, ::2] = big
    assert gpu.nms(wide[:, ::2], 0.4) == c_oracle.nms(big, 0.4)
    extra = np.concatenate([big, np.ones((120, 3), np.float32)], axis=1)     # extra columns are ignored
    assert gpu.nms(extra, 0.4) == c_oracle.nms(big, 0.4)


@pytest.mark.parametrize("T,N,C,thr", [(17, 300, 30, 0.3), (5, 64, 1, 0.5), (9, 33, 7, 0.3), (3, 1000, 4, 0.3),
                                      (4, 150, 9, 0.3), (4, 190, 9, 0.3), (4, 257, 9, 0.3), (4, 320, 9, 0.5), (4, 380, 9, 0.3),
                                        (3, 1025, 2, 0.3), (160, 2000, 3, 0.3), (2, 2048, 2, 0.5)])
def test_nms_frames_vs_oracle(T, N, C, thr):
    b, s = synth.boxes_scores(T, N, C, seed=T * 1000 + N)
    km, ki, kc = c_oracle.nms_frames(b, s, thr)
    dev = torch.device("cuda")
    db = torch.from_numpy(b.reshape(-1, 4)).to(dev)
    ds = torch.from_numpy(s.reshape(-1, C)).to(dev)
    seg = ops.seg_offsets_uniform(T, N, dev)
    keep_idx, keep_cnt, keep_mask, status = ops.nms_frames(db, ds, seg, thr, N, want_mask=True)
    assert ops.raise_for_status(status) == 0
    keep_cnt = keep_cnt.cpu().numpy()            # [C, T]
    assert np.array_equal(keep_cnt.T, kc)
    gi = keep_idx.cpu().numpy().reshape(C, T, N)
    gm = keep_mask.cpu().numpy().reshape(C, T, N)
    for t in range(T):
        for c in range(C):
            want = ki[t, c]
            want = np.where(want >= 0, want + t * N, -1)
            assert np.array_equal(gi[c, t], want), (t, c)
    assert np.array_equal(gm.transpose(1, 0, 2), km)
    # class-major score layout gives the same answer
    ds_cm = ds.t().contiguous()
    keep_idx2, keep_cnt2, _, _ = ops.nms_frames(db, ds_cm, seg, thr, N, class_major=True)
    assert torch.equal(keep_idx2, keep_idx) and torch.equal(keep_cnt2.cpu(), torch.from_numpy(keep_cnt))


@pytest.mark.parametrize("T,N,C,thr", [(6, 300, 5, 0.5), (6, 300, 5, 0.3), (6, 300, 5, 0.25), (5, 96, 3, 1.0 / 3.0),
                                      (3, 1200, 2, 0.5), (2, 2048, 2, 0.25), (4, 300, 4, 0.2)])
def test_nms_frames_integer_boxes_hit_the_threshold_exactly(T, N, C, thr):
    """SURVEY 8(d)'s "integer" variant: rounded coordinates make IoU a ratio of small integers, so IoU == thresh
    (>= suppresses, nms.pyx:64) and IoUs whose float32 quotient straddles the double threshold really occur -- the
    pairs the division-free filter must hand to the exact division.  Every frame also gets exact duplicates."""
    b, s = synth.boxes_scores(T, N, C, seed=77 * T + N, integer=True)
    b[:, 1::7] = b[:, 0::7][:, :b[:, 1::7].shape[1]]            # duplicates: IoU exactly 1
    # construct pairs with IoU exactly 1/2, 1/3, 1/4 and 3/10: [0,0,w-1,h-1] against a box of the same height
    b[:, 2] = (10, 10, 29, 19)                                  # 20 x 10 = 200
    b[:, 3] = (10, 10, 19, 19)                                  # inside it, 100: IoU 1/2
    b[:, 4] = (400, 10, 429, 19)                                # 300
    b[:, 5] = (400, 10, 409, 19)                                # 100 inside 300: IoU 1/3
    b[:, 6] = (700, 10, 739, 19)                                # 400
    b[:, 8] = (700, 10, 709, 19)                                # IoU 1/4
    b[:, 9] = (900, 300, 999, 309)                              # 1000
    b[:, 10] = (900, 300, 929, 309)                             # 300 inside 1000: IoU 3/10
    b[:, 11] = (900, 500, 949, 509)                             # 500
    b[:, 12] = (900, 500, 909, 509)                             # IoU 1/5
    km, ki, kc = c_oracle.nms_frames(b, s, thr)
    dev = torch.device("cuda")
    seg = ops.seg_offsets_uniform(T, N, dev)
    keep_idx, keep_cnt, keep_mask, status = ops.nms_frames(torch.from_numpy(b.reshape(-1, 4)).to(dev),
                                                           torch.from_numpy(s.reshape(-1, C)).to(dev), seg, thr, N,
                                                           want_mask=True, frame_major_out=True)
    assert ops.raise_for_status(status) == 0
    assert np.array_equal(keep_cnt.cpu().numpy(), kc)
    assert np.array_equal(keep_mask.cpu().numpy().reshape(T, C, N), km)
    gi = keep_idx.cpu().numpy().reshape(T, C, N)
    want = np.where(ki >= 0, ki + (np.arange(T) * N)[:, None, None], -1)
    assert np.array_equal(gi, want)
    # the constructed pair that sits exactly on this threshold is suppressed by its partner whenever the partner
    # ranks higher (>=, not >): check it through the oracle's own mask so the test cannot pass vacuously
    on_thr = {0.5: (2, 3), 0.25: (6, 8), 0.2: (11, 12)}.get(thr)
    if on_thr:
        i, j = on_thr
        hit = 0
        for t in range(T):
            for c in range(C):
                hi, lo = (i, j) if s[t, i, c] > s[t, j, c] else (j, i)
                if km[t, c, hi]:
                    assert not km[t, c, lo]
                    hit += 1
        assert hit > 0


def test_nms_frames_ragged():
    rng = np.random.default_rng(5)
    counts = np.asarray([0, 1, 40, 0, 300, 33, 64, 2], np.int32)
    T, nmax, C = len(counts), 300, 3
    b, s = synth.boxes_scores(T, nmax, C, seed=77)
    km, ki, kc = c_oracle.nms_frames(b, s, 0.3, counts)
    rows_b = np.concatenate([b[t, :counts[t]] for t in range(T)])
    rows_s = np.concatenate([s[t, :counts[t]] for t in range(T)])
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    dev = torch.device("cuda")
    keep_idx, keep_cnt, keep_mask, status = ops.nms_frames(
        torch.from_numpy(rows_b).to(dev), torch.from_numpy(rows_s).to(dev),
        torch.from_numpy(off).to(dev), 0.3, int(counts.max()), want_mask=True)
    assert ops.raise_for_status(status) == 0
    assert np.array_equal(keep_cnt.cpu().numpy().T, kc)
    gi = keep_idx.cpu().numpy()
    gm = keep_mask.cpu().numpy()
    for t in range(T):
        for c in range(C):
            n = counts[t]
            want = ki[t, c, :n]
            want = np.where(want >= 0, want + off[t], -1)
            assert np.array_equal(gi[c, off[t]:off[t] + n], want)
            assert np.array_equal(gm[c, off[t]:off[t] + n], km[t, c, :n])


def test_full_config2_properties():
    """BASELINE config 2 (1000 frames x 300 boxes x 30 classes): size-independent properties plus
    an oracle check on a sample of frames."""
    T, N, C, thr = 1000, 300, 30, 0.3
    b, s = synth.boxes_scores(T, N, C, seed=2)
    dev = torch.device("cuda")
    db = torch.from_numpy(b.reshape(-1, 4)).to(dev)
    ds = torch.from_numpy(s.reshape(-1, C)).to(dev)
    seg = ops.seg_offsets_uniform(T, N, dev)
    keep_idx, keep_cnt, keep_mask, status = ops.nms_frames(db, ds, seg, thr, N, want_mask=True)
    assert ops.raise_for_status(status) == 0
    km = keep_mask.view(C, T, N)
    assert torch.equal(km.sum(-1).to(torch.int32), keep_cnt)                  # counts == mask popcounts
    ki = keep_idx.view(C, T, N)
    valid = ki >= 0
    assert torch.equal(valid.sum(-1).to(torch.int32), keep_cnt)
    # kept rows listed in descending score
    rows = ki.clamp(min=0).long()
    sc = ds.t().reshape(C, T * N).gather(1, rows.reshape(C, -1)).view(C, T, N)
    sc = torch.where(valid, sc, torch.full_like(sc, -1.0))
    assert bool((sc[:, :, 1:] <= sc[:, :, :-1]).all())
    # idempotence: NMS of the survivors keeps every survivor (per class, on a few classes)
    for c in (0, 13, 29):
        ds_c = torch.where(km[c].reshape(-1).bool(), ds[:, c], torch.full_like(ds[:, c], -1.0))
        # suppressed rows get the lowest scores so they cannot influence survivors
        k2, c2, m2, _ = ops.nms_frames(db, ds_c.contiguous(), seg, thr, N, want_mask=True)
        assert bool((m2.view(T, N)[km[c].bool()] == 1).all())
    # oracle on a sample of frames
    sample = [0, 1, 499, 998, 999]
    okm, oki, okc = c_oracle.nms_frames(b[sample], s[sample], thr)
    assert np.array_equal(keep_cnt.cpu().numpy().T[sample], okc)
    assert np.array_equal(km.cpu().numpy().transpose(1, 0, 2)[sample], okm)


def test_nms_any_length_frames():
    """Frames longer than the bit-matrix kernels hold (> 2048 boxes) take the on-the-fly path."""
    rng = np.random.default_rng(41)
    d = helpers.unique_score_dets(rng, 5000, scale=900.0)
    assert gpu.nms(d, 0.3) == c_oracle.nms(d, 0.3)
    v = helpers.unique_score_dets(rng, 9000, with_frame=3, scale=900.0)
    assert gpu.vid_nms(v, 0.5) == c_oracle.vid_nms(v, 0.5)
    t = helpers.unique_score_dets(rng, 4, with_frame=3)[:, :5]
    assert gpu.track_det_nms(t, v, 0.3) == c_oracle.track_det_nms(t, v, 0.3)
    z = np.concatenate([d[:3000], np.asarray([[5, 5, 4, 4, 0.001], [7, 7, 6, 6, 0.0005]], np.float32)])
    with pytest.raises(ZeroDivisionError):
        gpu.nms(z, 0.3)


@pytest.mark.parametrize("n", [1100, 1500, 2046])
def test_zero_division_big_frames(n):
    """ZeroDivisionError semantics (visited pairs only, nms.pyx:64) in the 1025..2048-box kernel."""
    rng = np.random.default_rng(n)
    d = helpers.unique_score_dets(rng, n, scale=900.0)
    deg = np.asarray([[5, 5, 4, 4, 0.0], [7, 7, 6, 6, 0.0]], np.float32)
    for s0, s1 in ((0.001, 0.0005), (2.0, 1.5), (2.0, 0.0005)):
        deg[0, 4], deg[1, 4] = s0, s1
        z = np.concatenate([d[:n // 2], deg[:1], d[n // 2:], deg[1:]])
        try:
            want = c_oracle.nms(z, 0.3)
        except ZeroDivisionError:
            want = ZeroDivisionError
        if want is ZeroDivisionError:
            with pytest.raises(ZeroDivisionError):
                gpu.nms(z, 0.3)
        else:
            assert gpu.nms(z, 0.3) == want
    # one degenerate box alone has a positive union with every sane box: no error
    z = np.concatenate([d, deg[:1]])
    assert gpu.nms(z, 0.3) == c_oracle.nms(z, 0.3)
    # threshold 0: the top box suppresses everything, the zero-union pair is never visited
    deg[0, 4], deg[1, 4] = 0.001, 0.0005
    z = np.concatenate([d, deg])
    assert gpu.nms(z, 0.0) == c_oracle.nms(z, 0.0) == [int(np.argmax(z[:, 4]))]


def test_frame_major_layout_and_pipelined_postprocessor():
    """Frame-major outputs (ragged frames) and the chunk-pipelined host API equal the oracle."""
    from vdetlib_b200.vdet.video_det import VideoPostProcessor
    counts = np.asarray([5, 0, 64, 300, 33, 1, 128, 7], np.int32)
    T, nmax, C = len(counts), 300, 4
    b, s = synth.boxes_scores(T, nmax, C, seed=91)
    km, ki, kc = c_oracle.nms_frames(b, s, 0.3, counts)
    rows_b = np.concatenate([b[t, :counts[t]] for t in range(T)])
    rows_s = np.concatenate([s[t, :counts[t]] for t in range(T)])
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    dev = torch.device("cuda")
    keep_idx, keep_cnt, keep_mask, status = ops.nms_frames(
        torch.from_numpy(rows_b).to(dev), torch.from_numpy(rows_s).to(dev), torch.from_numpy(off).to(dev),
        0.3, int(counts.max()), want_mask=True, frame_major_out=True)
    assert ops.raise_for_status(status) == 0
    assert np.array_equal(keep_cnt.cpu().numpy(), kc)                       # [S, C]
    gi, gm = keep_idx.cpu().numpy(), keep_mask.cpu().numpy()
    for t in range(T):
        n = counts[t]
        for c in range(C):
            blk = off[t] * C + c * n
            want = np.where(ki[t, c, :n] >= 0, ki[t, c, :n] + off[t], -1)
            assert np.array_equal(gi[blk:blk + n], want)
            assert np.array_equal(gm[blk:blk + n], km[t, c, :n])
    # the public host API, pipelined over 3 uneven chunks
    T2, N2, C2 = 10, 120, 5
    b2, s2 = synth.boxes_scores(T2, N2, C2, seed=92)
    km2, ki2, kc2 = c_oracle.nms_frames(b2, s2, 0.3)
    ls, lb = c_oracle.link_f32(b2)
    for n_chunks in (1, 3, 10):
        pp = VideoPostProcessor(T2, N2, C2, 0.3, n_chunks=n_chunks)
        out = pp.run_host(b2, s2)
        assert np.array_equal(out.keep_mask(), km2) and np.array_equal(out["keep_cnt"], kc2)
        for t in range(T2):                                          # ordered keep lists, frame-local indices
            for c in range(C2):
                assert np.array_equal(out.keep_list(t, c), ki2[t, c, :kc2[t, c]]), (t, c)
        got = out["succ"][:(T2 - 1) * N2].reshape(T2 - 1, N2) - np.arange(1, T2)[:, None] * N2
        assert np.array_equal(got, ls) and np.array_equal(out["link_iou"][:(T2 - 1) * N2].reshape(T2 - 1, N2), lb)
        dev_out = pp.run_device(pp.d_boxes, pp.d_scores)
        assert np.array_equal(dev_out["keep_idx"].cpu().numpy(), np.where(ki2 >= 0, ki2 + (np.arange(T2) * N2)[:, None, None], -1))
        for _ in range(3):                                           # CUDA-graph replay of the same step
            pp.d_mask.zero_(); pp.d_succ.zero_()
            g_out = pp.run_device(pp.d_boxes, pp.d_scores, graph=True)
            assert np.array_equal(g_out["keep_mask"].cpu().numpy(), km2)
            assert np.array_equal(g_out["succ"].cpu().numpy(), out["succ"])


@pytest.mark.parametrize("graph", [False, True])
def test_video_postprocessor_two_steps_in_flight(graph):
    """submit_staged / collect: two double-buffered steps in flight, restaged inputs in between,
    eager stream pipeline and whole-step CUDA-graph replay -- same results as the oracle every time."""
    from vdetlib_b200.vdet.video_det import VideoPostProcessor
    T, N, C = 12, 150, 7
    pp = VideoPostProcessor(T, N, C, 0.3, n_chunks=4, n_stage=1)
    data = [synth.boxes_scores(T, N, C, seed=300 + k) for k in range(3)]
    want = []
    for b, s in data:
        km, _, kc = c_oracle.nms_frames(b, s, 0.3)
        ls, lb = c_oracle.link_f32(b)
        want.append((km, kc, ls, lb))

    def check(out, w):
        km, kc, ls, lb = w
        assert np.array_equal(out.keep_mask(), km) and np.array_equal(out["keep_cnt"], kc)
        got = out["succ"][:(T - 1) * N].reshape(T - 1, N) - np.arange(1, T)[:, None] * N
        assert np.array_equal(got, ls)
        assert np.array_equal(out["link_iou"][:(T - 1) * N].reshape(T - 1, N), lb)

    for rnd in range(3):                                     # slots alternate: 0,1 | 0,1 | 0,1
        b, s = data[rnd]
        pp.stage(b, s)
        t0 = pp.submit_staged(graph=graph)
        t1 = pp.submit_staged(graph=graph)                   # same staged inputs, the other slot
        with pytest.raises(RuntimeError):
            pp.submit_staged(graph=graph)                    # both slots busy
        with pytest.raises(RuntimeError):
            pp.stage(b, s)                                   # in-flight steps still read the staging buffers
        check(pp.collect(t0), want[rnd])
        # keep the pipeline full: a third step goes out before the second is collected
        t2 = pp.submit_staged(graph=graph)
        check(pp.collect(t1), want[rnd])
        check(pp.collect(t2), want[rnd])
        with pytest.raises(RuntimeError):
            pp.collect(t2)
    # the synchronous call still works afterwards, on slot 0, mixed with the device entry
    out = pp.run_host(*data[0])
    check(out, want[0])
    dev_out = pp.run_device(pp.d_boxes, pp.d_scores)
    assert np.array_equal(dev_out["keep_mask"].cpu().numpy(), want[0][0])


@pytest.mark.parametrize("graph", [False, True])
def test_video_postprocessor_streams_new_shards_from_pageable_memory(graph):
    """submit_host / collect with a NEW shard per step from ordinary (pageable) NumPy arrays, two steps in
    flight on per-slot staging buffers: every ticket returns the ordered keep lists and the link of ITS shard;
    a status flag raised by one shard does not poison the next one."""
    from vdetlib_b200.vdet.video_det import VideoPostProcessor
    T, N, C = 10, 200, 5
    pp = VideoPostProcessor(T, N, C, 0.3, n_chunks=3, want_bits=True)
    shards = [synth.boxes_scores(T, N, C, seed=900 + k) for k in range(5)]

    def check(out, b, s):
        km, ki, kc = c_oracle.nms_frames(b, s, 0.3)
        ls, lb = c_oracle.link_f32(b)
        assert np.array_equal(out["keep_cnt"], kc) and np.array_equal(out.keep_mask(), km)
        for t in range(T):
            for c in range(C):
                assert np.array_equal(out.keep_list(t, c), ki[t, c, :kc[t, c]]), (t, c)
        bits = np.unpackbits(out["keep_bits"].view(np.uint8), bitorder="little").reshape(T, C, -1)[:, :, :N]
        assert np.array_equal(bits, km)
        got = out["succ"][:(T - 1) * N].reshape(T - 1, N) - np.arange(1, T)[:, None] * N
        assert np.array_equal(got, ls) and np.all(out["succ"][(T - 1) * N:] == -1)
        assert np.array_equal(out["link_iou"][:(T - 1) * N].reshape(T - 1, N), lb)

    tickets = [pp.submit_host(*shards[0], graph=graph)]
    for k in range(1, len(shards)):
        tickets.append(pp.submit_host(*shards[k], graph=graph))
        check(pp.collect(tickets[k - 1]), *shards[k - 1])
    check(pp.collect(tickets[-1]), *shards[-1])
    # a degenerate shard raises ZeroDivisionError (nms.pyx:64) ...
    bad_b, bad_s = shards[0][0].copy(), shards[0][1].copy()
    bad_b[3, :2] = np.asarray([10.0, 10.0, 9.0, 20.0], np.float32)          # two boxes of width 0: union == 0
    bad_s[3, 0, :], bad_s[3, 1, :] = 0.9995, 0.9994
    with pytest.raises(ZeroDivisionError):
        pp.collect(pp.submit_host(bad_b, bad_s, graph=graph))
    # ... and the next clean shard on the same slot is clean again (status word reset per step)
    for k in range(2):
        check(pp.collect(pp.submit_host(*shards[k], graph=graph)), *shards[k])


def test_video_postprocessor_registered_caller_arrays():
    """register_host_arrays: caller-owned arrays pinned in place are uploaded from directly (no staging copy), eager
    and graph replay (one graph per source array), mixed with ordinary pageable shards; unregistering falls back."""
    from vdetlib_b200.vdet.video_det import VideoPostProcessor
    T, N, C = 8, 160, 4
    pp = VideoPostProcessor(T, N, C, 0.3, n_chunks=2)
    shards = [synth.boxes_scores(T, N, C, seed=970 + k) for k in range(3)]
    want = [c_oracle.nms_frames(b, s, 0.3) for b, s in shards]
    links = [c_oracle.link_f32(b) for b, _ in shards]
    pp.register_host_arrays(*[a for sh in shards[:2] for a in sh])          # the third shard stays pageable

    def check(out, k):
        km, ki, kc = want[k]
        assert np.array_equal(out["keep_cnt"], kc) and np.array_equal(out.keep_mask(), km)
        got = out["succ"][:(T - 1) * N].reshape(T - 1, N) - np.arange(1, T)[:, None] * N
        assert np.array_equal(got, links[k][0])

    for graph in (False, True):
        order = [0, 1, 2, 1, 0, 2, 0]
        t = pp.submit_host(*shards[order[0]], graph=graph)
        for prev, k in zip(order, order[1:]):
            t2 = pp.submit_host(*shards[k], graph=graph)
            check(pp.collect(t), prev)
            t = t2
        check(pp.collect(t), order[-1])
    assert pp._src[0][0].data_ptr() != pp.h_boxes_sets[0].data_ptr() or pp._src[1][0].data_ptr() != pp.h_boxes_sets[1].data_ptr()
    pp.unregister_host_arrays()
    assert not pp._registered
    assert pp._staged[0] is None and pp._staged[1] is not None       # slot 0 had a caller array, slot 1 the pageable shard
    with pytest.raises(RuntimeError):
        pp.run_staged(graph=True)                # slot 0: its staged shard was a caller array that is no longer pinned
    for k in (0, 1):
        check(pp.collect(pp.submit_host(*shards[k], graph=True)), k)
        assert pp._src[0][0].data_ptr() == pp.h_boxes_sets[0].data_ptr()
    with pytest.raises(ValueError):
        pp.register_host_arrays(shards[0][0].astype(np.float64))


def test_video_postprocessor_producer_writes_in_place(graph=True):
    """input_buffers / commit_inputs: the producer fills the pinned upload buffers itself (no staging copy); uniform
    and ragged shards, two steps in flight, each slot handing out its own buffers."""
    from vdetlib_b200.vdet.video_det import VideoPostProcessor
    T, N, C = 6, 200, 5
    pp = VideoPostProcessor(T, N, C, 0.3, n_chunks=3)
    shards = [synth.boxes_scores(T, N, C, seed=990 + k) for k in range(3)]
    tickets = []
    for k in range(2):
        hb, hs = pp.input_buffers()
        assert hb.shape == (T * N, 4) and hs.shape == (T * N, C)
        hb[:] = shards[k][0].reshape(-1, 4)
        hs[:] = shards[k][1].reshape(-1, C)
        pp.commit_inputs()
        tickets.append(pp.submit_staged(graph=(graph and k == 1)))
    with pytest.raises(RuntimeError):
        pp.input_buffers()                                   # both slots in flight: no buffer is the caller's
    for k in range(2):
        out = pp.collect(tickets[k])
        km, ki, kc = c_oracle.nms_frames(*shards[k], 0.3)
        assert np.array_equal(out["keep_cnt"], kc) and np.array_equal(out.keep_mask(), km)
        ls, lb = c_oracle.link_f32(shards[k][0])
        assert np.array_equal(out["succ"][:(T - 1) * N].reshape(T - 1, N) - np.arange(1, T)[:, None] * N, ls)
    # ragged, in place: packed rows from row 0
    counts = np.asarray([200, 0, 31, 7], np.int32)
    b, s = shards[2]
    hb, hs = pp.input_buffers()
    r = 0
    for t, n in enumerate(counts):
        hb[r:r + n] = b[t, :n]
        hs[r:r + n] = s[t, :n]
        r += n
    pp.commit_inputs(counts)
    out = pp.collect(pp.submit_staged())
    km, ki, kc = c_oracle.nms_frames(b[:4], s[:4], 0.3, counts)
    assert np.array_equal(out["keep_cnt"], kc)
    for t in range(4):
        for c in range(C):
            assert np.array_equal(out.keep_list(t, c), ki[t, c, :kc[t, c]]), (t, c)
    with pytest.raises(ValueError):
        pp.commit_inputs([N + 1])


def test_video_postprocessor_ragged_frames():
    """Ragged shards (packed rows + counts) through the staged pipeline, chunk edges balanced by rows."""
    from vdetlib_b200.vdet.video_det import VideoPostProcessor
    N, C = 300, 6
    pp = VideoPostProcessor(12, N, C, 0.3, n_chunks=4)
    for trial, counts in enumerate([[5, 0, 64, 300, 33, 1, 128, 7], [300] * 12, [0, 0, 9], [17]]):
        counts = np.asarray(counts, np.int32)
        T = len(counts)
        b, s = synth.boxes_scores(T, N, C, seed=950 + trial)
        km, ki, kc = c_oracle.nms_frames(b, s, 0.3, counts)
        rows_b = np.concatenate([b[t, :counts[t]] for t in range(T)])
        rows_s = np.concatenate([s[t, :counts[t]] for t in range(T)])
        off = np.concatenate([[0], np.cumsum(counts)])
        out = pp.collect(pp.submit_host(rows_b, rows_s, counts=counts))
        assert np.array_equal(out["keep_cnt"], kc)
        for t in range(T):
            for c in range(C):
                assert np.array_equal(out.keep_list(t, c), ki[t, c, :kc[t, c]]), (trial, t, c)
        for t in range(T):
            a, e = off[t], off[t + 1]
            if e == a:
                continue
            if t + 1 < T and counts[t + 1] > 0:
                iou = c_oracle.pair_iou_f32(b[t, :counts[t]], b[t + 1, :counts[t + 1]])
                assert np.array_equal(out["succ"][a:e], off[t + 1] + np.argmax(iou, axis=1))
                assert np.array_equal(out["link_iou"][a:e], iou.max(axis=1))
            else:
                assert np.all(out["succ"][a:e] == -1)


def test_compact_keep_matches_padded_blocks():
    """vdet_compact_keep: the padded frame-major keep blocks as one contiguous list (uint16 local / int32 row),
    prefix offsets and bit masks -- including a big-frame shard (2000-box frames)."""
    from vdetlib_b200 import _lib
    dev = torch.device("cuda")
    lib = _lib.load()
    for counts, C in (([300] * 7, 30), ([5, 0, 64, 300, 33, 1, 128, 7], 4), ([2000, 1500, 3], 3)):
        counts = np.asarray(counts, np.int32)
        T, nmax = len(counts), int(counts.max())
        b, s = synth.boxes_scores(T, nmax, C, seed=960)
        rows_b = np.concatenate([b[t, :counts[t]] for t in range(T)])
        rows_s = np.concatenate([s[t, :counts[t]] for t in range(T)])
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        d_off = torch.from_numpy(off).to(dev)
        keep_idx, keep_cnt, _, status = ops.nms_frames(torch.from_numpy(rows_b).to(dev), torch.from_numpy(rows_s).to(dev),
                                                       d_off, 0.3, nmax, frame_major_out=True)
        gi, gc = keep_idx.cpu().numpy(), keep_cnt.cpu().numpy()
        nb, W = T * C, (nmax + 31) // 32
        for dtype, tdt in ((_lib.KEEP_U16_LOCAL, torch.uint16), (_lib.KEEP_I32_ROW, torch.int32)):
            k_off = torch.empty(nb + 1, dtype=torch.int32, device=dev)
            k_off2 = torch.zeros(nb + 1, dtype=torch.int32).pin_memory()
            k_out = torch.zeros(int(gc.sum()) + 8, dtype=tdt).pin_memory()      # mapped pinned host memory
            k_bits = torch.empty((nb, W), dtype=torch.int32, device=dev)
            _lib.check(lib.vdet_compact_keep(keep_idx.data_ptr(), keep_cnt.data_ptr(), d_off.data_ptr(), T, nmax, C, dtype,
                                             k_off.data_ptr(), k_off2.data_ptr(), k_out.data_ptr(), k_bits.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "compact_keep")
            torch.cuda.synchronize()
            pre = np.concatenate([[0], np.cumsum(gc.reshape(-1))])
            assert np.array_equal(k_off.cpu().numpy(), pre) and np.array_equal(k_off2.numpy(), pre)
            got, bits = k_out.numpy(), k_bits.cpu().numpy().view(np.uint32)
            for t in range(T):
                n = counts[t]
                for c in range(C):
                    k = t * C + c
                    rows = gi[off[t] * C + c * n:off[t] * C + c * n + gc[t, c]]
                    want = rows - off[t] if dtype == _lib.KEEP_U16_LOCAL else rows
                    assert np.array_equal(got[pre[k]:pre[k + 1]].astype(np.int64), want), (t, c)
                    m = np.zeros(W * 32, np.uint8)
                    m[rows - off[t]] = 1
                    assert np.array_equal(np.unpackbits(bits[k].view(np.uint8), bitorder="little"), m), (t, c)
            assert np.all(got[pre[-1]:] == 0)                                   # nothing written beyond sum K


def test_sort_by_score_desc():
    rng = np.random.default_rng(6)
    for n in (1, 5, 2048, 2049, 100000):
        sc = rng.uniform(-1, 1, n).astype(np.float32)
        sc[rng.integers(0, n, max(n // 10, 1))] = 0.25                      # ties keep their input order
        ids = rng.integers(0, 1 << 40, n)
        so, io = ops.sort_by_score_desc(torch.from_numpy(sc).cuda(), torch.from_numpy(ids).cuda())
        order = np.argsort(-sc, kind="stable")
        assert np.array_equal(so.cpu().numpy(), sc[order]) and np.array_equal(io.cpu().numpy(), ids[order])
        # float64 keys (Python-float scores): values that collapse in float32 stay ordered
        sd = sc.astype(np.float64) + rng.integers(-3, 4, n) * 1e-12
        so, io = ops.sort_by_score_desc(torch.from_numpy(sd).cuda(), torch.from_numpy(ids).cuda())
        order = np.argsort(-sd, kind="stable")
        assert np.array_equal(so.cpu().numpy(), sd[order]) and np.array_equal(io.cpu().numpy(), ids[order])


@pytest.mark.parametrize("N,C", [(300, 6), (64, 3), (1000, 2), (1500, 3), (2048, 2)])
def test_tied_scores_follow_the_documented_rule(N, C):
    """Heavy score ties (saturated / quantised scores).  The reference's own tie order is
    NumPy-build dependent; the documented rule here is "descending score, then ascending row" and the
    C oracle implements the same rule -- this exercises the 64-bit sort fallback of the register
    kernel and the match.any ordinal of the big-frame kernel."""
    T = 4
    b, s = synth.boxes_scores(T, N, C, seed=N + C)
    rng = np.random.default_rng(N)
    s[:, :, 0] = np.round(s[:, :, 0] * 4) / 4                      # 5 distinct values
    s[:, :, 1] = (s[:, :, 1] > 0.5).astype(np.float32)             # saturated 0 / 1
    if C > 2:
        s[:, :, 2] = 0.75                                          # all equal
    s[1, ::7, :] = -0.0                                            # -0.0 ties with +0.0
    km, ki, kc = c_oracle.nms_frames(b, s, 0.3)
    dev = torch.device("cuda")
    keep_idx, keep_cnt, keep_mask, status = ops.nms_frames(
        torch.from_numpy(b.reshape(-1, 4)).to(dev), torch.from_numpy(s.reshape(-1, C)).to(dev),
        ops.seg_offsets_uniform(T, N, dev), 0.3, N, want_mask=True, frame_major_out=True)
    assert ops.raise_for_status(status) == 0
    assert np.array_equal(keep_cnt.cpu().numpy(), kc)
    assert np.array_equal(keep_mask.cpu().numpy().reshape(T, C, N), km)
    want = np.where(ki >= 0, ki + (np.arange(T) * N)[:, None, None], -1)
    assert np.array_equal(keep_idx.cpu().numpy().reshape(T, C, N), want)
    # the drop-in entries (single class) follow the same rule, including the global vid_nms order
    dets = np.concatenate([np.repeat(np.arange(T), N)[:, None].astype(np.float32), b.reshape(-1, 4), s[:, :, 0].reshape(-1, 1)], axis=1)
    assert gpu.vid_nms(dets, 0.3) == c_oracle.vid_nms(dets, 0.3)
    assert gpu.nms(dets[:N, 1:], 0.3) == c_oracle.nms(dets[:N, 1:], 0.3)
