"""BASELINE.json configs 3, 4, 5 at FULL per-GPU size on one device, through size-independent
properties plus oracle checks on sampled frames/rows (the oracle cannot finish the full sizes)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, oracle_np
from vdetlib_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_config3_link_5000x1000():
    """5000 frames x 1000 boxes tubelet linking (config 3; here the whole video on one GPU)."""
    T, N = 5000, 1000
    b, _ = synth.boxes_scores(T, N, 1, seed=3)
    db = torch.from_numpy(b.reshape(-1, 4)).to(DEV)
    seg = ops.seg_offsets_uniform(T, N, torch.device(DEV))
    succ, best = ops.link_frames(db, seg, N)
    succ_t = succ[:(T - 1) * N].view(T - 1, N).long()
    best_t = best[:(T - 1) * N].view(T - 1, N)
    # every successor lies in the next frame; IoU in [0, 1]
    lo = (torch.arange(1, T, device=DEV) * N)[:, None]
    assert bool(((succ_t >= lo) & (succ_t < lo + N)).all())
    assert bool(((best_t >= 0) & (best_t <= 1)).all())
    assert bool((succ[(T - 1) * N:] == -1).all())
    # best is the row maximum of the IoU matrix and succ its first arg-max (sampled frame pairs)
    for t in (0, 2499, 4998):
        m = ops.iou_matrix(db[t * N:(t + 1) * N], db[(t + 1) * N:(t + 2) * N])
        mx, am = m.max(dim=1)
        assert torch.equal(best_t[t], mx)
        first = (m == mx[:, None]).float().argmax(dim=1)
        assert torch.equal(succ_t[t] - (t + 1) * N, first)
    os_, ob = c_oracle.link_f32(b[:3])
    assert np.array_equal(succ_t[:2].cpu().numpy() - (np.arange(1, 3) * N)[:, None], os_)
    assert np.array_equal(best_t[:2].cpu().numpy(), ob)
    # sharding property: linking two halves with the boundary frame as halo gives the same answer
    h = T // 2
    s1, b1 = ops.link_frames(db[:h * N], ops.seg_offsets_uniform(h, N, torch.device(DEV)), N, db[h * N:(h + 1) * N])
    assert torch.equal(b1, best[:h * N])
    assert torch.equal(s1[(h - 1) * N:], succ[(h - 1) * N:h * N] - h * N)


def test_config4_temporal_30x10000():
    """Temporal smoothing of 30-class x 10000-frame tubelets (config 4: 256 tubelets per GPU)."""
    K, L = 256 * 30, 10000
    base = synth.score_rows(96, L, seed=44, missing_frac=0.05)
    x = torch.from_numpy(base).to(DEV).repeat(K // 96, 1).contiguous()
    x[7] = -1e5; x[7, 5000] = 0.5                                   # one row with a single valid score
    orig = x.clone()
    ops.raise_for_status(ops.score_completion_(x))
    assert bool((x > -10).all())
    keep = orig > -10
    assert torch.equal(x[keep], orig[keep])                          # valid scores untouched
    assert bool((x[7] == 0.5).all())
    # filled values stay inside the range of their row (interpolation / replication)
    assert bool((x.amax(1) <= orig.amax(1)).all()) and bool((x.amin(1) >= torch.where(keep, orig, torch.ones_like(orig)).amin(1)).all())
    for k in (0, 95, K - 1):
        assert np.allclose(x[k].cpu().numpy(), oracle_np.completion_row(base[k % 96].astype(np.float64)), rtol=0, atol=1e-5)
    m3, m5, m9 = (ops.temporal_maxpool(x, w) for w in (3, 5, 9))
    assert bool((m3 >= x).all()) and bool((m5 >= m3).all()) and bool((m9 >= m5).all())
    assert torch.equal(ops.temporal_maxpool(m3, 3), m5)
    assert np.array_equal(m9[3].cpu().numpy().astype(np.float64), oracle_np.temporal_maxpool_row(x[3].cpu().numpy().astype(np.float64), 9))
    taps = torch.from_numpy(synth.gaussian_taps(30, 9)).to(DEV)
    y = ops.temporal_conv1d(x, taps)
    # linearity: conv(a*x) == a*conv(x) for a power of two (exact in floating point)
    assert torch.equal(ops.temporal_conv1d(x * 2.0, taps), y * 2.0)
    want = oracle_np.temporal_conv1d(x[31:32].cpu().numpy(), synth.gaussian_taps(30, 9)[1:2])
    assert np.array_equal(y[31].cpu().numpy(), want[0])


def test_config5_video_2000x2000x30():
    """One video of config 5 (2000 frames x 2000 boxes x 30 classes) = one GPU's share."""
    T, N, C, thr = 2000, 2000, 30, 0.3
    dev = torch.device(DEV)
    db = torch.empty((T * N, 4), dtype=torch.float32, device=dev)
    ds = torch.empty((T * N, C), dtype=torch.float32, device=dev)
    host = {}
    for t0 in range(0, T, 250):                                     # generate in slabs (host memory)
        b, s = synth.boxes_scores(250, N, C, seed=500 + t0)
        db[t0 * N:(t0 + 250) * N] = torch.from_numpy(b.reshape(-1, 4)).to(dev)
        ds[t0 * N:(t0 + 250) * N] = torch.from_numpy(s.reshape(-1, C)).to(dev)
        if t0 in (0, 1750):
            host[t0] = (b[:2], s[:2])
    seg = ops.seg_offsets_uniform(T, N, dev)
    keep_idx, keep_cnt, keep_mask, status = ops.nms_frames(db, ds, seg, thr, N, want_mask=True, frame_major_out=True)
    assert ops.raise_for_status(status) == 0
    km = keep_mask.view(T, C, N)
    assert torch.equal(km.sum(-1).to(torch.int32), keep_cnt)
    ki = keep_idx.view(T, C, N)
    assert torch.equal((ki >= 0).sum(-1).to(torch.int32), keep_cnt)
    assert int(keep_cnt.min()) >= 1 and int(keep_cnt.max()) < N
    for t0, (b, s) in host.items():                                  # oracle on the slab's first two frames
        okm, oki, okc = c_oracle.nms_frames(b, s, thr)
        assert np.array_equal(keep_cnt[t0:t0 + 2].cpu().numpy(), okc)
        assert np.array_equal(km[t0:t0 + 2].cpu().numpy(), okm)
    succ, best = ops.link_frames(db, seg, N)
    assert bool(((best >= 0) & (best <= 1)).all())
    os_, ob = c_oracle.link_f32(host[0][0])
    assert np.array_equal(succ[:N].cpu().numpy() - N, os_[0]) and np.array_equal(best[:N].cpu().numpy(), ob[0])
