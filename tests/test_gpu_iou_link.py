"""GPU parity: dense IoU (f64 = utils/common.py:451-468 bit for bit; f32 = nms.pyx pair
arithmetic bit for bit), the suppression bit matrix, and the frame-to-frame link."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, oracle_np
from vdetlib_b200 import ops, synth
from vdetlib_b200.utils.common import iou as gpu_iou

import helpers

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_iou_f64_golden_and_types():
    g = helpers.golden_npz("arrays.npz")
    got = gpu_iou(g["iou_a_int"], g["iou_b_int"])
    assert got.dtype == np.float64 and np.array_equal(got, g["iou_int"])
    assert np.array_equal(gpu_iou(g["iou_a_f"], g["iou_b_f"]), g["iou_f"])
    assert np.array_equal(gpu_iou([[0, 0, 9, 9]], [[9, 0, 18, 9], [0, 0, 9, 9]]), [[10.0 / 190.0, 1.0]])
    with pytest.raises(IndexError):
        gpu_iou([[0, 0, 9, 9]], np.asarray([]))        # reference fails on the 1-D empty array too


@pytest.mark.parametrize("na,nb", [(1, 1), (1, 300), (37, 129), (64, 1024), (33, 1027), (200, 2050)])
def test_iou_matrix_vs_oracle(na, nb):
    rng = np.random.default_rng(na * 7 + nb)
    a = rng.uniform(0, 500, (na, 4)); a[:, 2:] += a[:, :2]
    b = rng.uniform(0, 500, (nb, 4)); b[:, 2:] += b[:, :2]
    want64 = oracle_np.iou(a, b)
    got64 = ops.iou_matrix(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV)).cpu().numpy()
    assert np.array_equal(got64, want64)
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    want32 = c_oracle.pair_iou_f32(a32, b32)
    got32 = ops.iou_matrix(torch.from_numpy(a32).to(DEV), torch.from_numpy(b32).to(DEV)).cpu().numpy()
    assert np.array_equal(got32, want32)
    # float32 scores within 1e-5 of the float64 reference (north_star tolerance)
    assert np.abs(got32.astype(np.float64) - oracle_np.iou(a32, b32)).max() <= 1e-5


def test_iou_matrix_large_roundtrip_property():
    """At roofline size (8192 x 8192) check symmetry and the diagonal instead of the oracle."""
    b, _ = synth.boxes_scores(1, 8192, 1, seed=5)
    x = torch.from_numpy(b[0]).to(DEV)
    m = ops.iou_matrix(x, x)
    assert torch.equal(m, m.t())
    assert bool((m.diagonal() == 1.0).all())
    sub = c_oracle.pair_iou_f32(b[0, :64], b[0, 4000:4100])
    assert np.array_equal(m[:64, 4000:4100].cpu().numpy(), sub)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 300, 1000])
def test_bitmask_vs_oracle(n):
    b, _ = synth.boxes_scores(1, n, 1, seed=n)
    for thr in (0.3, 0.5):
        mask, status = ops.iou_bitmask(torch.from_numpy(b[0]).to(DEV), thr)
        assert int(status.item()) == 0
        assert np.array_equal(mask.cpu().numpy().view(np.uint32), c_oracle.iou_bitmask(b[0], thr))


def _link_oracle_packed(b, counts):
    T, nmax = b.shape[:2]
    succ, best = c_oracle.link_f32(b, counts)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    ws, wb = [], []
    for t in range(T - 1):
        s = succ[t, :counts[t]].astype(np.int64)
        ws.append(np.where(s >= 0, s + off[t + 1], -1))
        wb.append(best[t, :counts[t]])
    return np.concatenate(ws) if ws else np.zeros(0, np.int64), np.concatenate(wb) if wb else np.zeros(0, np.float32), off


@pytest.mark.parametrize("T,N,integer", [(2, 1, False), (5, 63, False), (4, 300, False), (3, 1000, False), (3, 1500, False),
                                         (5, 300, True), (3, 700, True)])
def test_link_vs_oracle(T, N, integer):
    # integer coordinates: IoUs are ratios of small integers, so equal maxima occur (the FIRST arg-max wins)
    b, _ = synth.boxes_scores(T, N, 1, seed=T + N, integer=integer)
    counts = np.full(T, N, np.int32)
    want_s, want_b, off = _link_oracle_packed(b, counts)
    db = torch.from_numpy(b.reshape(-1, 4)).to(DEV)
    seg = ops.seg_offsets_uniform(T, N, torch.device(DEV))
    succ, best = ops.link_frames(db, seg, N)
    n_linked = (T - 1) * N
    assert np.array_equal(succ.cpu().numpy()[:n_linked], want_s)
    assert np.array_equal(best.cpu().numpy()[:n_linked], want_b)
    assert bool((succ[n_linked:] == -1).all()) and bool((best[n_linked:] == 0).all())   # no halo
    # within 1e-5 of the float64 IoU of the chosen successor
    f64 = oracle_np.iou(b[0], b[1])
    assert np.abs(best.cpu().numpy()[:N] - f64.max(axis=1)).max() <= 1e-5


def test_link_ragged_and_halo():
    counts = np.asarray([40, 0, 17, 300, 1, 64], np.int32)
    T, nmax = len(counts), 300
    b, _ = synth.boxes_scores(T + 1, nmax, 1, seed=9)
    want_s, want_b, off = _link_oracle_packed(b[:T], counts)
    rows = np.concatenate([b[t, :counts[t]] for t in range(T)])
    dev = torch.device(DEV)
    halo = b[T, :50]
    succ, best = ops.link_frames(torch.from_numpy(rows).to(dev), torch.from_numpy(off.astype(np.int32)).to(dev),
                                 int(counts.max()), torch.from_numpy(halo).to(dev))
    n_linked = int(off[T - 1])
    assert np.array_equal(succ.cpu().numpy()[:n_linked], want_s)
    assert np.array_equal(best.cpu().numpy()[:n_linked], want_b)
    # last frame vs the halo == a 2-frame link whose second frame is the halo
    two = np.zeros((2, nmax, 4), np.float32)
    two[0, :counts[-1]] = b[T - 1, :counts[-1]]
    two[1, :50] = halo
    hs, hb = c_oracle.link_f32(two, np.asarray([counts[-1], 50], np.int32))
    assert np.array_equal(succ.cpu().numpy()[n_linked:], hs[0, :counts[-1]])
    assert np.array_equal(best.cpu().numpy()[n_linked:], hb[0, :counts[-1]])


def _weird_boxes(n, seed):
    """Boxes that are NOT 'sane' for the branch-free division: huge coordinates, zero-area and
    inverted boxes (negative areas, zero unions) mixed into ordinary ones."""
    rng = np.random.default_rng(seed)
    b = helpers.unique_score_dets(rng, n, scale=400.0)[:, :4]
    b[::7] *= 2.0e4                                   # coordinates beyond 2^20
    b[3::11, 2] = b[3::11, 0] - 1.0                   # width exactly 0  -> area 0
    b[5::13, 2] = b[5::13, 0] - 30.0                  # inverted         -> negative width
    b[6::17] = b[6::17][:, [0, 1, 0, 1]] - np.asarray([0, 0, 1, 1], np.float32)   # 0 x 0 boxes
    return np.ascontiguousarray(b, np.float32)


def test_iou_link_generic_path_on_weird_boxes():
    a, b = _weird_boxes(150, 1), _weird_boxes(1100, 2)
    got = ops.iou_matrix(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV)).cpu().numpy()
    want = c_oracle.pair_iou_f32(a, b)
    assert np.array_equal(got, want, equal_nan=True)
    assert np.isnan(want).any() or np.isinf(want).any() or (want < 0).any()       # the case is really exercised
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    with np.errstate(all="ignore"):
        want64 = oracle_np.iou(a64, b64)
    got64 = ops.iou_matrix(torch.from_numpy(a64).to(DEV), torch.from_numpy(b64).to(DEV)).cpu().numpy()
    assert np.array_equal(got64, want64, equal_nan=True)
    # link over frames of weird boxes
    T, N = 4, 300
    fr = np.stack([_weird_boxes(N, 10 + t) for t in range(T)])
    want_s, want_b, off = _link_oracle_packed(fr, np.full(T, N, np.int32))
    succ, best = ops.link_frames(torch.from_numpy(fr.reshape(-1, 4)).to(DEV),
                                 ops.seg_offsets_uniform(T, N, torch.device(DEV)), N)
    assert np.array_equal(succ.cpu().numpy()[:(T - 1) * N], want_s)
    assert np.array_equal(best.cpu().numpy()[:(T - 1) * N], want_b, equal_nan=True)


def test_nms_generic_path_on_weird_boxes():
    """Thresholds outside the fast-filter range and boxes outside the sane range still match."""
    from vdetlib_b200.utils import cython_nms as gpu
    rng = np.random.default_rng(3)
    for thr in (0.0, 1e-8, 0.3, 0.999999, 1.0, 1.5, 3.0, -0.5):
        d = helpers.unique_score_dets(rng, 400)
        assert gpu.nms(d, thr) == c_oracle.nms(d, thr), thr
    d = helpers.unique_score_dets(rng, 500)
    d[::7, :4] *= 2.0e4
    d[5::13, 2] = d[5::13, 0] - 30.0          # inverted boxes: negative areas, unions can be <= 0
    for thr in (0.3, 0.5):
        try:
            want = c_oracle.nms(d, thr)
        except ZeroDivisionError:
            want = None
        if want is None:
            with pytest.raises(ZeroDivisionError):
                gpu.nms(d, thr)
        else:
            assert gpu.nms(d, thr) == want


def test_follow_links_and_gather_rows():
    """Links -> tubelet score rows (build-defined glue): against a direct NumPy walk."""
    T, N, C = 12, 90, 4
    b, s = synth.boxes_scores(T, N, C, seed=21)
    dev = torch.device(DEV)
    db = torch.from_numpy(b.reshape(-1, 4)).to(dev)
    ds = torch.from_numpy(s.reshape(-1, C)).to(dev)
    succ, best = ops.link_frames(db, ops.seg_offsets_uniform(T, N, dev), N)
    start = torch.arange(0, N, dtype=torch.int32, device=dev)
    for min_iou in (0.0, 0.2):
        rows = ops.follow_links(succ, best, start, T, min_iou).cpu().numpy()
        sn, bn = succ.cpu().numpy(), best.cpu().numpy()
        want = np.full((T, N), -1, np.int32)
        for k in range(N):
            r = k
            for t in range(T):
                want[t, k] = r
                if r >= 0:
                    r = sn[r] if (sn[r] >= 0 and bn[r] >= np.float32(min_iou)) else -1
        assert np.array_equal(rows, want)
        out = ops.gather_chain_scores(ds, torch.from_numpy(rows).to(dev)).cpu().numpy()
        sflat = s.reshape(-1, C)
        ref = np.where(want.T[:, None, :] >= 0, sflat[np.maximum(want.T, 0)].transpose(0, 2, 1), np.float32(-1e5))
        assert np.array_equal(out, ref)


@pytest.mark.parametrize("T,N,seed", [(6, 300, 1), (4, 512, 2), (3, 1000, 3), (3, 2000, 4), (5, 640, 5), (3, 2048, 6)])
def test_x_sorted_link_equals_the_full_scan(T, N, seed):
    """link_sorted.cu (frames sorted by x1, only the pairs that can overlap in x) against link.cu's full scan and the
    C port: ragged frames, duplicated boxes (equal IoUs: the FIRST arg-max in original order must win), boxes that
    touch nothing (successor = box 0, IoU 0), a halo with a device-side count, an insane box (falls back per pair)."""
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda")
    b, _ = synth.boxes_scores(T + 1, N, 1, seed=600 + seed)
    counts = rng.integers(max(N // 2, 1), N + 1, T).astype(np.int32)
    counts[rng.integers(0, T)] = N
    for t in range(1, T):                                        # duplicates in the next frame -> tied IoUs
        k = min(int(counts[t]) // 3, 40)
        src = rng.integers(0, counts[t], k)
        dst = rng.integers(0, counts[t], k)
        b[t, dst] = b[t, src]
    b[0, 0] = np.asarray([5000.0, 5000.0, 5010.0, 5010.0], np.float32)      # overlaps nothing
    rows = np.concatenate([b[t, :counts[t]] for t in range(T)])
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    halo_cap = N
    n_halo = int(rng.integers(1, N + 1))
    halo = b[T, :halo_cap].copy()
    d_rows, d_off = torch.from_numpy(rows).to(dev), torch.from_numpy(off).to(dev)
    d_halo = torch.from_numpy(halo).to(dev)
    d_cnt = torch.tensor([n_halo], dtype=torch.int32, device=dev)
    for insane in (False, True):
        if insane:
            rows2 = rows.copy()
            rows2[off[1] + 1] = np.asarray([10.0, 10.0, 9.0, 30.0], np.float32)      # zero width: not sane
            d_rows = torch.from_numpy(rows2).to(dev)
            frames = [rows2[off[t]:off[t + 1]] for t in range(T)]
        else:
            frames = [rows[off[t]:off[t + 1]] for t in range(T)]
        s1, i1 = ops.link_frames(d_rows, d_off, N, d_halo, halo_row_base=len(rows), halo_count=d_cnt)
        s0, i0 = ops.link_frames(d_rows, d_off, N, d_halo, halo_row_base=len(rows), halo_count=d_cnt, ws=False)
        assert np.array_equal(s1.cpu().numpy(), s0.cpu().numpy()), insane
        assert np.array_equal(i1.cpu().numpy().view(np.uint32), i0.cpu().numpy().view(np.uint32)), insane
        if not insane:
            su, io = s1.cpu().numpy(), i1.cpu().numpy()
            for t in range(T):
                nxt = frames[t + 1] if t + 1 < T else halo[:n_halo]
                base = off[t + 1] if t + 1 < T else len(rows)
                iou = c_oracle.pair_iou_f32(frames[t], nxt)
                assert np.array_equal(su[off[t]:off[t + 1]], base + np.argmax(iou, axis=1)), t
                assert np.array_equal(io[off[t]:off[t + 1]], iou.max(axis=1)), t
            assert su[0] == off[1] and io[0] == 0.0                      # the far-away box: first box of the next frame
