"""Test-only stand-ins that let VideoPostProcessor's HOST logic run without a GPU: streams and events that do
nothing (every copy on CPU tensors is immediate, so program order is the only order), ``pin_memory`` as the
identity, and the two launches of the staged step restated with the oracle -- ``ops.link_frames`` and the raw
C-ABI call ``vdet_nms_frames_f32`` (which receives plain addresses: here they point at CPU tensors and are read
back with ctypes).  What a run on these fakes checks is the bookkeeping of the pipeline: which buffer of which
slot is copied where, chunk offsets, tickets, the staging rules, frame-major views.  The asynchronous behaviour
and the kernels are the business of the ``-m gpu`` tests."""
import contextlib
import ctypes

import numpy as np
import torch

from oracle import c_oracle


class FakeStream(object):
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass


class FakeEvent(object):
    def __init__(self, *a, **k):
        self.recorded = 0

    def record(self, stream=None):
        self.recorded += 1

    def synchronize(self):
        pass


def _view(ptr, dtype, count):
    if count == 0:
        return np.zeros(0, dtype)
    buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype, count=count)


class FakeLib(object):
    """The C-ABI entries VideoPostProcessor calls directly, on host addresses."""

    def __init__(self):
        self.nms_calls = []

    def vdet_nms_frames_workspace_bytes(self, max_seg_len, n_classes, device):
        return 0

    def vdet_nms_frames_f32(self, boxes, box_ld, scores, ldr, ldc, seg, n_segs, max_len, row_ids, C, thresh,
                            keep_idx, keep_cnt, keep_mask, n_rows, layout, status, ws, ws_bytes, stream):
        assert box_ld == 4 and ldc == 1 and ldr == C and row_ids is None and layout == 1     # frame-major
        off = _view(seg, np.int32, n_segs + 1).copy()           # ABSOLUTE rows: base pointers are the shard's
        n = int(off[-1])
        assert n <= n_rows
        b = _view(boxes, np.float32, n * 4).reshape(n, 4)
        s = _view(scores, np.float32, n * C).reshape(n, C)
        ki = _view(keep_idx, np.int32, n * C)
        km = _view(keep_mask, np.uint8, n * C) if keep_mask else None
        kc = _view(keep_cnt, np.int32, n_segs * C).reshape(n_segs, C)
        self.nms_calls.append((int(boxes), n_segs))
        for f in range(n_segs):
            a, e = int(off[f]), int(off[f + 1])
            m = e - a
            for c in range(C):
                d = np.concatenate([b[a:e], s[a:e, c:c + 1]], axis=1).astype(np.float32)
                try:
                    k = np.asarray(c_oracle.nms(d, thresh), dtype=np.int64)
                except ZeroDivisionError:                       # nms.pyx:64 -> the device status word
                    _view(status, np.uint32, 1)[0] |= 1
                    k = np.zeros(0, np.int64)
                blk = a * C + c * m
                ki[blk:blk + m] = -1
                ki[blk:blk + len(k)] = a + k
                if km is not None:
                    km[blk:blk + m] = 0
                    km[blk + k] = 1
                kc[f, c] = len(k)
        return 0

    def vdet_compact_keep(self, keep_idx, keep_cnt, seg, n_segs, max_seg_len, C, out_dtype, keep_off, keep_off_mirror,
                          keep_out, keep_bits, stream):
        """include/vdet_b200.h: padded frame-major blocks -> one contiguous list + prefix offsets (+ bit masks)."""
        off = _view(seg, np.int32, n_segs + 1)
        cnt = _view(keep_cnt, np.int32, n_segs * C)
        pre = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
        for dst in (keep_off, keep_off_mirror):
            if dst:
                _view(dst, np.int32, n_segs * C + 1)[:] = pre
        ki = _view(keep_idx, np.int32, int(off[-1]) * C)
        out = _view(keep_out, np.uint16 if out_dtype == 0 else np.int32, max(int(pre[-1]), 1))
        words = (max_seg_len + 31) // 32
        bits = _view(keep_bits, np.uint32, n_segs * C * words).reshape(n_segs * C, words) if keep_bits else None
        for f in range(n_segs):
            a, m = int(off[f]), int(off[f + 1] - off[f])
            for c in range(C):
                k = f * C + c
                rows = ki[a * C + c * m:a * C + c * m + cnt[k]]
                out[pre[k]:pre[k + 1]] = (rows - a) if out_dtype == 0 else rows
                if bits is not None:
                    bits[k] = 0
                    for loc in rows - a:
                        bits[k, loc >> 5] |= np.uint32(1 << (loc & 31))
        return 0

    def vdet_last_error(self):
        return b""


def link_frames(boxes, seg_offsets, max_seg_len, halo=None, halo_row_base=0, out=None, halo_count=None, ws=None):
    """ops.link_frames on CPU tensors: FIRST arg-max IoU box of the next frame (halo for the last frame)."""
    b, off = boxes.numpy(), seg_offsets.numpy()
    n = b.shape[0]
    succ = np.full(n, -1, np.int32)
    best = np.zeros(n, np.float32)
    S = len(off) - 1
    for f in range(S):
        a, e = off[f], off[f + 1]
        if f < S - 1:
            nxt, base = b[off[f + 1]:off[f + 2]], int(off[f + 1])
        else:
            nxt, base = (halo.numpy() if halo is not None else np.zeros((0, 4), np.float32)), int(halo_row_base)
            if halo_count is not None:
                nxt = nxt[:max(0, min(int(halo_count.numpy().ravel()[0]), len(nxt)))]
        if e > a and len(nxt):
            iou = c_oracle.pair_iou_f32(b[a:e], nxt)
            succ[a:e] = base + np.argmax(iou, axis=1)
            best[a:e] = iou.max(axis=1)
    if out is not None:
        out[0][:n].copy_(torch.from_numpy(succ))
        out[1][:n].copy_(torch.from_numpy(best))
        return out
    return torch.from_numpy(succ), torch.from_numpy(best)


def _nms_frames_via(lib):
    def nms_frames(boxes, scores, seg_offsets, thresh, max_seg_len, row_ids=None, want_mask=False,
                   status=None, class_major=False, frame_major_out=False, out=None):
        """ops.nms_frames in the form the sharded device step uses it: frame-major, caller's buffers."""
        assert frame_major_out and out is not None and row_ids is None and not class_major
        n, C = scores.shape
        lib.vdet_nms_frames_f32(boxes.data_ptr(), 4, scores.data_ptr(), C, 1, seg_offsets.data_ptr(),
                                seg_offsets.numel() - 1, max_seg_len, None, C, float(thresh), out[0].data_ptr(),
                                out[1].data_ptr(), out[2].data_ptr(), n, 1, status.data_ptr(), None, 0, 0)
        return out[0], out[1], out[2], status
    return nms_frames


def install(monkeypatch):
    """Patch torch.cuda / ops / _lib so that VideoPostProcessor(device=cpu) runs; returns the FakeLib."""
    from vdetlib_b200 import _lib, ops
    lib = FakeLib()
    monkeypatch.setattr(ops, "nms_frames", _nms_frames_via(lib))
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(ops, "default_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(ops, "link_frames", link_frames)
    monkeypatch.setattr(ops, "link_workspace_bytes", lambda *a, **k: 64)
    real_load = _lib.load

    class _Both(object):                      # the real library for host-only entries, the fake for launches
        def __getattr__(self, name):
            return getattr(lib, name) if hasattr(lib, name) else getattr(real_load(), name)
    monkeypatch.setattr(_lib, "load", lambda: _Both())
    return lib
