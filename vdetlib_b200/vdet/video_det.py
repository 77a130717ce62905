"""Drop-in for the post-CNN part of vdetlib's ``vdet.video_det`` (reference vdet/video_det.py).

``apply_vid_nms`` keeps the reference signature and proto-in / proto-out behaviour (:51-61).
``VideoPostProcessor`` is the packed-tensor entry for whole videos: all 30 classes of every
frame in one launch on class-shared boxes, plus the frame-to-frame link (the path
BASELINE.json benchmarks); it avoids the proto-dict walk that dominates the reference's host time.
"""
import copy
import logging

import numpy as np
import torch

from .. import _lib, ops
from ..utils.cython_nms import vid_nms
from ..utils.protocol import det_score


def apply_vid_nms(det_proto, class_index, thres=0.3):
    """Per-frame NMS of one class over a whole det proto.  vdet/video_det.py:51-61.

    As in the reference the threshold is hard-coded to 0.3 (:57) -- ``thres`` is ignored -- and
    the returned detections are the SAME dict objects, in global descending-score order."""
    logging.info('Apply NMS on video: {}'.format(det_proto['video']))
    new_det = {}
    new_det['video'] = det_proto['video']
    boxes = np.asarray([[det['frame'], ] + list(det['bbox']) + [det_score(det, class_index), ]
                        for det in det_proto['detections']], dtype='float32').reshape(-1, 6)
    keep = vid_nms(boxes, thresh=0.3)
    new_det['detections'] = copy.copy([det_proto['detections'][i] for i in keep])
    logging.info("{} / {} windows kept.".format(len(new_det['detections']), len(det_proto['detections'])))
    return new_det


def threshold_topk_frames(scores, boxes, thresh=0.05, max_per_image=100):
    """Post-CNN per-class score floor + cap for a stack of frames (vdet/video_det.py:88-100).

    scores [T, R, C] float32 (class 0 = background), boxes [T, R, 4*C] float32.
    Returns ``all_boxes[cls][frame]`` = float32 [K,5] arrays like the reference's
    ``fast_rcnn_det_vid`` (entries of class 0 are empty lists)."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    boxes = np.asarray(boxes, dtype=np.float32)
    T, R, C = scores.shape
    dev_scores = torch.from_numpy(scores.reshape(T * R, C)).to(ops.default_device())
    seg = ops.seg_offsets_uniform(T, R, dev_scores.device)
    idx, cnt = ops.threshold_topk(dev_scores, seg, R, thresh, max_per_image)
    idx = idx.cpu().numpy()
    cnt = cnt.cpu().numpy()
    all_boxes = [[[] for _ in range(T)] for _ in range(C)]
    for t in range(T):
        for j in range(1, C):
            sel = idx[t, j, :cnt[t, j]]
            all_boxes[j][t] = np.hstack((boxes[t, sel, j * 4:(j + 1) * 4],
                                         scores[t, sel, j][:, np.newaxis])).astype(np.float32, copy=False)
    return all_boxes


class _Slot(object):
    """Device inputs / outputs, pinned result buffers and the completion event of ONE in-flight step."""

    def __init__(self, pp, stage_set=0):
        T, N, C, dev = pp.T, pp.N, pp.C, pp.device
        rows = T * N
        self.stage_set = stage_set                          # which pinned upload buffers this slot reads
        self.h_boxes, self.h_scores = pp.h_boxes_sets[stage_set], pp.h_scores_sets[stage_set]
        self.d_boxes = torch.empty((rows, 4), dtype=torch.float32, device=dev)
        self.d_scores = torch.empty((rows, C), dtype=torch.float32, device=dev)
        self.d_idx = torch.empty(rows * C, dtype=torch.int32, device=dev)
        self.d_mask = torch.empty(rows * C, dtype=torch.uint8, device=dev)
        self.d_cnt = torch.empty((T, C), dtype=torch.int32, device=dev)
        self.d_succ = torch.empty(rows, dtype=torch.int32, device=dev)
        self.d_iou = torch.empty(rows, dtype=torch.float32, device=dev)
        self.h_mask = torch.empty(rows * C, dtype=torch.uint8).pin_memory()
        self.h_cnt = torch.empty((T, C), dtype=torch.int32).pin_memory()
        self.h_succ = torch.empty(rows, dtype=torch.int32).pin_memory()
        self.h_iou = torch.empty(rows, dtype=torch.float32).pin_memory()
        self.status = ops.new_status(dev)
        self.h_status = torch.zeros(1, dtype=self.status.dtype).pin_memory()
        self.ev_boxes = torch.cuda.Event()
        self.ev_link = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.busy = False
        self.graph = None                                   # whole-step CUDA graph (single rank)
        self.launch_stream = torch.cuda.Stream(device=dev)  # the graph of this slot is launched here
        # frames of more than 1024 boxes keep their bit matrix in a global scratch slot per CTA
        ws_bytes = _lib.load().vdet_nms_frames_workspace_bytes(N, C, dev.index or 0) if N > 1024 else 0
        self.ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
        ws_ptr = self.ws.data_ptr() if ws_bytes else None
        # per pipeline chunk: views of every buffer + the prebuilt argument list of the NMS launch
        self.chunks = []
        for f0, f1 in pp.chunks:
            r0, r1 = f0 * N, f1 * N
            d_sc = self.d_scores[r0:r1]
            d_idx, d_mask, d_cnt = self.d_idx[r0 * C:r1 * C], self.d_mask[r0 * C:r1 * C], self.d_cnt[f0:f1]
            seg = pp.chunk_seg[f1 - f0]
            nms_args = (self.d_boxes[r0:r1].data_ptr(), 4, d_sc.data_ptr(), C, 1, seg.data_ptr(), f1 - f0, N, None, C,
                        pp.nms_thresh, d_idx.data_ptr(), d_cnt.data_ptr(), d_mask.data_ptr(), r1 - r0,
                        _lib.LAYOUT_FRAME_MAJOR, self.status.data_ptr(), ws_ptr, ws_bytes)
            self.chunks.append({
                "d_scores": d_sc, "h_scores": self.h_scores[r0:r1], "nms_args": nms_args,
                "d_mask": d_mask, "h_mask": self.h_mask[r0 * C:r1 * C], "d_cnt": d_cnt, "h_cnt": self.h_cnt[f0:f1],
                "ev_in": torch.cuda.Event(), "ev_nms": torch.cuda.Event()})

    def host_views(self, pp):
        return {"keep_mask": self.h_mask.numpy().reshape(pp.T, pp.C, pp.N), "keep_cnt": self.h_cnt.numpy(),
                "succ": self.h_succ.numpy(), "link_iou": self.h_iou.numpy()}


class VideoPostProcessor(object):
    """NMS (all classes) + frame-to-frame link for one video shard of fixed shape.

    Input: boxes [T, N, 4] float32 and scores [T, N, C] float32 on the HOST (any array-like;
    copied through pinned staging buffers) or already on the device.  Output (device tensors from
    :meth:`run_device`, host arrays from :meth:`run_host` / :meth:`run_staged` / :meth:`collect`),
    frame-major:

      keep_mask [T, C, N] uint8   1 = detection survives per-frame NMS for that class
      keep_cnt  [T, C]    int32   survivors per (frame, class)
      keep_idx  [T, C, N] int32   surviving rows of the frame in descending score, -1 padded
                                  (device only; not copied back by default)
      succ      [T*N]     int32   packed row of the best-IoU box in the next frame (-1: none)
      link_iou  [T*N]     float32 that IoU

    ``halo`` (boxes of the first frame of the NEXT shard, [H,4]) links the shard's last frame
    across a shard boundary (see vdetlib_b200.dist).

    The host path is pipelined twice over.  Inside a step the shard is cut into ``n_chunks`` frame
    ranges; chunk k's host->device copy (copy stream), NMS (compute stream) and result read-back
    (read-back stream) overlap with the neighbouring chunks', so PCIe in both directions and the SMs
    are busy at the same time (frame-major outputs make every chunk a contiguous byte range).
    Across steps, :meth:`submit_staged` / :meth:`collect` keep up to ``n_slots`` steps in flight, each
    with its own device and result buffers: the upload of step k+1 starts while the last NMS chunk
    and read-back of step k are still running, so the host->device link -- the bound of this path --
    never idles.  On a single rank a whole step can also be replayed from one CUDA graph
    (``graph=True``): one launch per step instead of ~50 stream operations.
    """

    def __init__(self, n_frames, n_boxes, n_classes, nms_thresh=0.3, device=None, n_chunks=8, n_slots=2, n_stage=1):
        self.T, self.N, self.C = int(n_frames), int(n_boxes), int(n_classes)
        self.nms_thresh = float(nms_thresh)
        self.device = device or ops.default_device()
        T, N, C = self.T, self.N, self.C
        rows = T * N
        dev = self.device
        self.seg_offsets = ops.seg_offsets_uniform(T, N, dev)
        # pinned upload buffers: one set shared by every slot (n_stage=1: the staged shard can be submitted any
        # number of times), or one set per slot (n_stage=n_slots: shard k+1 is staged while shard k is in flight)
        n_slots = max(1, int(n_slots))
        if int(n_stage) not in (1, n_slots):
            raise ValueError("n_stage must be 1 or n_slots")
        self.n_stage = int(n_stage)
        self.h_boxes_sets = [torch.empty((rows, 4), dtype=torch.float32).pin_memory() for _ in range(self.n_stage)]
        self.h_scores_sets = [torch.empty((rows, C), dtype=torch.float32).pin_memory() for _ in range(self.n_stage)]
        self.h_boxes, self.h_scores = self.h_boxes_sets[0], self.h_scores_sets[0]
        # frame ranges of the pipeline chunks
        n_chunks = max(1, min(int(n_chunks), T))
        edges = [round(k * T / n_chunks) for k in range(n_chunks + 1)]
        self.chunks = [(edges[k], edges[k + 1]) for k in range(n_chunks) if edges[k + 1] > edges[k]]
        self.chunk_seg = {f1 - f0: ops.seg_offsets_uniform(f1 - f0, N, dev) for f0, f1 in self.chunks}
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self.slots = [_Slot(self, k % self.n_stage) for k in range(n_slots)]
        self._next_slot = 0
        self._graphs = {}
        self._lib = _lib.load()

    # slot 0 doubles as the buffer set of the synchronous / device-resident entry points
    d_boxes = property(lambda self: self.slots[0].d_boxes)
    d_scores = property(lambda self: self.slots[0].d_scores)
    d_idx = property(lambda self: self.slots[0].d_idx)
    d_mask = property(lambda self: self.slots[0].d_mask)
    d_cnt = property(lambda self: self.slots[0].d_cnt)
    d_succ = property(lambda self: self.slots[0].d_succ)
    d_iou = property(lambda self: self.slots[0].d_iou)
    h_mask = property(lambda self: self.slots[0].h_mask)
    h_cnt = property(lambda self: self.slots[0].h_cnt)
    h_succ = property(lambda self: self.slots[0].h_succ)
    h_iou = property(lambda self: self.slots[0].h_iou)
    status = property(lambda self: self.slots[0].status)
    h_status = property(lambda self: self.slots[0].h_status)

    # bytes crossing PCIe per staged step
    @property
    def h2d_bytes(self):
        return self.h_boxes.numel() * 4 + self.h_scores.numel() * 4

    @property
    def d2h_bytes(self):
        sl = self.slots[0]
        return sl.h_mask.numel() + sl.h_cnt.numel() * 4 + sl.h_succ.numel() * 4 + sl.h_iou.numel() * 4

    def _views(self, out):
        T, N, C = self.T, self.N, self.C
        return {"keep_idx": out[0].view(T, C, N), "keep_cnt": out[1], "keep_mask": out[2].view(T, C, N)}

    def _launch(self, d_boxes, d_scores, halo):
        out = ops.nms_frames(d_boxes, d_scores, self.seg_offsets, self.nms_thresh, self.N, want_mask=True,
                             status=self.status, frame_major_out=True, out=(self.d_idx, self.d_cnt, self.d_mask))
        ops.link_frames(d_boxes, self.seg_offsets, self.N, halo, out=(self.d_succ, self.d_iou))
        return out

    def run_device(self, d_boxes, d_scores, halo=None, graph=False):
        """Device tensors in ([T*N,4], [T*N,C]), device tensors out; asynchronous.

        ``graph=True`` replays the two launches (NMS, link) from a CUDA graph captured on first use for
        this pair of input buffers -- no per-step launch overhead or inter-kernel gap.  Outputs always
        live in the processor's own buffers (slot 0; valid until the next call)."""
        if not graph:
            out = self._launch(d_boxes, d_scores, halo)
        else:
            key = (d_boxes.data_ptr(), d_scores.data_ptr(), 0 if halo is None else halo.data_ptr())
            g = self._graphs.get(key)
            if g is None:
                self._launch(d_boxes, d_scores, halo)                    # warm up outside the capture
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch(d_boxes, d_scores, halo)
                self._graphs[key] = g
            g.replay()
            out = (self.d_idx, self.d_cnt, self.d_mask)
        res = self._views(out)
        res.update(succ=self.d_succ, link_iou=self.d_iou)
        return res

    def stage(self, boxes, scores):
        """Copy host arrays into the pinned upload buffers of the NEXT step to be submitted (not part of the
        timed region) with streaming stores: lines written with ordinary stores stay dirty in the CPU caches and
        the copy engine then reads them at roughly half the PCIe rate (profiles/r01_pcie.md).
        With one staging set (``n_stage=1``) every step in flight reads these buffers: collect all outstanding
        steps before restaging.  With a set per slot only the slot about to be reused must have been collected."""
        b = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
        s = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1, self.C)
        if b.shape[0] != self.h_boxes.shape[0] or s.shape[0] != self.h_scores.shape[0]:
            raise ValueError("stage: expected %d rows" % self.h_boxes.shape[0])
        if not any(sl.busy for sl in self.slots):
            self._next_slot = 0
        target = self.slots[self._next_slot]
        if any(sl.busy and sl.stage_set == target.stage_set for sl in self.slots):
            raise RuntimeError("stage: a submitted step still reads the staging buffers; collect() it first")
        ops.host_copy_stream(target.h_boxes, b)
        ops.host_copy_stream(target.h_scores, s)

    def _enqueue(self, sl, halo, halo_fn, fork):
        """Stream operations of one staged step on slot ``sl``.  The boxes (4 floats/row) go first, so
        the link -- which needs no scores -- and its D2H run under the upload of the scores
        (C floats/row); each score chunk is followed by its NMS launch and the download of its keep
        mask.  ``fork``: branch the copy streams off the current stream and join them again (what a
        stream capture needs); otherwise the copy streams run free and ``sl.done`` marks the end."""
        N = self.N
        cur = torch.cuda.current_stream()
        s_in, s_out = self.s_in, self.s_out
        if fork:
            s_in.wait_stream(cur)
            s_out.wait_stream(cur)
        with torch.cuda.stream(s_in):
            sl.d_boxes.copy_(sl.h_boxes, non_blocking=True)
            sl.ev_boxes.record(s_in)
            for ch in sl.chunks:
                ch["d_scores"].copy_(ch["h_scores"], non_blocking=True)
                ch["ev_in"].record(s_in)
        cur.wait_event(sl.ev_boxes)
        if halo_fn is not None:
            halo = halo_fn(sl.d_boxes[:N])
        ops.link_frames(sl.d_boxes, self.seg_offsets, N, halo, out=(sl.d_succ, sl.d_iou))
        sl.ev_link.record(cur)
        s_out.wait_event(sl.ev_link)
        with torch.cuda.stream(s_out):
            sl.h_succ.copy_(sl.d_succ, non_blocking=True)
            sl.h_iou.copy_(sl.d_iou, non_blocking=True)
        nms = self._lib.vdet_nms_frames_f32
        stream_ptr = cur.cuda_stream
        for ch in sl.chunks:
            cur.wait_event(ch["ev_in"])
            _lib.check(nms(*ch["nms_args"], stream_ptr), "nms_frames")
            ch["ev_nms"].record(cur)
            s_out.wait_event(ch["ev_nms"])
            with torch.cuda.stream(s_out):
                ch["h_mask"].copy_(ch["d_mask"], non_blocking=True)
                ch["h_cnt"].copy_(ch["d_cnt"], non_blocking=True)
        with torch.cuda.stream(s_out):
            sl.h_status.copy_(sl.status, non_blocking=True)
        if fork:
            cur.wait_stream(s_in)
            cur.wait_stream(s_out)
        else:
            sl.done.record(s_out)

    def _capture(self, sl):
        """Capture the whole staged step of slot ``sl`` (copies on three streams + kernels) into one graph."""
        with torch.cuda.stream(sl.launch_stream):
            self._enqueue(sl, None, None, fork=True)      # warm-up: kernel attributes are set outside the capture
        sl.launch_stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=sl.launch_stream):
            self._enqueue(sl, None, None, fork=True)
        sl.graph = g

    def submit_staged(self, halo=None, halo_fn=None, graph=False):
        """Enqueue one step on the next free slot and return a ticket for :meth:`collect`; does not block.
        Up to ``n_slots`` steps may be in flight.  ``halo_fn(d_first_frame_boxes)`` (optional) is called
        on the compute stream once the boxes are on the device and returns the halo tensor (boundary
        exchange).  ``graph=True`` (no halo) replays the step from a CUDA graph captured on first use."""
        k = self._next_slot
        sl = self.slots[k]
        if sl.busy:
            raise RuntimeError("submit_staged: all %d slots are in flight; collect() the oldest step first"
                               % len(self.slots))
        if graph and halo is None and halo_fn is None:
            if sl.graph is None:
                self._capture(sl)
            with torch.cuda.stream(sl.launch_stream):
                sl.graph.replay()
                sl.done.record(sl.launch_stream)
        else:
            self._enqueue(sl, halo, halo_fn, fork=False)
        sl.busy = True
        self._next_slot = (k + 1) % len(self.slots)
        return k

    def collect(self, ticket):
        """Wait for the step behind ``ticket``; returns host views of its pinned result buffers
        (valid until the slot is submitted again) after translating the status word."""
        sl = self.slots[ticket]
        if not sl.busy:
            raise RuntimeError("collect: ticket %r is not in flight" % (ticket,))
        sl.done.synchronize()
        sl.busy = False
        ops.raise_for_status_word(int(sl.h_status.item()))
        return sl.host_views(self)

    def run_staged(self, halo=None, halo_fn=None, graph=False):
        """One synchronous step from the staged inputs: submit + collect (on slot 0 when nothing is in flight,
        so that ``d_boxes`` / ``d_scores`` hold the step's inputs afterwards)."""
        if not any(sl.busy for sl in self.slots):
            self._next_slot = 0
        return self.collect(self.submit_staged(halo, halo_fn, graph))

    def run_host(self, boxes, scores, halo=None):
        """The user-facing call: host arrays in, host arrays out."""
        self.stage(boxes, scores)
        return self.run_staged(halo)
