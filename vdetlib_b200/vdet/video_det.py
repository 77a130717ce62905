"""Drop-in for the post-CNN part of vdetlib's ``vdet.video_det`` (reference vdet/video_det.py).

``apply_vid_nms`` keeps the reference signature and proto-in / proto-out behaviour (:51-61).
``VideoPostProcessor`` is the packed-tensor entry for whole videos: all 30 classes of every
frame in one launch on class-shared boxes, plus the frame-to-frame link (the path
BASELINE.json benchmarks); it avoids the proto-dict walk that dominates the reference's host time.
"""
import copy
import logging
import time

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
    dets = det_proto['detections']

    def score_of(det):
        # det_score (utils/protocol.py:323-327) scans the det's score list for class_index; protos written by
        # vdetlib list the classes in order, so the entry at position class_index - 1 is usually the one
        sc = det['scores']
        if 0 < class_index <= len(sc) and sc[class_index - 1]['class_index'] == class_index:
            return sc[class_index - 1]['score']
        return det_score(det, class_index)
    boxes = np.empty((len(dets), 6), dtype='float32')              # the [M,6] float32 matrix of :53-56, column-wise
    if len(dets):
        boxes[:, 0] = [det['frame'] for det in dets]
        boxes[:, 1:5] = [det['bbox'] for det in dets]
        boxes[:, 5] = [score_of(det) for det in dets]
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


class StepResult(dict):
    """Host results of one step (views of the slot's pinned buffers, valid until the slot is reused):

      keep_off  [T*C + 1] int32  prefix offsets of the keep lists in (frame, class) order
      keep_idx  [sum K]   uint16 kept boxes of every (frame, class) as indices WITHIN the frame, each list in
                                 descending score -- the order utils/nms.pyx:43-66 returns per problem
      keep_cnt  [T, C]    int32  survivors per (frame, class) (= diff of keep_off)
      keep_bits [T*C, W]  uint32 (only with ``want_bits``) bit i of block (t, c): box i of frame t kept
      succ      [rows]    int32  packed row of the best-IoU box in the next frame; -1: none; >= rows:
                                 rows + index into the next shard's first frame (the chain leaves this shard)
      link_iou  [rows]    float32 that IoU
    """

    def keep_list(self, t, c):
        """Kept boxes of frame ``t``, class ``c`` (indices within the frame, descending score)."""
        k = t * self["n_classes"] + c
        return self["keep_idx"][self["keep_off"][k]:self["keep_off"][k + 1]]

    def keep_mask(self):
        """uint8 [T, C, N] mask rebuilt on the host from the keep lists (uniform frames; a convenience for
        tests and small inputs -- the lists ARE the result)."""
        T, C, N = self["n_frames"], self["n_classes"], self["max_boxes"]
        cnt = self["keep_cnt"].reshape(-1)
        blk = np.repeat(np.arange(T * C, dtype=np.int64), cnt)
        m = np.zeros((T * C, N), np.uint8)
        m[blk, self["keep_idx"].astype(np.int64)] = 1
        return m.reshape(T, C, N)


class _Slot(object):
    """Pinned upload buffers, device inputs / outputs, pinned result buffers and the completion event of ONE
    in-flight step."""

    def __init__(self, pp, stage_set=0):
        T, N, C, dev = pp.T, pp.N, pp.C, pp.device
        rows = T * N
        self.stage_set = stage_set                          # which pinned upload buffers this slot reads
        self.h_boxes, self.h_scores = pp.h_boxes_sets[stage_set], pp.h_scores_sets[stage_set]
        self.h_seg = pp.h_seg_sets[stage_set]
        self.d_boxes = torch.empty((rows, 4), dtype=torch.float32, device=dev)
        self.d_scores = torch.empty((rows, C), dtype=torch.float32, device=dev)
        self.d_seg = ops.seg_offsets_uniform(T, N, dev)     # rewritten per step for ragged shards
        self.d_idx = torch.empty(rows * C, dtype=torch.int32, device=dev)
        self.d_mask = torch.empty(rows * C, dtype=torch.uint8, device=dev)      # device-resident entry only
        self.d_cnt = torch.empty((T, C), dtype=torch.int32, device=dev)
        self.d_off = torch.empty(T * C + 1, dtype=torch.int32, device=dev)
        self.d_succ = torch.empty(rows, dtype=torch.int32, device=dev)
        self.d_iou = torch.empty(rows, dtype=torch.float32, device=dev)
        # results: the keep lists and their offsets are written by the compaction kernels straight into
        # these pinned (device-mapped) buffers; succ / link_iou / status come home by copy engine
        self.h_keep = torch.empty(rows * C, dtype=torch.uint16).pin_memory()
        self.h_off = torch.zeros(T * C + 1, dtype=torch.int32).pin_memory()
        self.h_bits = torch.zeros((T * C, pp.W), dtype=torch.int32).pin_memory() if pp.want_bits else None
        self.h_succ = torch.empty(rows, dtype=torch.int32).pin_memory()
        self.h_iou = torch.empty(rows, dtype=torch.float32).pin_memory()
        self.status = ops.new_status(dev)
        self.h_status = torch.zeros(1, dtype=self.status.dtype).pin_memory()
        self.ev_boxes = torch.cuda.Event()
        self.ev_link = torch.cuda.Event()
        self.ev_nms = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.busy = False
        self.shape = None                                   # (n_frames, rows) of the step in flight
        self.graphs = {}                                    # whole-step CUDA graphs (uniform frames), per upload source
        self.launch_stream = torch.cuda.Stream(device=dev)  # the graph of this slot is launched here
        # frames of more than 1024 boxes keep their bit matrix in a global scratch slot per CTA
        self.ws_bytes = _lib.load().vdet_nms_frames_workspace_bytes(N, C, dev.index or 0) if N > 1024 else 0
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev) if self.ws_bytes else None
        self.ev_in = [torch.cuda.Event() for _ in range(pp.n_chunks)]
        # fixed workspace of the x-sorted link (its address is baked into the slot's CUDA graphs)
        self.link_ws = torch.empty(ops.link_workspace_bytes(rows, T, N), dtype=torch.uint8, device=dev)


class VideoPostProcessor(object):
    """NMS (all classes) + frame-to-frame link for one video shard of up to ``n_frames`` frames of up to
    ``n_boxes`` boxes.

    Input: boxes [T, N, 4] and scores [T, N, C] float32 on the HOST -- any C-contiguous arrays, pageable
    memory included; every step copies them through the processor's pinned upload buffers -- or already on
    the device (:meth:`run_device`).  Ragged shards pass packed rows [sum n_t, 4] / [sum n_t, C] plus
    ``counts`` [T'] (boxes per frame, T' <= n_frames, every n_t <= n_boxes).  Host output: a
    :class:`StepResult` -- the ordered keep lists of every (frame, class), i.e. what ``apply_vid_nms`` returns
    for every class (vdet/video_det.py:51-61, utils/nms.pyx:43-66), and the link.

    ``halo`` (boxes of the first frame of the NEXT shard, [H,4], optional ``halo_count`` on the device) links
    the shard's last frame across a shard boundary (see vdetlib_b200.dist); halo successors are reported as
    ``rows + index``.

    The host path is pipelined twice over.  Inside a step the score block is cut into ``n_chunks`` frame
    ranges; chunk k's host->device copy (copy stream) overlaps the NMS of the chunks before it (compute
    stream), the link and its read-back run under the score upload, and the keep lists go home from the
    compaction kernel itself (stores to mapped pinned memory: sum K is data dependent, a copy-engine D2H would
    need it on the host first).  Across steps, :meth:`submit_host` / :meth:`collect` keep up to ``n_slots``
    steps in flight, each with its own pinned, device and result buffers: while the host streams shard k+1
    into its upload buffers, shard k is uploading / computing.  Uniform shards replay the whole step from one
    CUDA graph per slot (one launch instead of ~40 stream operations).
    """

    def __init__(self, n_frames, n_boxes, n_classes, nms_thresh=0.3, device=None, n_chunks=8, n_slots=2, n_stage=None,
                 want_bits=False, stage_threads=0):
        self.T, self.N, self.C = int(n_frames), int(n_boxes), int(n_classes)
        self.nms_thresh = float(nms_thresh)
        self.device = device or ops.default_device()
        self.want_bits = bool(want_bits)
        self.stage_threads = int(stage_threads)
        self.W = (self.N + 31) // 32
        T, N, C = self.T, self.N, self.C
        rows = T * N
        dev = self.device
        self.seg_offsets = ops.seg_offsets_uniform(T, N, dev)
        self._uniform_off = np.arange(0, (T + 1) * N, N, dtype=np.int32)
        # pinned upload buffers: one set per slot (shard k+1 is staged while shard k is in flight), or one set
        # shared by every slot (n_stage=1: a staged shard can be submitted any number of times)
        n_slots = max(1, int(n_slots))
        n_stage = n_slots if n_stage is None else int(n_stage)
        if n_stage not in (1, n_slots):
            raise ValueError("n_stage must be 1 or n_slots")
        self.n_stage = n_stage
        self.h_boxes_sets = [torch.empty((rows, 4), dtype=torch.float32).pin_memory() for _ in range(n_stage)]
        self.h_scores_sets = [torch.empty((rows, C), dtype=torch.float32).pin_memory() for _ in range(n_stage)]
        self.h_seg_sets = [torch.from_numpy(self._uniform_off.copy()).pin_memory() for _ in range(n_stage)]
        self._staged = [None] * n_stage                     # per staging set: (n_frames, rows, uniform) of its shard
        self._src = [(self.h_boxes_sets[k], self.h_scores_sets[k]) for k in range(n_stage)]   # what the step uploads from
        self._registered = {}                               # data pointer -> (array kept alive, pinned tensor view)
        self.h_boxes, self.h_scores = self.h_boxes_sets[0], self.h_scores_sets[0]
        # frame ranges of the pipeline chunks (uniform shards; ragged ones balance rows per step)
        self.n_chunks = max(1, min(int(n_chunks), T))
        self.chunks = self._chunk_edges(self._uniform_off, T)
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self.s_link = torch.cuda.Stream(device=dev)         # the link of the device-resident step runs beside the NMS,
        self.s_nms = torch.cuda.Stream(device=dev, priority=-1)    # which has the higher priority
        self.slots = [_Slot(self, k % self.n_stage) for k in range(n_slots)]
        self._next_slot = 0
        self._graphs = {}
        self._lib = _lib.load()
        # host-side wall clock of the staged path, accumulated per step (what bounds the end-to-end rate on a
        # host-bound box): [stage copy, enqueue / graph launch, wait in collect], in seconds
        self.host_s = [0.0, 0.0, 0.0]
        self.host_steps = 0

    def _chunk_edges(self, off, n_frames):
        """Frame ranges whose row counts are as equal as the frame boundaries allow."""
        rows = int(off[n_frames])
        k = max(1, min(self.n_chunks, n_frames))
        targets = (np.arange(1, k) * rows) // k
        cuts = np.searchsorted(off[:n_frames + 1], targets, side="left")
        edges = sorted(set([0, n_frames] + [int(c) for c in cuts]))
        return [(edges[i], edges[i + 1]) for i in range(len(edges) - 1) if edges[i + 1] > edges[i]]

    # slot 0 doubles as the buffer set of the synchronous / device-resident entry points
    d_boxes = property(lambda self: self.slots[0].d_boxes)
    d_scores = property(lambda self: self.slots[0].d_scores)
    d_idx = property(lambda self: self.slots[0].d_idx)
    d_mask = property(lambda self: self.slots[0].d_mask)
    d_cnt = property(lambda self: self.slots[0].d_cnt)
    d_succ = property(lambda self: self.slots[0].d_succ)
    d_iou = property(lambda self: self.slots[0].d_iou)
    status = property(lambda self: self.slots[0].status)
    h_status = property(lambda self: self.slots[0].h_status)

    # bytes crossing PCIe per full-size step: uploads are fixed, the keep lists are data dependent
    @property
    def h2d_bytes(self):
        return self.h_boxes.numel() * 4 + self.h_scores.numel() * 4

    def d2h_bytes(self, result):
        """Bytes the step behind ``result`` sent home: keep lists + offsets + link + status (+ bit masks)."""
        rows = int(result["succ"].shape[0])
        b = 2 * int(result["keep_off"][-1]) + 4 * int(result["keep_off"].shape[0]) + 8 * rows + 4
        if "keep_bits" in result:
            b += 4 * int(result["keep_bits"].size)
        return b

    def _views(self, out):
        T, N, C = self.T, self.N, self.C
        return {"keep_idx": out[0].view(T, C, N), "keep_cnt": out[1], "keep_mask": out[2].view(T, C, N)}

    def _launch(self, d_boxes, d_scores, halo, halo_count=None):
        """NMS and link of one device-resident shard.  The two kernels are independent: the link is forked onto a
        side stream BEHIND the NMS launch, so its CTAs fill the SMs that the persistent NMS grid leaves idle in its
        last, partially filled round (1000 frames over 592 CTA slots) instead of running after it."""
        cur = torch.cuda.current_stream()
        self.status.zero_()
        self.s_nms.wait_stream(cur)
        self.s_link.wait_stream(cur)
        # the NMS goes to a HIGH-priority stream and is launched first: whichever grid the hardware starts first, SM
        # resources that become free go to NMS CTAs as long as any are pending, and the link takes what is left
        with torch.cuda.stream(self.s_nms):
            out = ops.nms_frames(d_boxes, d_scores, self.seg_offsets, self.nms_thresh, self.N, want_mask=True,
                                 status=self.status, frame_major_out=True, out=(self.d_idx, self.d_cnt, self.d_mask))
        with torch.cuda.stream(self.s_link):
            ops.link_frames(d_boxes, self.seg_offsets, self.N, halo, halo_row_base=self.T * self.N,
                            out=(self.d_succ, self.d_iou), halo_count=halo_count, ws=self.slots[0].link_ws)
        cur.wait_stream(self.s_nms)
        cur.wait_stream(self.s_link)
        return out

    def run_device(self, d_boxes, d_scores, halo=None, graph=False):
        """Device tensors in ([T*N,4], [T*N,C]), device tensors out (padded keep_idx [T,C,N], keep_cnt,
        keep_mask, succ, link_iou); asynchronous.

        ``graph=True`` replays the launches (NMS, link) from a CUDA graph captured on first use for
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

    # ---- staging ---------------------------------------------------------------------------
    def _target_slot(self):
        if not any(sl.busy for sl in self.slots):
            self._next_slot = 0
        return self.slots[self._next_slot]

    def register_host_arrays(self, *arrays):
        """Pin caller-owned C-contiguous float32 NumPy arrays IN PLACE (cudaHostRegister) so that later
        ``stage`` / ``submit_host`` calls given exactly these arrays upload straight from them -- no staging copy,
        i.e. a third of the host-memory traffic of the general path (the producer's write + one DMA read instead
        of + a read and a write of the copy).  For producers that fill a fixed ring of output buffers.  The
        arrays are kept alive by the processor and must not be modified between ``submit_host`` and
        ``collect`` of a step that uses them.  ``unregister_host_arrays`` undoes it."""
        cudart = torch.cuda.cudart()
        for a in arrays:
            if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]):
                raise ValueError("register_host_arrays: C-contiguous float32 NumPy arrays only")
            if a.ctypes.data in self._registered:
                continue
            rc = cudart.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError("cudaHostRegister failed: %r" % (rc,))
            self._registered[a.ctypes.data] = (a, torch.from_numpy(a))

    def unregister_host_arrays(self, *arrays):
        if any(sl.busy for sl in self.slots):
            raise RuntimeError("unregister_host_arrays: collect() the steps in flight first")
        for sl in self.slots:
            sl.graphs.clear()                               # captured uploads may point at these arrays
        cudart = torch.cuda.cudart()
        for a in (arrays or [v[0] for v in list(self._registered.values())]):
            if self._registered.pop(a.ctypes.data, None) is not None:
                cudart.cudaHostUnregister(a.ctypes.data)
        # a staging set that still names an array which is no longer pinned has nothing staged any more
        live = set(v[1].data_ptr() for v in self._registered.values())
        for k, (sb, ss) in enumerate(self._src):
            own = (self.h_boxes_sets[k].data_ptr(), self.h_scores_sets[k].data_ptr())
            if (sb.data_ptr(), ss.data_ptr()) != own and not (sb.data_ptr() in live and ss.data_ptr() in live):
                self._src[k] = (self.h_boxes_sets[k], self.h_scores_sets[k])
                self._staged[k] = None

    def stage(self, boxes, scores, counts=None):
        """Copy host arrays (pageable or not) into the pinned upload buffers of the NEXT step to be submitted,
        with streaming stores over several host threads: lines written with ordinary stores stay dirty in the
        CPU caches and the copy engine then reads them at roughly half the PCIe rate (profiles/r01_pcie.md).
        With one staging set (``n_stage=1``) every step in flight reads these buffers: collect all outstanding
        steps before restaging.  With a set per slot only the slot about to be reused must have been collected."""
        target = self._free_target("stage")
        b = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
        s = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1, self.C)
        shape = self._describe(target, counts, b.shape[0], s.shape[0])
        n_frames, rows, uniform = shape
        rb, rs = self._registered.get(b.ctypes.data), self._registered.get(s.ctypes.data)
        if rb is not None and rs is not None and rb[0].nbytes >= b.nbytes and rs[0].nbytes >= s.nbytes:
            # the caller's own arrays, pinned in place: the step uploads from them directly
            self._src[target.stage_set] = (rb[1].view(-1, 4), rs[1].view(-1, self.C))
        else:
            if rows:
                ops.host_copy_stream(target.h_boxes[:rows], b, self.stage_threads)
                ops.host_copy_stream(target.h_scores[:rows], s, self.stage_threads)
            self._src[target.stage_set] = (target.h_boxes, target.h_scores)
        self._staged[target.stage_set] = shape

    def _free_target(self, who):
        target = self._target_slot()
        if any(sl.busy and sl.stage_set == target.stage_set for sl in self.slots):
            raise RuntimeError("%s: a submitted step still reads the staging buffers; collect() it first" % who)
        return target

    def _describe(self, target, counts, box_rows, score_rows):
        """Validate the shard's shape and write its segment table; returns (n_frames, rows, uniform)."""
        if counts is None:
            n_frames, rows, uniform = self.T, self.T * self.N, True
            if box_rows != rows or score_rows != rows:
                raise ValueError("stage: expected %d rows" % rows)
            if self._staged[target.stage_set] is not None and not self._staged[target.stage_set][2]:
                target.h_seg.numpy()[:] = self._uniform_off
        else:
            cnt = np.asarray(counts, dtype=np.int64).reshape(-1)
            n_frames, rows, uniform = int(cnt.shape[0]), int(cnt.sum()), False
            if n_frames < 1 or n_frames > self.T or (cnt < 0).any() or int(cnt.max()) > self.N:
                raise ValueError("stage: counts must describe 1..%d frames of 0..%d boxes" % (self.T, self.N))
            if box_rows != rows or score_rows != rows:
                raise ValueError("stage: counts add up to %d rows, got %d / %d" % (rows, box_rows, score_rows))
            seg = target.h_seg.numpy()
            seg[0] = 0
            np.cumsum(cnt, out=seg[1:n_frames + 1])
            seg[n_frames + 1:] = rows
        return n_frames, rows, uniform

    def input_buffers(self):
        """Producer-side zero copy: NumPy views ``(boxes [T*N, 4], scores [T*N, C])`` of the pinned upload buffers
        of the NEXT step to be submitted.  A producer that writes its detections straight into them (packed rows
        from row 0; the rest is ignored) and then calls :meth:`commit_inputs` skips the staging copy altogether --
        one pass over host memory (the DMA read) instead of three.  The views stay valid for the life of the
        processor; they belong to the caller until ``commit_inputs`` + ``submit_staged`` and again after the
        ``collect`` of that step."""
        target = self._free_target("input_buffers")
        return target.h_boxes.numpy(), target.h_scores.numpy()

    def commit_inputs(self, counts=None):
        """Declare the buffers handed out by :meth:`input_buffers` filled: a full uniform shard, or ``counts[t]``
        boxes per frame in packed rows.  Follow with :meth:`submit_staged`."""
        target = self._free_target("commit_inputs")
        rows = self.T * self.N if counts is None else int(np.asarray(counts, dtype=np.int64).sum())
        shape = self._describe(target, counts, rows, rows)
        self._src[target.stage_set] = (target.h_boxes, target.h_scores)
        self._staged[target.stage_set] = shape

    # ---- one step on the streams -----------------------------------------------------------
    def _enqueue(self, sl, shape, halo, halo_count, fork):
        """Stream operations of one staged step on slot ``sl``.  The boxes (4 floats/row) go first, so the
        link -- which needs no scores -- and its D2H run under the upload of the scores (C floats/row); each
        score chunk is followed by its NMS launch; the compaction of the keep lists writes them home.
        ``fork``: branch the copy streams off the current stream and join them again (what a stream capture
        needs); otherwise the copy streams run free and ``sl.done`` marks the end."""
        n_frames, rows, uniform = shape
        N, C = self.N, self.C
        cur = torch.cuda.current_stream()
        s_in, s_out = self.s_in, self.s_out
        off = self._uniform_off if uniform else sl.h_seg.numpy()
        chunks = self.chunks if uniform else self._chunk_edges(off, n_frames)
        h_boxes, h_scores = self._src[sl.stage_set]
        if fork:
            s_in.wait_stream(cur)
            s_out.wait_stream(cur)
        sl.status.zero_()                                   # a flagged shard must not poison the next one
        with torch.cuda.stream(s_in):
            sl.d_seg.copy_(sl.h_seg, non_blocking=True)       # 4 KB: the step's segment table (uniform or ragged)
            if rows:
                sl.d_boxes[:rows].copy_(h_boxes[:rows], non_blocking=True)
            sl.ev_boxes.record(s_in)
            for k, (f0, f1) in enumerate(chunks):
                r0, r1 = int(off[f0]), int(off[f1])
                if r1 > r0:
                    sl.d_scores[r0:r1].copy_(h_scores[r0:r1], non_blocking=True)
                sl.ev_in[k].record(s_in)
        cur.wait_event(sl.ev_boxes)
        seg = sl.d_seg[:n_frames + 1]
        ops.link_frames(sl.d_boxes[:rows], seg, N, halo, halo_row_base=rows, out=(sl.d_succ, sl.d_iou),
                        halo_count=halo_count, ws=sl.link_ws)
        sl.ev_link.record(cur)
        s_out.wait_event(sl.ev_link)
        with torch.cuda.stream(s_out):
            sl.h_succ[:rows].copy_(sl.d_succ[:rows], non_blocking=True)
            sl.h_iou[:rows].copy_(sl.d_iou[:rows], non_blocking=True)
        nms = self._lib.vdet_nms_frames_f32
        stream_ptr = cur.cuda_stream
        ws_ptr = sl.ws.data_ptr() if sl.ws is not None else None
        for k, (f0, f1) in enumerate(chunks):
            cur.wait_event(sl.ev_in[k])
            # absolute row offsets: every chunk passes the shard's base pointers and its own slice of the
            # segment table / count table
            _lib.check(nms(sl.d_boxes.data_ptr(), 4, sl.d_scores.data_ptr(), C, 1, sl.d_seg[f0:].data_ptr(), f1 - f0, N,
                           None, C, self.nms_thresh, sl.d_idx.data_ptr(), sl.d_cnt[f0:].data_ptr(), None, rows,
                           _lib.LAYOUT_FRAME_MAJOR, sl.status.data_ptr(), ws_ptr, sl.ws_bytes, stream_ptr), "nms_frames")
        _lib.check(self._lib.vdet_compact_keep(sl.d_idx.data_ptr(), sl.d_cnt.data_ptr(), sl.d_seg.data_ptr(), n_frames, N, C,
                                               _lib.KEEP_U16_LOCAL, sl.d_off.data_ptr(), sl.h_off.data_ptr(),
                                               sl.h_keep.data_ptr(), sl.h_bits.data_ptr() if sl.h_bits is not None else None,
                                               stream_ptr), "compact_keep")
        sl.ev_nms.record(cur)
        s_out.wait_event(sl.ev_nms)
        with torch.cuda.stream(s_out):
            sl.h_status.copy_(sl.status, non_blocking=True)
        if fork:
            cur.wait_stream(s_in)
            cur.wait_stream(s_out)
        else:
            sl.done.record(s_out)

    def _capture(self, sl, shape, halo, halo_count):
        """Capture the whole staged step of slot ``sl`` (copies on three streams + kernels) into one graph."""
        with torch.cuda.stream(sl.launch_stream):
            self._enqueue(sl, shape, halo, halo_count, fork=True)     # warm-up: kernel attributes are set outside the capture
        sl.launch_stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=sl.launch_stream):
            self._enqueue(sl, shape, halo, halo_count, fork=True)
        return g

    def submit_staged(self, halo=None, halo_fn=None, graph=False, halo_count=None):
        """Enqueue one step on the next free slot from its staged inputs and return a ticket for
        :meth:`collect`; does not block.  Up to ``n_slots`` steps may be in flight.

        ``halo_fn(slot_index, h_first_frame_boxes)`` (optional) runs the boundary exchange for this step
        before anything else is enqueued and returns ``(halo, halo_count)`` device tensors (fixed addresses
        per slot, so the step's graph can hold them).  ``graph=True`` replays the step from a CUDA graph
        captured on first use (uniform shards only; ragged ones always take the eager stream path)."""
        k = self._next_slot
        sl = self.slots[k]
        if sl.busy:
            raise RuntimeError("submit_staged: all %d slots are in flight; collect() the oldest step first"
                               % len(self.slots))
        shape = self._staged[sl.stage_set]
        if shape is None:
            raise RuntimeError("submit_staged: nothing staged; call stage() first")
        n_frames, rows, uniform = shape
        src_boxes, src_scores = self._src[sl.stage_set]
        if graph and uniform:
            with torch.cuda.stream(sl.launch_stream):
                if halo_fn is not None:
                    halo, halo_count = halo_fn(k, src_boxes[:int(self._uniform_off[1])])
                # one graph per (upload source, halo buffers): the pinned staging set, or a registered caller array
                key = (src_boxes.data_ptr(), src_scores.data_ptr(), 0 if halo is None else halo.data_ptr(),
                       0 if halo_count is None else halo_count.data_ptr())
                g = sl.graphs.get(key)
                if g is None:
                    if len(sl.graphs) >= 16:
                        sl.graphs.clear()
                    g = sl.graphs[key] = self._capture(sl, shape, halo, halo_count)
                g.replay()
                sl.done.record(sl.launch_stream)
        else:
            if halo_fn is not None:
                first = int(sl.h_seg[1].item()) if not uniform else self.N
                halo, halo_count = halo_fn(k, src_boxes[:first])
            self._enqueue(sl, shape, halo, halo_count, fork=False)
        sl.shape = shape
        sl.busy = True
        self._next_slot = (k + 1) % len(self.slots)
        return k

    def submit_host(self, boxes, scores, counts=None, halo=None, halo_fn=None, graph=True, halo_count=None):
        """The user-facing asynchronous call: stage the host arrays (this is where the caller's memory is read;
        the arrays may be reused as soon as this returns) and enqueue the step.  Returns a ticket."""
        t0 = time.perf_counter()
        self.stage(boxes, scores, counts)
        t1 = time.perf_counter()
        ticket = self.submit_staged(halo, halo_fn, graph, halo_count)
        self.host_s[0] += t1 - t0
        self.host_s[1] += time.perf_counter() - t1
        self.host_steps += 1
        return ticket

    def collect(self, ticket):
        """Wait for the step behind ``ticket``; returns a :class:`StepResult` of host views (valid until the
        slot is submitted again) after translating the status word."""
        sl = self.slots[ticket]
        if not sl.busy:
            raise RuntimeError("collect: ticket %r is not in flight" % (ticket,))
        t0 = time.perf_counter()
        sl.done.synchronize()
        self.host_s[2] += time.perf_counter() - t0
        sl.busy = False
        ops.raise_for_status_word(int(sl.h_status.item()))
        n_frames, rows, uniform = sl.shape
        nb = n_frames * self.C
        off = sl.h_off.numpy()[:nb + 1]
        res = StepResult(keep_off=off, keep_idx=sl.h_keep.numpy()[:int(off[nb])],
                         keep_cnt=np.diff(off).reshape(n_frames, self.C),
                         succ=sl.h_succ.numpy()[:rows], link_iou=sl.h_iou.numpy()[:rows],
                         n_frames=n_frames, n_classes=self.C, max_boxes=self.N)
        if sl.h_bits is not None:
            res["keep_bits"] = sl.h_bits.numpy()[:nb].view(np.uint32)
        return res

    def run_staged(self, halo=None, halo_fn=None, graph=False, halo_count=None):
        """One synchronous step from the staged inputs: submit + collect (on slot 0 when nothing is in flight,
        so that ``d_boxes`` / ``d_scores`` hold the step's inputs afterwards)."""
        if not any(sl.busy for sl in self.slots):
            self._next_slot = 0
        return self.collect(self.submit_staged(halo, halo_fn, graph, halo_count))

    def run_host(self, boxes, scores, counts=None, halo=None, graph=False):
        """The user-facing synchronous call: host arrays in, host results out."""
        self.stage(boxes, scores, counts)
        return self.run_staged(halo, graph=graph)
