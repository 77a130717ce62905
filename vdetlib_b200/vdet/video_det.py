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

from .. import ops
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
    dev_scores = torch.from_numpy(scores.reshape(T * R, C)).cuda()
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


class VideoPostProcessor(object):
    """NMS (all classes) + frame-to-frame link for one video shard of fixed shape.

    Input: boxes [T, N, 4] float32 and scores [T, N, C] float32 on the HOST (any array-like;
    copied through pinned staging buffers) or already on the device.  Output (device tensors from
    :meth:`run_device`, host arrays from :meth:`run_host` / :meth:`run_staged`), frame-major:

      keep_mask [T, C, N] uint8   1 = detection survives per-frame NMS for that class
      keep_cnt  [T, C]    int32   survivors per (frame, class)
      keep_idx  [T, C, N] int32   surviving rows of the frame in descending score, -1 padded
                                  (device only; not copied back by default)
      succ      [T*N]     int32   packed row of the best-IoU box in the next frame (-1: none)
      link_iou  [T*N]     float32 that IoU

    ``halo`` (boxes of the first frame of the NEXT shard, [H,4]) links the shard's last frame
    across a shard boundary (see vdetlib_b200.dist).

    The host path is pipelined: the shard is cut into ``n_chunks`` frame ranges; chunk k's
    host->device copy (copy stream), NMS (compute stream) and result read-back (read-back stream)
    overlap with the neighbouring chunks', so PCIe in both directions and the SMs are busy at the
    same time.  Frame-major outputs make every chunk a contiguous byte range.
    """

    def __init__(self, n_frames, n_boxes, n_classes, nms_thresh=0.3, device=None, n_chunks=8):
        self.T, self.N, self.C = int(n_frames), int(n_boxes), int(n_classes)
        self.nms_thresh = float(nms_thresh)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        T, N, C = self.T, self.N, self.C
        rows = T * N
        dev = self.device
        self.seg_offsets = ops.seg_offsets_uniform(T, N, dev)
        self.d_boxes = torch.empty((rows, 4), dtype=torch.float32, device=dev)
        self.d_scores = torch.empty((rows, C), dtype=torch.float32, device=dev)
        self.d_idx = torch.empty(rows * C, dtype=torch.int32, device=dev)
        self.d_mask = torch.empty(rows * C, dtype=torch.uint8, device=dev)
        self.d_cnt = torch.empty((T, C), dtype=torch.int32, device=dev)
        self.d_succ = torch.empty(rows, dtype=torch.int32, device=dev)
        self.d_iou = torch.empty(rows, dtype=torch.float32, device=dev)
        self.h_boxes = torch.empty((rows, 4), dtype=torch.float32).pin_memory()
        self.h_scores = torch.empty((rows, C), dtype=torch.float32).pin_memory()
        self.h_mask = torch.empty(rows * C, dtype=torch.uint8).pin_memory()
        self.h_cnt = torch.empty((T, C), dtype=torch.int32).pin_memory()
        self.h_succ = torch.empty(rows, dtype=torch.int32).pin_memory()
        self.h_iou = torch.empty(rows, dtype=torch.float32).pin_memory()
        self.status = ops.new_status(dev)
        self.h_status = torch.zeros(1, dtype=self.status.dtype).pin_memory()
        # frame ranges of the pipeline chunks
        n_chunks = max(1, min(int(n_chunks), T))
        edges = [round(k * T / n_chunks) for k in range(n_chunks + 1)]
        self.chunks = [(edges[k], edges[k + 1]) for k in range(n_chunks) if edges[k + 1] > edges[k]]
        self.chunk_seg = {f1 - f0: ops.seg_offsets_uniform(f1 - f0, N, dev) for f0, f1 in self.chunks}
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self._graphs = {}

    # bytes crossing PCIe per run_staged() call
    @property
    def h2d_bytes(self):
        return self.h_boxes.numel() * 4 + self.h_scores.numel() * 4

    @property
    def d2h_bytes(self):
        return self.h_mask.numel() + self.h_cnt.numel() * 4 + self.h_succ.numel() * 4 + self.h_iou.numel() * 4

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
        live in the processor's own buffers (valid until the next call)."""
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
        """Copy host arrays into the pinned upload buffers (not part of the timed region) with
        streaming stores: lines written with ordinary stores stay dirty in the CPU caches and the
        copy engine then reads them at roughly half the PCIe rate (profiles/r01_pcie.md)."""
        b = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
        s = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1, self.C)
        if b.shape[0] != self.h_boxes.shape[0] or s.shape[0] != self.h_scores.shape[0]:
            raise ValueError("stage: expected %d rows" % self.h_boxes.shape[0])
        ops.host_copy_stream(self.h_boxes, b)
        ops.host_copy_stream(self.h_scores, s)

    def run_staged(self, halo=None, halo_fn=None):
        """Pipelined H2D (pinned buffers) -> link + NMS per chunk -> D2H; synchronises and returns
        host views.  The boxes (4 floats/row) go first, so the link -- which needs no scores --
        and its D2H run under the upload of the scores (C floats/row); each score chunk is
        followed by its NMS launch and the download of its keep mask.  The step is bound by the
        upload; what remains after its last byte is one chunk's NMS and mask download.
        ``halo_fn(d_first_frame_boxes)`` (optional) is called on the compute stream once the
        boxes are on the device and returns the halo tensor (boundary exchange)."""
        N, C = self.N, self.C
        cur = torch.cuda.current_stream()
        self.s_in.wait_stream(cur)
        self.s_out.wait_stream(cur)
        ev_in = []
        with torch.cuda.stream(self.s_in):
            self.d_boxes.copy_(self.h_boxes, non_blocking=True)
            ev_boxes = torch.cuda.Event()
            ev_boxes.record(self.s_in)
            for f0, f1 in self.chunks:
                r0, r1 = f0 * N, f1 * N
                self.d_scores[r0:r1].copy_(self.h_scores[r0:r1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_in)
                ev_in.append(ev)
        cur.wait_event(ev_boxes)
        if halo_fn is not None:
            halo = halo_fn(self.d_boxes[:N])
        ops.link_frames(self.d_boxes, self.seg_offsets, N, halo, out=(self.d_succ, self.d_iou))
        ev = torch.cuda.Event()
        ev.record(cur)
        self.s_out.wait_event(ev)
        with torch.cuda.stream(self.s_out):
            self.h_succ.copy_(self.d_succ, non_blocking=True)
            self.h_iou.copy_(self.d_iou, non_blocking=True)
        for k, (f0, f1) in enumerate(self.chunks):
            r0, r1 = f0 * N, f1 * N
            cur.wait_event(ev_in[k])
            ops.nms_frames(self.d_boxes[r0:r1], self.d_scores[r0:r1], self.chunk_seg[f1 - f0], self.nms_thresh, N,
                           want_mask=True, status=self.status, frame_major_out=True,
                           out=(self.d_idx[r0 * C:r1 * C], self.d_cnt[f0:f1], self.d_mask[r0 * C:r1 * C]))
            ev = torch.cuda.Event()
            ev.record(cur)
            self.s_out.wait_event(ev)
            with torch.cuda.stream(self.s_out):
                self.h_mask[r0 * C:r1 * C].copy_(self.d_mask[r0 * C:r1 * C], non_blocking=True)
                self.h_cnt[f0:f1].copy_(self.d_cnt[f0:f1], non_blocking=True)
        with torch.cuda.stream(self.s_out):
            self.h_status.copy_(self.status, non_blocking=True)
        self.s_out.synchronize()
        ops.raise_for_status_word(int(self.h_status.item()))
        return {"keep_mask": self.h_mask.numpy().reshape(self.T, C, N), "keep_cnt": self.h_cnt.numpy(),
                "succ": self.h_succ.numpy(), "link_iou": self.h_iou.numpy()}

    def run_host(self, boxes, scores, halo=None):
        """The user-facing call: host arrays in, host arrays out."""
        self.stage(boxes, scores)
        return self.run_staged(halo)
