"""Drop-in for the tubelet-proposal part of vdetlib's ``vdet.track`` (reference vdet/track.py).

``greedily_track_from_det`` (:122-186) and ``greedily_track_from_raw_dets`` (:189-252) keep
their signatures and control flow: detections are ranked once, the best surviving detection
seeds ``track_method`` (the caller's tracker -- the MATLAB FCNT/TLD trackers of the reference are
external and out of scope), and detections overlapping the new tubelet are suppressed.  The
suppression step -- one ``track_det_nms`` call per tracked box in the reference (:172-183) -- is
one kernel launch per tubelet here, with ``det_info`` and ``keep`` resident on the GPU.

Differences, stated: the reference restarts a MATLAB engine and retries when ``track_method``
raises (:161-168); here the exception propagates.
"""
import logging
from collections import defaultdict
from operator import itemgetter

import numpy as np
import torch

from .. import ops


class _GreedyState(object):
    """GPU-resident state of one greedy tracking run."""

    def __init__(self, det_info):
        self.m = len(det_info)
        self.det_info = torch.from_numpy(np.ascontiguousarray(det_info, dtype=np.float32)).to(ops.default_device())
        self.keep = torch.ones(max(self.m, 1), dtype=torch.uint8, device=self.det_info.device)
        self.status = ops.new_status(self.det_info.device)
        if self.m:
            self.row_ids, self.seg_offsets, seg_frame, self.max_len = ops.segment_by_frame(self.det_info[:, 0])
            self.seg_of = {float(f): s for s, f in enumerate(seg_frame.cpu().tolist())}
        else:
            self.seg_of = {}

    def suppress(self, new_tracks, nms_thres):
        """vdet/track.py:172-183: apply every box of every new tracklet, in order."""
        if not self.m:
            return
        for tracklet in new_tracks:
            # boxes of one launch must lie on distinct frames; split where a frame repeats
            batch, seen = [], set()
            for box in tracklet:
                if box['frame'] in seen:
                    self._launch(batch, nms_thres)
                    batch, seen = [], set()
                batch.append(box)
                seen.add(box['frame'])
            self._launch(batch, nms_thres)
        ops.raise_for_status(self.status)

    def _launch(self, boxes, nms_thres):
        if not boxes:
            return
        t = np.asarray([[b['frame']] + list(b['bbox']) for b in boxes], dtype=np.float32)   # :178
        seg = np.asarray([self.seg_of.get(float(f), -1) for f in t[:, 0]], dtype=np.int32)
        dev = self.det_info.device
        ops.track_nms_step(self.det_info, self.seg_offsets, self.row_ids,
                           torch.from_numpy(np.ascontiguousarray(t[:, 1:5])).to(dev),
                           torch.from_numpy(seg).to(dev), nms_thres, self.keep, self.status)

    def keep_host(self):
        return self.keep[:self.m].cpu().numpy().astype(bool).tolist()


def _nms_thres(opts):
    if hasattr(opts, 'nms_thres') and opts.nms_thres is not None:
        return opts.nms_thres
    return 0.3


def _greedy_loop(vid_proto, state, n_dets, anchor_of, score_of, track_method, opts, nms_thres):
    """The outer loop both greedy trackers share (vdet/track.py:141-185 and :206-251): take the best detection that
    is still alive, stop below ``opts.thres`` or at ``opts.max_tracks``, track it with the external tracker, then
    drop every detection the new tracklets suppress.  ``anchor_of(i) -> (frame, bbox)``, ``score_of(i)``.

    The reference retries a failed ``track_method`` call once after restarting its MATLAB engine (:159-168).  The
    engine is out of scope here; the control flow is kept through ``opts.on_tracker_error`` (optional callable,
    e.g. one that restarts the caller's tracker backend): when the tracker raises, the hook runs and the call is
    retried once; without a hook the exception propagates."""
    keep = [True] * n_dets
    cur_top_det_id = 0
    tracks = []
    while np.any(keep) and len(tracks) < opts.max_tracks:
        while cur_top_det_id < len(keep) and not keep[cur_top_det_id]:
            cur_top_det_id += 1
        if cur_top_det_id == len(keep):
            break
        top = cur_top_det_id
        cur_top_det_id += 1
        if score_of(top) < opts.thres:
            logging.info("Upon low confidence: total {} tracks".format(len(tracks)))
            break
        logging.info("tracking top No.{} in {}".format(len(tracks), vid_proto['video']))
        anchor_frame_id, anchor_bbox = anchor_of(top)
        try:
            new_tracks = track_method(vid_proto, anchor_frame_id, anchor_bbox, opts)
        except Exception:
            hook = getattr(opts, 'on_tracker_error', None)
            if hook is None:
                raise
            hook(opts)
            new_tracks = track_method(vid_proto, anchor_frame_id, anchor_bbox, opts)
        tracks.extend(new_tracks)
        logging.info("Applying nms between new tracks ({}) and detections.".format(len(new_tracks)))
        state.suppress(new_tracks, nms_thres)
        keep = state.keep_host()
        logging.info("{} / {} boxes kept.".format(np.sum(keep), len(keep)))
    return tracks


def greedily_track_from_det(vid_proto, det_proto, track_method, score_fun, opts):
    '''greedily track top detections and supress detections
       that have large overlaps with tracked boxes  (vdet/track.py:122-186)'''
    nms_thres = _nms_thres(opts)
    assert vid_proto['video'] == det_proto['video']
    track_proto = {}
    track_proto['video'] = vid_proto['video']
    track_proto['method'] = track_method.__name__

    dets = sorted(det_proto['detections'], key=lambda x: score_fun(x), reverse=True)
    det_info = np.asarray([[det['frame'], ] + list(det['bbox']) + [score_fun(det), ]
                           for det in dets], dtype=np.float32).reshape(-1, 6)
    track_proto['tracks'] = _greedy_loop(
        vid_proto, _GreedyState(det_info), len(dets),
        lambda i: (dets[i]['frame'], list(map(int, dets[i]['bbox']))), lambda i: score_fun(dets[i]),
        track_method, opts, nms_thres)
    return track_proto


def greedily_track_from_raw_dets(vid_proto, det_info, track_method, class_idx, opts):
    '''greedily track top detections and supress detections
       that have large overlaps with tracked boxes  (vdet/track.py:189-252)'''
    nms_thres = _nms_thres(opts)
    track_proto = {}
    track_proto['video'] = vid_proto['video']
    track_proto['method'] = track_method.__name__

    det_info = np.asarray(sorted(det_info[:, [0, 1, 2, 3, 4, 4 + class_idx]],
                                 key=itemgetter(5), reverse=True), dtype=np.float32).reshape(-1, 6)
    track_proto['tracks'] = _greedy_loop(
        vid_proto, _GreedyState(det_info), len(det_info),
        lambda i: (int(det_info[i][0]), list(map(int, det_info[i][1:5]))), lambda i: det_info[i][-1],
        track_method, opts, nms_thres)
    return track_proto
