"""Drop-in for the post-CNN part of vdetlib's ``vdet.tubelet_cls`` (reference vdet/tubelet_cls.py).

Same function names, arguments, defaults, return protos, in-place behaviour and exceptions as
the reference; the per-box NumPy ``iou`` calls and the Python while-loops are replaced by one
batched kernel launch per call (libvdet_b200.so).  Scores and boxes cross to the GPU as
float64 so the results are the reference's Python-float results bit for bit.

Out of scope here (need Caffe + images): the CNN/SVM scoring functions, reference :53-260.
"""
import copy
import logging
from collections import defaultdict

import numpy as np
import torch

from .. import _lib, ops
from ..utils.protocol import tubelets_proto_from_tracks_proto
from .dataset import imagenet_vdet_classes


# --------------------------------------------------------------------------------------
# packing helpers (host side of the proto <-> tensor boundary)
# --------------------------------------------------------------------------------------
def _pack_rows(rows, fill=0.0):
    """list of 1-D float sequences -> (float64 tensor [K, Lmax] on GPU, int32 lengths [K])."""
    K = len(rows)
    lens = np.asarray([len(r) for r in rows], dtype=np.int32)
    L = int(lens.max()) if K else 0
    L_pad = max(L, 1)
    buf = np.full((K, L_pad), fill, dtype=np.float64)
    for i, r in enumerate(rows):
        if len(r):
            buf[i, :len(r)] = r
    dev = ops.default_device()
    return torch.from_numpy(buf).to(dev), torch.from_numpy(lens).to(dev), lens


def _pool_on_gpu(tub_boxes, tub_frames, frame_dets, overlap_thres, mode):
    """tub_boxes: list of bbox lists; tub_frames: their frame ids; frame_dets: dict
    frame -> (det_boxes ndarray [N,4], det_scores ndarray [N]).  Returns (arg_local, score)
    numpy arrays; arg_local indexes the frame's det arrays, -1 = no det selected."""
    frames = sorted(frame_dets.keys())
    seg_of = {f: s for s, f in enumerate(frames)}
    counts = [len(frame_dets[f][0]) for f in frames]
    seg_off = np.zeros(len(frames) + 1, dtype=np.int64)
    np.cumsum(counts, out=seg_off[1:])
    if len(frames):
        det_boxes = np.concatenate([np.asarray(frame_dets[f][0], dtype=np.float64).reshape(-1, 4) for f in frames])
        det_scores = np.concatenate([np.asarray(frame_dets[f][1], dtype=np.float64).reshape(-1) for f in frames])
    else:
        det_boxes = np.zeros((0, 4)); det_scores = np.zeros((0,))
    tb = np.asarray(tub_boxes, dtype=np.float64).reshape(-1, 4)
    seg = np.asarray([seg_of.get(f, -1) for f in tub_frames], dtype=np.int32)
    dev = ops.default_device()
    arg, score = ops.spatial_maxpool(
        torch.from_numpy(tb).to(dev), torch.from_numpy(seg).to(dev),
        torch.from_numpy(det_boxes).to(dev), torch.from_numpy(det_scores).to(dev),
        torch.from_numpy(seg_off.astype(np.int32)).to(dev), overlap_thres, mode)
    arg = arg.cpu().numpy().astype(np.int64)
    score = score.cpu().numpy()
    base = np.where(seg >= 0, seg_off[np.maximum(seg, 0)], 0)
    arg_local = np.where(arg >= 0, arg - base, -1)
    return arg_local, score


# --------------------------------------------------------------------------------------
# score completion / temporal max-pool / temporal convolution
# --------------------------------------------------------------------------------------
def do_score_completion(score_proto):
    """Fill runs of ``det_score <= -10`` in every tubelet, IN PLACE.  vdet/tubelet_cls.py:284-303.

    A tubelet with no valid score raises IndexError like the reference (:295)."""
    tubelets = [t for t in score_proto['tubelets'] if len(t['boxes'])]
    if not tubelets:
        return
    rows = [[b['det_score'] for b in t['boxes']] for t in tubelets]
    dev, lens_dev, lens = _pack_rows(rows)
    status = ops.score_completion_(dev, lens_dev)
    ops.raise_for_status(status)
    out = dev.cpu().numpy()
    for t, row, n in zip(tubelets, out, lens):
        for b, v in zip(t['boxes'], row[:n]):
            if not (b['det_score'] > -10):
                b['det_score'] = float(v)


def score_proto_temporal_maxpool(score_proto, window_size):
    """Sliding max of det_score over ``window_size`` frames.  vdet/tubelet_cls.py:386-414.

    As in the reference the returned proto is a SHALLOW copy (:393): the input's tubelet boxes
    are updated in place.  ValueError for an even window (:389-390) or a gt tubelet (:396-397)."""
    if window_size == 1:
        return score_proto
    if window_size % 2 != 1:
        raise ValueError('Window size must be odd!')
    new_score_proto = copy.copy(score_proto)
    new_score_proto['method'] += '_temporal_maxpool_{}'.format(window_size)
    tubelets = new_score_proto['tubelets']
    rows = []
    for k, tubelet in enumerate(tubelets):
        if tubelet['gt'] == 1:
            # the reference has already rewritten the tubelets before this one (:395-412)
            _maxpool_rows(tubelets[:k], rows, window_size)
            raise ValueError('Dangerous: Score file contains gt tracks!')
        rows.append([b['det_score'] for b in tubelet['boxes']])
    _maxpool_rows(tubelets, rows, window_size)
    return new_score_proto


def _maxpool_rows(tubelets, rows, window_size):
    if not tubelets:
        return
    dev, lens_dev, lens = _pack_rows(rows, fill=-1e5)
    out = ops.temporal_maxpool(dev, window_size, lens_dev).cpu().numpy()
    for t, row, n in zip(tubelets, out, lens):
        for b, v in zip(t['boxes'], row[:n]):
            b['det_score'] = float(v)


class _Blob(object):
    """The two things score_conv_cls touches on a Caffe blob: ``shape`` / ``reshape`` and ``data``."""

    def __init__(self, channels):
        self.data = np.zeros((1, channels, 1, 1), dtype=np.float32)

    @property
    def shape(self):
        return self.data.shape

    def reshape(self, *shape):
        if tuple(shape) != self.data.shape:
            self.data = np.zeros(shape, dtype=np.float32)


CONV_CHANNELS = ('det_scores', 'track_scores', 'anchors', 'abs_anchors', 'gt_overlaps', 'labels', 'all_scores', 'feats')


class TemporalConvNet(object):
    """Stand-in for the Caffe net of ``score_conv_cls`` (vdet/tubelet_cls.py:15-51).

    The reference feeds per-tubelet 1-D channels as ``(1, C, 1, L)`` blobs to an external Caffe model whose
    prototxt and weights are not part of vdetlib (the arithmetic is **parity unpinned**).  What IS
    pinned by the reference is the marshalling -- which channels exist, how they are computed from the score
    proto, their blob shapes, and that ``probs[:, 1, :]`` becomes ``conv_score`` -- and this class offers
    exactly that surface: ``blobs`` (name -> blob with ``shape`` / ``reshape`` / ``data``), ``forward()``
    returning ``{'probs': (1, 2, 1, L)}``.  The model behind it is explicit: ``taps[name]`` is an odd-length
    temporal filter per channel of blob ``name`` ([w] or [C, w]); the depthwise responses (one launch of
    ``vdet_temporal_conv1d`` over all channels) are summed with ``bias`` into a logit z per frame and
    ``probs = softmax([0, z])``, i.e. ``probs[:, 1] = 1 / (1 + exp(-z))``.
    """

    def __init__(self, taps, bias=0.0, pad_mode="zero"):
        self.taps = {}
        for k, v in taps.items():
            if k not in CONV_CHANNELS:
                raise KeyError(k)
            self.taps[k] = np.atleast_2d(np.asarray(v, dtype=np.float32))      # [C, w]
        self.bias = float(bias)
        self.pad_mode = pad_mode
        self.blobs = {k: _Blob(v.shape[0]) for k, v in self.taps.items()}

    def logits(self, rows, lengths=None):
        """rows: dict name -> float32 CUDA tensor [R * C_name, Lmax] (R tubelets, channel-minor); returns z [R, Lmax]."""
        z = None
        for name, taps in self.taps.items():
            x = rows[name]
            C = taps.shape[0]
            lens = None if lengths is None else lengths.repeat_interleave(C)
            y = ops.temporal_conv1d(x, torch.from_numpy(taps).to(x.device), self.pad_mode, lens)
            y = y.view(-1, C, x.shape[1]).sum(dim=1)
            z = y if z is None else z + y
        return z + self.bias

    def forward(self):
        """Caffe's ``net.forward()`` for the blobs as they are now (one tubelet)."""
        dev = ops.default_device()
        rows = {k: torch.from_numpy(np.ascontiguousarray(b.data[0, :, 0, :])).to(dev) for k, b in self.blobs.items()}
        z = self.logits(rows).cpu().numpy().astype(np.float32)[0]
        p1 = (1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(np.float32)
        return {'probs': np.stack([1.0 - p1, p1]).reshape(1, 2, 1, -1)}


def _conv_channels(tubelet, blob_names):
    """The per-tubelet channel dict of vdet/tubelet_cls.py:19-41."""
    track = {}
    boxes = tubelet['boxes']
    track['length'] = len(boxes)
    track['gt'] = tubelet['gt']
    track['mean_iou'] = np.mean([[x['gt_overlap'] for x in boxes]])
    track['det_scores'] = [x['det_score'] for x in boxes]
    track['track_scores'] = [x['track_score'] for x in boxes]
    track['anchors'] = [x['anchor'] * 1. / track['length'] for x in boxes]         # :28
    track['abs_anchors'] = [abs(a) for a in track['anchors']]                       # :30
    track['gt_overlaps'] = [x['gt_overlap'] for x in boxes]
    track['labels'] = [1 if iou >= 0.5 else 0 for iou in track['gt_overlaps']]      # :33
    # skip memory heavy features if possible
    if 'all_scores' in blob_names:
        track['all_scores'] = [x['all_score'] for x in boxes]
    if 'feats' in blob_names:
        track['feats'] = [x['feat'] for x in boxes]
    return track


def score_conv_cls(score_proto, net):
    """Temporal-convolution rescoring: writes ``conv_score`` per box.  vdet/tubelet_cls.py:15-51.

    ``net`` is anything with Caffe's surface (``blobs[name].shape / reshape / data``, ``forward()`` returning
    ``probs``) -- a real Caffe net works unchanged, one forward per tubelet like the reference.  With a
    :class:`TemporalConvNet` all tubelets are evaluated in ONE batch on the GPU (ragged rows): same channels,
    same result as its own per-tubelet ``forward()``."""
    new_score_proto = copy.copy(score_proto)
    blob_names = set(net.blobs.keys())
    if isinstance(net, TemporalConvNet) and all(len(t['boxes']) for t in new_score_proto['tubelets']) \
            and new_score_proto['tubelets']:
        tubelets = new_score_proto['tubelets']
        tracks = [_conv_channels(t, blob_names) for t in tubelets]
        lens = np.asarray([tr['length'] for tr in tracks], dtype=np.int32)
        Lmax = int(lens.max())
        dev = ops.default_device()
        rows = {}
        for name in blob_names.intersection(tracks[0].keys()):
            C = net.blobs[name].shape[1]
            buf = np.zeros((len(tracks), C, Lmax), dtype=np.float32)
            for k, tr in enumerate(tracks):
                buf[k, :, :lens[k]] = np.asarray(tr[name], dtype='float32').reshape(lens[k], -1).T if C > 1 \
                    else np.asarray(tr[name], dtype='float32')
            rows[name] = torch.from_numpy(buf.reshape(-1, Lmax)).to(dev)
        z = net.logits(rows, torch.from_numpy(lens).to(dev)).cpu().numpy().astype(np.float32)
        p1 = (1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(np.float32)
        for t, row in zip(tubelets, p1):
            for box, prob in zip(t['boxes'], row):
                box['conv_score'] = float(prob)
        return new_score_proto
    for tubelet in new_score_proto['tubelets']:
        track = _conv_channels(tubelet, blob_names)
        for blob_name in blob_names.intersection(set(track.keys())):
            num_channels = net.blobs[blob_name].shape[1]
            net.blobs[blob_name].reshape(1, num_channels, 1, track['length'])                  # :43-45
            net.blobs[blob_name].data[...] = np.asarray(track[blob_name], dtype='float32')      # :46
        blobs_out = net.forward()
        probs = blobs_out['probs'][:, 1, :]                                                     # :48
        for box, prob in zip(tubelet['boxes'], probs.ravel()):
            box['conv_score'] = float(prob)
    return new_score_proto


# --------------------------------------------------------------------------------------
# spatial max-pooling of detections onto tubelets
# --------------------------------------------------------------------------------------
def _spatial_max_pooling(vid_proto, tubelets_proto, frame_dets, overlap_thres):
    """vdet/tubelet_cls.py:324-347 / :509-532 for every frame of the video at once."""
    tub_boxes, tub_frames, where = [], [], []
    by_frame = defaultdict(list)
    for i, tubelet in enumerate(tubelets_proto):
        for j, box in enumerate(tubelet['boxes']):
            by_frame[box['frame']].append((i, j))
    for frame in vid_proto['frames']:
        fid = frame['frame']
        if fid not in frame_dets:
            continue
        for i, j in by_frame[fid]:
            cur = tubelets_proto[i]['boxes'][j]
            if cur is None:
                continue
            tub_boxes.append(cur['bbox'])
            tub_frames.append(fid)
            where.append((i, j))
    if not where:
        return
    arg, score = _pool_on_gpu(tub_boxes, tub_frames, frame_dets, overlap_thres, _lib.POOL_ARGMAX_SCORE)
    for (i, j), fid, a, s in zip(where, tub_frames, arg, score):
        cur = tubelets_proto[i]['boxes'][j]
        if a >= 0:
            cur['det_score'] = float(s)
            cur['bbox'] = frame_dets[fid][0][a].tolist()
        else:
            logging.debug("Tubelet %d has no overlapping dets (IOU > %s).", i, overlap_thres)
            cur['det_score'] = float(-1e5)


def dets_spatial_max_pooling(vid_proto, track_proto, det_proto, class_idx, overlap_thres=0.7):
    """Spatial max-pooling of a det proto onto tubelets + score completion.
    vdet/tubelet_cls.py:305-350."""
    assert vid_proto['video'] == track_proto['video']
    score_proto = {}
    score_proto['video'] = vid_proto['video']
    score_proto['method'] = "spatial_max_pooling_IOU_{}".format(overlap_thres)
    tubelets_proto = tubelets_proto_from_tracks_proto(track_proto['tracks'], class_idx)
    logging.info("Sampling dets in {} for {}...".format(vid_proto['video'], imagenet_vdet_classes[class_idx]))

    frame_to_det_idx = defaultdict(list)
    dets = det_proto['detections']
    for i, det in enumerate(dets):
        frame_to_det_idx[det['frame']].append(i)
    frame_dets = {}
    for fid, idx in frame_to_det_idx.items():
        frame_dets[fid] = (np.asarray([dets[i]['bbox'] for i in idx]),
                           np.asarray([dets[i]['scores'][class_idx - 1]['score'] for i in idx]))   # :329
    _spatial_max_pooling(vid_proto, tubelets_proto, frame_dets, overlap_thres)
    score_proto['tubelets'] = tubelets_proto
    do_score_completion(score_proto)
    return score_proto


def raw_dets_spatial_max_pooling(vid_proto, track_proto, frame_to_det, class_idx, overlap_thres=0.7):
    """Same with raw per-frame arrays ``frame_to_det[frame] = (boxes [N,4], zs [N,C])``.
    vdet/tubelet_cls.py:493-535."""
    assert vid_proto['video'] == track_proto['video']
    score_proto = {}
    score_proto['video'] = vid_proto['video']
    score_proto['method'] = "spatial_max_pooling_IOU_{}".format(overlap_thres)
    tubelets_proto = tubelets_proto_from_tracks_proto(track_proto['tracks'], class_idx)
    logging.info("Sampling dets in {} for {}...".format(vid_proto['video'], imagenet_vdet_classes[class_idx]))
    frame_dets = {}
    for fid, (det_boxes, det_scores) in frame_to_det.items():
        if det_boxes.size == 0:
            continue                                                     # :513
        frame_dets[fid] = (det_boxes, det_scores[:, class_idx - 1].ravel())   # :514
    _spatial_max_pooling(vid_proto, tubelets_proto, frame_dets, overlap_thres)
    score_proto['tubelets'] = tubelets_proto
    do_score_completion(score_proto)
    return score_proto


def rcnn_sampling_dets_scoring(vid_proto, track_proto, det_proto, net, class_idx, rcnn_model, overlap_thres=0.7,
                               save_feat=False, save_all_sc=False, score_column=None):
    """Score tubelet boxes with a CNN + SVM and let overlapping detections out-score them.
    vdet/tubelet_cls.py:196-260, same positional arguments.

    The CNN / SVM are out of scope (Caffe + images); they enter as callables with the meaning of the
    reference's calls: ``net(frame_path, boxes) -> features [P, F]`` stands for
    ``googlenet_features(imread(frame_path), boxes, net, 'pool5')`` (:213, :241) and
    ``rcnn_model(features) -> scores [P, K]`` for ``svm_scores(features, svm_from_rcnn_model(rcnn_model))``
    (:198, :214).  ``score_column``: the column of ``scores`` that holds ``class_idx``
    (reference: ``index_vdet_to_det[class_idx] - 1`` of a 200-column DET model, :215-218).

    The pooling epilogue (:221-247) is one batched launch for the whole video: per tubelet box the detections of
    its frame with IoU > ``overlap_thres`` (strict), FIRST arg-max of their class score (``det_score``, i.e.
    looked up by class_index), ``-inf`` when none; the detection replaces score and bbox only if its score is
    strictly greater than the box's own CNN score."""
    tubelets_proto = tubelets_proto_from_tracks_proto(track_proto['tracks'], class_idx)
    logging.info("Scoring {} for {}...".format(vid_proto['video'], imagenet_vdet_classes[class_idx]))
    if score_column is None:
        raise ValueError("rcnn_sampling_dets_scoring: score_column (the SVM column of class_idx) is required")
    frame_to_det_idx = defaultdict(list)
    dets = det_proto['detections']
    for i, det in enumerate(dets):
        frame_to_det_idx[det['frame']].append(i)
    by_frame = defaultdict(list)
    for i, tubelet in enumerate(tubelets_proto):
        seen = set()
        for j, box in enumerate(tubelet['boxes']):
            if box['frame'] not in seen:                       # tubelet_box_at_frame: the FIRST box of that frame
                seen.add(box['frame'])
                by_frame[box['frame']].append((i, j))
    tub_boxes, tub_frames, where, own = [], [], [], []
    frame_dets = {}
    for frame in vid_proto['frames']:
        fid = frame['frame']
        entries = by_frame.get(fid, [])
        if not entries:
            continue
        path = str(vid_proto['root_path'] + '/' + frame['path']) if 'root_path' in vid_proto else frame['path']
        valid_boxes = np.asarray([tubelets_proto[i]['boxes'][j]['bbox'] for i, j in entries])
        features = net(path, valid_boxes)
        scores = np.asarray(rcnn_model(features))
        cls_scores = scores[:, score_column]
        idx = frame_to_det_idx.get(fid, [])
        if idx:
            from ..utils.protocol import det_score
            frame_dets[fid] = (np.asarray([dets[k]['bbox'] for k in idx]),
                               np.asarray([det_score(dets[k], class_idx) for k in idx]))
        for (i, j), sc, feat, all_sc in zip(entries, cls_scores, features, scores):
            tub_boxes.append(tubelets_proto[i]['boxes'][j]['bbox'])
            tub_frames.append(fid)
            where.append((i, j))
            own.append((sc, feat, all_sc, path))
    if not where:
        return tubelets_proto
    arg, score = _pool_on_gpu(tub_boxes, tub_frames, frame_dets, overlap_thres, _lib.POOL_ARGMAX_SCORE)
    for (i, j), fid, a, s, (sc, feat, all_sc, path) in zip(where, tub_frames, arg, score, own):
        cur = tubelets_proto[i]['boxes'][j]
        max_score = frame_dets[fid][1][a] if a >= 0 else -np.inf        # :236-238
        if max_score > sc:                                                # :239
            max_box = frame_dets[fid][0][a].tolist()
            cur['det_score'] = max_score
            cur['bbox'] = max_box
            if save_feat or save_all_sc:
                max_feat = net(path, [max_box])
                if save_feat:
                    cur['feat'] = np.asarray(max_feat).ravel().tolist()
                if save_all_sc:
                    cur['all_score'] = np.asarray(rcnn_model(max_feat)).ravel().tolist()
        else:
            cur['det_score'] = sc
            if save_feat:
                cur['feat'] = np.asarray(feat).ravel().tolist()
            if save_all_sc:
                cur['all_score'] = np.asarray(all_sc).ravel().tolist()
    return tubelets_proto


def anchor_propagate(vid_proto, track_proto, det_proto, class_idx):
    """Give every box of a tubelet the class score of the det that overlaps its anchor box most.
    vdet/tubelet_cls.py:353-383."""
    assert vid_proto['video'] == track_proto['video']
    score_proto = {}
    score_proto['video'] = vid_proto['video']
    score_proto['method'] = "anchor_propagate"
    tubelets_proto = tubelets_proto_from_tracks_proto(track_proto['tracks'], class_idx)
    logging.info("Propagating anchor scores in {} for {}...".format(vid_proto['video'],
                 imagenet_vdet_classes[class_idx]))
    dets = det_proto['detections']
    frame_to_det_idx = defaultdict(list)
    for i, det in enumerate(dets):
        frame_to_det_idx[det['frame']].append(i)
    anchors = []
    for tubelet in tubelets_proto:
        anchor_box = [box for box in tubelet['boxes'] if box['anchor'] == 0]
        assert len(anchor_box) == 1
        anchors.append(anchor_box[0])
    frame_dets = {}
    for a in anchors:
        fid = a['frame']
        if fid in frame_dets:
            continue
        idx = frame_to_det_idx[fid]
        if len(idx) == 0:
            # the reference's iou() fails on the empty 1-D det array (common.py:455)
            raise IndexError("too many indices for array")
        frame_dets[fid] = (np.asarray([dets[i]['bbox'] for i in idx]),
                           np.asarray([dets[i]['scores'][class_idx - 1]['score'] for i in idx]))
    if anchors:
        arg, _ = _pool_on_gpu([a['bbox'] for a in anchors], [a['frame'] for a in anchors], frame_dets,
                              0.0, _lib.POOL_ARGMAX_IOU)
        for tubelet, a, k in zip(tubelets_proto, anchors, arg):
            anchor_score = frame_dets[a['frame']][1][k]
            for box in tubelet['boxes']:
                box['det_score'] = anchor_score
    score_proto['tubelets'] = tubelets_proto
    return score_proto


# --------------------------------------------------------------------------------------
# interpolation of strided tubelets
# --------------------------------------------------------------------------------------
_INTERP_FIELDS = ('x1', 'y1', 'x2', 'y2', 'det_score', 'anchor')


def score_proto_interpolation(score_proto, vid_proto):
    """Perform interpolation on score protocols if only part of the tracks are available.
    vdet/tubelet_cls.py:430-490: every tubelet with >= 2 boxes is densified over
    [min frame, max frame] (stretched to frame 1 / the last frame when it starts at 2 / ends one
    before the end, :472-475) by linear interpolation of bbox, det_score and anchor; linear
    extrapolation beyond the ends.  ValueError for gt tubelets (:447-448)."""
    new_score_proto = {}
    new_score_proto['video'] = score_proto['video']
    new_score_proto['method'] = score_proto['method'] + '_interpolation'
    max_frames = len(vid_proto['frames'])
    tubelets_proto = [None] * len(score_proto['tubelets'])
    todo = []
    for t_i, tubelet in enumerate(score_proto['tubelets']):
        if tubelet['gt'] == 1:
            raise ValueError('Dangerous: Score file contains gt tracks!')
        if len(tubelet['boxes']) < 2:
            tubelets_proto[t_i] = copy.copy(tubelet)
            continue
        todo.append(t_i)
    if todo:
        xs, ys, knot_off, dense_first, dense_off = [], [[] for _ in _INTERP_FIELDS], [0], [], [0]
        for t_i in todo:
            boxes = score_proto['tubelets'][t_i]['boxes']
            idx = np.asarray([b['frame'] for b in boxes])
            if len(set(idx.tolist())) != len(idx):
                # two boxes on one frame: interp1d divides by a zero knot spacing (nan / inf with a warning); the
                # result is undefined in the reference, refused here
                raise ValueError('score_proto_interpolation: tubelet {} has two boxes on the same frame'.format(t_i))
            order = np.argsort(idx, kind='mergesort')          # interp1d sorts its x (assume_sorted=False)
            vals = [[b['bbox'][0] for b in boxes], [b['bbox'][1] for b in boxes], [b['bbox'][2] for b in boxes],
                    [b['bbox'][3] for b in boxes], [b['det_score'] for b in boxes], [b['anchor'] for b in boxes]]
            xs.extend(idx[order].tolist())
            for f in range(6):
                ys[f].extend(np.asarray(vals[f], dtype=np.float64)[order].tolist())
            knot_off.append(len(xs))
            min_idx, max_idx = int(idx.min()), int(idx.max())
            if min_idx == 2:
                min_idx = 1
            if max_idx == max_frames - 1:
                max_idx = max_frames
            dense_first.append(min_idx)
            dense_off.append(dense_off[-1] + max_idx - min_idx + 1)
        dev = ops.default_device()
        out = ops.tubelet_interpolate(
            torch.tensor(xs, dtype=torch.float64, device=dev), torch.tensor(ys, dtype=torch.float64, device=dev),
            torch.tensor(knot_off, dtype=torch.int32, device=dev), torch.tensor(dense_first, dtype=torch.int32, device=dev),
            torch.tensor(dense_off, dtype=torch.int32, device=dev)).cpu().numpy()
        for k, t_i in enumerate(todo):
            tubelet = score_proto['tubelets'][t_i]
            new_tubelet = {key: tubelet[key] for key in ['gt', 'class', 'class_index']}
            cols = out[:, dense_off[k]:dense_off[k + 1]]
            new_tubelet['boxes'] = [
                {'frame': dense_first[k] + q, 'det_score': float(cols[4, q]), 'anchor': float(cols[5, q]),
                 'bbox': [float(cols[0, q]), float(cols[1, q]), float(cols[2, q]), float(cols[3, q])]}
                for q in range(cols.shape[1])]
            tubelets_proto[t_i] = new_tubelet
    new_score_proto['tubelets'] = tubelets_proto
    return new_score_proto
