"""Packed (struct-of-arrays) protos and their on-disk container -- SURVEY 8f row 4.

The reference keeps every detection / tubelet box as a JSON dict and reads it back with
``json.load`` (+ gzip) -- utils/protocol.py:209-236 -- and its raw-detection loaders walk one
``.mat`` file per frame (``load_frame_to_det`` / ``load_det_info``, utils/protocol.py:528-555).
Once the kernels take a millisecond, that dict walk IS the wall time.  This module keeps the proto
API (``unpack_proto`` returns dicts equal to the originals, key order included) while giving the
tensor entry points the same data as flat arrays:

  * ``pack_proto`` / ``unpack_proto``: det, track and score protos <-> columns
    (numbers -> float64 columns with an exact int/float type map, fixed-length number lists such as
    ``bbox`` -> 2-D columns, a det's ``scores`` list -> one [M, C] matrix + a class table, strings
    and anything irregular -> a JSON column); protos of any other shape are carried as JSON.
  * ``save_packed`` / ``load_packed``: one little-endian file ``VDETPK01 | header | 64-byte aligned
    raw arrays``; ``load_packed(mmap=True)`` maps the arrays instead of reading them, so a det
    matrix goes from the page cache to the pinned staging buffer in one streaming copy.
  * ``PackedDets``: the hot path's view of a packed det proto / raw detections: frame-grouped
    float32 boxes + scores + segment offsets (what ``ops.nms_frames`` and
    ``VideoPostProcessor.stage`` consume) and the reference's ``det_info`` / ``frame_to_det`` forms.

Everything here is host-side NumPy; nothing imports torch until ``PackedDets.device_tensors``.
"""
import json
import os
import struct

import numpy as np

MAGIC = b"VDETPK01"
_ALIGN = 64


# ------------------------------------------------------------------------------------------
# container
# ------------------------------------------------------------------------------------------
def save_packed(path, meta, arrays):
    """Write ``meta`` (JSON-serialisable dict) and ``arrays`` (name -> ndarray) to ``path``.

    Layout: MAGIC (8 B) | header length u64 | header JSON | zero padding to 64 B | array 0 | pad | ...
    The header lists dtype, shape and absolute byte offset of every array."""
    names = sorted(arrays)
    blobs = []
    for n in names:
        a = np.asarray(arrays[n])
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        blobs.append(np.ascontiguousarray(a))

    def header(offsets):
        return json.dumps({"meta": meta, "arrays": {
            n: {"dtype": b.dtype.str, "shape": list(b.shape), "offset": o}
            for n, b, o in zip(names, blobs, offsets)}}, separators=(",", ":")).encode("utf-8")

    # the header's own length decides the offsets it lists: iterate to the fixed point
    offsets = [0] * len(blobs)
    for _ in range(8):
        base = 16 + len(header(offsets))
        pos, new = _round_up(base), []
        for b in blobs:
            new.append(pos)
            pos = _round_up(pos + b.nbytes)
        if new == offsets:
            break
        offsets = new
    else:
        raise RuntimeError("save_packed: header layout did not converge")
    head = header(offsets)
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(head)))
        f.write(head)
        for b, o in zip(blobs, offsets):
            f.write(b"\0" * (o - f.tell()))
            f.write(b.tobytes() if b.ndim == 0 else memoryview(b.reshape(-1).view(np.uint8)))
    os.replace(tmp, path)


def load_packed(path, mmap=True):
    """Read a container written by :func:`save_packed`; returns ``(meta, arrays)``.
    ``mmap=True`` maps the arrays read-only (no copy, pages come in on first touch)."""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("%s: not a VDETPK01 container" % path)
        (hlen,) = struct.unpack("<Q", f.read(8))
        head = json.loads(f.read(hlen).decode("utf-8"))
        arrays = {}
        if mmap:
            whole = np.memmap(path, dtype=np.uint8, mode="r")
        for name, d in head["arrays"].items():
            dt = np.dtype(d["dtype"])
            count = int(np.prod(d["shape"], dtype=np.int64))
            if mmap:
                raw = whole[d["offset"]:d["offset"] + count * dt.itemsize]
                arrays[name] = raw.view(dt).reshape(d["shape"])
            else:
                f.seek(d["offset"])
                arrays[name] = np.frombuffer(f.read(count * dt.itemsize), dtype=dt).reshape(d["shape"])
    return head["meta"], arrays


def _round_up(x):
    return (x + _ALIGN - 1) // _ALIGN * _ALIGN


# ------------------------------------------------------------------------------------------
# records <-> columns
# ------------------------------------------------------------------------------------------
def _is_num(v):
    return (type(v) is int or type(v) is float)


def _is_score_list(v):
    return (type(v) is list and len(v) > 0 and
            all(type(s) is dict and list(s.keys()) == ["class", "class_index", "score"] and _is_num(s["score"])
                for s in v))


def _encode_records(records, prefix, arrays):
    """Columnar form of a list of dicts.  Returns the JSON part; numeric columns go to ``arrays``
    under ``prefix + key``.  Every record's key ORDER is kept (one order per distinct key tuple)."""
    n = len(records)
    orders, order_of = [], {}
    order_id = np.zeros(n, dtype=np.int32)
    for i, r in enumerate(records):
        if type(r) is not dict:
            raise TypeError("record %d is not a dict" % i)
        k = tuple(r.keys())
        if k not in order_of:
            order_of[k] = len(orders)
            orders.append(list(k))
        order_id[i] = order_of[k]
    desc = {"n": n, "orders": orders, "cols": {}}
    if len(orders) > 1:
        arrays[prefix + "__order"] = order_id
    keys = []
    for o in orders:
        keys += [k for k in o if k not in keys]
    for key in keys:
        present = np.fromiter((key in r for r in records), dtype=bool, count=n)
        vals = [r[key] for r in records if key in r]
        col = {"all": bool(present.all())}                        # (who carries the key follows from the orders)
        name = prefix + key
        if all(_is_num(v) for v in vals):
            col["t"] = "num"
            arrays[name] = np.asarray(vals, dtype=np.float64).reshape(len(vals))
            _int_map(vals, name, col, arrays)
        elif (vals and all(type(v) is list and len(v) == len(vals[0]) and all(_is_num(x) for x in v) for v in vals)
              and len(vals[0]) > 0):
            col["t"] = "vec"
            arrays[name] = np.asarray(vals, dtype=np.float64).reshape(len(vals), len(vals[0]))
            _int_map([x for v in vals for x in v], name, col, arrays)
        elif vals and all(_is_score_list(v) for v in vals) and _same_class_table(vals):
            col["t"] = "scores"
            col["classes"] = [s["class"] for s in vals[0]]
            col["class_index"] = [s["class_index"] for s in vals[0]]
            flat = [s["score"] for v in vals for s in v]
            arrays[name] = np.asarray(flat, dtype=np.float64).reshape(len(vals), len(vals[0]))
            _int_map(flat, name, col, arrays)
        else:
            col["t"] = "json"
            col["v"] = vals
        desc["cols"][key] = col
    return desc


def _same_class_table(vals):
    first = [(s["class"], s["class_index"]) for s in vals[0]]
    return all(len(v) == len(first) and all((s["class"], s["class_index"]) == f for s, f in zip(v, first))
               for v in vals)


def _int_map(flat, name, col, arrays):
    """Which numbers were Python ints (JSON ``3`` vs ``3.0``): all / none / a per-element bitmap.
    Integers that float64 cannot hold exactly are refused rather than silently rounded."""
    ints = np.fromiter((type(x) is int for x in flat), dtype=bool, count=len(flat))
    if ints.any():
        big = [x for x in flat if type(x) is int and abs(x) > (1 << 53)]
        if big:
            raise OverflowError("integer %d does not fit the packed float64 column %s" % (big[0], name))
    if ints.all():
        col["int"] = "all"
    elif not ints.any():
        col["int"] = "none"
    else:
        col["int"] = "map"
        arrays[name + "__int"] = np.packbits(ints)


def _column_values(col, name, arrays):
    """Python values of one column, in record order (only the records that carry the key)."""
    t = col["t"]
    if t == "json":
        return col["v"]
    a = np.asarray(arrays[name])
    flat = a.reshape(-1).tolist()                               # Python floats
    mode = col["int"]
    if mode == "all":
        flat = [int(x) for x in flat]
    elif mode == "map":
        ints = np.unpackbits(np.asarray(arrays[name + "__int"]), count=len(flat)).astype(bool).tolist()
        flat = [int(x) if i else x for x, i in zip(flat, ints)]
    if t == "num":
        return flat
    w = a.shape[1]
    rows = [flat[i * w:(i + 1) * w] for i in range(a.shape[0])]
    if t == "vec":
        return rows
    names, idx = col["classes"], col["class_index"]
    return [[{"class": c, "class_index": k, "score": s} for c, k, s in zip(names, idx, row)] for row in rows]


def _decode_records(desc, prefix, arrays):
    n = desc["n"]
    orders = desc["orders"]
    order_id = (np.asarray(arrays[prefix + "__order"]).tolist() if len(orders) > 1 else [0] * n)
    iters = {}
    for key, col in desc["cols"].items():
        iters[key] = iter(_column_values(col, prefix + key, arrays))
    out = []
    for i in range(n):
        out.append({k: next(iters[k]) for k in orders[order_id[i]]})
    return out


# ------------------------------------------------------------------------------------------
# protos
# ------------------------------------------------------------------------------------------
def _ragged(groups):
    """list of lists of records -> (offsets int64 [G+1], flat record list)."""
    off = np.zeros(len(groups) + 1, dtype=np.int64)
    np.cumsum([len(g) for g in groups], out=off[1:])
    return off, [r for g in groups for r in g]


def pack_proto(proto):
    """det / track / score proto (utils/protocol.py:32-154) -> ``(meta, arrays)``.  Any other
    JSON-style object is carried in ``meta`` unchanged (kind "json")."""
    arrays = {}
    try:
        if type(proto) is dict and type(proto.get("detections")) is list:
            top = {k: v for k, v in proto.items() if k != "detections"}
            meta = {"kind": "det", "top": top, "top_order": list(proto.keys()),
                    "detections": _encode_records(proto["detections"], "det.", arrays)}
            return meta, arrays
        if type(proto) is dict and type(proto.get("tracks")) is list and all(type(t) is list for t in proto["tracks"]):
            off, flat = _ragged(proto["tracks"])
            arrays["track.__offsets"] = off
            top = {k: v for k, v in proto.items() if k != "tracks"}
            meta = {"kind": "track", "top": top, "top_order": list(proto.keys()),
                    "boxes": _encode_records(flat, "track.", arrays)}
            return meta, arrays
        if (type(proto) is dict and type(proto.get("tubelets")) is list and
                all(type(t) is dict and type(t.get("boxes")) is list for t in proto["tubelets"])):
            off, flat = _ragged([t["boxes"] for t in proto["tubelets"]])
            arrays["tub.__offsets"] = off
            heads = [{k: (None if k == "boxes" else v) for k, v in t.items()} for t in proto["tubelets"]]
            top = {k: v for k, v in proto.items() if k != "tubelets"}
            meta = {"kind": "score", "top": top, "top_order": list(proto.keys()),
                    "tubelets": _encode_records(heads, "tubhead.", arrays),
                    "boxes": _encode_records(flat, "tub.", arrays)}
            return meta, arrays
    except (TypeError, OverflowError):
        arrays.clear()
    return {"kind": "json", "value": proto}, {}


def unpack_proto(meta, arrays):
    """Inverse of :func:`pack_proto`: a proto equal to the original (values, types, key order)."""
    kind = meta["kind"]
    if kind == "json":
        return meta["value"]
    if kind == "det":
        body = {"detections": _decode_records(meta["detections"], "det.", arrays)}
    elif kind == "track":
        flat = _decode_records(meta["boxes"], "track.", arrays)
        off = np.asarray(arrays["track.__offsets"]).tolist()
        body = {"tracks": [flat[a:b] for a, b in zip(off[:-1], off[1:])]}
    elif kind == "score":
        flat = _decode_records(meta["boxes"], "tub.", arrays)
        off = np.asarray(arrays["tub.__offsets"]).tolist()
        heads = _decode_records(meta["tubelets"], "tubhead.", arrays)
        for h, a, b in zip(heads, off[:-1], off[1:]):
            h["boxes"] = flat[a:b]
        body = {"tubelets": heads}
    else:
        raise ValueError("unknown packed proto kind %r" % kind)
    out = {}
    for k in meta["top_order"]:
        out[k] = body[k] if k in body else meta["top"][k]
    return out


def proto_dump_packed(obj, file_path):
    """``proto_dump`` (utils/protocol.py:223-236) into the packed container."""
    meta, arrays = pack_proto(obj)
    save_packed(file_path, meta, arrays)


def proto_load_packed(file_path, mmap=True):
    """``proto_load`` (utils/protocol.py:209-220) from the packed container: dicts out."""
    meta, arrays = load_packed(file_path, mmap=mmap)
    return unpack_proto(meta, arrays)


# ------------------------------------------------------------------------------------------
# the hot path's view of detections
# ------------------------------------------------------------------------------------------
class PackedDets(object):
    """Detections of one video as flat arrays: ``frames`` int64 [M] (1-based frame ids),
    ``boxes`` float64 [M,4], ``scores`` float64 [M,C], in the order of the source."""

    def __init__(self, video, frames, boxes, scores, classes=None, class_index=None):
        self.video = video
        self.frames = np.asarray(frames, dtype=np.int64).reshape(-1)
        self.boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
        scores = np.asarray(scores, dtype=np.float64)
        self.scores = scores if scores.ndim == 2 else scores.reshape(self.boxes.shape[0], -1 if scores.size else 0)
        if not (self.frames.shape[0] == self.boxes.shape[0] == self.scores.shape[0]):
            raise ValueError("PackedDets: frames / boxes / scores disagree on the number of detections")
        self.classes = list(classes) if classes is not None else None
        self.class_index = list(class_index) if class_index is not None else list(range(self.scores.shape[1]))
        self.boxes_are_int = False      # every bbox number of the source proto was a Python int

    # ---- constructors ----------------------------------------------------------------------
    @classmethod
    def from_packed(cls, meta, arrays):
        """From a packed det proto (``load_packed`` of a file written by ``proto_dump_packed``)."""
        if meta.get("kind") == "dets":
            out = cls(meta["video"], arrays["frames"], arrays["boxes"], arrays["scores"],
                      meta.get("classes"), meta.get("class_index"))
            out.boxes_are_int = bool(meta.get("boxes_are_int", False))
            return out
        if meta.get("kind") != "det":
            raise ValueError("not a packed det proto")
        cols = meta["detections"]["cols"]
        for key, want in (("frame", "num"), ("bbox", "vec"), ("scores", "scores")):
            if key not in cols or cols[key]["t"] != want or not cols[key]["all"]:
                raise ValueError("packed det proto has no regular %r column" % key)
        sc = cols["scores"]
        out = cls(meta["top"].get("video"), arrays["det.frame"], arrays["det.bbox"], arrays["det.scores"],
                  sc["classes"], sc["class_index"])
        out.boxes_are_int = cols["bbox"]["int"] == "all"
        return out

    @classmethod
    def from_det_proto(cls, det_proto):
        return cls.from_packed(*pack_proto(det_proto))

    @classmethod
    def from_frame_to_det(cls, video, frame_to_det):
        """From the reference's raw-detection dict frame -> (boxes [N,4], zs [N,C])
        (``load_frame_to_det``, utils/protocol.py:528-540), frames in ascending order."""
        frames = sorted(frame_to_det)
        counts = [len(frame_to_det[f][0]) for f in frames]
        fr = np.repeat(np.asarray(frames, dtype=np.int64), counts)
        live = [f for f, c in zip(frames, counts) if c]
        if not live:
            return cls(video, fr, np.zeros((0, 4)), np.zeros((0, 0)))
        return cls(video, fr, np.concatenate([np.asarray(frame_to_det[f][0], dtype=np.float64).reshape(-1, 4) for f in live]),
                   np.concatenate([np.asarray(frame_to_det[f][1], dtype=np.float64) for f in live]))

    # ---- reference forms -----------------------------------------------------------------
    def to_det_info(self):
        """``load_det_info``'s array: rows [frame, x1, y1, x2, y2, score_0 ...] float64 (:542-555)."""
        return np.concatenate([self.frames[:, None].astype(np.float64), self.boxes, self.scores], axis=1)

    def to_frame_to_det(self):
        off, order, seg_frames = self.frame_segments()
        return {int(f): (self.boxes[order[a:b]], self.scores[order[a:b]])
                for f, a, b in zip(seg_frames, off[:-1], off[1:])}

    def to_det_proto(self):
        """A det proto (utils/protocol.py:77-110) with the packed values (frames as ints, bboxes and
        scores as Python floats); ``classes`` must be known."""
        if self.classes is None:
            raise ValueError("to_det_proto: class names unknown")
        names, idx = self.classes, self.class_index
        return {"video": self.video, "detections": [
            {"frame": int(f), "bbox": b, "scores": [{"class": c, "class_index": k, "score": s}
                                                    for c, k, s in zip(names, idx, row)]}
            for f, b, row in zip(self.frames.tolist(), self.boxes.tolist(), self.scores.tolist())]}

    # ---- tensor forms ----------------------------------------------------------------------
    def frame_segments(self, as_f32=False):
        """Stable grouping by frame: ``(seg_offsets int64 [S+1], order int64 [M], seg_frames [S])``;
        ``order[seg_offsets[s]:seg_offsets[s+1]]`` are the source rows of frame ``seg_frames[s]``.
        ``as_f32``: group on the float32 cast of the frame ids, as ``vid_nms`` sees them -- ``apply_vid_nms`` puts
        the frame id into a float32 matrix (vdet/video_det.py:53-56), so ids that collide there (beyond 2^24) are
        ONE frame to the suppression."""
        keys = np.asarray(self.frames, dtype=np.float32) if as_f32 else self.frames
        order = np.argsort(keys, kind="stable")
        seg_frames, counts = np.unique(keys, return_counts=True)
        off = np.zeros(len(seg_frames) + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        return off, order, seg_frames

    def grouped_f32(self):
        """Frame-grouped float32 arrays for the kernels: ``(boxes [M,4], scores [M,C], seg_offsets
        int32 [S+1], order int64 [M], max_seg_len)`` -- the cast ``apply_vid_nms`` performs on its
        ``[M,6]`` matrix (vdet/video_det.py:53-56)."""
        off, order, _ = self.frame_segments(as_f32=True)
        if off[-1] >= (1 << 31):
            raise OverflowError("more than 2^31 detections")
        return (np.ascontiguousarray(self.boxes[order], dtype=np.float32),
                np.ascontiguousarray(self.scores[order], dtype=np.float32),
                off.astype(np.int32), order, int(np.diff(off).max()) if len(off) > 1 else 0)

    def uniform_f32(self):
        """``(boxes [T,N,4], scores [T,N,C])`` float32 when every frame holds the same number of
        detections (what ``VideoPostProcessor.stage`` takes); ValueError otherwise."""
        b, s, off, _, n = self.grouped_f32()
        T = len(off) - 1
        if T == 0 or not np.all(np.diff(off) == n):
            raise ValueError("frames hold different numbers of detections")
        return b.reshape(T, n, 4), s.reshape(T, n, -1)

    def device_tensors(self, device=None):
        """Frame-grouped CUDA tensors ``(boxes, scores, seg_offsets, max_seg_len, order)``."""
        import torch
        from .. import ops
        dev = device or ops.default_device()
        b, s, off, order, n = self.grouped_f32()
        return (torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(off).to(dev), n, order)

    # ---- file --------------------------------------------------------------------------------
    def save(self, path):
        save_packed(path, {"kind": "dets", "video": self.video, "classes": self.classes,
                           "class_index": self.class_index, "boxes_are_int": bool(self.boxes_are_int)},
                    {"frames": self.frames, "boxes": self.boxes, "scores": self.scores})

    @classmethod
    def load(cls, path, mmap=True):
        return cls.from_packed(*load_packed(path, mmap=mmap))


def packed_vid_nms(dets, thresh=0.3, device=None):
    """``apply_vid_nms`` for EVERY class of a :class:`PackedDets` in one kernel launch.

    Returns a list with one int64 array per class column: the rows (indices into the packed
    source order, i.e. into ``det_proto['detections']``) that ``vid_nms`` keeps for the matrix
    ``[frame, bbox, score_c]`` cast to float32 -- suppression per frame, keep order = global
    descending score (utils/nms.pyx:71-125; ties: ascending row, the documented rule)."""
    import torch
    from .. import ops
    M, C = dets.scores.shape
    if M == 0:
        return [np.zeros(0, dtype=np.int64) for _ in range(C)]
    boxes, scores, seg, max_len, order = dets.device_tensors(device)
    _, _, keep_mask, status = ops.nms_frames(boxes, scores, seg, thresh, max_len, want_mask=True)
    ops.raise_for_status(status)
    src = torch.from_numpy(order).to(boxes.device)                    # grouped row -> source row
    # vid_nms sorts ALL rows by score first (stable on the source order), so the kept rows come out
    # in descending score with ties by ascending SOURCE row
    src_sorted, perm = torch.sort(src)
    out = []
    for c in range(C):
        kept = keep_mask[c][perm].bool()                              # in source-row order
        rows = src_sorted[kept]
        sc = scores[:, c][perm][kept].contiguous()
        _, ids = ops.sort_by_score_desc(sc, rows.contiguous())
        out.append(ids.cpu().numpy())
    return out


# ------------------------------------------------------------------------------------------
# raw detections on disk (.mat per frame) -- utils/protocol.py:528-555
# ------------------------------------------------------------------------------------------
def _frame_mat_files(vid_proto, det_dir):
    """(frame id, path) of every frame's score file that exists -- the lookup both reference
    loaders share (:531-537, :546-551): ``<basename>.mat`` first, then ``<path>.mat``."""
    for frame in vid_proto['frames']:
        basename = os.path.splitext(frame['path'])[0]
        score_file = os.path.join(det_dir, basename + '.mat')
        if not os.path.isfile(score_file):
            score_file = os.path.join(det_dir, frame['path'] + '.mat')
        if os.path.isfile(score_file):
            yield frame['frame'], score_file


def load_frame_to_det(vid_proto, det_dir):
    """utils/protocol.py:528-540: frame id -> (boxes, zs) straight from the per-frame .mat files."""
    import scipy.io as sio
    frame_to_det = {}
    for frame_id, score_file in _frame_mat_files(vid_proto, det_dir):
        d = sio.loadmat(score_file)
        frame_to_det[frame_id] = (d['boxes'], d['zs'])
    return frame_to_det


def load_det_info(vid_proto, det_dir):
    """utils/protocol.py:542-555: ``[[frame_id, x1, y1, x2, y2, scores...], ...]`` as a float64 array
    (frames without boxes are skipped, :552).  One concatenate per frame instead of a Python list per box."""
    import scipy.io as sio
    rows = []
    for frame_id, score_file in _frame_mat_files(vid_proto, det_dir):
        d = sio.loadmat(score_file)
        if d['boxes'].size == 0:
            continue
        n = min(len(d['boxes']), len(d['zs']))                       # zip() stops at the shorter one
        rows.append(np.concatenate([np.full((n, 1), frame_id, dtype=np.float64),
                                    np.asarray(d['boxes'][:n], dtype=np.float64),
                                    np.asarray(d['zs'][:n], dtype=np.float64)], axis=1))
    if not rows:
        return np.asarray([])
    return np.concatenate(rows, axis=0)


def pack_det_dir(vid_proto, det_dir, out_path=None):
    """Read a directory of per-frame .mat files ONCE and return (and optionally save) the
    :class:`PackedDets` side-car; later runs ``PackedDets.load`` it instead of T ``loadmat`` calls."""
    dets = PackedDets.from_frame_to_det(vid_proto['video'], load_frame_to_det(vid_proto, det_dir))
    if out_path is not None:
        dets.save(out_path)
    return dets


# ------------------------------------------------------------------------------------------
# the temporal kernels' view of a score proto
# ------------------------------------------------------------------------------------------
class PackedTubelets(object):
    """A score proto (utils/protocol.py:111-154) as ragged arrays: ``offsets`` int64 [K+1] and one flat
    column per numeric box field (``det_score``, ``frame``, ``anchor``, ``track_score`` ..., ``bbox`` as
    [n,4]).  The temporal stages of vdet/tubelet_cls.py run on ``padded(key)`` = [K, Lmax] rows +
    lengths -- exactly what the dict adapters build with a Python loop per call -- and write back with
    ``set_padded``; ``to_score_proto()`` returns the dict form with the updated values.

    In-place stages (same arithmetic, kernels and exceptions as the dict adapters of
    vdetlib_b200.vdet.tubelet_cls): ``complete_scores_`` (:284-303), ``temporal_maxpool_`` (:386-414),
    ``conv_scores_`` (:15-51 with a TemporalConvNet)."""

    def __init__(self, meta, arrays):
        if meta.get("kind") != "score":
            raise ValueError("not a packed score proto")
        self.meta = json.loads(json.dumps(meta))                  # private copy: stages edit it
        self.arrays = dict(arrays)
        self.offsets = np.asarray(arrays["tub.__offsets"], dtype=np.int64)

    @classmethod
    def from_score_proto(cls, score_proto):
        meta, arrays = pack_proto(score_proto)
        return cls(meta, arrays)

    @classmethod
    def load(cls, path, mmap=True):
        return cls(*load_packed(path, mmap=mmap))

    def save(self, path):
        save_packed(path, self.meta, self.arrays)

    def to_score_proto(self):
        return unpack_proto(self.meta, self.arrays)

    # ---- columns -----------------------------------------------------------------------------
    @property
    def n_tubelets(self):
        return len(self.offsets) - 1

    @property
    def lengths(self):
        return np.diff(self.offsets).astype(np.int32)

    def _col(self, key):
        col = self.meta["boxes"]["cols"].get(key)
        if col is None or col["t"] not in ("num", "vec") or not col["all"]:
            raise KeyError("score proto has no regular numeric box field %r" % key)
        return col

    def column(self, key):
        """Flat float64 values of a box field, tubelet after tubelet (read-only view)."""
        self._col(key)
        return np.asarray(self.arrays["tub." + key])

    def head(self, key):
        """Per-tubelet values of a tubelet field (``class_index``, ``gt`` ...)."""
        col = self.meta["tubelets"]["cols"][key]
        return _column_values(col, "tubhead." + key, self.arrays)

    def padded(self, key="det_score", fill=0.0):
        """``(rows float64 [K', Lmax], lengths int32 [K'], index int64 [K'])`` over the tubelets that
        have at least one box (what the dict adapters process); padding = ``fill``."""
        vals = self.column(key).astype(np.float64, copy=False)
        lens = self.lengths
        index = np.nonzero(lens > 0)[0]
        lens = lens[index]
        L = int(lens.max()) if len(lens) else 0
        rows = np.full((len(index), max(L, 1)), fill, dtype=np.float64)
        if len(index):
            col_id = np.arange(max(L, 1))[None, :]
            take = col_id < lens[:, None]
            rows[take] = vals                                       # row-major: tubelet after tubelet
        return rows, lens, index

    def set_padded(self, key, rows, lens, where=None):
        """Scatter padded rows back into the flat column ``key`` (created if the proto has none, placed
        last in every box -- ``conv_score``).  ``where`` (bool [K', Lmax]) restricts the update."""
        rows = np.asarray(rows, dtype=np.float64)
        take = np.arange(rows.shape[1])[None, :] < np.asarray(lens)[:, None]
        name = "tub." + key
        cols = self.meta["boxes"]["cols"]
        if key not in cols:
            if int(take.sum()) != int(self.offsets[-1]):
                raise ValueError("a new box field must be written for every box")
            cols[key] = {"all": True, "t": "num", "int": "none"}
            for order in self.meta["boxes"]["orders"]:
                order.append(key)
            self.arrays[name] = rows[take].copy()
            return
        self._col(key)
        flat = np.array(self.arrays[name], dtype=np.float64, copy=True)   # mmap / shared -> private
        new = rows[take]
        if where is None:
            flat[...] = new
            changed = np.ones(flat.shape, dtype=bool)
        else:
            changed = np.asarray(where, dtype=bool)[take]
            flat[changed] = new[changed]
        self.arrays[name] = flat
        # the written values are floats: fix the int/float type map of the column
        col = cols[key]
        if col["int"] == "all":
            ints = ~changed
        elif col["int"] == "map":
            ints = np.unpackbits(np.asarray(self.arrays[name + "__int"]), count=flat.size).astype(bool) & ~changed
        else:
            ints = None
        if ints is not None:
            self.arrays.pop(name + "__int", None)
            if ints.all():
                col["int"] = "all"
            elif not ints.any():
                col["int"] = "none"
            else:
                col["int"] = "map"
                self.arrays[name + "__int"] = np.packbits(ints)

    @classmethod
    def for_pooling(cls, vid_proto, track_proto, class_idx, overlap_thres=0.7):
        """The score-proto shell dets_spatial_max_pooling starts from (vdet/tubelet_cls.py:306-311): video,
        method name, tubelets of ``track_proto`` with ``det_score = -1e5`` (utils/protocol.py:448-464)."""
        from .protocol import tubelets_proto_from_tracks_proto
        assert vid_proto['video'] == track_proto['video']
        return cls.from_score_proto({
            'video': vid_proto['video'], 'method': "spatial_max_pooling_IOU_{}".format(overlap_thres),
            'tubelets': tubelets_proto_from_tracks_proto(track_proto['tracks'], class_idx)})

    def _write_rows(self, key, rows, values, ints):
        """Overwrite rows ``rows`` of column ``key`` (num: [n], vec: [n,w]) keeping the int/float type map exact."""
        col = self._col(key)
        name = "tub." + key
        arr = np.array(self.arrays[name], dtype=np.float64, copy=True)
        arr[rows] = values
        self.arrays[name] = arr
        width = 1 if arr.ndim == 1 else arr.shape[1]
        if col["int"] == ("all" if ints else "none"):
            return
        old = col["int"]
        if old == "map":
            m = np.unpackbits(np.asarray(self.arrays[name + "__int"]), count=arr.size).astype(bool)
        else:
            m = np.full(arr.size, old == "all")
        m = m.reshape(-1, width)
        m[rows] = bool(ints)
        m = m.reshape(-1)
        self.arrays.pop(name + "__int", None)
        if m.all():
            col["int"] = "all"
        elif not m.any():
            col["int"] = "none"
        else:
            col["int"] = "map"
            self.arrays[name + "__int"] = np.packbits(m)

    # ---- stages (GPU) ------------------------------------------------------------------------
    def pool_dets_(self, dets, class_idx, vid_proto, overlap_thres=0.7, device=None):
        """Spatial max-pooling of packed detections onto the tubelets, the body of dets_spatial_max_pooling /
        raw_dets_spatial_max_pooling (vdet/tubelet_cls.py:324-347, :509-532) without a dict walk over the
        detections: for every tubelet box on a frame of ``vid_proto`` that has detections, among the dets with
        IoU > overlap_thres the FIRST arg-max of the class score (column ``class_idx - 1``, positional like
        :329 / :514) gives ``det_score`` and ``bbox``; none -> ``det_score = -1e5``, bbox unchanged.  Follow with
        ``complete_scores_()`` as the reference does (:349, :534)."""
        from .. import _lib, ops
        off, order, seg_frames = dets.frame_segments()
        if not len(seg_frames) or not int(self.offsets[-1]):
            return self
        vid_frames = np.asarray([f['frame'] for f in vid_proto['frames']], dtype=np.int64)
        tub_frames = self.column('frame').astype(np.int64)
        seg = np.searchsorted(seg_frames, tub_frames)
        seg[seg >= len(seg_frames)] = 0
        has = (seg_frames[seg] == tub_frames) & np.isin(tub_frames, vid_frames)
        sel = np.nonzero(has)[0]
        if not len(sel):
            return self
        det_boxes = np.ascontiguousarray(dets.boxes[order])
        det_scores = np.ascontiguousarray(dets.scores[order, class_idx - 1])
        arg, score = ops.spatial_maxpool(
            self._to_dev(self.column('bbox')[sel].astype(np.float64), device), self._to_dev(seg[sel].astype(np.int32), device),
            self._to_dev(det_boxes, device), self._to_dev(det_scores, device), self._to_dev(off.astype(np.int32), device),
            overlap_thres, _lib.POOL_ARGMAX_SCORE)
        arg, score = arg.cpu().numpy().astype(np.int64), score.cpu().numpy()
        hit = arg >= 0
        new_score = np.where(hit, score, -1e5)
        self._write_rows('det_score', sel, new_score, ints=False)
        if hit.any():
            self._write_rows('bbox', sel[hit], det_boxes[arg[hit]], ints=dets.boxes_are_int)
        return self

    @staticmethod
    def _to_dev(a, device):
        import torch
        from .. import ops
        dev = device if device is not None else ops.default_device()
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def complete_scores_(self, device=None):
        """do_score_completion (vdet/tubelet_cls.py:284-303): runs of ``det_score <= -10`` are filled from
        their valid neighbours; IndexError when a tubelet has no valid score at all (:295)."""
        from .. import ops
        rows, lens, _ = self.padded("det_score")
        if not len(lens):
            return self
        missing = ~(rows > -10)                                     # before the kernel: rows may alias the tensor
        dev_rows = self._to_dev(rows, device)
        status = ops.score_completion_(dev_rows, self._to_dev(lens, device))
        ops.raise_for_status(status)
        self.set_padded("det_score", dev_rows.cpu().numpy(), lens, where=missing)
        return self

    def temporal_maxpool_(self, window_size, device=None):
        """score_proto_temporal_maxpool (vdet/tubelet_cls.py:386-414).  ValueError for an even window or
        a gt tubelet -- raised BEFORE anything is written (the reference has rewritten the tubelets in
        front of the gt one by then, :395-412)."""
        from .. import ops
        if window_size == 1:
            return self
        if window_size % 2 != 1:
            raise ValueError('Window size must be odd!')
        if "gt" in self.meta["tubelets"]["cols"] and any(g == 1 for g in self.head("gt")):
            raise ValueError('Dangerous: Score file contains gt tracks!')
        self.meta["top"]["method"] = self.meta["top"]["method"] + '_temporal_maxpool_{}'.format(window_size)
        rows, lens, _ = self.padded("det_score", fill=-1e5)
        if len(lens):
            out = ops.temporal_maxpool(self._to_dev(rows, device), window_size, self._to_dev(lens, device))
            self.set_padded("det_score", out.cpu().numpy(), lens)
        return self

    def conv_scores_(self, net, device=None):
        """score_conv_cls (vdet/tubelet_cls.py:15-51) with a TemporalConvNet: writes ``conv_score``."""
        from .. import ops
        lens_all = self.lengths
        lens = lens_all[lens_all > 0]                               # empty tubelets have no box to score
        if not len(lens):
            return self
        lens_dev = self._to_dev(lens, device)
        total = None
        length = np.repeat(lens.astype(np.float64), lens)           # tubelet length per box
        for name, taps in net.taps.items():
            if name == 'det_scores':
                flat = self.column('det_score')
            elif name == 'track_scores':
                flat = self.column('track_score')
            elif name == 'anchors':
                flat = self.column('anchor') * 1. / length                          # :28
            elif name == 'abs_anchors':
                flat = np.abs(self.column('anchor') * 1. / length)                  # :30
            elif name == 'gt_overlaps':
                flat = self.column('gt_overlap')
            else:
                raise KeyError(name)
            rows = np.zeros((len(lens), int(lens.max())), dtype=np.float64)
            rows[np.arange(rows.shape[1])[None, :] < lens[:, None]] = flat
            y = ops.temporal_conv1d(self._to_dev(rows, device), self._to_dev(np.asarray(taps, np.float64).reshape(1, -1), device),
                                    net.pad_mode, lens_dev)
            total = y if total is None else total + y
        self.set_padded("conv_score", (total + net.bias).cpu().numpy(), lens)
        return self
