"""Packed protos and raw-detection loaders -- host side, no GPU.

Pins: proto round trips against the committed golden protos (which the reference's own functions
produced); ``load_det_info`` / ``load_frame_to_det`` against the reference's functions run here by
oracle/gen_golden.py (tests/golden/det_mat.npz) and, when /root/reference is mounted, live."""
import copy
import json
import os

import numpy as np
import pytest

from oracle import gen_golden, oracle_np, ref_py2
from vdetlib_b200.utils import packed, protocol

import helpers


def _protos():
    with open(os.path.join(helpers.GOLDEN, "protos.json")) as f:
        return json.load(f)


def _same(a, b):
    """Equal values, equal types (3 vs 3.0), equal key order."""
    return a == b and json.dumps(a) == json.dumps(b)


def _all_golden_protos():
    g = _protos()
    out = [("det", g["det"]), ("track", g["track"]), ("vid", g["vid"]), ("annot", g["annot"]),
           ("interp_in", g["interp_in"]), ("anchor_track", g["anchor_track"])]
    out += [(k, v) for k, v in g["out"].items() if isinstance(v, dict)]
    return out


def test_round_trip_of_every_golden_proto(tmp_path):
    kinds = set()
    for name, proto in _all_golden_protos():
        meta, arrays = packed.pack_proto(proto)
        kinds.add(meta["kind"])
        assert _same(packed.unpack_proto(meta, arrays), proto), name
        path = str(tmp_path / (name + ".vdetpk"))
        packed.save_packed(path, meta, arrays)
        for mm in (True, False):
            assert _same(packed.unpack_proto(*packed.load_packed(path, mmap=mm)), proto), name
    assert kinds == {"det", "track", "score", "json"}


def test_columns_hold_the_numbers():
    det = _protos()["det"]
    meta, arrays = packed.pack_proto(det)
    cols = meta["detections"]["cols"]
    assert cols["frame"]["t"] == "num" and cols["frame"]["int"] == "all"
    assert cols["bbox"]["t"] == "vec" and cols["scores"]["t"] == "scores" and cols["hash"]["t"] == "json"
    M = len(det["detections"])
    assert arrays["det.bbox"].shape == (M, 4) and arrays["det.scores"].shape[0] == M
    assert arrays["det.bbox"].dtype == np.float64 and arrays["det.scores"].dtype == np.float64
    for i in (0, M // 2, M - 1):
        d = det["detections"][i]
        assert arrays["det.bbox"][i].tolist() == d["bbox"]
        assert arrays["det.scores"][i].tolist() == [s["score"] for s in d["scores"]]
    sp = _protos()["out"]["smp_1"]
    meta, arrays = packed.pack_proto(sp)
    off = arrays["tub.__offsets"]
    assert off.tolist() == np.cumsum([0] + [len(t["boxes"]) for t in sp["tubelets"]]).tolist()
    flat = [b["det_score"] for t in sp["tubelets"] for b in t["boxes"]]
    assert arrays["tub.det_score"].tolist() == flat


def test_irregular_records_survive():
    proto = {"video": "v", "detections": [
        {"frame": 1, "bbox": [1, 2, 3, 4], "scores": [{"class": "a", "class_index": 1, "score": 0.5}]},
        {"frame": 2, "bbox": [1.5, 2, 3, 4.0], "scores": [{"class": "a", "class_index": 1, "score": 1}], "hash": "x"},
        {"bbox": [0.0, 0.0, 1.0, 1.0], "frame": 3.0, "scores": [{"class": "a", "class_index": 1, "score": -1e5}],
         "flag": True, "note": None, "nested": {"k": [1, 2]}},
        {"frame": 4, "bbox": [1, 2, 3, 4], "scores": [{"class": "a", "class_index": 1, "score": float("-inf")}]},
    ], "extra": {"anything": [1, "two", 3.0]}}
    meta, arrays = packed.pack_proto(proto)
    assert meta["kind"] == "det" and meta["detections"]["cols"]["bbox"]["int"] == "map"
    assert _same(packed.unpack_proto(meta, arrays), proto)
    # different class tables / ragged bboxes fall back to JSON columns, still exact
    proto2 = copy.deepcopy(proto)
    proto2["detections"][1]["scores"][0]["class"] = "b"
    proto2["detections"][0]["bbox"] = [1, 2, 3]
    meta, arrays = packed.pack_proto(proto2)
    assert meta["detections"]["cols"]["scores"]["t"] == "json" and meta["detections"]["cols"]["bbox"]["t"] == "json"
    assert _same(packed.unpack_proto(meta, arrays), proto2)
    # an integer float64 cannot hold -> the whole proto is carried as JSON rather than rounded
    proto3 = {"video": "v", "detections": [{"frame": (1 << 60) + 1, "bbox": [0, 0, 1, 1]}]}
    meta, arrays = packed.pack_proto(proto3)
    assert meta["kind"] == "json" and not arrays and _same(packed.unpack_proto(meta, arrays), proto3)
    # empty protos
    for empty in ({"video": "v", "detections": []}, {"video": "v", "method": "m", "tracks": []},
                  {"video": "v", "method": "m", "tracks": [[], []]}, {"video": "v", "method": "m", "tubelets": []},
                  {"video": "v", "method": "m", "tubelets": [{"gt": 0, "boxes": []}]}, [], 3):
        assert _same(packed.unpack_proto(*packed.pack_proto(empty)), empty)


def test_random_records_round_trip():
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies
    num = st.one_of(st.integers(-10 ** 6, 10 ** 6), st.floats(allow_nan=False, width=64), st.just(-1e5))
    value = st.one_of(num, st.lists(num, min_size=4, max_size=4), st.text(max_size=4), st.none(), st.booleans())
    record = st.dictionaries(st.sampled_from(["frame", "bbox", "det_score", "anchor", "hash", "z"]), value, max_size=6)

    @hyp.settings(max_examples=150, deadline=None)
    @hyp.given(st.lists(st.lists(record, max_size=5), max_size=4))
    def run(tracks):
        proto = {"video": "v", "method": "m", "tracks": tracks}
        meta, arrays = packed.pack_proto(proto)
        assert meta["kind"] == "track"
        assert _same(packed.unpack_proto(meta, arrays), proto)
    run()


def test_container_layout(tmp_path):
    path = str(tmp_path / "c.vdetpk")
    arrays = {"a": np.arange(7, dtype=np.float64), "b": np.arange(6, dtype=np.int32).reshape(2, 3),
              "e": np.zeros((0, 4), np.float32), "big": np.arange(5, dtype=">i4"), "m": np.array([True, False, True])}
    packed.save_packed(path, {"hello": [1, 2, {"x": None}]}, arrays)
    raw = open(path, "rb").read()
    assert raw[:8] == b"VDETPK01"
    hlen = int.from_bytes(raw[8:16], "little")
    head = json.loads(raw[16:16 + hlen])
    for name, d in head["arrays"].items():
        assert d["offset"] % 64 == 0 and d["offset"] >= 16 + hlen
    meta, got = packed.load_packed(path)
    assert meta == {"hello": [1, 2, {"x": None}]}
    for k, v in arrays.items():
        assert np.array_equal(got[k], v) and got[k].shape == v.shape and got[k].dtype.itemsize == v.dtype.itemsize
    assert isinstance(got["a"].base, np.memmap) or isinstance(got["a"], np.memmap) or got["a"].base is not None
    assert not got["a"].flags.writeable
    (tmp_path / "bad").write_bytes(b"not a container")
    with pytest.raises(ValueError):
        packed.load_packed(str(tmp_path / "bad"))


def test_proto_load_dump_take_the_packed_path(tmp_path):
    det = _protos()["det"]
    p = str(tmp_path / "det.vdetpk")
    protocol.proto_dump(det, p)
    assert open(p, "rb").read(8) == b"VDETPK01"
    assert _same(protocol.proto_load(p), det)
    # a side-car next to the JSON file wins, like the reference's .gz preference
    j = str(tmp_path / "x.det")
    protocol.proto_dump({"video": "json version", "detections": []}, j)
    assert protocol.proto_load(j)["video"] == "json version"
    protocol.proto_dump(det, j + ".vdetpk")
    assert _same(protocol.proto_load(j), det)
    # ... but only while it is not older than the file it stands for (a regenerated JSON must win)
    import os
    import time
    st = os.stat(j + ".vdetpk")
    os.utime(j, (st.st_atime, st.st_mtime + 5))
    assert protocol.proto_load(j)["video"] == "json version"
    os.utime(j + ".vdetpk", (time.time() + 10, time.time() + 10))
    assert _same(protocol.proto_load(j), det)
    # and the reference formats still work
    protocol.proto_dump(det, str(tmp_path / "y.det.gz"))
    assert _same(protocol.proto_load(str(tmp_path / "y.det")), det)


def test_packed_dets_views():
    g = _protos()
    det = g["det"]
    pd = packed.PackedDets.from_det_proto(det)
    M = len(det["detections"])
    assert pd.video == det["video"] and pd.frames.tolist() == [d["frame"] for d in det["detections"]]
    # the float32 matrix apply_vid_nms builds (vdet/video_det.py:53-56), for every class at once
    order = pd.frame_segments()[1]
    boxes32, scores32, seg, order2, max_len = pd.grouped_f32()
    assert np.array_equal(order, order2) and seg[-1] == M and max_len == np.diff(seg).max()
    for c, cls in enumerate(pd.class_index):
        want = np.asarray([[d["frame"]] + list(d["bbox"]) + [protocol.det_score(d, cls)] for d in det["detections"]],
                          dtype="float32")
        assert np.array_equal(boxes32, want[order, 1:5]) and np.array_equal(scores32[:, c], want[order, 5])
    fr = pd.frames[order]
    assert np.all(np.diff(fr) >= 0)
    for s in range(len(seg) - 1):
        rows = order[seg[s]:seg[s + 1]]
        assert np.all(np.diff(rows) > 0) and len(set(pd.frames[rows].tolist())) == 1     # stable inside a frame
    b3, s3 = pd.uniform_f32()
    assert b3.shape[1:] == (max_len, 4) and s3.shape == (len(seg) - 1, max_len, len(pd.class_index))
    # back to a det proto: same numbers, the fields the packed view carries
    back = pd.to_det_proto()
    for d0, d1 in zip(det["detections"], back["detections"]):
        assert d1["frame"] == d0["frame"] and d1["bbox"] == d0["bbox"] and d1["scores"] == d0["scores"]
    # det_info / frame_to_det forms
    info = pd.to_det_info()
    assert info.shape == (M, 5 + len(pd.class_index)) and info.dtype == np.float64
    assert info[:, 0].tolist() == pd.frames.tolist()
    f2d = pd.to_frame_to_det()
    pd2 = packed.PackedDets.from_frame_to_det("v", f2d)
    assert np.array_equal(pd2.boxes, pd.boxes[order]) and np.array_equal(pd2.scores, pd.scores[order])
    ragged = packed.PackedDets("v", [1, 1, 2], np.zeros((3, 4)), np.zeros((3, 2)))
    with pytest.raises(ValueError):
        ragged.uniform_f32()


def test_packed_dets_file(tmp_path):
    pd = packed.PackedDets.from_det_proto(_protos()["det"])
    p = str(tmp_path / "dets.vdetpk")
    pd.save(p)
    for mm in (True, False):
        q = packed.PackedDets.load(p, mmap=mm)
        assert q.video == pd.video and q.classes == pd.classes and q.class_index == pd.class_index
        assert np.array_equal(q.frames, pd.frames) and np.array_equal(q.boxes, pd.boxes)
        assert np.array_equal(q.scores, pd.scores)
    # a packed det PROTO file opens as PackedDets too (no dict is built)
    p2 = str(tmp_path / "proto.vdetpk")
    packed.proto_dump_packed(_protos()["det"], p2)
    q = packed.PackedDets.load(p2)
    assert np.array_equal(q.scores, pd.scores) and q.classes == pd.classes


def _check_against(det_info, f2d, want):
    assert det_info.dtype == np.float64 and np.array_equal(det_info, want["det_info"])
    assert sorted(f2d) == want["frames_with_file"].tolist()
    for f, (b, z) in f2d.items():
        assert np.array_equal(b, want["f2d_boxes_%d" % f]) and b.dtype == want["f2d_boxes_%d" % f].dtype
        assert np.array_equal(z, want["f2d_zs_%d" % f]) and z.dtype == want["f2d_zs_%d" % f].dtype


def test_mat_loaders_match_the_reference_golden(tmp_path):
    want = helpers.golden_npz("det_mat.npz")
    vid, mats = gen_golden.det_mat_case()
    gen_golden.write_det_mats(vid, mats, str(tmp_path))
    info = packed.load_det_info(vid, str(tmp_path))
    f2d = packed.load_frame_to_det(vid, str(tmp_path))
    _check_against(info, f2d, want)
    # the side-car: read the directory once, then no .mat is touched again
    side = str(tmp_path / "dets.vdetpk")
    pd = packed.pack_det_dir(vid, str(tmp_path), side)
    again = packed.PackedDets.load(side)
    assert np.array_equal(again.to_det_info(), np.asarray(sorted(info.tolist(), key=lambda r: r[0])))
    assert np.array_equal(again.boxes, pd.boxes)
    got = again.to_frame_to_det()
    for f in got:
        assert np.array_equal(got[f][0], f2d[f][0].astype(np.float64)) and np.array_equal(got[f][1], f2d[f][1].astype(np.float64))
    # no file at all -> the reference's empty array
    empty = packed.load_det_info(vid, str(tmp_path / "nothing_here"))
    assert empty.shape == (0,) and empty.dtype == np.float64


@pytest.mark.skipif(not ref_py2.available(), reason="/root/reference not mounted")
def test_mat_loaders_match_the_live_reference(tmp_path):
    R = ref_py2.RefFunctions()
    vid, mats = gen_golden.det_mat_case()
    gen_golden.write_det_mats(vid, mats, str(tmp_path))
    ref_info = np.asarray(R.load_det_info(vid, str(tmp_path)))
    ref_f2d = R.load_frame_to_det(vid, str(tmp_path))
    want = {"det_info": ref_info, "frames_with_file": np.asarray(sorted(ref_f2d))}
    for f, (b, z) in ref_f2d.items():
        want["f2d_boxes_%d" % f], want["f2d_zs_%d" % f] = b, z
    _check_against(packed.load_det_info(vid, str(tmp_path)), packed.load_frame_to_det(vid, str(tmp_path)), want)


def test_greedy_tracker_input_from_packed_dets():
    """``to_det_info`` is the array greedily_track_from_raw_dets consumes (vdet/track.py:186-205): fed to
    the oracle's restatement it reproduces the golden tracks the reference produced from the raw arrays."""
    g = _protos()
    pd = packed.PackedDets.from_det_proto(g["det"])
    opts = helpers.Opts(max_tracks=4, thres=0.6, nms_thres=None)
    tp, _ = oracle_np.greedily_track_from_raw_dets(g["vid"], pd.to_det_info(), helpers.fake_tracker, 3, opts)
    assert tp == g["out"]["greedy_raw"]


def test_grouped_f32_groups_on_the_float32_frame_id():
    """The kernels' view groups frames on the float32 cast of the frame id, like the [M,6] float32 matrix of
    apply_vid_nms (vdet/video_det.py:53-56): ids that collide in float32 are one frame to vid_nms."""
    big = 1 << 24
    det = {"video": "v", "detections": [
        {"frame": f, "bbox": [1, 2, 30, 40], "hash": "h", "scores": [{"class": "a", "class_index": 1, "score": 0.5}]}
        for f in (3, big, big + 1, 3, big + 2)]}
    pd = packed.PackedDets.from_det_proto(det)
    off64, order64, seg64 = pd.frame_segments()
    assert len(seg64) == 4                                           # exact ids: four frames
    b, s, off, order, n = pd.grouped_f32()
    assert off.tolist() == [0, 2, 4, 5] and n == 2                   # float32: 2^24 and 2^24 + 1 are one frame
    assert order.tolist() == [0, 3, 1, 2, 4]
