"""PackedTubelets (vdetlib_b200.utils.packed): the host side of the packed temporal path, on CPU.

The kernels are replaced by the NumPy oracle (fake ``ops`` on CPU tensors), so what is checked here is
the packing / scatter / type-map / exception logic around them, against the golden rows the reference's
own do_score_completion and score_proto_temporal_maxpool produced (tests/golden/arrays.npz).  The same
stages with the real kernels are the dict adapters' tests in tests/test_gpu_temporal_tubelet.py."""
import copy
import json

import numpy as np
import pytest
import torch

from oracle import oracle_np
from vdetlib_b200 import ops
from vdetlib_b200.utils import packed

import helpers


@pytest.fixture
def fake_ops(monkeypatch):
    def completion(scores, lengths=None, miss_thr=-10.0, status=None):
        a, lens = scores.numpy(), lengths.numpy()
        st = 0
        for i in range(a.shape[0]):
            try:
                a[i, :lens[i]] = oracle_np.completion_row(a[i, :lens[i]], miss_thr)
            except IndexError:
                st |= 2
        return torch.tensor([st], dtype=torch.int32)

    def maxpool(scores, window, lengths=None, pad=-1e5, out=None):
        a, lens = scores.numpy(), lengths.numpy()
        o = np.full_like(a, pad)
        for i in range(a.shape[0]):
            o[i, :lens[i]] = oracle_np.temporal_maxpool_row(a[i, :lens[i]], window, pad)
        return torch.from_numpy(o)

    def conv(x, taps, pad_mode="zero", lengths=None, out=None):
        a, lens = x.numpy(), lengths.numpy()
        o = np.zeros_like(a)
        for i in range(a.shape[0]):
            o[i, :lens[i]] = oracle_np.temporal_conv1d(a[i:i + 1, :lens[i]], taps.numpy(), pad_mode)[0]
        return torch.from_numpy(o)

    monkeypatch.setattr(ops, "score_completion_", completion)
    monkeypatch.setattr(ops, "temporal_maxpool", maxpool)
    monkeypatch.setattr(ops, "temporal_conv1d", conv)
    return torch.device("cpu")


def _proto_from_rows(rows):
    """The score proto oracle/gen_golden.py fed to the reference (section 4)."""
    return {'video': 'v', 'method': 'm',
            'tubelets': [{'gt': 0, 'boxes': [{'det_score': float(v)} for v in r]} for r in rows]}


def _rows_of(proto):
    return [[b['det_score'] for b in t['boxes']] for t in proto['tubelets']]


def test_views_and_scatter():
    sp = helpers.golden_protos()["out"]["smp_1"]
    pt = packed.PackedTubelets.from_score_proto(sp)
    assert pt.n_tubelets == len(sp['tubelets'])
    assert pt.lengths.tolist() == [len(t['boxes']) for t in sp['tubelets']]
    assert pt.head("class_index") == [t['class_index'] for t in sp['tubelets']]
    rows, lens, index = pt.padded("det_score", fill=-7.0)
    for k, i in enumerate(index):
        want = [b['det_score'] for b in sp['tubelets'][i]['boxes']]
        assert rows[k, :lens[k]].tolist() == want and np.all(rows[k, lens[k]:] == -7.0)
    assert pt.column("bbox").shape == (int(pt.offsets[-1]), 4)
    with pytest.raises(KeyError):
        pt.column("hash")                                        # strings are not a numeric column
    # writing the same values back changes nothing (types included); a restricted write touches only `where`
    pt.set_padded("det_score", rows, lens)
    assert json.dumps(pt.to_score_proto()) == json.dumps(sp)
    where = np.zeros(rows.shape, bool)
    where[0, 0] = True
    pt.set_padded("det_score", rows + 1.0, lens, where=where)
    got = pt.to_score_proto()
    first = sp['tubelets'][index[0]]['boxes'][0]['det_score']
    assert got['tubelets'][index[0]]['boxes'][0]['det_score'] == first + 1.0
    got['tubelets'][index[0]]['boxes'][0]['det_score'] = first
    assert got == sp
    # frames are ints in the proto: a float write flips exactly the written entries of the type map
    fr, lens, _ = pt.padded("frame")
    where = np.zeros(fr.shape, bool)
    where[0, 1] = True
    pt.set_padded("frame", fr + 0.5, lens, where=where)
    boxes0 = pt.to_score_proto()['tubelets'][index[0]]['boxes']
    assert type(boxes0[0]['frame']) is int and type(boxes0[1]['frame']) is float


def test_completion_and_maxpool_match_the_reference_golden(fake_ops):
    g = helpers.golden_npz("arrays.npz")
    rows_in = g["rows_in"]
    pt = packed.PackedTubelets.from_score_proto(_proto_from_rows(rows_in))
    pt.complete_scores_(device=fake_ops)
    done = np.asarray(_rows_of(pt.to_score_proto()))
    assert np.array_equal(done, g["rows_completed"])
    # only the missing entries were written: valid scores keep their exact values
    valid = rows_in > -10
    assert np.array_equal(done[valid], rows_in[valid])
    for w in (3, 5, 9, 201):
        q = packed.PackedTubelets.from_score_proto(pt.to_score_proto())
        q.temporal_maxpool_(w, device=fake_ops)
        out = q.to_score_proto()
        assert np.array_equal(np.asarray(_rows_of(out)), g["rows_maxpool_%d" % w])
        assert out['method'] == 'm_temporal_maxpool_%d' % w
    same = packed.PackedTubelets.from_score_proto(pt.to_score_proto()).temporal_maxpool_(1, device=fake_ops)
    assert same.to_score_proto() == pt.to_score_proto()


def test_ragged_tubelets_and_errors(fake_ops):
    rng = np.random.default_rng(3)
    proto = {'video': 'v', 'method': 'm', 'tubelets': []}
    for k, n in enumerate([5, 0, 17, 1, 9]):
        sc = rng.uniform(-1, 1, n)
        sc[rng.random(n) < 0.4] = -1e5
        if n:
            sc[rng.integers(0, n)] = 0.25                           # at least one valid score
        proto['tubelets'].append({'gt': 0, 'class_index': 3, 'boxes': [
            {'frame': f + 1, 'bbox': [f, f, f + 10, f + 10], 'det_score': float(v), 'anchor': f - n // 2,
             'track_score': 0.5, 'hash': 'h%d' % f} for f, v in enumerate(sc)]})
    want = copy.deepcopy(proto)
    for t in want['tubelets']:
        if t['boxes']:
            done = oracle_np.completion_row([b['det_score'] for b in t['boxes']])
            for b, v in zip(t['boxes'], done):
                if not (b['det_score'] > -10):
                    b['det_score'] = float(v)
    pt = packed.PackedTubelets.from_score_proto(proto).complete_scores_(device=fake_ops)
    assert pt.to_score_proto() == want and json.dumps(pt.to_score_proto()) == json.dumps(want)
    # temporal convolution writes a new last field into every box
    from vdetlib_b200.vdet.tubelet_cls import TemporalConvNet
    net = TemporalConvNet({'det_scores': [0.25, 0.5, 0.25], 'anchors': [0.0, 1.0, 0.0], 'abs_anchors': [1.0]}, bias=0.125)
    pt.conv_scores_(net, device=fake_ops)
    got = pt.to_score_proto()
    for t_in, t_out in zip(want['tubelets'], got['tubelets']):
        n = len(t_in['boxes'])
        if not n:
            assert t_out['boxes'] == []
            continue
        ds = np.asarray([[b['det_score'] for b in t_in['boxes']]])
        an = np.asarray([[b['anchor'] * 1. / n for b in t_in['boxes']]])
        y = (oracle_np.temporal_conv1d(ds, np.asarray([[0.25, 0.5, 0.25]]))
             + oracle_np.temporal_conv1d(an, np.asarray([[0.0, 1.0, 0.0]]))
             + oracle_np.temporal_conv1d(np.abs(an), np.asarray([[1.0]]))) + 0.125
        for b_in, b_out, v in zip(t_in['boxes'], t_out['boxes'], y[0]):
            assert list(b_out.keys()) == list(b_in.keys()) + ['conv_score'] and b_out['conv_score'] == float(v)
    # the reference's exceptions
    bad = _proto_from_rows(np.asarray([[0.1, -1e5, 0.3], [-1e5, -1e5, -1e5]]))
    with pytest.raises(IndexError):
        packed.PackedTubelets.from_score_proto(bad).complete_scores_(device=fake_ops)
    pt2 = packed.PackedTubelets.from_score_proto(want)
    with pytest.raises(ValueError):
        pt2.temporal_maxpool_(4, device=fake_ops)
    gt = copy.deepcopy(want)
    gt['tubelets'][2]['gt'] = 1
    pt3 = packed.PackedTubelets.from_score_proto(gt)
    with pytest.raises(ValueError):
        pt3.temporal_maxpool_(3, device=fake_ops)
    assert pt3.to_score_proto() == gt                                # nothing was written


def test_file_round_trip(tmp_path, fake_ops):
    sp = helpers.golden_protos()["out"]["smp_3"]
    p = str(tmp_path / "s.vdetpk")
    packed.proto_dump_packed(sp, p)
    pt = packed.PackedTubelets.load(p)                               # memory-mapped columns
    pt.temporal_maxpool_(3, device=fake_ops)                         # writes go to private copies
    pt.save(str(tmp_path / "t.vdetpk"))
    again = packed.proto_load_packed(str(tmp_path / "t.vdetpk"))
    assert again == pt.to_score_proto() and again['method'].endswith('_temporal_maxpool_3')
    assert packed.proto_load_packed(p) == sp                         # the source file is untouched
    with pytest.raises(ValueError):
        packed.PackedTubelets(*packed.pack_proto(helpers.golden_protos()["det"]))


def test_packed_spatial_max_pooling_matches_the_reference_golden(monkeypatch):
    """The whole tubelet-scoring chain on packed data -- pool_dets_ + complete_scores_ -- against the protos the
    reference's dets_spatial_max_pooling / raw_dets_spatial_max_pooling produced (no dict walk over the dets)."""
    import kernel_double
    dev = kernel_double.install(monkeypatch)
    p = helpers.golden_protos()
    vid, det, trk, out = p["vid"], p["det"], p["track"], p["out"]
    dets = packed.PackedDets.from_det_proto(det)
    assert dets.boxes_are_int
    for cls in (1, 3):
        for suffix, thr in (("", 0.7), ("_05", 0.5)):
            pt = packed.PackedTubelets.for_pooling(vid, copy.deepcopy(trk), cls, thr)
            pt.pool_dets_(dets, cls, vid, thr, device=dev).complete_scores_(device=dev)
            got = pt.to_score_proto()
            want = out["smp_%d%s" % (cls, suffix)]
            assert got == want and json.dumps(got) == json.dumps(want)
    # raw detections (frame -> (boxes, zs) arrays): float boxes come back as floats
    from vdetlib_b200 import synth
    boxes, scores = synth.boxes_scores(8, 40, 5, seed=4000, integer=True, frame_offset=1e-4)
    f2d = {t + 1: (boxes[t].astype(np.float64), scores[t].astype(np.float64)) for t in range(8)}
    f2d[9] = (np.zeros((0, 4)), np.zeros((0, 5)))                    # a frame without boxes is skipped (:513)
    raw = packed.PackedDets.from_frame_to_det(vid['video'], f2d)
    pt = packed.PackedTubelets.for_pooling(vid, copy.deepcopy(trk), 2)
    pt.pool_dets_(raw, 2, vid, device=dev).complete_scores_(device=dev)
    got = pt.to_score_proto()
    assert got == out["raw_smp_2"] and json.dumps(got) == json.dumps(out["raw_smp_2"])
    # the same through the side-car files
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        dets.save(os.path.join(d, "dets.vdetpk"))
        again = packed.PackedDets.load(os.path.join(d, "dets.vdetpk"))
        pt = packed.PackedTubelets.for_pooling(vid, copy.deepcopy(trk), 1)
        pt.pool_dets_(again, 1, vid, device=dev).complete_scores_(device=dev)
        assert pt.to_score_proto() == out["smp_1"] and json.dumps(pt.to_score_proto()) == json.dumps(out["smp_1"])
