"""VideoPostProcessor's pipeline bookkeeping on the CPU (tests/cuda_fake.py): slots, tickets, chunk offsets,
staging rules, frame-major views -- the same assertions as the GPU test of the staged path, plus the
per-slot staging mode (a NEW shard staged while the previous one is in flight).  Kernels = the oracle."""
import numpy as np
import pytest

from oracle import c_oracle
from vdetlib_b200 import synth

import cuda_fake


@pytest.fixture
def fake(monkeypatch):
    return cuda_fake.install(monkeypatch)


def _want(b, s, thr=0.3):
    km, _, kc = c_oracle.nms_frames(b, s, thr)
    ls, lb = c_oracle.link_f32(b)
    return km, kc, ls, lb


def _want_lists(b, s, thr=0.3):
    """The reference's result per (frame, class): kept boxes in descending score (utils/nms.pyx:43-66)."""
    T, N, C = s.shape
    return [[np.asarray(c_oracle.nms(np.concatenate([b[t], s[t, :, c:c + 1]], axis=1).astype(np.float32), thr))
             for c in range(C)] for t in range(T)]


def _check(out, want, T, N):
    km, kc, ls, lb = want
    assert np.array_equal(out.keep_mask(), km) and np.array_equal(out["keep_cnt"], kc)
    assert out["keep_idx"].dtype == np.uint16 and out["keep_off"][-1] == kc.sum() == len(out["keep_idx"])
    got = out["succ"][:(T - 1) * N].reshape(T - 1, N) - np.arange(1, T)[:, None] * N
    assert np.array_equal(got, ls)
    assert np.array_equal(out["link_iou"][:(T - 1) * N].reshape(T - 1, N), lb)
    assert np.all(out["succ"][(T - 1) * N:] == -1)                          # last frame: no halo


@pytest.mark.parametrize("n_chunks", [1, 3, 10])
def test_run_host_over_uneven_chunks(fake, n_chunks):
    from vdetlib_b200.vdet.video_det import VideoPostProcessor
    T, N, C = 10, 60, 4
    b, s = synth.boxes_scores(T, N, C, seed=92)
    pp = VideoPostProcessor(T, N, C, 0.3, n_chunks=n_chunks)
    assert [f1 - f0 for f0, f1 in pp.chunks] and sum(f1 - f0 for f0, f1 in pp.chunks) == T
    out = pp.run_host(b, s)
    _check(out, _want(b, s), T, N)
    lists = _want_lists(b, s)
    for t in range(T):
        for c in range(C):
            assert np.array_equal(out.keep_list(t, c), lists[t][c]), (t, c)          # ordered, frame-local
    assert len(fake.nms_calls) == len(pp.chunks)
    assert pp.h2d_bytes == T * N * (4 + C) * 4
    assert pp.d2h_bytes(out) == 2 * int(out["keep_cnt"].sum()) + 4 * (T * C + 1) + T * N * 8 + 4


def test_two_steps_in_flight_shared_staging(fake):
    """The flow of tests/test_gpu_nms.py::test_video_postprocessor_two_steps_in_flight (eager streams)."""
    from vdetlib_b200.vdet.video_det import VideoPostProcessor
    T, N, C = 12, 50, 3
    pp = VideoPostProcessor(T, N, C, 0.3, n_chunks=4, n_stage=1)
    data = [synth.boxes_scores(T, N, C, seed=300 + k) for k in range(2)]
    for b, s in data:
        want = _want(b, s)
        pp.stage(b, s)
        t0 = pp.submit_staged()
        t1 = pp.submit_staged()
        with pytest.raises(RuntimeError):
            pp.submit_staged()                                   # both slots busy
        with pytest.raises(RuntimeError):
            pp.stage(b, s)                                       # in-flight steps read the one staging set
        _check(pp.collect(t0), want, T, N)
        t2 = pp.submit_staged()
        _check(pp.collect(t1), want, T, N)
        _check(pp.collect(t2), want, T, N)
        with pytest.raises(RuntimeError):
            pp.collect(t2)
    out = pp.run_host(*data[0])
    _check(out, _want(*data[0]), T, N)
    assert np.array_equal(pp.d_boxes.numpy().reshape(T, N, 4), data[0][0])      # slot 0 holds the step's inputs


def test_per_slot_staging_streams_new_shards(fake):
    """n_stage = n_slots: shard k+1 is staged and submitted while shard k is still in flight; every ticket
    returns the results of ITS shard."""
    from vdetlib_b200.vdet.video_det import VideoPostProcessor
    T, N, C = 8, 40, 3
    pp = VideoPostProcessor(T, N, C, 0.3, n_chunks=2, n_slots=2, n_stage=2)
    shards = [synth.boxes_scores(T, N, C, seed=500 + k) for k in range(5)]
    wants = [_want(b, s) for b, s in shards]
    pp.stage(*shards[0])
    tickets = [pp.submit_staged()]
    for k in range(1, len(shards)):
        pp.stage(*shards[k])                                     # the other slot's buffers: allowed while k-1 flies
        tickets.append(pp.submit_staged())
        with pytest.raises(RuntimeError):
            pp.stage(*shards[k])                                 # both slots in flight now
        _check(pp.collect(tickets[k - 1]), wants[k - 1], T, N)
    _check(pp.collect(tickets[-1]), wants[-1], T, N)
    # staging sets are really separate buffers, and slot k reads set k
    assert pp.h_boxes_sets[0].data_ptr() != pp.h_boxes_sets[1].data_ptr()
    assert [sl.stage_set for sl in pp.slots] == [0, 1]
    with pytest.raises(ValueError):
        VideoPostProcessor(T, N, C, 0.3, n_slots=2, n_stage=3)
    with pytest.raises(ValueError):
        pp.stage(shards[0][0][:-1], shards[0][1][:-1])


def test_gpu_test_body_of_the_staged_path_on_the_harness(fake):
    """The assertions of tests/test_gpu_nms.py::test_video_postprocessor_two_steps_in_flight (eager streams),
    verbatim, with the launches swapped for the oracle."""
    import test_gpu_nms
    test_gpu_nms.test_video_postprocessor_two_steps_in_flight(False)


def test_ragged_shards_and_status_reset(fake):
    """Ragged frames through stage / submit / collect (VERDICT r01 missing #1): packed rows + counts, fewer frames
    than the capacity, empty frames, chunk edges balanced by rows; then a uniform shard on the same slots."""
    from vdetlib_b200.vdet.video_det import VideoPostProcessor
    T, N, C = 9, 48, 3
    pp = VideoPostProcessor(T, N, C, 0.3, n_chunks=3, want_bits=True)
    rng = np.random.default_rng(7)
    for trial, counts in enumerate([[5, 0, 48, 17, 1, 33], [48] * 9, [0, 0, 7], [1]]):
        counts = np.asarray(counts, np.int32)
        Tr = len(counts)
        b, s = synth.boxes_scores(Tr, N, C, seed=700 + trial)
        rows_b = np.concatenate([b[t, :counts[t]] for t in range(Tr)])
        rows_s = np.concatenate([s[t, :counts[t]] for t in range(Tr)])
        off = np.concatenate([[0], np.cumsum(counts)])
        t = pp.submit_host(rows_b, rows_s, counts=counts)
        out = pp.collect(t)
        assert out["keep_cnt"].shape == (Tr, C) and out["succ"].shape == (int(off[-1]),)
        for f in range(Tr):
            for c in range(C):
                d = np.concatenate([b[f, :counts[f]], s[f, :counts[f], c:c + 1]], axis=1).astype(np.float32)
                want = np.asarray(c_oracle.nms(d, 0.3), dtype=np.int64)
                assert np.array_equal(out.keep_list(f, c), want), (trial, f, c)
                bits = out["keep_bits"][f * C + c]
                got = [i for i in range(counts[f]) if (bits[i >> 5] >> (i & 31)) & 1]
                assert got == sorted(want.tolist())
        # link: frame f -> frame f+1 in packed rows; empty next frame or last frame -> -1
        for f in range(Tr):
            a, e = off[f], off[f + 1]
            if e == a:
                continue
            if f + 1 < Tr and counts[f + 1] > 0:
                iou = c_oracle.pair_iou_f32(b[f, :counts[f]], b[f + 1, :counts[f + 1]])
                assert np.array_equal(out["succ"][a:e], off[f + 1] + np.argmax(iou, axis=1))
                assert np.array_equal(out["link_iou"][a:e], iou.max(axis=1))
            else:
                assert np.all(out["succ"][a:e] == -1)
    with pytest.raises(ValueError):
        pp.stage(rows_b, rows_s, counts=[N + 1])
    with pytest.raises(ValueError):
        pp.stage(rows_b, rows_s, counts=[2, 2])                  # rows do not add up
    # a uniform shard after ragged ones: the segment table is restored
    b, s = synth.boxes_scores(T, N, C, seed=801)
    _check(pp.run_host(b, s), _want(b, s), T, N)


def test_gpu_test_bodies_of_the_streaming_and_ragged_paths_on_the_harness(fake):
    """tests/test_gpu_nms.py's streaming (new shard per step, status reset) and ragged tests, verbatim, eager streams."""
    import test_gpu_nms
    test_gpu_nms.test_video_postprocessor_streams_new_shards_from_pageable_memory(False)
    test_gpu_nms.test_video_postprocessor_ragged_frames()
    test_gpu_nms.test_video_postprocessor_producer_writes_in_place(False)
