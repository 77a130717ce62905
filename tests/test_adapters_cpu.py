"""The reference-named adapters (vdetlib_b200.utils.*, vdetlib_b200.vdet.*) end to end WITHOUT a GPU: the
operators are replaced by the oracle-backed CPU restatements of tests/kernel_double.py, and the assertions
are the very ones of the ``-m gpu`` tests (their functions are called from here), i.e. the golden protos the
reference's own functions produced.  This covers the adapters' host logic -- proto walking, packing, index
bookkeeping, in-place / shallow-copy rules, exception types -- in the ``-m "not gpu"`` tier; the kernels
themselves are only ever checked on the GPU."""
import numpy as np
import pytest

import kernel_double
import test_gpu_nms as G_nms
import test_gpu_packed as G_packed
import test_gpu_temporal_tubelet as G_tub


@pytest.fixture(autouse=True)
def cpu_kernels(monkeypatch):
    kernel_double.install(monkeypatch)


def test_cython_nms_module_golden_and_errors():
    G_nms.test_nms_golden_all()
    G_nms.test_vid_and_track_golden()
    G_nms.test_config1_300_boxes_thr05()
    G_nms.test_known_answers_and_errors()
    G_nms.test_strided_and_wide_input()


def test_iou_adapter_golden():
    import helpers
    from vdetlib_b200.utils import common
    g = helpers.golden_npz("arrays.npz")
    got = common.iou(g["iou_a_int"], g["iou_b_int"])
    assert got.dtype == np.float64 and np.array_equal(got, g["iou_int"])
    assert np.array_equal(common.iou(g["iou_a_f"].tolist(), g["iou_b_f"]), g["iou_f"])      # array-likes
    with pytest.raises(IndexError):
        common.iou([1, 2, 3, 4], g["iou_b_int"])


def test_spatial_max_pooling_and_anchor_propagate_golden():
    G_tub.test_spatial_max_pooling_protos_golden()


def test_score_proto_functions():
    G_tub.test_score_proto_functions()


def test_apply_vid_nms_and_image_nms_golden():
    G_tub.test_vid_nms_and_image_nms_protos_golden()


def test_greedy_tracking_golden_and_keep_state():
    G_tub.test_greedy_tracking_golden()
    G_tub.test_greedy_tracking_keep_state_vs_oracle()


def test_threshold_topk_frames():
    G_tub.test_threshold_topk_vs_oracle()


def test_score_proto_interpolation_golden():
    G_tub.test_score_proto_interpolation_golden()


def test_overlap_merge_top_golden():
    G_tub.test_overlap_merge_top_golden()


def test_packed_vid_nms_golden():
    G_packed.test_packed_vid_nms_golden()
    G_packed.test_packed_vid_nms_synthetic(5, 77, 4, True)


def test_round2_call_sites_golden():
    """rcnn_sampling_dets_scoring, score_conv_cls (marshalling + TemporalConvNet), threshold/top-k: adapters and the
    NumPy restatements against what the reference's own functions produced (tests/golden/r02.json)."""
    G_tub.test_rcnn_sampling_dets_scoring_golden()
    G_tub.test_score_conv_cls_channel_marshalling_golden()
    G_tub.test_score_conv_cls_temporal_conv_net_batched()
    G_tub.test_threshold_topk_pinned_to_fast_rcnn_det_vid()


def test_bounded_completion_on_the_double():
    """The assertions of the GPU test of vdet_score_completion_bounded with the oracle-backed stand-in: checks the
    test's own bounds bookkeeping (and the stand-in) against whole-row completion."""
    G_tub.test_bounded_completion_equals_the_unsharded_rows(np.float64)
