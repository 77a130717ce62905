// nms_plan.h -- which CTA shape a vdet_nms_frames_f32 launch takes (host logic only: plain C++, no CUDA, so that
// tests/test_nms_plan_cpu.py can compile it with g++ and check it against the measured table).
//
// Two shapes exist for staged frames of at most 320 boxes (nms_frames.cuh): the default (8 warps, 4 CTAs per SM) and
// the wide one (10 warps, 3 CTAs per SM, every class staged at once).  Per SM they hold about the same number of
// warps; what differs is how a launch QUANTISES: the grid is persistent, a launch of n frames is n / slots rounds,
// and the last, partially filled round costs between a quarter of a round (a handful of frames, each alone on its
// SM) and a whole one (from about two thirds full).  Measured on B200, 300 boxes x 30 classes
// (profiles/r02_nms_shapes_T.jsonl, ms default / wide): 300 frames 0.121 / 0.117, 600: 0.186 / 0.183,
// 800: 0.216 / 0.232, 1000: 0.294 / 0.278, 1184: 0.298 / 0.341, 2000: 0.490 / 0.504, 3000: 0.753 / 0.791;
// 1000 frames x 20 classes 0.237 / 0.214, x 12 classes 0.173 / 0.183, x 8 classes 0.133 / 0.141.
#pragma once

namespace vdet {

constexpr int NMS_PLAN_WARPS_DEFAULT = 8, NMS_PLAN_CTAS_DEFAULT = 4;
constexpr int NMS_PLAN_WARPS_WIDE = 10, NMS_PLAN_CTAS_WIDE = 3;

// Cost of the last round of a persistent launch as a fraction of a full round, by how full it is.
inline double nms_tail_cost(const double fill) {
    if (fill <= 0.0) return 0.0;
    if (fill <= 0.35) return 0.25 + 0.7 * fill;
    if (fill >= 0.68) return 1.0;
    return 0.495 + (fill - 0.35) * (0.505 / 0.33);
}

// Relative cost of n_frames on `slots` resident CTAs when one round costs `round_cost`.
inline double nms_launch_cost(const int n_frames, const int slots, const double round_cost) {
    if (n_frames <= 0 || slots <= 0) return 0.0;
    const int full = n_frames / slots, rem = n_frames % slots;
    return round_cost * ((double)full + nms_tail_cost((double)rem / (double)slots));
}

// true: take the wide shape (the caller has already checked that it fits).  A round costs one class per warp and
// turn -- ceil(C / warps) turns -- and 4 % more in the wide shape (30 resident warps per SM hide a little less
// latency than 32: the steady state of long launches favours the default by 3-5 %).  Launches that fill less than
// half of the default grid are cut into class ranges by the kernel's work planner and stay with the default shape.
inline bool nms_prefer_wide(const int n_frames, const int n_classes, const int sm_count) {
    const int slots_d = sm_count * NMS_PLAN_CTAS_DEFAULT, slots_w = sm_count * NMS_PLAN_CTAS_WIDE;
    if (n_frames <= 0 || n_classes < NMS_PLAN_WARPS_WIDE || 2 * n_frames <= slots_d) return false;
    const double round_d = (double)((n_classes + NMS_PLAN_WARPS_DEFAULT - 1) / NMS_PLAN_WARPS_DEFAULT);
    const double round_w = 1.04 * (double)((n_classes + NMS_PLAN_WARPS_WIDE - 1) / NMS_PLAN_WARPS_WIDE);
    return nms_launch_cost(n_frames, slots_w, round_w) < nms_launch_cost(n_frames, slots_d, round_d);
}

}  // namespace vdet
