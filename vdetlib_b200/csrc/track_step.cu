// track_step.cu -- the suppression step of the greedy tubelet proposal.
//
// Replaces the Python loop of vdet/track.py:172-183 (= :238-249): for every box of a freshly
// tracked tubelet, the surviving detections of that frame go through
// track_det_nms (utils/nms.pyx:128-189):
//   round 1 (:163-183)  drop dets with IoU(det, track box) >= thresh,
//   round 2 (:186-187)  greedy NMS among the rest, in descending score,
// and `keep[]` is cleared for everything not returned.  The reference does this with one tiny
// Cython call per tracked box; here one warp handles one tracked box and the per-video state
// (det_info sorted by score, keep[]) never leaves the GPU.
//
// Round 2 walks the frame's score-ordered list once; a kept candidate's box is broadcast and
// every lane tests its own later candidates on the fly -- no bit matrix is materialised because
// only K_kept x n/32 pair tests are ever needed.  Frames of up to 4096 detections.
#include "common.cuh"

namespace vdet {

constexpr int TS_MAX_FRAME = 4096;
constexpr int TS_WARPS = 4;
constexpr int TS_MAX_CHUNKS = TS_MAX_FRAME / 32;

__device__ __forceinline__ float4 det_box(const float* __restrict__ det_info, int row) {
    const float* d = det_info + (int64_t)row * 6;
    return make_float4(__ldg(d + 1), __ldg(d + 2), __ldg(d + 3), __ldg(d + 4));
}

// Alive sets live in shared memory, one word per 32-candidate chunk of the frame's score-ordered
// list (bit l of word c = candidate c*32+l), so frames of up to 4096 detections need no register
// arrays.  A kept candidate's box is broadcast and every lane tests its own later candidates on the
// fly; the dead lanes of a chunk are collected with one ballot and cleared by lane 0.
__global__ void __launch_bounds__(TS_WARPS * 32) track_nms_step_kernel(const float* __restrict__ det_info,
                                                                       const int32_t* __restrict__ seg_offsets,
                                                                       const int32_t* __restrict__ row_ids,
                                                                       const float4* __restrict__ track_boxes,
                                                                       const int32_t* __restrict__ track_seg, int q0,
                                                                       int q1, float T, uint8_t* __restrict__ keep,
                                                                       uint32_t* status) {
    __shared__ uint32_t s_alive[TS_WARPS][TS_MAX_CHUNKS];
    __shared__ uint32_t s_entry[TS_WARPS][TS_MAX_CHUNKS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = q0 + blockIdx.x * TS_WARPS + warp;
    if (qi >= q1) return;
    const int seg = track_seg[qi];
    if (seg < 0) return;
    const int off = seg_offsets[seg];
    const int n = seg_offsets[seg + 1] - off;
    if (n > TS_MAX_FRAME) { if (lane == 0) atomicOr(status, 0x80000000u); return; }
    const float4 tb = track_boxes[qi];
    const float ta = area_f32(tb);
    const int chunks = (n + 31) >> 5;
    uint32_t* alive = s_alive[warp];
    uint32_t* entry = s_entry[warp];

    // round 1 (nms.pyx:163-183) on the detections that are still kept
    bool zd = false;
    for (int c = 0; c < chunks; ++c) {
        const int e = c * 32 + lane;
        bool was = false, ok = false;
        if (e < n) {
            const int row = row_ids[off + e];
            if (keep[row]) {
                was = true;
                const float4 b = det_box(det_info, row);
                float inter, uni;
                inter_union_f32(b, area_f32(b), tb, ta, inter, uni);
                if (uni == 0.0f) zd = true;
                else ok = !iou_ge(inter, uni, T);
            }
        }
        const unsigned bw = __ballot_sync(FULL, was), bo = __ballot_sync(FULL, ok);
        if (lane == 0) { entry[c] = bw; alive[c] = bo; }
    }
    __syncwarp();
    if (__any_sync(FULL, zd)) { if (lane == 0) atomicOr(status, VDET_STATUS_ZERO_DIVISION); return; }

    // round 2 (nms.pyx:186-187): greedy NMS among the alive candidates, in list order (= descending score)
    for (int e = 0; e < n; ++e) {
        if (!((alive[e >> 5] >> (e & 31)) & 1u)) continue;            // warp-uniform (broadcast read)
        const float4 bi = det_box(det_info, row_ids[off + e]);
        const float ai = area_f32(bi);
        for (int c = e >> 5; c < chunks; ++c) {
            const int e2 = c * 32 + lane;
            bool dead = false;
            if (e2 > e && e2 < n && ((alive[c] >> lane) & 1u)) {
                const float4 bj = det_box(det_info, row_ids[off + e2]);
                float inter, uni;
                inter_union_f32(bi, ai, bj, area_f32(bj), inter, uni);
                if (uni == 0.0f) zd = true;
                else dead = iou_ge(inter, uni, T);
            }
            const unsigned bd = __ballot_sync(FULL, dead);
            if (bd && lane == 0) alive[c] &= ~bd;
            __syncwarp();
        }
    }
    if (__any_sync(FULL, zd) && lane == 0) atomicOr(status, VDET_STATUS_ZERO_DIVISION);
    // everything that entered alive but was not returned is cleared (track.py:181-183)
    for (int c = 0; c < chunks; ++c) {
        const uint32_t drop = entry[c] & ~alive[c];
        if ((drop >> lane) & 1u) keep[row_ids[off + c * 32 + lane]] = 0;
    }
}

}  // namespace vdet

using namespace vdet;

extern "C" int vdet_track_nms_step_f32(const float* det_info, int64_t m,
                                       const int32_t* seg_offsets, const int32_t* row_ids, int n_segs,
                                       const float* track_boxes, const int32_t* track_seg, int q,
                                       double thresh, uint8_t* keep, uint32_t* status, void* stream) {
    VDET_REQUIRE(m >= 0 && n_segs >= 0 && q >= 0, "track_nms_step: negative size");
    VDET_REQUIRE(((uintptr_t)track_boxes & 15) == 0, "track_nms_step: track_boxes must be 16-byte aligned");
    if (q == 0 || m == 0 || n_segs == 0) return VDET_OK;
    // Boxes are applied in array order; the caller guarantees that boxes inside one call lie on
    // distinct frames (one tracklet), so one launch per call keeps the reference's sequencing.
    const float T = thresh_ceil_f32(thresh);
    track_nms_step_kernel<<<(unsigned)((q + TS_WARPS - 1) / TS_WARPS), TS_WARPS * 32, 0, (cudaStream_t)stream>>>(
        det_info, seg_offsets, row_ids, (const float4*)track_boxes, track_seg, 0, q, T, keep, status);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}
