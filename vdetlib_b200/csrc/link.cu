// link.cu -- frame-to-frame tubelet link (K7 of SURVEY 2.1; build-defined).
//
// The reference never links boxes across frames by IoU (it delegates association to the
// external FCNT/TLD trackers, vdet/track.py:18-106); BASELINE.json's north_star asks for a
// batched IoU link.  Semantics are restated from the reference's own pieces: the pair IoU of
// utils/nms.pyx:57-64 (float32) and the FIRST-arg-max rule of np.argmax as used in
// vdet/tubelet_cls.py:375-376.
//
// Mapping: one thread owns one box i of frame t (its box and area live in registers) and
// walks the boxes of frame t+1, which the CTA stages once in shared memory as float4 + area
// (coalesced 16-byte global reads, warp-uniform LDS.128 broadcasts in the loop).  No
// reduction is needed: each thread keeps its own running (best, arg).
#include "common.cuh"

namespace vdet {

constexpr int LINK_THREADS = 64;
constexpr int LINK_STAGE = 1024;     // boxes of frame t+1 staged per pass

__global__ void __launch_bounds__(LINK_THREADS) link_frames_kernel(const float4* __restrict__ boxes,
                                                                   const int32_t* __restrict__ seg_offsets,
                                                                   int n_segs, const float4* __restrict__ halo,
                                                                   int n_halo, int halo_row_base,
                                                                   int32_t* __restrict__ succ,
                                                                   float* __restrict__ best_iou) {
    __shared__ float4 s_box[LINK_STAGE];
    __shared__ float s_area[LINK_STAGE];
    const int seg = blockIdx.x;
    const int off = seg_offsets[seg];
    const int n = seg_offsets[seg + 1] - off;
    const int i = blockIdx.y * LINK_THREADS + threadIdx.x;
    if ((int)blockIdx.y * LINK_THREADS >= n) return;          // whole CTA beyond this frame
    const bool last = (seg == n_segs - 1);
    const float4* nxt = last ? halo : boxes + seg_offsets[seg + 1];
    const int m = last ? n_halo : (seg_offsets[seg + 2] - seg_offsets[seg + 1]);
    const int out_base = last ? halo_row_base : seg_offsets[seg + 1];

    const bool active = i < n;
    const float4 bi = active ? __ldg(boxes + off + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float ai = area_f32(bi);
    float best = -1.0f;        // every valid IoU is >= 0, so the first valid j always wins
    int arg = -1;
    const bool own_sane = box_sane(bi);          // inactive lanes hold the (sane) zero box
    for (int j0 = 0; j0 < m; j0 += LINK_STAGE) {
        const int cnt = (m - j0) < LINK_STAGE ? (m - j0) : LINK_STAGE;
        __syncthreads();
        bool ok = own_sane;
        for (int e = threadIdx.x; e < cnt; e += LINK_THREADS) {
            const float4 b = __ldg(nxt + j0 + e);
            s_box[e] = b;
            s_area[e] = area_f32(b);
            ok = ok && box_sane(b);
        }
        // barrier + vote: when every box in play is sane (common.cuh) the branch-free division is
        // exact and unions cannot be zero; otherwise the generic IEEE path runs.  Same bits.
        if (__syncthreads_and(ok)) {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                float inter, uni;
                inter_union_f32(bi, ai, s_box[j], s_area[j], inter, uni);
                const float v = div_sane(inter, uni);
                if (v > best) { best = v; arg = j0 + j; }      // strict '>' keeps the FIRST maximum
            }
        } else {
#pragma unroll 2
            for (int j = 0; j < cnt; ++j) {
                float inter, uni;
                inter_union_f32(bi, ai, s_box[j], s_area[j], inter, uni);
                const float v = iou_quotient(inter, uni);
                // NaN (0/0) and union==0 never win
                if (uni != 0.0f && (arg < 0 || v > best)) { best = v; arg = j0 + j; }
            }
        }
    }
    if (active) {
        succ[off + i] = arg < 0 ? -1 : out_base + arg;
        best_iou[off + i] = arg < 0 ? 0.0f : best;
    }
}

}  // namespace vdet

using namespace vdet;

extern "C" int vdet_link_frames_f32(const float* boxes, const int32_t* seg_offsets, int n_segs,
                                    int max_seg_len, const float* halo_boxes, int n_halo, int halo_row_base,
                                    int32_t* succ, float* best_iou, int64_t n_rows, void* stream) {
    VDET_REQUIRE(n_segs >= 0 && max_seg_len >= 0 && n_halo >= 0 && n_rows >= 0, "link_frames: negative size");
    if (n_segs == 0 || n_rows == 0 || max_seg_len == 0) return VDET_OK;
    VDET_REQUIRE(((uintptr_t)boxes & 15) == 0 && ((uintptr_t)halo_boxes & 15) == 0,
                 "link_frames: boxes must be 16-byte aligned");
    VDET_REQUIRE(max_seg_len <= 65535 * LINK_THREADS, "link_frames: frame too long");
    dim3 grid((unsigned)n_segs, (unsigned)((max_seg_len + LINK_THREADS - 1) / LINK_THREADS));
    link_frames_kernel<<<grid, LINK_THREADS, 0, (cudaStream_t)stream>>>(
        (const float4*)boxes, seg_offsets, n_segs, (const float4*)halo_boxes, n_halo, halo_row_base, succ, best_iou);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}
