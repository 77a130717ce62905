// link.cu -- frame-to-frame tubelet link (K7 of SURVEY 2.1; build-defined).
//
// The reference never links boxes across frames by IoU (it delegates association to the
// external FCNT/TLD trackers, vdet/track.py:18-106); BASELINE.json's north_star asks for a
// batched IoU link.  Semantics are restated from the reference's own pieces: the pair IoU of
// utils/nms.pyx:57-64 (float32) and the FIRST-arg-max rule of np.argmax as used in
// vdet/tubelet_cls.py:375-376.
//
// Mapping: one thread owns ROWS boxes of frame t (boxes and areas in registers) and walks the
// boxes of frame t+1, which the CTA stages once in shared memory as float4 + area (coalesced
// 16-byte global reads, warp-uniform LDS.128 broadcasts in the loop, each broadcast shared by the
// thread's ROWS pair evaluations).  No reduction is needed: each thread keeps its own running
// (best, arg) per row.  Rows per CTA (THREADS x ROWS) are chosen per launch so that a frame pads
// to as few row slots as possible and, among equals, is staged by as few CTAs as possible: a
// 2000-box frame is staged by 4 CTAs (512 rows each) instead of 32.
#include <stdlib.h>

#include "common.cuh"

namespace vdet {

constexpr int LINK_STAGE = 1024;     // boxes of frame t+1 staged per pass

template <int THREADS, int ROWS>
__global__ void __launch_bounds__(THREADS) link_frames_kernel(const float4* __restrict__ boxes,
                                                              const int32_t* __restrict__ seg_offsets,
                                                              int n_segs, const float4* __restrict__ halo,
                                                              int n_halo, const int32_t* __restrict__ n_halo_dev,
                                                              int halo_row_base,
                                                              int32_t* __restrict__ succ,
                                                              float* __restrict__ best_iou) {
    __shared__ float4 s_box[LINK_STAGE];
    __shared__ __align__(8) float s_area[LINK_STAGE];
    const int seg = blockIdx.x;
    const int off = seg_offsets[seg];
    const int n = seg_offsets[seg + 1] - off;
    const int row0 = blockIdx.y * (THREADS * ROWS);
    if (row0 >= n) return;                                     // whole CTA beyond this frame
    const bool last = (seg == n_segs - 1);
    const float4* nxt = last ? halo : boxes + seg_offsets[seg + 1];
    int m;
    if (last) {
        // the neighbour's box count may only be known on the device (ragged frames: it arrives with the
        // boundary all-gather); n_halo is then the capacity of the halo buffer
        m = n_halo;
        if (n_halo_dev != nullptr) {
            const int md = *n_halo_dev;
            m = md < 0 ? 0 : (md < n_halo ? md : n_halo);
        }
    } else {
        m = seg_offsets[seg + 2] - seg_offsets[seg + 1];
    }
    const int out_base = last ? halo_row_base : seg_offsets[seg + 1];

    float4 bi[ROWS];
    float ai[ROWS], best[ROWS];
    int arg[ROWS];
    bool own_sane = true;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int i = row0 + r * THREADS + threadIdx.x;
        bi[r] = i < n ? __ldg(boxes + off + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        ai[r] = area_f32(bi[r]);
        best[r] = -1.0f;        // every valid IoU is >= 0, so the first valid j always wins
        arg[r] = -1;
        own_sane = own_sane && box_sane(bi[r]);              // inactive rows hold the (sane) zero box
    }
    for (int j0 = 0; j0 < m; j0 += LINK_STAGE) {
        const int cnt = (m - j0) < LINK_STAGE ? (m - j0) : LINK_STAGE;
        __syncthreads();
        bool ok = own_sane;
        for (int e = threadIdx.x; e < cnt; e += THREADS) {
            const float4 b = __ldg(nxt + j0 + e);
            s_box[e] = b;
            s_area[e] = area_f32(b);
            ok = ok && box_sane(b);
        }
        // barrier + vote: when every box in play is sane (common.cuh) the branch-free division is
        // exact and unions cannot be zero; otherwise the generic IEEE path runs.  Same bits.
        if (__syncthreads_and(ok)) {
            // sane boxes: two pair IoUs per packed FADD2 / FMUL2 / FFMA2 sequence (common.cuh) -- two of the
            // thread's rows against one staged box, or (one row per thread) one row against two staged boxes
            if (ROWS >= 2) {
#pragma unroll 4
                for (int j = 0; j < cnt; ++j) {
                    const float4 bj = s_box[j];
                    const float aj = s_area[j];
                    const f32x2 aj2 = pk2(aj, aj);
#pragma unroll
                    for (int r = 0; r + 1 < ROWS; r += 2) {
                        f32x2 inter, uni, nuni;
                        inter_union_f32x2(bj, aj2, bi[r], bi[r + 1], pk2(ai[r], ai[r + 1]), inter, uni, nuni);
                        float v0, v1;
                        upk2(div_sane2(inter, nuni), v0, v1);
                        if (v0 > best[r]) { best[r] = v0; arg[r] = j0 + j; }      // strict '>' keeps the FIRST maximum
                        if (v1 > best[r + 1]) { best[r + 1] = v1; arg[r + 1] = j0 + j; }
                    }
                }
            } else {
                const f32x2 ai2 = pk2(ai[0], ai[0]);
                int j = 0;
#pragma unroll 4
                for (; j + 1 < cnt; j += 2) {
                    const float2 aj = *reinterpret_cast<const float2*>(s_area + j);
                    f32x2 inter, uni, nuni;
                    inter_union_f32x2(bi[0], ai2, s_box[j], s_box[j + 1], pk2(aj.x, aj.y), inter, uni, nuni);
                    float v0, v1;
                    upk2(div_sane2(inter, nuni), v0, v1);
                    if (v0 > best[0]) { best[0] = v0; arg[0] = j0 + j; }
                    if (v1 > best[0]) { best[0] = v1; arg[0] = j0 + j + 1; }
                }
                if (j < cnt) {
                    float inter, uni;
                    inter_union_f32(bi[0], ai[0], s_box[j], s_area[j], inter, uni);
                    const float v = div_sane(inter, uni);
                    if (v > best[0]) { best[0] = v; arg[0] = j0 + j; }
                }
            }
        } else {
#pragma unroll 2
            for (int j = 0; j < cnt; ++j) {
                const float4 bj = s_box[j];
                const float aj = s_area[j];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    float inter, uni;
                    inter_union_f32(bi[r], ai[r], bj, aj, inter, uni);
                    const float v = iou_quotient(inter, uni);
                    // NaN (0/0) and union==0 never win
                    if (uni != 0.0f && (arg[r] < 0 || v > best[r])) { best[r] = v; arg[r] = j0 + j; }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int i = row0 + r * THREADS + threadIdx.x;
        if (i < n) {
            succ[off + i] = arg[r] < 0 ? -1 : out_base + arg[r];
            best_iou[off + i] = arg[r] < 0 ? 0.0f : best[r];
        }
    }
}

template <int THREADS, int ROWS>
static int launch_link(const float* boxes, const int32_t* seg_offsets, int n_segs, int max_seg_len,
                       const float* halo_boxes, int n_halo, const int32_t* n_halo_dev, int halo_row_base,
                       int32_t* succ, float* best_iou, cudaStream_t st) {
    constexpr int RPC = THREADS * ROWS;
    dim3 grid((unsigned)n_segs, (unsigned)((max_seg_len + RPC - 1) / RPC));
    link_frames_kernel<THREADS, ROWS><<<grid, THREADS, 0, st>>>(
        (const float4*)boxes, seg_offsets, n_segs, (const float4*)halo_boxes, n_halo, n_halo_dev, halo_row_base,
        succ, best_iou);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

// link_sorted.cu: the same link with the pairs that cannot overlap in x left out (frames of up to 2048 boxes)
size_t link_sorted_ws_bytes(int64_t n_rows, int n_segs, int n_halo);
int launch_link_frames_sorted(const float* boxes, const int32_t* seg_offsets, int n_segs, int max_seg_len,
                              const float* halo_boxes, int n_halo, const int32_t* n_halo_dev, int halo_row_base,
                              int32_t* succ, float* best_iou, int64_t n_rows, void* ws, cudaStream_t st);

}  // namespace vdet

using namespace vdet;

extern "C" size_t vdet_link_workspace_bytes(int64_t n_rows, int n_segs, int n_halo) {
    return link_sorted_ws_bytes(n_rows > 0 ? n_rows : 0, n_segs > 0 ? n_segs : 0, n_halo > 0 ? n_halo : 0);
}

extern "C" int vdet_link_frames_f32(const float* boxes, const int32_t* seg_offsets, int n_segs,
                                    int max_seg_len, const float* halo_boxes, int n_halo,
                                    const int32_t* n_halo_dev, int halo_row_base,
                                    int32_t* succ, float* best_iou, int64_t n_rows,
                                    void* ws, size_t ws_bytes, void* stream) {
    VDET_REQUIRE(n_segs >= 0 && max_seg_len >= 0 && n_halo >= 0 && n_rows >= 0, "link_frames: negative size");
    if (n_segs == 0 || n_rows == 0 || max_seg_len == 0) return VDET_OK;
    VDET_REQUIRE(((uintptr_t)boxes & 15) == 0 && ((uintptr_t)halo_boxes & 15) == 0,
                 "link_frames: boxes must be 16-byte aligned");
    VDET_REQUIRE(max_seg_len <= 65535 * 64, "link_frames: frame too long");
    // with a workspace: x1-sorted frames, only the pairs that can overlap in x are evaluated (same results)
    static const bool no_sort = getenv("VDET_LINK_NO_SORT") != nullptr && atoi(getenv("VDET_LINK_NO_SORT")) != 0;
    // (measured: 2500 x 1000 boxes 1.83 -> 1.11 ms, 500 x 2000 boxes 1.53 -> 0.88 ms; at 300 boxes per frame a warp's rows
    // span most of the frame in x and sorting only breaks even, so short frames keep the plain scan)
    if (ws != nullptr && !no_sort && max_seg_len <= 2048 && n_halo <= 2048 && max_seg_len >= 512 &&
        ws_bytes >= link_sorted_ws_bytes(n_rows, n_segs, halo_boxes ? n_halo : 0))
        return launch_link_frames_sorted(boxes, seg_offsets, n_segs, max_seg_len, halo_boxes, halo_boxes ? n_halo : 0,
                                         n_halo_dev, halo_row_base, succ, best_iou, n_rows, ws, (cudaStream_t)stream);
    // rows per CTA: fewest padded row slots, then fewest CTAs per frame
    const int cand[4] = {64, 128, 256, 512};
    int best_rpc = 64;
    long best_pad = -1;
    for (int k = 0; k < 4; ++k) {
        const long pad = ((long)max_seg_len + cand[k] - 1) / cand[k] * cand[k];
        if (best_pad < 0 || pad <= best_pad) { best_pad = pad; best_rpc = cand[k]; }
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (best_rpc) {
        case 64:  return launch_link<64, 1>(boxes, seg_offsets, n_segs, max_seg_len, halo_boxes, n_halo, n_halo_dev, halo_row_base, succ, best_iou, st);
        case 128: return launch_link<64, 2>(boxes, seg_offsets, n_segs, max_seg_len, halo_boxes, n_halo, n_halo_dev, halo_row_base, succ, best_iou, st);
        case 256: return launch_link<128, 2>(boxes, seg_offsets, n_segs, max_seg_len, halo_boxes, n_halo, n_halo_dev, halo_row_base, succ, best_iou, st);
        default:  return launch_link<128, 4>(boxes, seg_offsets, n_segs, max_seg_len, halo_boxes, n_halo, n_halo_dev, halo_row_base, succ, best_iou, st);
    }
}
