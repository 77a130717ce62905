// tubelet_chain.cu -- turn frame-to-frame links into tubelet score rows (build-defined glue between
// the link of SURVEY 8a row 15 and the temporal smoothing of rows 10-12; the reference gets its
// tubelets from external trackers, vdet/track.py:18-106).
//
//   follow_links : chain k starts at packed row start[k] and follows succ[] frame by frame; a chain
//                  ends at succ == -1, at succ >= n_rows (a halo successor: the chain continues on
//                  the next shard, include/vdet_b200.h) or when the link IoU drops below min_iou.  One thread per
//                  chain (pointer chasing is inherently serial along the frame axis; chains run in
//                  parallel), rows written frame-major so that the T stores of a warp's 32 chains
//                  coalesce.
//   gather_rows  : out[k, c, t] = scores[chain_rows[t, k], c]  (or `missing` after the chain ended):
//                  the [tubelet x class, frame] rows that completion / max-pool / conv consume.
#include "common.cuh"

namespace vdet {

__global__ void __launch_bounds__(128) follow_links_kernel(const int32_t* __restrict__ succ,
                                                           const float* __restrict__ link_iou, int64_t n_rows,
                                                           const int32_t* __restrict__ start, int n_chains,
                                                           int n_frames, float min_iou,
                                                           int32_t* __restrict__ chain_rows) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_chains) return;
    int row = start[k];
    if (row >= n_rows) row = -1;
    for (int t = 0; t < n_frames; ++t) {
        chain_rows[(int64_t)t * n_chains + k] = row;
        if (row >= 0) {
            const int nxt = __ldg(succ + row);
            row = (nxt >= 0 && nxt < n_rows && __ldg(link_iou + row) >= min_iou) ? nxt : -1;
        }
    }
}

__global__ void __launch_bounds__(256) gather_chain_scores_kernel(const float* __restrict__ scores, int n_classes,
                                                                  const int32_t* __restrict__ chain_rows,
                                                                  int n_chains, int n_frames, float missing,
                                                                  float* __restrict__ out) {
    // grid: (frame tiles, chains); thread = one frame of one chain, loops over classes
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (t >= n_frames) return;
    const int row = chain_rows[(int64_t)t * n_chains + k];
    float* dst = out + (int64_t)k * n_classes * n_frames + t;
    if (row < 0) {
        for (int c = 0; c < n_classes; ++c) dst[(int64_t)c * n_frames] = missing;
    } else {
        const float* src = scores + (int64_t)row * n_classes;
        for (int c = 0; c < n_classes; ++c) dst[(int64_t)c * n_frames] = __ldg(src + c);
    }
}

}  // namespace vdet

using namespace vdet;

extern "C" int vdet_follow_links(const int32_t* succ, const float* link_iou, int64_t n_rows, const int32_t* start,
                                 int n_chains, int n_frames, float min_iou, int32_t* chain_rows, void* stream) {
    VDET_REQUIRE(n_chains >= 0 && n_frames >= 0 && n_rows >= 0, "follow_links: negative size");
    if (n_chains == 0 || n_frames == 0) return VDET_OK;
    follow_links_kernel<<<(unsigned)((n_chains + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        succ, link_iou, n_rows, start, n_chains, n_frames, min_iou, chain_rows);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

extern "C" int vdet_gather_chain_scores_f32(const float* scores, int n_classes, const int32_t* chain_rows,
                                            int n_chains, int n_frames, float missing, float* out, void* stream) {
    VDET_REQUIRE(n_chains >= 0 && n_frames >= 0 && n_classes >= 1, "gather_chain_scores: bad size");
    VDET_REQUIRE(n_chains <= 65535, "gather_chain_scores: more than 65535 chains per call");
    if (n_chains == 0 || n_frames == 0) return VDET_OK;
    dim3 grid((unsigned)((n_frames + 255) / 256), (unsigned)n_chains);
    gather_chain_scores_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(scores, n_classes, chain_rows, n_chains,
                                                                       n_frames, missing, out);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}
