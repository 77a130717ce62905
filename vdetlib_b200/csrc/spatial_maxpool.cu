// spatial_maxpool.cu -- fused IoU -> threshold -> arg-max (K3 of SURVEY 2.1).
//
// Replaces the inner body of dets_spatial_max_pooling / raw_dets_spatial_max_pooling
// (vdet/tubelet_cls.py:330-347, :515-532) and the anchor lookup of anchor_propagate
// (:375-377): for one tubelet box, IoU against every detection of its frame in float64
// exactly as utils/common.py:451-468, then
//   ARGMAX_SCORE: among dets with IoU > thresh (strict, :334) the FIRST arg-max class score
//                 (np.argmax, :337), or "none" (score -1e5, box unchanged, :341-345);
//   ARGMAX_IOU  : the FIRST arg-max IoU (:376), no threshold.
// One warp per tubelet box; lanes stride over the frame's detections (coalesced reads), each
// lane keeps its own running best in ascending index order, and a shuffle tree merges the 32
// candidates with the tie rule "lower index wins".
#include "common.cuh"

namespace vdet {

__device__ __forceinline__ double sm_area(double x1, double y1, double x2, double y2) {
    return __dmul_rn(__dadd_rn(__dsub_rn(x2, x1), 1.0), __dadd_rn(__dsub_rn(y2, y1), 1.0));
}

template <typename BT, typename ST>
__global__ void __launch_bounds__(256) spatial_maxpool_kernel(const BT* __restrict__ tub_boxes,
                                                              const int32_t* __restrict__ tub_seg, int64_t P,
                                                              const BT* __restrict__ det_boxes,
                                                              const ST* __restrict__ det_scores, int64_t score_ld,
                                                              const int32_t* __restrict__ seg_offsets, int n_segs,
                                                              double thresh, int mode,
                                                              int32_t* __restrict__ out_arg,
                                                              double* __restrict__ out_score) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const int seg = tub_seg[p];
    int off = 0, n = 0;
    if (seg >= 0 && seg < n_segs) { off = seg_offsets[seg]; n = seg_offsets[seg + 1] - off; }
    const double ax1 = (double)tub_boxes[p * 4 + 0], ay1 = (double)tub_boxes[p * 4 + 1];
    const double ax2 = (double)tub_boxes[p * 4 + 2], ay2 = (double)tub_boxes[p * 4 + 3];
    const double aa = sm_area(ax1, ay1, ax2, ay2);

    double best = 0.0;
    int arg = 0x7fffffff;                      // "none"
    for (int j = lane; j < n; j += 32) {
        const BT* b = det_boxes + (int64_t)(off + j) * 4;
        const double bx1 = (double)b[0], by1 = (double)b[1], bx2 = (double)b[2], by2 = (double)b[3];
        const double ix1 = dmax_np(ax1, bx1), ix2 = dmin_np(ax2, bx2);
        const double iy1 = dmax_np(ay1, by1), iy2 = dmin_np(ay2, by2);
        const double iw = dmax_np(0.0, __dadd_rn(__dsub_rn(ix2, ix1), 1.0));
        const double ih = dmax_np(0.0, __dadd_rn(__dsub_rn(iy2, iy1), 1.0));
        const double inter = __dmul_rn(iw, ih);
        const double ovr = __ddiv_rn(inter, __dsub_rn(__dadd_rn(aa, sm_area(bx1, by1, bx2, by2)), inter));
        if (mode == VDET_POOL_ARGMAX_SCORE) {
            if (ovr > thresh) {
                const double s = (double)det_scores[(int64_t)(off + j) * score_ld];
                if (arg == 0x7fffffff || s > best) { best = s; arg = j; }
            }
        } else {
            if (arg == 0x7fffffff || ovr > best) { best = ovr; arg = j; }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double ob = __shfl_xor_sync(FULL, best, d);
        const int oa = __shfl_xor_sync(FULL, arg, d);
        const bool take = (oa != 0x7fffffff) && (arg == 0x7fffffff || ob > best || (ob == best && oa < arg));
        if (take) { best = ob; arg = oa; }
    }
    if (lane == 0) {
        if (arg == 0x7fffffff) {
            out_arg[p] = -1;
            out_score[p] = -1e5;
        } else {
            out_arg[p] = off + arg;
            out_score[p] = (mode == VDET_POOL_ARGMAX_IOU)
                               ? (double)det_scores[(int64_t)(off + arg) * score_ld]
                               : best;          // ARGMAX_SCORE: the score; MAX_IOU: the IoU itself
        }
    }
}

template <typename BT, typename ST>
static int launch_pool(const void* tub_boxes, const int32_t* tub_seg, int64_t p, const void* det_boxes,
                       const void* det_scores, int64_t score_ld, const int32_t* seg_offsets, int n_segs,
                       double thresh, int mode, int32_t* out_arg, double* out_score, cudaStream_t st) {
    const unsigned grid = (unsigned)((p + 7) / 8);
    spatial_maxpool_kernel<BT, ST><<<grid, 256, 0, st>>>((const BT*)tub_boxes, tub_seg, p, (const BT*)det_boxes,
                                                         (const ST*)det_scores, score_ld, seg_offsets, n_segs,
                                                         thresh, mode, out_arg, out_score);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

}  // namespace vdet

using namespace vdet;

extern "C" int vdet_spatial_maxpool(const void* tub_boxes, const int32_t* tub_seg, int64_t p,
                                    const void* det_boxes, int box_dtype,
                                    const void* det_scores, int64_t score_ld, int score_dtype,
                                    const int32_t* det_seg_offsets, int n_segs,
                                    double thresh, int mode,
                                    int32_t* out_arg, double* out_score, void* stream) {
    VDET_REQUIRE(p >= 0 && n_segs >= 0, "spatial_maxpool: negative size");
    VDET_REQUIRE(mode == VDET_POOL_ARGMAX_SCORE || mode == VDET_POOL_ARGMAX_IOU || mode == VDET_POOL_MAX_IOU,
                 "spatial_maxpool: bad mode");
    VDET_REQUIRE((box_dtype | 1) == 1 && (score_dtype | 1) == 1, "spatial_maxpool: bad dtype");
    if (p == 0) return VDET_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define POOL_ARGS tub_boxes, tub_seg, p, det_boxes, det_scores, score_ld, det_seg_offsets, n_segs, thresh, mode, out_arg, out_score, st
    if (box_dtype == VDET_DTYPE_F32 && score_dtype == VDET_DTYPE_F32) return launch_pool<float, float>(POOL_ARGS);
    if (box_dtype == VDET_DTYPE_F32 && score_dtype == VDET_DTYPE_F64) return launch_pool<float, double>(POOL_ARGS);
    if (box_dtype == VDET_DTYPE_F64 && score_dtype == VDET_DTYPE_F32) return launch_pool<double, float>(POOL_ARGS);
    return launch_pool<double, double>(POOL_ARGS);
#undef POOL_ARGS
}
