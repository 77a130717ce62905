// topk.cu -- post-CNN per-class score floor + cap.
//
// Replaces the per-frame, per-class NumPy loop of vdet/video_det.py:88-100:
//     inds = where(scores[:, j] > thresh); if len(inds) > max_per_image: keep argsort(-s)[:max]
// One warp per (frame, class j >= 1).  The warp first counts the candidates with ballots; when
// they fit (count <= k) it emits them in ascending row order with a ballot prefix; otherwise
// it sorts (descending score, ties by ascending row) with the register bitonic network shared
// with the NMS kernel and emits the first k.
#include "common.cuh"
#include "warp_sort.cuh"

namespace vdet {

template <int NPER>
__global__ void __launch_bounds__(256) threshold_topk_kernel(const float* __restrict__ scores,
                                                             const int32_t* __restrict__ seg_offsets, int n_segs,
                                                             int C, float thresh, int k,
                                                             int32_t* __restrict__ idx_out,
                                                             int32_t* __restrict__ cnt_out) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= (int64_t)n_segs * C) return;
    const int seg = (int)(item / C), c = (int)(item - (int64_t)seg * C);
    int32_t* out = idx_out + item * k;
    if (c == 0) {                                  // background column is skipped (video_det.py:89)
        for (int e = lane; e < k; e += 32) out[e] = -1;
        if (lane == 0) cnt_out[item] = 0;
        return;
    }
    const int off = seg_offsets[seg];
    const int n = seg_offsets[seg + 1] - off;
    const float* col = scores + (int64_t)off * C + c;

    uint64_t key[NPER];
    int count = 0;
#pragma unroll
    for (int r = 0; r < NPER; ++r) {
        const int e = r * 32 + lane;
        key[r] = ~0ull;
        bool hit = false;
        if (e < n) {
            const float s = __ldg(col + (int64_t)e * C);
            hit = s > thresh;
            if (hit) key[r] = ((uint64_t)f32_key_desc(s) << 32) | (uint32_t)e;
        }
        count += __popc(__ballot_sync(FULL, hit));
    }
    if (count <= k) {
        int base = 0;
#pragma unroll
        for (int r = 0; r < NPER; ++r) {
            const bool hit = key[r] != ~0ull;
            const unsigned b = __ballot_sync(FULL, hit);
            if (hit) out[base + __popc(b & lanemask_lt())] = r * 32 + lane;
            base += __popc(b);
        }
        for (int e = count + lane; e < k; e += 32) out[e] = -1;
        if (lane == 0) cnt_out[item] = count;
    } else {
        warp_bitonic_sort<NPER>(key, lane);
#pragma unroll
        for (int r = 0; r < NPER; ++r) {
            const int pos = lane * NPER + r;
            if (pos < k) out[pos] = (int32_t)(uint32_t)key[r];
        }
        if (lane == 0) cnt_out[item] = k;
    }
}


// Frames with more than 1024 rows: same logic, the (key,index) array lives in shared memory
// (one warp = one (frame, class), 4 warps per CTA, warp-synchronous bitonic sort).
constexpr int TK_BIG_WARPS = 4;

__global__ void __launch_bounds__(TK_BIG_WARPS * 32) threshold_topk_big_kernel(const float* __restrict__ scores,
                                                                               const int32_t* __restrict__ seg_offsets,
                                                                               int n_segs, int C, float thresh, int k,
                                                                               int npad, int32_t* __restrict__ idx_out,
                                                                               int32_t* __restrict__ cnt_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw) + (size_t)warp * npad;
    const int64_t item = (int64_t)blockIdx.x * TK_BIG_WARPS + warp;
    if (item >= (int64_t)n_segs * C) return;
    const int seg = (int)(item / C), c = (int)(item - (int64_t)seg * C);
    int32_t* out = idx_out + item * k;
    if (c == 0) {
        for (int e = lane; e < k; e += 32) out[e] = -1;
        if (lane == 0) cnt_out[item] = 0;
        return;
    }
    const int off = seg_offsets[seg];
    const int n = seg_offsets[seg + 1] - off;
    const float* col = scores + (int64_t)off * C + c;
    int count = 0;
    for (int base = 0; base < npad; base += 32) {
        const int e = base + lane;
        bool hit = false;
        uint64_t key = ~0ull;
        if (e < n) {
            const float s = __ldg(col + (int64_t)e * C);
            hit = s > thresh;
            if (hit) key = ((uint64_t)f32_key_desc(s) << 32) | (uint32_t)e;
        }
        keys[e] = key;
        count += __popc(__ballot_sync(FULL, hit));
    }
    __syncwarp();
    if (count <= k) {
        int base_out = 0;
        for (int base = 0; base < n; base += 32) {
            const int e = base + lane;
            const bool hit = e < n && keys[e] != ~0ull;
            const unsigned b = __ballot_sync(FULL, hit);
            if (hit) out[base_out + __popc(b & lanemask_lt())] = e;
            base_out += __popc(b);
        }
        for (int e = count + lane; e < k; e += 32) out[e] = -1;
        if (lane == 0) cnt_out[item] = count;
        return;
    }
    for (int size = 2; size <= npad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < (npad >> 1); t += 32) {
                const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
                const int j = i + stride;
                const uint64_t a = keys[i], b = keys[j];
                const bool up = (i & size) == 0;
                if ((a > b) == up) { keys[i] = b; keys[j] = a; }
            }
            __syncwarp();
        }
    }
    for (int e = lane; e < k; e += 32) out[e] = (int32_t)(uint32_t)keys[e];
    if (lane == 0) cnt_out[item] = k;
}

template <int NPER>
static int launch_topk(const float* scores, const int32_t* seg_offsets, int n_segs, int C, float thresh, int k,
                       int32_t* idx_out, int32_t* cnt_out, cudaStream_t st) {
    const int64_t items = (int64_t)n_segs * C;
    threshold_topk_kernel<NPER><<<(unsigned)((items + 7) / 8), 256, 0, st>>>(scores, seg_offsets, n_segs, C, thresh,
                                                                             k, idx_out, cnt_out);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

}  // namespace vdet

using namespace vdet;

extern "C" int vdet_threshold_topk_f32(const float* scores, const int32_t* seg_offsets, int n_segs,
                                       int max_seg_len, int n_classes, float thresh, int k,
                                       int32_t* idx_out, int32_t* cnt_out, void* stream) {
    VDET_REQUIRE(n_segs >= 0 && max_seg_len >= 0 && n_classes >= 1 && k >= 1, "threshold_topk: bad size");
    if (n_segs == 0) return VDET_OK;
    if (max_seg_len > 1024) {
        int npad = 2048;
        while (npad < max_seg_len) npad <<= 1;
        const size_t smem = (size_t)TK_BIG_WARPS * npad * sizeof(uint64_t);
        if (smem > max_dynamic_smem(threshold_topk_big_kernel)) {
            set_error("threshold_topk: max_seg_len %d is not supported by this build", max_seg_len);
            return VDET_ERR_UNSUPPORTED;
        }
        VDET_CUDA(allow_dynamic_smem(threshold_topk_big_kernel, smem));
        const int64_t items = (int64_t)n_segs * n_classes;
        threshold_topk_big_kernel<<<(unsigned)((items + TK_BIG_WARPS - 1) / TK_BIG_WARPS), TK_BIG_WARPS * 32, smem,
                                    (cudaStream_t)stream>>>(scores, seg_offsets, n_segs, n_classes, thresh, k, npad,
                                                            idx_out, cnt_out);
        VDET_LAUNCH_CHECK();
        return VDET_OK;
    }
    int nper = 1;
    while (32 * nper < max_seg_len) nper <<= 1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (nper) {
        case 1:  return launch_topk<1>(scores, seg_offsets, n_segs, n_classes, thresh, k, idx_out, cnt_out, st);
        case 2:  return launch_topk<2>(scores, seg_offsets, n_segs, n_classes, thresh, k, idx_out, cnt_out, st);
        case 4:  return launch_topk<4>(scores, seg_offsets, n_segs, n_classes, thresh, k, idx_out, cnt_out, st);
        case 8:  return launch_topk<8>(scores, seg_offsets, n_segs, n_classes, thresh, k, idx_out, cnt_out, st);
        case 16: return launch_topk<16>(scores, seg_offsets, n_segs, n_classes, thresh, k, idx_out, cnt_out, st);
        default: return launch_topk<32>(scores, seg_offsets, n_segs, n_classes, thresh, k, idx_out, cnt_out, st);
    }
}
