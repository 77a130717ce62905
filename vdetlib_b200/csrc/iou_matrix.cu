// iou_matrix.cu -- dense pairwise IoU (K2 of SURVEY 2.1).
//
// Replaces utils/common.py:451-468 (`iou`, float64) and provides the float32 matrix with the
// pair arithmetic of utils/nms.pyx:57-64 -- the kernel BASELINE.json's "% HBM peak on IoU
// kernel" refers to.  The kernel is a pure streaming writer: inputs are 16*(A+B) bytes, the
// output is 4*A*B (f32) or 8*A*B (f64) bytes, so every design choice is about keeping
// 128-byte-coalesced stores in flight while the FP32 pipes produce one IoU per 4 bytes:
//   * a CTA owns a [TILE_R x TILE_C] output tile; the TILE_C b-boxes and their areas are
//     staged once in shared memory (float4), the TILE_R a-boxes/areas as warp-uniform values;
//   * each thread keeps 4 b-boxes in registers and walks the tile's rows, so per output the
//     inner loop is the ~22 FP32/FMNMX instructions of the IoU itself plus 1/4 store;
//   * stores are 16-byte vectors when the row pitch allows (nb % 4 == 0), else four 4-byte
//     stores laid out so that a warp still writes 128 contiguous bytes per instruction;
//   * stores use the streaming (evict-first) policy: the matrix is written once, never read.
#include "common.cuh"

namespace vdet {

constexpr int IOU_THREADS = 256;
constexpr int IOU_TILE_C = IOU_THREADS * 4;   // 1024 columns per CTA
#ifndef VDET_IOU_STORE
#define VDET_IOU_STORE 0          // 0: st.global.cs (evict first), 1: default policy, 2: st.global.wt
#endif
constexpr int IOU_TILE_R = 32;                // rows per CTA (float64 kernel; float32: 32 or 128, chosen per launch)
constexpr int IOU_TILE_R_BIG = 128;           // float32, tall matrices: 16384^2 runs at 91.5 % of the measured HBM peak with
                                              // 128-row tiles against 87.6 % with 32 (90.4 % with 64, 78.8 % with 16)

__device__ __forceinline__ void iou_store4(float4* p, const float4 v) {
#if VDET_IOU_STORE == 1
    *p = v;
#elif VDET_IOU_STORE == 2
    __stwt(p, v);
#else
    __stcs(p, v);
#endif
}

template <bool VEC, bool FAST, int TILE_R>
__device__ __forceinline__ void iou_matrix_f32_body(const float4* __restrict__ a, int64_t na,
                                                                     const float4* __restrict__ b, int64_t nb,
                                                                     float* __restrict__ out) {
    __shared__ float4 s_a[TILE_R];
    __shared__ float s_aa[TILE_R];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * IOU_TILE_C;
    const int64_t r0 = (int64_t)blockIdx.y * TILE_R;
    // this thread's 4 columns (re-read here: cheap, L1-resident after the sanity pass)
    int64_t col[4];
    float4 bb[4];
    float ba[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        col[k] = VEC ? (c0 + 4 * tid + k) : (c0 + warp * 128 + k * 32 + lane);
        bb[k] = col[k] < nb ? __ldg(b + col[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
        ba[k] = area_f32(bb[k]);
        // keep the area in its register: without this nvcc rematerialises it (4 FADD + 1 FMUL)
        // inside the row loop -- seen as 8.3 instead of 6 FADD per IoU in the first ncu capture
        asm volatile("" : "+f"(ba[k]));
    }
    if (tid < TILE_R) {
        const int64_t r = r0 + tid;
        const float4 v = r < na ? __ldg(a + r) : make_float4(0.f, 0.f, 0.f, 0.f);
        s_a[tid] = v;
        s_aa[tid] = area_f32(v);
    }
    __syncthreads();
    const int rows = (int)((na - r0) < TILE_R ? (na - r0) : TILE_R);
    // FAST (every box of the tile is sane): the four IoUs of a thread and row are two PACKED pairs -- FADD2 /
    // FMUL2 / FFMA2 evaluate both halves with one instruction each (common.cuh), ~14 instead of ~22
    // instructions per IoU, which is what moves this kernel from issue bound to HBM-write bound.
    const f32x2 ba01 = pk2(ba[0], ba[1]), ba23 = pk2(ba[2], ba[3]);
    auto row_of_four = [&](const float4 av, const float aa, float (&v)[4]) {
        if (FAST && VEC) {
            const f32x2 aa2 = pk2(aa, aa);
            f32x2 inter, uni, nuni;
            inter_union_f32x2(av, aa2, bb[0], bb[1], ba01, inter, uni, nuni);
            upk2(div_sane2(inter, nuni), v[0], v[1]);
            inter_union_f32x2(av, aa2, bb[2], bb[3], ba23, inter, uni, nuni);
            upk2(div_sane2(inter, nuni), v[2], v[3]);
        } else {
            // (the scalar-store layout -- row pitch not a multiple of 4 -- measured slower with packed pairs: 0.268 ms
            // against 0.230 at 16381^2, so it keeps the scalar sequence)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float inter, uni;
                inter_union_f32(av, aa, bb[k], ba[k], inter, uni);
                v[k] = FAST ? div_sane(inter, uni) : iou_quotient(inter, uni);
            }
        }
    };
    if (VEC && rows == TILE_R && c0 + IOU_TILE_C <= nb) {
        // interior tile: no bounds checks, one 16-byte streaming store per thread and row
        float* row = out + r0 * nb + c0 + 4 * tid;
#pragma unroll 4
        for (int r = 0; r < TILE_R; ++r, row += nb) {
            float v[4];
            row_of_four(s_a[r], s_aa[r], v);
            iou_store4(reinterpret_cast<float4*>(row), make_float4(v[0], v[1], v[2], v[3]));
        }
        return;
    }
#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
        float v[4];
        row_of_four(s_a[r], s_aa[r], v);
        float* row = out + (r0 + r) * nb;
        if (VEC) {
            if (col[3] < nb) {
                __stcs(reinterpret_cast<float4*>(row + col[0]), make_float4(v[0], v[1], v[2], v[3]));
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (col[k] < nb) __stcs(row + col[k], v[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (col[k] < nb) __stcs(row + col[k], v[k]);
        }
    }
}

// The CTA first votes whether every box it touches is "sane" (common.cuh): if so the whole tile
// uses the branch-free 6-instruction division, else the generic IEEE path.  Same bits either way.
template <bool VEC, int TILE_R>
__global__ void __launch_bounds__(IOU_THREADS) iou_matrix_f32_kernel(const float4* __restrict__ a, int64_t na,
                                                                     const float4* __restrict__ b, int64_t nb,
                                                                     float* __restrict__ out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * IOU_TILE_C;
    const int64_t r0 = (int64_t)blockIdx.y * TILE_R;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t col = VEC ? (c0 + 4 * tid + k) : (c0 + warp * 128 + k * 32 + lane);
        if (col < nb) ok = ok && box_sane(__ldg(b + col));
    }
    if (tid < TILE_R && r0 + tid < na) ok = ok && box_sane(__ldg(a + r0 + tid));
    if (__syncthreads_and(ok)) iou_matrix_f32_body<VEC, true, TILE_R>(a, na, b, nb, out);
    else iou_matrix_f32_body<VEC, false, TILE_R>(a, na, b, nb, out);
}

// utils/common.py:451-468 in float64, operation for operation.
__device__ __forceinline__ double area_f64(const double x1, const double y1, const double x2, const double y2) {
    return __dmul_rn(__dadd_rn(__dsub_rn(x2, x1), 1.0), __dadd_rn(__dsub_rn(y2, y1), 1.0));
}

__device__ __forceinline__ double pair_iou_f64(const double ax1, const double ay1, const double ax2,
                                               const double ay2, const double aa, const double bx1,
                                               const double by1, const double bx2, const double by2,
                                               const double ba) {
    const double ix1 = dmax_np(ax1, bx1), ix2 = dmin_np(ax2, bx2);
    const double iy1 = dmax_np(ay1, by1), iy2 = dmin_np(ay2, by2);
    const double iw = dmax_np(0.0, __dadd_rn(__dsub_rn(ix2, ix1), 1.0));
    const double ih = dmax_np(0.0, __dadd_rn(__dsub_rn(iy2, iy1), 1.0));
    const double inter = __dmul_rn(iw, ih);
    return __ddiv_rn(inter, __dsub_rn(__dadd_rn(aa, ba), inter));
}

constexpr int IOU64_TILE_C = IOU_THREADS * 2;

__global__ void __launch_bounds__(IOU_THREADS) iou_matrix_f64_kernel(const double* __restrict__ a, int64_t na,
                                                                     const double* __restrict__ b, int64_t nb,
                                                                     double* __restrict__ out) {
    __shared__ double s_a[IOU_TILE_R][5];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * IOU64_TILE_C;
    const int64_t r0 = (int64_t)blockIdx.y * IOU_TILE_R;
    if (tid < IOU_TILE_R) {
        const int64_t r = r0 + tid;
        double x1 = 0, y1 = 0, x2 = 0, y2 = 0;
        if (r < na) { x1 = a[r * 4]; y1 = a[r * 4 + 1]; x2 = a[r * 4 + 2]; y2 = a[r * 4 + 3]; }
        s_a[tid][0] = x1; s_a[tid][1] = y1; s_a[tid][2] = x2; s_a[tid][3] = y2;
        s_a[tid][4] = area_f64(x1, y1, x2, y2);
    }
    int64_t col[2];
    double bx1[2], by1[2], bx2[2], by2[2], ba[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        col[k] = c0 + warp * 64 + k * 32 + lane;
        bx1[k] = by1[k] = bx2[k] = by2[k] = 0.0;
        if (col[k] < nb) {
            const double2 lo = __ldg(reinterpret_cast<const double2*>(b + col[k] * 4));
            const double2 hi = __ldg(reinterpret_cast<const double2*>(b + col[k] * 4 + 2));
            bx1[k] = lo.x; by1[k] = lo.y; bx2[k] = hi.x; by2[k] = hi.y;
        }
        ba[k] = area_f64(bx1[k], by1[k], bx2[k], by2[k]);
    }
    __syncthreads();
    const int rows = (int)((na - r0) < IOU_TILE_R ? (na - r0) : IOU_TILE_R);
    for (int r = 0; r < rows; ++r) {
        const double ax1 = s_a[r][0], ay1 = s_a[r][1], ax2 = s_a[r][2], ay2 = s_a[r][3], aa = s_a[r][4];
        double* row = out + (r0 + r) * nb;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const double v = pair_iou_f64(ax1, ay1, ax2, ay2, aa, bx1[k], by1[k], bx2[k], by2[k], ba[k]);
            if (col[k] < nb) __stcs(row + col[k], v);
        }
    }
}

// Suppression bit matrix of one frame (the shared-memory structure of nms_frames_kernel,
// exported for tests / diagnostics).  One warp per (32-row block, 32-column word).
__global__ void __launch_bounds__(256) iou_bitmask_kernel(const float4* __restrict__ boxes, int n, float T,
                                                          uint32_t* __restrict__ mask, uint32_t* status) {
    const int lane = threadIdx.x & 31;
    const int W = (n + 31) >> 5;
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= W * W) return;
    const int rb = tile / W, cb = tile - rb * W;
    const int i = rb * 32 + lane;
    const float4 bi = i < n ? __ldg(boxes + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float ai = area_f32(bi);
    uint32_t word = 0;
    bool zero = false;
    for (int jj = 0; jj < 32; ++jj) {
        const int j = cb * 32 + jj;
        if (j >= n) break;
        const float4 bj = __ldg(boxes + j);
        float inter, uni;
        inter_union_f32(bi, ai, bj, area_f32(bj), inter, uni);
        if (iou_ge(inter, uni, T)) word |= (1u << jj);
        zero |= (uni == 0.0f) && (i != j) && (i < n);
    }
    if (i < n) mask[(int64_t)i * W + cb] = word;
    if (__any_sync(FULL, zero) && lane == 0) atomicOr(status, VDET_STATUS_ZERO_DIVISION);
}

}  // namespace vdet

using namespace vdet;

extern "C" int vdet_iou_matrix_f32(const float* a, int64_t na, const float* b, int64_t nb,
                                   float* out, void* stream) {
    VDET_REQUIRE(na >= 0 && nb >= 0, "iou_matrix: negative size");
    if (na == 0 || nb == 0) return VDET_OK;
    VDET_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0, "iou_matrix: boxes must be 16-byte aligned");
    // tall matrices take 128-row tiles (fewer, longer CTAs: less per-CTA set-up per byte written); short ones keep
    // 32 rows so that a few hundred rows still spread over the SMs
    const int tile_r = na >= 2048 ? IOU_TILE_R_BIG : IOU_TILE_R;
    const int64_t gx = (nb + IOU_TILE_C - 1) / IOU_TILE_C, gy = (na + tile_r - 1) / tile_r;
    dim3 grid((unsigned)gx, (unsigned)gy);
    VDET_REQUIRE(gy <= 65535, "iou_matrix: more than 2M rows per call");
    const bool vec = (nb % 4 == 0) && (((uintptr_t)out & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const float4 *a4 = (const float4*)a, *b4 = (const float4*)b;
    if (tile_r == IOU_TILE_R_BIG) {
        if (vec) iou_matrix_f32_kernel<true, IOU_TILE_R_BIG><<<grid, IOU_THREADS, 0, st>>>(a4, na, b4, nb, out);
        else iou_matrix_f32_kernel<false, IOU_TILE_R_BIG><<<grid, IOU_THREADS, 0, st>>>(a4, na, b4, nb, out);
    } else {
        if (vec) iou_matrix_f32_kernel<true, IOU_TILE_R><<<grid, IOU_THREADS, 0, st>>>(a4, na, b4, nb, out);
        else iou_matrix_f32_kernel<false, IOU_TILE_R><<<grid, IOU_THREADS, 0, st>>>(a4, na, b4, nb, out);
    }
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

extern "C" int vdet_iou_matrix_f64(const double* a, int64_t na, const double* b, int64_t nb,
                                   double* out, void* stream) {
    VDET_REQUIRE(na >= 0 && nb >= 0, "iou_matrix: negative size");
    if (na == 0 || nb == 0) return VDET_OK;
    VDET_REQUIRE(((uintptr_t)b & 15) == 0, "iou_matrix_f64: boxes must be 16-byte aligned");
    const int64_t gx = (nb + IOU64_TILE_C - 1) / IOU64_TILE_C, gy = (na + IOU_TILE_R - 1) / IOU_TILE_R;
    VDET_REQUIRE(gy <= 65535, "iou_matrix_f64: more than 2M rows per call");
    dim3 grid((unsigned)gx, (unsigned)gy);
    iou_matrix_f64_kernel<<<grid, IOU_THREADS, 0, (cudaStream_t)stream>>>(a, na, b, nb, out);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

extern "C" int vdet_iou_bitmask_f32(const float* boxes, int n, double thresh, uint32_t* mask,
                                    uint32_t* status, void* stream) {
    VDET_REQUIRE(n >= 0, "iou_bitmask: negative size");
    if (n == 0) return VDET_OK;
    VDET_REQUIRE(((uintptr_t)boxes & 15) == 0, "iou_bitmask: boxes must be 16-byte aligned");
    const int W = (n + 31) / 32;
    const int64_t tiles = (int64_t)W * W;
    iou_bitmask_kernel<<<(unsigned)((tiles + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)boxes, n, thresh_ceil_f32(thresh), mask, status);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}
