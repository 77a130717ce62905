// primitives.cu -- exclusive scan and stable LSD radix sort (hand-written; no CUB/Thrust).
//
// Used by the vid_nms pipeline (utils/nms.pyx:71-125): rows are grouped by frame with a
// stable sort of the frame column, and the kept rows are put into the reference's global
// descending-score order (nms.pyx:80,97-125) with a stable sort of the score keys.
#include "primitives.cuh"

namespace vdet {

// =========================================================================================
// exclusive scan
// =========================================================================================
constexpr int SC_T = 256;
constexpr int SC_I = 8;
constexpr int SC_TILE = SC_T * SC_I;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// Exclusive scan of one value per thread across the block; returns the exclusive prefix and
// writes the block total to *block_total (same value in every thread).
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* block_total) {
    __shared__ uint32_t s_warp[SC_T / 32];
    __shared__ uint32_t s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SC_T / 32 ? s_warp[lane] : 0u;
        const uint32_t wi = warp_incl_scan(w, lane);
        if (lane < SC_T / 32) s_warp[lane] = wi - w;
        if (lane == SC_T / 32 - 1) s_total = wi;
    }
    __syncthreads();
    *block_total = s_total;
    return s_warp[warp] + incl - v;
}

__global__ void __launch_bounds__(SC_T) k_scan_reduce(const uint32_t* __restrict__ in, int64_t n,
                                                      uint32_t* __restrict__ partial) {
    const int64_t base = (int64_t)blockIdx.x * SC_TILE;
    uint32_t s = 0;
    for (int i = threadIdx.x; i < SC_TILE; i += SC_T) {
        const int64_t g = base + i;
        if (g < n) s += in[g];
    }
    uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SC_T) k_scan_tile(const uint32_t* in, uint32_t* out, int64_t n,
                                                    const uint32_t* __restrict__ tile_base,
                                                    uint32_t* __restrict__ total) {
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_I;
    uint32_t v[SC_I];
    uint32_t s = 0;
#pragma unroll
    for (int q = 0; q < SC_I; ++q) {
        v[q] = (base + q < n) ? in[base + q] : 0u;
        s += v[q];
    }
    uint32_t tot;
    uint32_t run = block_excl_scan(s, &tot) + (tile_base ? tile_base[blockIdx.x] : 0u);
    if (total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
        *total = tot + (tile_base ? tile_base[blockIdx.x] : 0u);
#pragma unroll
    for (int q = 0; q < SC_I; ++q) {
        if (base + q < n) out[base + q] = run;
        run += v[q];
    }
}

static inline int64_t scan_tiles(int64_t n) { return (n + SC_TILE - 1) / SC_TILE; }

size_t scan_scratch_elems(int64_t n) {
    size_t tot = 64;
    while (n > SC_TILE) {
        n = scan_tiles(n);
        tot += align_up((size_t)n, 64);
    }
    return tot;
}

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total,
                       uint32_t* scratch, cudaStream_t st) {
    if (n <= 0) {
        if (total) VDET_CUDA(cudaMemsetAsync(total, 0, sizeof(uint32_t), st));
        return VDET_OK;
    }
    const int64_t nb = scan_tiles(n);
    if (nb == 1) {
        k_scan_tile<<<1, SC_T, 0, st>>>(in, out, n, nullptr, total);
        VDET_LAUNCH_CHECK();
        return VDET_OK;
    }
    uint32_t* partial = scratch;
    k_scan_reduce<<<(unsigned)nb, SC_T, 0, st>>>(in, n, partial);
    VDET_LAUNCH_CHECK();
    int rc = exclusive_scan_u32(partial, partial, nb, nullptr, scratch + align_up((size_t)nb, 64), st);
    if (rc != VDET_OK) return rc;
    k_scan_tile<<<(unsigned)nb, SC_T, 0, st>>>(in, out, n, partial, total);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

// =========================================================================================
// stable LSD radix sort, 8 bits per pass
// =========================================================================================
constexpr int RS_T = 256;
constexpr int RS_WARPS = RS_T / 32;
constexpr int RS_I = 8;
constexpr int RS_TILE = RS_T * RS_I;
constexpr int RS_BINS = 256;

static inline int64_t radix_tiles(int64_t n) { return (n + RS_TILE - 1) / RS_TILE; }

size_t radix_scratch_elems(int64_t n) {
    const int64_t nt = radix_tiles(n > 0 ? n : 1);
    return align_up((size_t)(RS_BINS * nt), 64) + scan_scratch_elems(RS_BINS * nt);
}

__global__ void __launch_bounds__(RS_T) k_radix_hist(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                     uint32_t mask, uint32_t* __restrict__ hist, int64_t ntiles) {
    __shared__ uint32_t s_hist[RS_BINS];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int i = threadIdx.x; i < RS_TILE; i += RS_T) {
        const int64_t g = base + i;
        if (g < n) atomicAdd(&s_hist[(keys[g] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = s_hist[threadIdx.x];
}

__global__ void __launch_bounds__(RS_T) k_radix_scatter(const uint32_t* __restrict__ keys_in,
                                                        const uint32_t* __restrict__ vals_in,
                                                        uint32_t* __restrict__ keys_out,
                                                        uint32_t* __restrict__ vals_out, int64_t n, int shift,
                                                        uint32_t mask, const uint32_t* __restrict__ hist_scanned,
                                                        int64_t ntiles) {
    __shared__ uint32_t s_cnt[RS_WARPS][RS_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_T) (&s_cnt[0][0])[i] = 0;
    __syncthreads();

    uint32_t key[RS_I], val[RS_I], dig[RS_I], loc[RS_I];
    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (RS_I * 32);
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < RS_I; ++j) {
        const int64_t g = wbase + j * 32 + lane;
        const bool valid = g < n;
        key[j] = valid ? keys_in[g] : 0u;
        val[j] = valid ? vals_in[g] : 0u;
        dig[j] = valid ? ((key[j] >> shift) & mask) : 0xffffu;      // invalid lanes only match each other
        const unsigned peers = __match_any_sync(FULL, dig[j]);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && valid) {
            old = s_cnt[warp][dig[j]];
            s_cnt[warp][dig[j]] = old + __popc(peers);
        }
        old = __shfl_sync(FULL, old, leader);
        loc[j] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    {   // thread d turns the per-warp counts of digit d into global bases
        const int d = threadIdx.x;
        uint32_t run = hist_scanned[(int64_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = s_cnt[w][d];
            s_cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_I; ++j) {
        const int64_t g = wbase + j * 32 + lane;
        if (g < n) {
            const uint32_t o = s_cnt[warp][dig[j]] + loc[j];
            keys_out[o] = key[j];
            vals_out[o] = val[j];
        }
    }
}

int radix_sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt,
                     int64_t n, int begin_bit, int end_bit, uint32_t* scratch, cudaStream_t st) {
    if (n <= 1 || end_bit <= begin_bit) return 0;
    const int64_t nt = radix_tiles(n);
    uint32_t* hist = scratch;
    uint32_t* scan_scratch = scratch + align_up((size_t)(RS_BINS * nt), 64);
    int flip = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        const int bits = (end_bit - shift) < 8 ? (end_bit - shift) : 8;
        const uint32_t mask = (1u << bits) - 1u;
        uint32_t* kin = flip ? keys_alt : keys;
        uint32_t* vin = flip ? vals_alt : vals;
        uint32_t* kout = flip ? keys : keys_alt;
        uint32_t* vout = flip ? vals : vals_alt;
        k_radix_hist<<<(unsigned)nt, RS_T, 0, st>>>(kin, n, shift, mask, hist, nt);
        VDET_LAUNCH_CHECK();
        int rc = exclusive_scan_u32(hist, hist, RS_BINS * nt, nullptr, scan_scratch, st);
        if (rc != VDET_OK) return rc;
        k_radix_scatter<<<(unsigned)nt, RS_T, 0, st>>>(kin, vin, kout, vout, n, shift, mask, hist, nt);
        VDET_LAUNCH_CHECK();
        flip ^= 1;
    }
    return flip;
}

}  // namespace vdet
