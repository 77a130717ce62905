// warp_sort.cuh -- register-resident warp bitonic sort of 32*NPER 64-bit keys.
#pragma once
#include "common.cuh"

namespace vdet {

// ---- warp-level bitonic sort of 32*NPER 64-bit keys, blocked layout (position = lane*NPER+r)
template <int NPER>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t (&k)[NPER], const int lane) {
#pragma unroll
    for (int size = 2; size <= 32 * NPER; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= NPER) {
                const int lstride = stride / NPER;
                const bool lower = (lane & lstride) == 0;
#pragma unroll
                for (int r = 0; r < NPER; ++r) {
                    const int q = lane * NPER + r;
                    const bool up = (q & size) == 0;
                    const uint64_t other = __shfl_xor_sync(FULL, k[r], lstride);
                    // keys are distinct (the index is part of the key), so "other > k" == !(other < k)
                    const bool lt = other < k[r];
                    k[r] = (lt == (up == lower)) ? other : k[r];
                }
            } else {
#pragma unroll
                for (int r = 0; r < NPER; ++r) {
                    const int r2 = r ^ stride;
                    if (r2 > r) {
                        const int q = lane * NPER + r;
                        const bool up = (q & size) == 0;
                        const uint64_t a = k[r], b = k[r2];
                        const bool sw = ((a > b) == up);
                        k[r] = sw ? b : a;
                        k[r2] = sw ? a : b;
                    }
                }
            }
        }
    }
}

// 32-bit keys only: a compare-exchange is one VIMNMX per output (min or max picked by a
// predicate), against ~7 ALU instructions for the 64-bit (key,index) network above.
// One merge stage family: all strides below `size` of a block of `size` elements, direction from
// bit `size` of the position (DESC flips every direction).
template <int NPER, bool DESC>
__device__ __forceinline__ void warp_bitonic_stage_u32(uint32_t (&k)[NPER], const int lane, const int size) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
        if (stride >= NPER) {
            const int lstride = stride / NPER;
            const bool lower = (lane & lstride) == 0;
            const bool up = (((lane * NPER) & size) == 0) != DESC;
            const bool keep_min = (up == lower);
#pragma unroll
            for (int r = 0; r < NPER; ++r) {
                const uint32_t other = __shfl_xor_sync(FULL, k[r], lstride);
                k[r] = keep_min ? min(k[r], other) : max(k[r], other);
            }
        } else {
#pragma unroll
            for (int r = 0; r < NPER; ++r) {
                const int r2 = r ^ stride;
                if (r2 > r) {
                    const int q = lane * NPER + r;
                    const bool up = ((q & size) == 0) != DESC;
                    const uint32_t a = k[r], b = k[r2];
                    k[r] = up ? min(a, b) : max(a, b);
                    k[r2] = up ? max(a, b) : min(a, b);
                }
            }
        }
    }
}

template <int NPER, bool DESC = false>
__device__ __forceinline__ void warp_bitonic_sort_u32(uint32_t (&k)[NPER], const int lane) {
#pragma unroll
    for (int size = 2; size <= 32 * NPER; size <<= 1) warp_bitonic_stage_u32<NPER, DESC>(k, lane, size);
}

// 64*NPER keys in two register arrays (position = h*32*NPER + lane*NPER + r): the halves are sorted
// in opposite directions, one compare-exchange per register pair makes both halves bitonic with
// lo <= hi, and one merge stage family finishes each half.  (A single 64-register array makes
// nvcc keep the keys in local memory.)
template <int NPER>
__device__ __forceinline__ void warp_bitonic_sort2_u32(uint32_t (&lo)[NPER], uint32_t (&hi)[NPER], const int lane) {
    warp_bitonic_sort_u32<NPER, false>(lo, lane);
    warp_bitonic_sort_u32<NPER, true>(hi, lane);
#pragma unroll
    for (int r = 0; r < NPER; ++r) {
        const uint32_t a = lo[r], b = hi[r];
        lo[r] = min(a, b);
        hi[r] = max(a, b);
    }
    // strides 16*NPER .. 1 inside each half, every direction "up" (bit 32*NPER of lane*NPER+r is clear)
    warp_bitonic_stage_u32<NPER, false>(lo, lane, 32 * NPER);
    warp_bitonic_stage_u32<NPER, false>(hi, lane, 32 * NPER);
}

}  // namespace vdet
