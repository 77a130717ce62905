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
// other/mine compare-exchange with a runtime direction: take `other` iff (other < mine) != flip.
// One ISETP (the direction predicate rides on its XOR input) and one SEL.
__device__ __forceinline__ uint32_t pick_u32(const uint32_t other, const uint32_t mine, const uint32_t flip) {
    uint32_t out;
    asm("{\n\t.reg .pred pf, pt;\n\tsetp.ne.u32 pf, %3, 0;\n\tsetp.lt.u32.xor pt, %1, %2, pf;\n\t"
        "selp.u32 %0, %1, %2, pt;\n\t}" : "=r"(out) : "r"(other), "r"(mine), "r"(flip));
    return out;
}

// in-lane compare-exchange with a runtime direction: (a, b) -> (min, max) when flip == 0, (max, min) otherwise.
__device__ __forceinline__ void cswap_u32(uint32_t& a, uint32_t& b, const uint32_t flip) {
    uint32_t lo, hi;
    asm("{\n\t.reg .pred pf, pt;\n\tsetp.ne.u32 pf, %4, 0;\n\tsetp.lt.u32.xor pt, %3, %2, pf;\n\t"
        "selp.u32 %0, %3, %2, pt;\n\tselp.u32 %1, %2, %3, pt;\n\t}"
        : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(flip));
    a = lo;
    b = hi;
}

template <int NPER, bool DESC>
__device__ __forceinline__ void warp_bitonic_stage_u32(uint32_t (&k)[NPER], const int lane, const int size) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
        if (stride >= NPER) {
            // partner in another lane; which side keeps the minimum depends on the lane.  Written as
            // compare + select (ISETP with the lane predicate folded in, SEL): "min or max picked by a
            // runtime predicate" costs two VIMNMX and a SEL.  Equal keys: either choice is the same value.
            const int lstride = stride / NPER;
            const bool lower = (lane & lstride) == 0;
            // the last level (size == all keys) has one direction for the whole warp
            const bool up = (size >= 32 * NPER) ? !DESC : ((((lane * NPER) & size) == 0) != DESC);
            const uint32_t flip = (up == lower) ? 0u : 1u;            // 0: keep the minimum
#pragma unroll
            for (int r = 0; r < NPER; ++r) {
                const uint32_t other = __shfl_xor_sync(FULL, k[r], lstride);
                k[r] = pick_u32(other, k[r], flip);
            }
        } else if (size >= NPER && size < 32 * NPER) {
            // both keys in this lane, direction from the lane: one compare, two selects
            const bool up = (((lane * NPER) & size) == 0) != DESC;
            const uint32_t flip = up ? 0u : 1u;
#pragma unroll
            for (int r = 0; r < NPER; ++r) {
                const int r2 = r ^ stride;
                if (r2 > r) {
                    cswap_u32(k[r], k[r2], flip);
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < NPER; ++r) {
                const int r2 = r ^ stride;
                if (r2 > r) {
                    const bool up = (size >= 32 * NPER) ? !DESC : (((r & size) == 0) != DESC);   // compile-time direction
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
