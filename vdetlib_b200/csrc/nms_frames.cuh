// nms_frames.cu -- per-frame greedy NMS on class-shared boxes (K1 + K2-bitmask of SURVEY 2.1).
//
// Replaces utils/nms.pyx:17-125 (nms / vid_nms inner loops) and the per-class
// apply_vid_nms passes of vdet/video_det.py:51-61.
//
// One CTA owns one frame at a time (persistent grid-stride over frames):
//   A. the frame's boxes are staged in shared memory (float4, coalesced) with their areas,
//      and -- when the score block is box-major [n, C] -- the frame's scores are transposed
//      into shared memory once, coalesced;
//   B. the CTA builds the frame's suppression bit matrix in ORIGINAL index space,
//      bit (i,j) = (IoU_f32(i,j) >= T).  Geometry is class independent (a det proto has one
//      bbox and C class scores, utils/protocol.py:77-110), so the 30 classes share it;
//   C. each warp takes classes round-robin: it sorts the class's scores (descending, ties by
//      ascending row) with a register-resident warp bitonic network, then walks the order
//      once: candidate i is kept iff its bit in the warp's `removed` set is clear, and a
//      kept candidate ORs its mask row into the set (one word per lane).
// The result is exactly the keep list of nms.pyx:43-66 for every (frame, class).
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "warp_sort.cuh"

// Build switches (tools/build_variant.py builds the other setting as a separate library for A/B timing):
//   VDET_TILE_PACKED   the 32x32 bit-matrix tile evaluates two columns per step on packed float32 pairs
//                      (FADD2 / FMUL2); 0 = the scalar tile of round 1, kept for A/B timing.
#ifndef VDET_TILE_PACKED
#define VDET_TILE_PACKED 1
#endif
//   VDET_NMS_CTAS_PER_SM   resident CTAs per SM the register budget of the <= 512-key variants is compiled for
//                      (4 = 64 registers per thread; 5 = 48 registers with ~350 bytes of spills per thread:
//                      measured 0.375 ms against 0.325 on config 2)
#ifndef VDET_NMS_CTAS_PER_SM
#define VDET_NMS_CTAS_PER_SM 4
#endif

namespace vdet {

// The rank search reads the sorted keys through NON-volatile asm loads, which the compiler may interleave across
// the elements of a lane (a volatile load keeps them in program order: one probe chain at a time; measured
// 0.3497 ms against 0.3528 on config 2).  Ordering after the key stores comes from a data dependence: every probe
// address contains a token that is defined after the __syncwarp().
__device__ __forceinline__ uint32_t order_token() {
    uint32_t t;
    asm volatile("mov.u32 %0, 0;" : "=r"(t) : : "memory");
    return t;
}
__device__ __forceinline__ uint32_t lds_u32_search(const uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Lower bound of GQ keys at once in a sorted array of 32*NP keys stored at SKEWED word addresses
// a(q) = q + (q >> 5) from byte address base_b (the stored range ends at acap_b = address of a(cap), cap a
// multiple of 32; slots [n, cap) hold the padding key 0xffffffff, which is never < key).  The search walks the
// skewed byte addresses directly: while the steps are multiples of 32, pos is one too and
// a(pos + step - 1) = a(pos) + step + step/32 - 2; the last five steps stay inside one 32-block, where a() is
// linear.  Only the big steps can leave the stored range and need a bound test; a big step is never taken onto
// cap itself, so the result is min(lower bound, cap - 1) -- the lower bound itself for a key that is in the
// array.  The GQ probe chains are independent and branch free, so they interleave.  Returns the byte address of
// that skewed slot per key.
template <int NP, int GQ>
__device__ __forceinline__ void lb_search(const uint32_t base_b, const uint32_t acap_b, const uint32_t (&key)[GQ],
                                          uint32_t (&ap)[GQ]) {
#pragma unroll
    for (int q = 0; q < GQ; ++q) ap[q] = base_b;
#pragma unroll
    for (int step = 16 * NP; step >= 32; step >>= 1) {
#pragma unroll
        for (int q = 0; q < GQ; ++q) {
            const uint32_t a = ap[q] + 4u * (uint32_t)(step + (step >> 5) - 2);
            const bool inside = a + 8u < acap_b;                 // the step would land INSIDE the stored range
            const uint32_t v = lds_u32_search(inside ? a : base_b);
            if (inside && v < key[q]) ap[q] += 4u * (uint32_t)(step + (step >> 5));
        }
    }
#pragma unroll
    for (int step = (NP > 1 ? 16 : 16 * NP); step > 0; step >>= 1) {
#pragma unroll
        for (int q = 0; q < GQ; ++q)
            if (lds_u32_search(ap[q] + 4u * (uint32_t)(step - 1)) < key[q]) ap[q] += 4u * (uint32_t)step;
    }
}
// word offset of a skewed address -> position: a(q) = q + q/32  =>  q = a - a/33
__device__ __forceinline__ uint32_t unskew(const uint32_t words) { return words - words / 33u; }


// CTA shapes of the <= 1024-box variants: 8 warps, four CTAs per SM (64 registers per thread), or -- for frames of at
// most 320 boxes whose classes are all staged at once -- 10 warps, three CTAs per SM: 30 classes are three per warp
// with no chunk barrier in between (config 2: 0.294 -> 0.277 ms, profiles/r02_nms_variants.md); nms_plan.h picks.
constexpr int NMS_THREADS = 256;
constexpr int NMS_THREADS_WIDE = 320;
constexpr int NMS_CTAS_WIDE = 3;

struct NmsFramesParams {
    const float* boxes; int box_ld; int box_vec;
    const float* scores; int64_t score_ldr, score_ldc;
    const int32_t* seg_offsets; int n_segs;
    const int32_t* row_ids;
    int n_classes;
    float thresh_f32;
    int32_t* keep_idx; int32_t* keep_cnt; uint8_t* keep_mask;
    int64_t n_rows;
    uint32_t* status;
    int nb;        // padded frame capacity (multiple of 32)
    int stage;     // 1: scores transposed into shared memory
    uint32_t* gmask;   // big-frame variant: per-CTA bit-matrix slots in global memory
    uint16_t* gcnt;    // big-frame variant: per-warp tie counters (global scratch; ties are the cold path)
    int npad;          // big-frame variant: power-of-two sort length >= nb
    int fast_filter;   // 1: division-free threshold filter allowed (2^-20 <= T <= 2)
    float thresh_hi, thresh_lo;   // T(1 +- 2^-21) for that filter
    int cls_chunk;     // classes staged in shared memory at a time (>= n_classes: all at once)
    int so_words;      // per-warp order scratch: (nb/32)*33 words
    int frame_major;   // output layout (VDET_LAYOUT_*)
    // work items: frames [0, split_from) are one item each; every later frame is cut into `nsplit`
    // class ranges (each item rebuilds the frame's bit matrix) so that the last, partially filled
    // round of the persistent grid still occupies every CTA slot
    int split_from, nsplit, n_items;
};

// 32x32 bit-matrix transpose across the warp (lane = row): five block-swap steps, each one
// shuffle + shift + bit-select.  out[L] bit r == in[r] bit L.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, const int lane) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const uint32_t m = (s == 16) ? 0x0000ffffu : (s == 8) ? 0x00ff00ffu : (s == 4) ? 0x0f0f0f0fu
                         : (s == 2) ? 0x33333333u : 0x55555555u;     // positions with bit s clear
        const uint32_t y = __shfl_xor_sync(FULL, x, s);
        const bool upper = (lane & s) == 0;
        const uint32_t t = upper ? (y << s) : (y >> s);
        const uint32_t keep = upper ? m : ~m;
        x = (x & keep) | (t & ~keep);
    }
    return x;
}

// One 32x32 tile of the suppression bit matrix: lane = row i (box in registers), the 32 columns
// of block `cb` are broadcast from shared memory.  Returns this lane's word for row i and, in
// `tword`, the transposed word (row cb*32+lane, columns = this row block) -- IoU is symmetric
// bit for bit (max/min/add commute), so only tiles with cb >= rb are evaluated.
//
// FAST: the threshold test avoids the IEEE division.  fl(inter/uni) >= T holds iff
// inter/uni >= m for a midpoint m in [T(1-2^-24), T].  With Thi = fl(T(1+2^-21)) and
// Tlo = fl(T(1-2^-21)) (host, any rounding): inter > fl(Thi*uni) >= T*uni(1+2^-21)(1-2^-24)^2
// > T*uni proves the test true; inter < fl(Tlo*uni) <= T*uni(1-2^-21)(1+2^-24)^2 < T(1-2^-24)*uni
// proves it false.  Pairs in between (or with uni <= 0 / NaN) set `uncertain`, and the caller
// redoes the tile with the exact division (FAST = false).  Results are identical to the exact
// path by construction.  The host enables FAST only for 2^-20 <= T <= 2, and unions outside
// (1e-30, 1e30) are "uncertain", so T*uni can neither overflow nor go subnormal.
// SANE (CTA-uniform: every box of the frame passes box_sane): unions lie in [2^-48, 2^43] and are
// never zero (uni >= the larger area, rounding is monotone), so the range and zero tests go.
template <bool FAST, bool SANE>
__device__ __forceinline__ uint32_t mask_tile(const float4 bi, const float ai, const float4* __restrict__ sbox,
                                              const float* __restrict__ sarea, const int cb, const float T,
                                              const float Thi, const float Tlo,
                                              const int lane, uint32_t& tword, bool& zero, bool& uncertain) {
    uint32_t word = 0;
    bool z = false, unc = false;
    if (FAST && VDET_TILE_PACKED) {
        // two columns per step on packed float32 pairs (FADD2 / FMUL2, common.cuh): same bits, ~30 % fewer
        // instructions per pair
        const f32x2 ai2 = pk2(ai, ai), Thi2 = pk2(Thi, Thi), Tlo2 = pk2(Tlo, Tlo);
#pragma unroll
        for (int jj = 0; jj < 32; jj += 2) {
            const int j = cb * 32 + jj;
            const float2 aj = *reinterpret_cast<const float2*>(sarea + j);
            f32x2 inter2, uni2, nuni2;
            inter_union_f32x2(bi, ai2, sbox[j], sbox[j + 1], pk2(aj.x, aj.y), inter2, uni2, nuni2);
            float i0, i1, hi0, hi1, lo0, lo1;
            upk2(inter2, i0, i1);
            upk2(mul2(Thi2, uni2), hi0, hi1);
            upk2(mul2(Tlo2, uni2), lo0, lo1);
            const bool sup0 = i0 > hi0, sup1 = i1 > hi1;
            unc |= (!sup0 && !(i0 < lo0)) || (!sup1 && !(i1 < lo1));
            if (!SANE) {
                float u0, u1;
                upk2(uni2, u0, u1);
                unc |= !(u0 > 1e-30f && u0 < 1e30f) || !(u1 > 1e-30f && u1 < 1e30f);
                z |= (u0 == 0.0f) || (u1 == 0.0f);
            }
            if (sup0) word |= (1u << jj);
            if (sup1) word |= (2u << jj);
        }
    } else {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            const int j = cb * 32 + jj;
            const float4 bj = sbox[j];
            const float aj = sarea[j];
            float inter, uni;
            inter_union_f32(bi, ai, bj, aj, inter, uni);
            bool sup;
            if (FAST) {
                sup = inter > __fmul_rn(Thi, uni);
                unc |= !sup && !(inter < __fmul_rn(Tlo, uni));
                if (!SANE) unc |= !(uni > 1e-30f && uni < 1e30f);
            } else {
                sup = iou_ge(inter, uni, T);
            }
            if (!SANE) z |= (uni == 0.0f);
            if (sup) word |= (1u << jj);
        }
    }
    tword = warp_transpose32(word, lane);
    zero = z;
    uncertain = unc;
    return word;
}

// The tile with the cheapest admissible test; an uncertain pair anywhere redoes it exactly.
template <bool SANE>
__device__ __forceinline__ uint32_t mask_tile_auto(const bool fast, const float4 bi, const float ai,
                                                   const float4* __restrict__ sbox, const float* __restrict__ sarea,
                                                   const int cb, const float T, const float Thi, const float Tlo,
                                                   const int lane, uint32_t& tword, bool& zero) {
    bool unc;
    if (fast) {
        const uint32_t word = mask_tile<true, SANE>(bi, ai, sbox, sarea, cb, T, Thi, Tlo, lane, tword, zero, unc);
        if (!__any_sync(FULL, unc)) return word;
    }
    return mask_tile<false, false>(bi, ai, sbox, sarea, cb, T, Thi, Tlo, lane, tword, zero, unc);
}

// Exact ZeroDivisionError test of nms.pyx:64 (cold path, only for frames that contain a
// zero-union pair at all): the pair (ci, j) is visited by the reference iff j comes later in the
// score order and is not yet removed when ci is kept.  `so` is the warp's striped order scratch.
static __device__ __noinline__ void zero_division_check(const uint32_t* so, int ngroups, const float4* sbox,
                                                 const float* sarea, uint32_t rem, uint32_t ci, int pos, int n,
                                                 int lane, uint32_t* status) {
    const float4 bi = sbox[ci];
    const float ai = sarea[ci];
    bool zd = false;
    for (int g2 = 0; g2 < ngroups; ++g2) {
        const int pos2 = g2 * 32 + lane;
        const uint32_t j = pos2 < n ? so[g2 * 33 + lane] : 0u;
        const uint32_t wj = __shfl_sync(FULL, rem, (int)((j >> 5) & 31));
        if (pos2 > pos && pos2 < n && !((wj >> (j & 31)) & 1u)) {
            float inter, uni;
            inter_union_f32(bi, ai, sbox[j], sarea[j], inter, uni);
            zd |= (uni == 0.0f);
        }
    }
    if (__any_sync(FULL, zd) && lane == 0) atomicOr(status, VDET_STATUS_ZERO_DIVISION);
}

// STAGE: the frame's scores are transposed into shared memory (as sort keys) in phase A; a
// compile-time switch, so the per-element key fetch carries no trace of the other path.
// NPB > 0: the class's keys are sorted as TWO arrays, A = the first 32*NPER elements and B = the next 32*NPB
// (NPB = NPER/4), and an element's rank is the sum of its lower bounds in both.  A 300-box frame then
// sorts 256 + 64 keys (330 compare-exchanges per lane) instead of padding to a 512-key network (720): the
// network was the largest single item of the per-class work, the two extra probe chains cost a
// third of what it saves.
template <int NPER, int NPB, bool STAGE, int THREADS = NMS_THREADS>
__global__ void __launch_bounds__(THREADS, (NPER <= 16 ? (THREADS > NMS_THREADS ? NMS_CTAS_WIDE : VDET_NMS_CTAS_PER_SM) : 1))
nms_frames_kernel(const NmsFramesParams p) {
    constexpr int NMS_THREADS = THREADS;                 // (shadows the default CTA shape)
    constexpr int NMS_WARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NB = p.nb;
    const int W = NB >> 5;          // mask words per row (<= 32 in this variant)
    const int WS = W | 1;           // odd row stride: column writes of phase B are conflict free
    const int SST = NB + 1;         // odd class stride of the staged score block
    float4* sbox = reinterpret_cast<float4*>(smem_raw);
    float* sarea = reinterpret_cast<float*>(sbox + NB);
    int32_t* srow = reinterpret_cast<int32_t*>(sarea + NB);
    uint32_t* smask = reinterpret_cast<uint32_t*>(srow + NB);
    uint32_t* sord = smask + (size_t)NB * WS;                       // [NMS_WARPS][so_words] order scratch
    uint32_t* sscore = sord + NMS_WARPS * p.so_words;             // staged scores, already as sort keys
    __shared__ int s_zero_union;
    __shared__ int s_next_class;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int C = p.n_classes;
    const float T = p.thresh_f32;

    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        int seg = item, c_begin = 0, c_end = C;
        if (item >= p.split_from) {
            const int q = item - p.split_from;
            seg = p.split_from + q / p.nsplit;
            const int part = q - (seg - p.split_from) * p.nsplit;
            c_begin = (int)((int64_t)part * C / p.nsplit);
            c_end = (int)((int64_t)(part + 1) * C / p.nsplit);
        }
        const int off = p.seg_offsets[seg];
        const int n = p.seg_offsets[seg + 1] - off;
        if (n > NB) {   // caller's max_seg_len was wrong: refuse loudly instead of truncating
            if (tid == 0) atomicOr(p.status, 0x80000000u);
            continue;
        }
        // ---- A: stage boxes, areas, original row ids (and scores) ------------------------
        if (tid == 0) s_zero_union = 0;
        if (tid == 0) s_next_class = c_begin;
        bool all_sane = true;
        for (int e = tid; e < NB; e += NMS_THREADS) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            int32_t row = -1;
            if (e < n) {
                row = p.row_ids ? p.row_ids[off + e] : off + e;
                b = load_box(p.boxes, row, p.box_ld, p.box_vec);
            }
            sbox[e] = b;
            sarea[e] = area_f32(b);
            srow[e] = row;
            all_sane &= box_sane(b);
        }
        const bool sane = __syncthreads_and(all_sane) != 0;     // CTA-uniform: cheaper pair test
        // Scores of classes [c0, c1) -> shared memory as sort keys, class-major (sscore[(c - c0) * SST + r]).
        // The whole item is staged at once when it fits (cls_chunk >= classes of the item); otherwise
        // in chunks, which lets four CTAs share an SM instead of three.
        auto stage_scores = [&](const int c0, const int c1) {
            const int CH = c1 - c0;
            const int total = n * CH;
            if (p.row_ids == nullptr && p.score_ldc == 1) {
                // contiguous rows: coalesced reads along the class axis, transposed conflict-free writes
                const float* src = p.scores + (int64_t)off * p.score_ldr + c0;
                int r = tid / CH, c = tid - r * CH;                // one division, then incremental
                const int dr = NMS_THREADS / CH, dc = NMS_THREADS - dr * CH;
                for (int f = tid; f < total; f += NMS_THREADS) {
                    sscore[c * SST + r] = f32_key_desc(__ldg(src + (int64_t)r * p.score_ldr + c));
                    r += dr; c += dc;
                    if (c >= CH) { c -= CH; ++r; }
                }
            } else {
                for (int f = tid; f < total; f += NMS_THREADS) {
                    const int r = f / CH, c = f - r * CH;
                    sscore[c * SST + r] = f32_key_desc(__ldg(p.scores + (int64_t)srow[r] * p.score_ldr + (int64_t)(c0 + c) * p.score_ldc));
                }
            }
        };
        const int chunk = STAGE ? p.cls_chunk : (c_end - c_begin);
        if (STAGE) stage_scores(c_begin, min(c_begin + chunk, c_end));
        // ---- B: suppression bit matrix, original index space, upper-triangular tiles -------
        {
            const int Wn = (n + 31) >> 5;          // blocks actually populated by this frame
            int t = 0;
            for (int rb = 0; rb < Wn; ++rb) {
                for (int cb = rb; cb < Wn; ++cb, ++t) {
                    if ((NMS_WARPS & (NMS_WARPS - 1)) == 0 ? ((t & (NMS_WARPS - 1)) != warp) : (t % NMS_WARPS != warp)) continue;
                    const int i = rb * 32 + lane;
                    const float4 bi = sbox[i];
                    const float ai = sarea[i];
                    uint32_t tword;
                    bool zero;
                    const uint32_t word0 =
                        sane ? mask_tile_auto<true>(p.fast_filter, bi, ai, sbox, sarea, cb, T, p.thresh_hi, p.thresh_lo, lane, tword, zero)
                             : mask_tile_auto<false>(p.fast_filter, bi, ai, sbox, sarea, cb, T, p.thresh_hi, p.thresh_lo, lane, tword, zero);
                    uint32_t word = word0;
                    // columns / rows beyond the frame never suppress and are never visited
                    const int cvalid = n - cb * 32, rvalid = n - rb * 32;
                    if (cvalid < 32) word &= (1u << cvalid) - 1u;
                    if (rvalid < 32) tword &= (1u << rvalid) - 1u;
                    smask[i * WS + cb] = word;
                    if (cb != rb) smask[(cb * 32 + lane) * WS + rb] = tword;
                    // (a spurious flag from padding or the diagonal only enables the exact check)
                    if (__any_sync(FULL, zero) && lane == 0) s_zero_union = 1;
                }
            }
        }
        __syncthreads();
        const bool check_zero = (s_zero_union != 0);
        const int Wn = (n + 31) >> 5;

        // ---- C: per class: order by score + greedy walk (one warp per class) ---------------
        uint32_t* so = sord + warp * p.so_words;          // this warp's order scratch (skewed)
        const int cap = Wn * 32;                          // sorted positions >= cap are padding
        for (int c0 = c_begin; c0 < c_end; c0 += chunk) {
        const int c1 = min(c0 + chunk, c_end);
        if (STAGE && c0 != c_begin) {
            __syncthreads();                              // every warp is done with the previous chunk's keys
            stage_scores(c0, c1);
            if (tid == 0) s_next_class = c0;
            __syncthreads();
        }
        // the warps take the classes of the chunk from a shared counter: the cost of a class follows its keep count
        // (one trip of the walk per kept box), and a round-robin deal leaves the CTA waiting for its unluckiest
        // warp at every chunk barrier (config 2: 0.325 -> 0.294 ms)
        for (;;) {
            int c = 0;
            if (lane == 0) c = atomicAdd(&s_next_class, 1);
            c = __shfl_sync(FULL, c, 0);
            if (c >= c1) break;
            const uint32_t* sc_smem = sscore + (c - c0) * SST;
            const float* sc_glob = p.scores + (int64_t)c * p.score_ldc;
            auto score_key = [&](const int e) -> uint32_t {
                return STAGE ? sc_smem[e] : f32_key_desc(__ldg(sc_glob + (int64_t)srow[e] * p.score_ldr));
            };
            // -- order: so[skew(pos)] = index of the pos-th highest score (ties: lower index first).
            // Fast path: sort the 32-bit score keys alone, then every element finds its rank by
            // binary search in the sorted keys.  Equal keys (tied scores) make ranks ambiguous, so
            // a tie anywhere in the problem takes the 64-bit (key,index) network instead.
            bool ordered = false;
            if (NPB == 0) {
                uint32_t k32[NPER];
#pragma unroll
                for (int r = 0; r < NPER; ++r) {
                    const int e = r * 32 + lane;              // striped: conflict-free / coalesced
                    k32[r] = e < n ? score_key(e) : 0xffffffffu;
                }
                warp_bitonic_sort_u32<NPER>(k32, lane);       // blocked: position lane*NPER + r
                bool tie = false;
#pragma unroll
                for (int r = 0; r + 1 < NPER; ++r) tie |= (k32[r] == k32[r + 1]) && (lane * NPER + r + 1 < n);
                const uint32_t nxt = __shfl_down_sync(FULL, k32[0], 1);
                tie |= (lane < 31) && (k32[NPER - 1] == nxt) && ((lane + 1) * NPER < n);
                if (!__any_sync(FULL, tie)) {
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < NPER; ++r) {
                        const int pp = lane * NPER + r;
                        if (pp < cap) so[pp + (pp >> 5)] = k32[r];
                    }
                    __syncwarp();
                    // rank = lower bound of the element's key among the sorted keys (lb_search), four elements
                    // of a lane at a time; lanes without an element search for the padding key (never stored)
                    uint32_t rank[NPER];          // byte address of the element's (skewed) sorted slot
                    const uint32_t so_b = smem_addr_u32(so) + order_token();
                    const uint32_t acap_b = so_b + 4u * (uint32_t)(cap + (cap >> 5));
                    constexpr int GQ = NPER < 4 ? NPER : 4;
#pragma unroll
                    for (int g4 = 0; g4 < NPER; g4 += GQ) {
                        if (g4 * 32 < cap) {                                   // warp-uniform
                            uint32_t key4[GQ], ap4[GQ];
#pragma unroll
                            for (int q = 0; q < GQ; ++q) {
                                const int e = (g4 + q) * 32 + lane;
                                key4[q] = e < n ? score_key(e) : 0xffffffffu;
                            }
                            lb_search<NPER, GQ>(so_b, acap_b, key4, ap4);
#pragma unroll
                            for (int q = 0; q < GQ; ++q) rank[g4 + q] = ap4[q];
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < NPER; ++r) {
                        const int e = r * 32 + lane;
                        if (e < n) sts_u32(rank[r], (uint32_t)e);
                    }
                    ordered = true;
                }
            } else {
                constexpr int NPBX = NPB > 0 ? NPB : 1;
                constexpr int NA = 32 * NPER;
                uint32_t ka[NPER], kb[NPBX];
#pragma unroll
                for (int r = 0; r < NPER; ++r) {
                    const int e = r * 32 + lane;
                    ka[r] = e < n ? score_key(e) : 0xffffffffu;
                }
#pragma unroll
                for (int r = 0; r < NPBX; ++r) {
                    const int e = NA + r * 32 + lane;
                    kb[r] = e < n ? score_key(e) : 0xffffffffu;
                }
                warp_bitonic_sort_u32<NPER>(ka, lane);        // A: position lane*NPER + r
                warp_bitonic_sort_u32<NPBX>(kb, lane);        // B: position NA + lane*NPB + r
                bool tie = false;
#pragma unroll
                for (int r = 0; r + 1 < NPER; ++r) tie |= (ka[r] == ka[r + 1]) && (lane * NPER + r + 1 < n);
#pragma unroll
                for (int r = 0; r + 1 < NPBX; ++r) tie |= (kb[r] == kb[r + 1]) && (NA + lane * NPBX + r + 1 < n);
                const uint32_t nxa = __shfl_down_sync(FULL, ka[0], 1), nxb = __shfl_down_sync(FULL, kb[0], 1);
                tie |= (lane < 31) && (ka[NPER - 1] == nxa) && ((lane + 1) * NPER < n);
                tie |= (lane < 31) && (kb[NPBX - 1] == nxb) && (NA + (lane + 1) * NPBX < n);
                if (!__any_sync(FULL, tie)) {
                    const int capA = cap < NA ? cap : NA;
                    const int capB = cap > NA ? cap - NA : 0;
                    uint32_t* soB = so + (NA + NPER);                           // behind A's skewed range
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < NPER; ++r) {
                        const int pp = lane * NPER + r;
                        if (pp < capA) so[pp + (pp >> 5)] = ka[r];
                    }
#pragma unroll
                    for (int r = 0; r < NPBX; ++r) {
                        const int pp = lane * NPBX + r;
                        if (pp < capB) soB[pp + (pp >> 5)] = kb[r];
                    }
                    __syncwarp();
                    const uint32_t so_b = smem_addr_u32(so) + order_token();
                    const uint32_t sob_b = so_b + 4u * (uint32_t)(NA + NPER);
                    const uint32_t acapA_b = so_b + 4u * (uint32_t)(capA + (capA >> 5));
                    const uint32_t acapB_b = sob_b + 4u * (uint32_t)(capB + (capB >> 5));
                    uint32_t rank[NPER + NPBX];   // byte address of the element's slot in the merged order
                    bool xtie = false;            // equal keys across the two arrays
                    constexpr int GA = NPER < 4 ? NPER : 4, GB = NPBX < 4 ? NPBX : 4;
                    // lb_search returns min(lower bound, cap - 1) (enough for a key that IS in the array); a key of
                    // the OTHER array can exceed every stored key, so the slot it found is probed once more:
                    // smaller -> the bound is one further; equal -> a tie across the arrays
                    auto merged_slot = [&](const uint32_t apA, const uint32_t apB, const uint32_t foreign,
                                           const uint32_t key, const bool live) -> uint32_t {
                        const uint32_t r = unskew((apA - so_b) >> 2) + unskew((apB - sob_b) >> 2) + (foreign < key ? 1u : 0u);
                        xtie |= live && (foreign == key);
                        return so_b + 4u * (r + (r >> 5));
                    };
#pragma unroll
                    for (int g4 = 0; g4 < NPER; g4 += GA) {                     // elements of A
                        if (g4 * 32 < cap) {
                            uint32_t key4[GA], apA[GA], apB[GA];
#pragma unroll
                            for (int q = 0; q < GA; ++q) {
                                const int e = (g4 + q) * 32 + lane;
                                key4[q] = e < n ? score_key(e) : 0xffffffffu;
                                apB[q] = sob_b;
                            }
                            lb_search<NPER, GA>(so_b, acapA_b, key4, apA);
                            if (capB > 0) lb_search<NPBX, GA>(sob_b, acapB_b, key4, apB);
#pragma unroll
                            for (int q = 0; q < GA; ++q) {
                                const uint32_t other = capB > 0 ? lds_u32_search(apB[q]) : 0xffffffffu;
                                rank[g4 + q] = merged_slot(apA[q], apB[q], other, key4[q], (g4 + q) * 32 + lane < n);
                            }
                        }
                    }
#pragma unroll
                    for (int g4 = 0; g4 < NPBX; g4 += GB) {                     // elements of B
                        if (NA + g4 * 32 < cap) {
                            uint32_t key4[GB], apA[GB], apB[GB];
#pragma unroll
                            for (int q = 0; q < GB; ++q) {
                                const int e = NA + (g4 + q) * 32 + lane;
                                key4[q] = e < n ? score_key(e) : 0xffffffffu;
                            }
                            lb_search<NPER, GB>(so_b, acapA_b, key4, apA);
                            lb_search<NPBX, GB>(sob_b, acapB_b, key4, apB);
#pragma unroll
                            for (int q = 0; q < GB; ++q) {
                                const uint32_t other = lds_u32_search(apA[q]);
                                rank[NPER + g4 + q] = merged_slot(apA[q], apB[q], other, key4[q], NA + (g4 + q) * 32 + lane < n);
                            }
                        }
                    }
                    __syncwarp();
                    if (!__any_sync(FULL, xtie)) {
#pragma unroll
                        for (int r = 0; r < NPER; ++r) {
                            const int e = r * 32 + lane;
                            if (e < n) sts_u32(rank[r], (uint32_t)e);
                        }
#pragma unroll
                        for (int r = 0; r < NPBX; ++r) {
                            const int e = NA + r * 32 + lane;
                            if (e < n) sts_u32(rank[NPER + r], (uint32_t)e);
                        }
                        ordered = true;
                    }
                }
            }
            if (!ordered) {
                constexpr int NP2 = NPB > 0 ? 2 * NPER : NPER;          // power of two >= NPER + NPB
                uint64_t key[NP2];
#pragma unroll
                for (int r = 0; r < NP2; ++r) {
                    const int e = r * 32 + lane;
                    key[r] = e < n ? (((uint64_t)score_key(e) << 32) | (uint32_t)e) : ~0ull;
                }
                warp_bitonic_sort<NP2>(key, lane);
                __syncwarp();
#pragma unroll
                for (int r = 0; r < NP2; ++r) {
                    const int pp = lane * NP2 + r;
                    if (pp < cap) so[pp + (pp >> 5)] = (uint32_t)key[r];
                }
            }
            __syncwarp();

            // -- greedy walk, 32 candidates per step: every lane tests its own candidate against the
            // removed set, a ballot gives the alive ones; the lowest alive lane is by construction
            // the next kept box, its mask row is OR-ed in and kills later lanes of the same group.
            // The inner loop is the kernel's hottest code (one trip per kept box): everything that does
            // not depend on the kept box is hoisted -- the lane's column of the mask (mrow), the
            // word / bit of the lane's own candidate (isrc, ibit) -- and the exact ZeroDivisionError
            // test lives in a separate copy of the loop that only frames with a zero-union pair take.
            uint32_t rem = 0;        // lane w: word w of the removed set
            int cnt = 0;
            const int64_t blk = p.frame_major ? ((int64_t)off * C + (int64_t)c * n) : ((int64_t)c * p.n_rows + off);
            int32_t* out_idx = p.keep_idx + blk;
            uint8_t* out_m = p.keep_mask ? p.keep_mask + blk : nullptr;
            const unsigned lt = lanemask_lt();
            const uint32_t mrow = smem_addr_u32(smask + (lane < Wn ? lane : 0));   // lanes beyond the row read word 0 ...
            const uint32_t lmask = lane < Wn ? 0xffffffffu : 0u;                   // ... and drop it
            const uint32_t row_bytes = (uint32_t)WS * 4u;
#pragma unroll 1
            for (int g = 0; g < Wn; ++g) {
                const bool valid = (g * 32 + lane) < n;
                const uint32_t i = valid ? so[g * 33 + lane] : 0u;
                const int isrc = (int)(i >> 5);
                const uint32_t ibit = 1u << (i & 31);
                const uint32_t w = __shfl_sync(FULL, rem, isrc);
                unsigned alive = __ballot_sync(FULL, valid && !(w & ibit));
                unsigned kgrp = 0;   // lanes of this group whose candidate is kept (warp-uniform)
                if (!check_zero) {
                    while (alive) {
                        const unsigned below = alive - 1u;                        // lowest alive lane = next kept box
                        const uint32_t ci = __shfl_sync(FULL, i, __ffs(alive) - 1);
                        const uint32_t roww = lds_u32(mrow + ci * row_bytes) & lmask;
                        rem |= roww;
                        kgrp |= alive & ~below;
                        const uint32_t wv = __shfl_sync(FULL, roww, isrc);
                        alive = alive & below & ~__ballot_sync(FULL, (wv & ibit) != 0u);
                    }
                } else {
                    while (alive) {
                        const int l = __ffs(alive) - 1;
                        const uint32_t ci = __shfl_sync(FULL, i, l);
                        zero_division_check(so, Wn, sbox, sarea, rem, ci, g * 32 + l, n, lane, p.status);
                        const uint32_t roww = lds_u32(mrow + ci * row_bytes) & lmask;
                        rem |= roww;
                        kgrp |= (1u << l);
                        const uint32_t wv = __shfl_sync(FULL, roww, isrc);
                        alive &= ~(__ballot_sync(FULL, (wv & ibit) != 0u) | (1u << l));
                    }
                }
                // outputs of this group: kept rows in walk (= descending score) order, byte mask
                const bool mine = (kgrp >> lane) & 1u;
                if (mine) out_idx[cnt + __popc(kgrp & lt)] = srow[i];
                if (out_m && valid) out_m[i] = (uint8_t)mine;
                cnt += __popc(kgrp);
            }
#pragma unroll 4
            for (int g = cnt >> 5; g < Wn; ++g) {                 // -1 padding of the frame's unused slots
                const int e = g * 32 + lane;
                if (e >= cnt && e < n) out_idx[e] = -1;
            }
            if (lane == 0) p.keep_cnt[p.frame_major ? ((int64_t)seg * C + c) : ((int64_t)c * p.n_segs + seg)] = cnt;
            __syncwarp();
        }
        }                  // class chunks
        __syncthreads();   // smem is reused by the next frame
    }
}

template <int NPER, int NPB, bool STAGE, int THREADS = NMS_THREADS>
static int launch_nms_frames_t(const NmsFramesParams& p, size_t smem, int grid, cudaStream_t st) {
    if (smem > max_dynamic_smem(nms_frames_kernel<NPER, NPB, STAGE, THREADS>)) {
        set_error("nms_frames: %zu bytes of shared memory needed", smem);
        return VDET_ERR_UNSUPPORTED;
    }
    VDET_CUDA(allow_dynamic_smem(nms_frames_kernel<NPER, NPB, STAGE, THREADS>, smem));
    nms_frames_kernel<NPER, NPB, STAGE, THREADS><<<grid, THREADS, smem, st>>>(p);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

// defined in nms_frames_split.cu / nms_frames_big.cu (separate translation units: they compile in parallel)
int launch_nms_frames_split(int nper, int npb, int threads, const NmsFramesParams& p, size_t smem, int grid, cudaStream_t st);
int launch_nms_frames_big(const NmsFramesParams& p, int grid, cudaStream_t st);
size_t nms_frames_big_ws_bytes(int grid, int nb);

}  // namespace vdet
