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
#include <stdlib.h>

#include "common.cuh"
#include "warp_sort.cuh"

// Build switches (tools/build_variant.py builds the other setting as a separate library for A/B timing):
//   VDET_EXP_PURE_LDS  (1 = default since round 2: 0.3497 ms against 0.3528 on config 2) the rank search reads the sorted keys through NON-volatile asm loads, which the compiler may
//                      interleave across the elements of a lane (the volatile form keeps them in program order: one
//                      probe chain at a time).  Ordering after the key stores comes from a data dependence: every
//                      probe address contains a token that is defined after the __syncwarp().
#ifndef VDET_EXP_PURE_LDS
#define VDET_EXP_PURE_LDS 1
#endif
//   VDET_TILE_PACKED   the 32x32 bit-matrix tile evaluates two columns per step on packed float32 pairs
//                      (FADD2 / FMUL2); 0 = the scalar tile of round 1, kept for A/B timing.
#ifndef VDET_TILE_PACKED
#define VDET_TILE_PACKED 1
#endif

namespace vdet {

#if VDET_EXP_PURE_LDS
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
#else
__device__ __forceinline__ uint32_t order_token() { return 0u; }
__device__ __forceinline__ uint32_t lds_u32_search(const uint32_t addr) { return lds_u32(addr); }
#endif

constexpr int NMS_THREADS = 256;
constexpr int NMS_WARPS = NMS_THREADS / 32;

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
__device__ __noinline__ void zero_division_check(const uint32_t* so, int ngroups, const float4* sbox,
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
template <int NPER, bool STAGE>
__global__ void __launch_bounds__(NMS_THREADS, (NPER <= 16 ? 4 : 1)) nms_frames_kernel(const NmsFramesParams p) {
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
                    if ((t & (NMS_WARPS - 1)) != warp) continue;
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
            __syncthreads();
        }
        for (int c = c0 + warp; c < c1; c += NMS_WARPS) {
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
            {
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
                    // Rank = lower bound of the element's key among the sorted keys, searched directly in
                    // SKEWED addresses a(q) = q + (q >> 5) so that no step pays for the skew: while the
                    // steps are multiples of 32, pos is one too and a(pos + step - 1) = a(pos) + step +
                    // step/32 - 2; the last five steps stay inside one 32-block, where a() is linear.
                    // Slots [n, cap) hold the padding key 0xffffffff (never < key), so only the big
                    // steps can leave the stored range [0, cap) and need a bound test; a big step is
                    // never taken onto cap itself (slot cap-1 holds the padding or the largest key).
                    // Addresses are 32-bit shared-memory BYTE addresses (one LDS with an immediate offset
                    // per probe instead of index arithmetic + scaling).
                    uint32_t rank[NPER];          // byte address of the element's (skewed) sorted slot
                    const uint32_t so_b = smem_addr_u32(so) + order_token();
                    const uint32_t acap_b = so_b + 4u * (uint32_t)(cap + (cap >> 5));
#if VDET_EXP_PURE_LDS
                    // four probe chains at a time, branch-free inside a group so that they interleave; lanes
                    // without an element search for the padding key (their result is never stored)
                    constexpr int GQ = NPER < 4 ? NPER : 4;
#pragma unroll
                    for (int g4 = 0; g4 < NPER; g4 += GQ) {
                        if (g4 * 32 < cap) {                                   // warp-uniform
                            uint32_t key4[GQ], ap4[GQ];
#pragma unroll
                            for (int q = 0; q < GQ; ++q) {
                                const int e = (g4 + q) * 32 + lane;
                                key4[q] = e < n ? score_key(e) : 0xffffffffu;
                                ap4[q] = so_b;
                            }
#pragma unroll
                            for (int step = 16 * NPER; step >= 32; step >>= 1) {
#pragma unroll
                                for (int q = 0; q < GQ; ++q) {
                                    const uint32_t a = ap4[q] + 4u * (uint32_t)(step + (step >> 5) - 2);
                                    const uint32_t v = lds_u32_search(a < acap_b ? a : so_b);
                                    if (a < acap_b && v < key4[q]) ap4[q] += 4u * (uint32_t)(step + (step >> 5));
                                }
                            }
#pragma unroll
                            for (int step = (NPER > 1 ? 16 : 16 * NPER); step > 0; step >>= 1) {
#pragma unroll
                                for (int q = 0; q < GQ; ++q)
                                    if (lds_u32_search(ap4[q] + 4u * (uint32_t)(step - 1)) < key4[q]) ap4[q] += 4u * (uint32_t)step;
                            }
#pragma unroll
                            for (int q = 0; q < GQ; ++q) rank[g4 + q] = ap4[q];
                        }
                    }
#else
#pragma unroll
                    for (int r = 0; r < NPER; ++r) {
                        const int e = r * 32 + lane;
                        uint32_t ap = so_b;
                        if (e < n) {
                            const uint32_t key = score_key(e);
#pragma unroll
                            for (int step = 16 * NPER; step >= 32; step >>= 1) {
                                const uint32_t a = ap + 4u * (uint32_t)(step + (step >> 5) - 2);
                                if (a < acap_b && lds_u32_search(a) < key) ap += 4u * (uint32_t)(step + (step >> 5));
                            }
#pragma unroll
                            for (int step = (NPER > 1 ? 16 : 16 * NPER); step > 0; step >>= 1)
                                if (lds_u32_search(ap + 4u * (uint32_t)(step - 1)) < key) ap += 4u * (uint32_t)step;
                        }
                        rank[r] = ap;
                    }
#endif
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < NPER; ++r) {
                        const int e = r * 32 + lane;
                        if (e < n) sts_u32(rank[r], (uint32_t)e);
                    }
                    ordered = true;
                }
            }
            if (!ordered) {
                uint64_t key[NPER];
#pragma unroll
                for (int r = 0; r < NPER; ++r) {
                    const int e = r * 32 + lane;
                    key[r] = e < n ? (((uint64_t)score_key(e) << 32) | (uint32_t)e) : ~0ull;
                }
                warp_bitonic_sort<NPER>(key, lane);
                __syncwarp();
#pragma unroll
                for (int r = 0; r < NPER; ++r) {
                    const int pp = lane * NPER + r;
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

// Work items for a persistent grid of `slots` CTAs over n_segs frames (see NmsFramesParams).
static void plan_items(NmsFramesParams& p, int slots) {
    p.split_from = p.n_segs; p.nsplit = 1; p.n_items = p.n_segs;
    if (slots <= 0 || p.n_classes < 2) return;
    // Only launches that cannot fill the grid once are split (a single image with 30 classes, a short
    // clip): measured on config 2, splitting the LAST round of a multi-round launch does not pay --
    // the CTAs of a thin last round already run alone on their SMs and finish early.
    if (p.n_segs * 2 > slots) return;
    const int rem = p.n_segs;
    int ns = slots / rem;
    if (ns > p.n_classes) ns = p.n_classes;
    if (ns > 8) ns = 8;
    if (ns < 2) return;
    p.split_from = p.n_segs - rem;
    p.nsplit = ns;
    p.n_items = p.split_from + rem * ns;
}

static size_t nms_smem_bytes(int nb, int nper, int n_classes, bool stage) {
    const int W = nb / 32, WS = W | 1;
    size_t b = (size_t)nb * (sizeof(float4) + sizeof(float) + sizeof(int32_t));
    b += (size_t)nb * WS * sizeof(uint32_t);
    b += (size_t)NMS_WARPS * (nb / 32) * 33 * sizeof(uint32_t);
    if (stage) b += (size_t)n_classes * (nb + 1) * sizeof(float);
    return b;
}

template <int NPER, bool STAGE>
static int launch_nms_frames_t(const NmsFramesParams& p, size_t smem, int grid, cudaStream_t st) {
    if (smem > max_dynamic_smem(nms_frames_kernel<NPER, STAGE>)) {
        set_error("nms_frames: %zu bytes of shared memory needed", smem);
        return VDET_ERR_UNSUPPORTED;
    }
    VDET_CUDA(allow_dynamic_smem(nms_frames_kernel<NPER, STAGE>, smem));
    nms_frames_kernel<NPER, STAGE><<<grid, NMS_THREADS, smem, st>>>(p);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

template <int NPER>
static int launch_nms_frames(const NmsFramesParams& p, size_t smem, int grid, cudaStream_t st) {
    return p.stage ? launch_nms_frames_t<NPER, true>(p, smem, grid, st)
                   : launch_nms_frames_t<NPER, false>(p, smem, grid, st);
}


// ==========================================================================================
// Big-frame variant: 1024 < max frame length <= 2048 (BASELINE config 5: 2000 boxes/frame).
// Same three phases; what changes is where things live and how the pairs are enumerated:
//   * the bit matrix (N x N/32 words = 500 KB at N=2000) does not fit in shared memory: every
//     persistent CTA owns a slot in a global scratch buffer (L2 resident), in ORIGINAL index space;
//   * phase A additionally sorts the frame's boxes by x1 (CTA-wide bitonic sort of (key, index) in
//     shared memory) and stages them in that order.  Phase B then sweeps, for every box i of the
//     sorted order, only the later boxes j whose x1 does not exceed x2_i (+ a margin): a pair that
//     does not overlap in x has inter == 0 and can never reach a positive threshold, and in sorted
//     order those pairs are a contiguous tail that is cut off with one compare per 32 candidates.
//     On BASELINE's synthetic frames 22 % of the pairs overlap in x, so the sweep evaluates 4.6x
//     fewer pairs than the N^2/2 of the tiled version.  One warp owns row i, the lanes take 32
//     consecutive j; the few set bits (~22 per row) are scattered to the matrix in original index
//     space, both (i,j) and (j,i), with red.global.or on a slot that was zero-filled first.
//     Frames with an insane box or a threshold <= 0 sweep every j > i (same code, no cut-off);
//   * the per-class order is built per warp in shared memory: register bitonic sort of the 32-bit
//     score keys, rank by binary search (two elements per lane in flight, keys prefetched one trip
//     ahead), and -- only when scores tie -- a stable ordinal among equal keys from match.any
//     ballots over the elements in index order (the radix-sort ranking trick), which reproduces
//     "descending score, then ascending row" without a 64-bit network; each element writes its
//     index straight to its slot of the 16-bit order array;
//   * the removed set takes two words per lane (one 8-byte load per mask row); mask rows are read
//     from the CTA's global slot with ld.global.cg, and the rows of the boxes a step keeps are
//     fetched four at a time so that their L2 latencies overlap.
// ==========================================================================================
constexpr int BIG_MAX = 2048;
constexpr int BIG_THREADS = 512;         // 16 warps: one class each in phase C, 128 registers per thread
constexpr int BIG_WARPS = BIG_THREADS / 32;
constexpr int BIG_NPER = BIG_MAX / 32;   // keys per lane of the register sort
constexpr int BIG_SK_LD = BIG_MAX + 64;  // per-warp key scratch, skewed by one word per 32

// Skew of the per-warp key scratch: the sorted keys leave the register network in blocked
// layout (position half*1024 + lane*32 + r), so an unskewed store would put all 32 lanes on one bank.
__device__ __forceinline__ int skw(const int q) { return q + (q >> 5); }

// CTA-wide bitonic sort (ascending) of npow2 64-bit keys in shared memory.
__device__ __forceinline__ void cta_bitonic_sort_u64(uint64_t* s, const int npow2, const int tid) {
    for (int size = 2; size <= npow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < (npow2 >> 1); t += BIG_THREADS) {
                const int i = 2 * t - (t & (stride - 1));
                const int j = i + stride;
                const bool up = (i & size) == 0;
                const uint64_t a = s[i], b = s[j];
                if ((a > b) == up) { s[i] = b; s[j] = a; }
            }
        }
    }
    __syncthreads();
}

// Removed-set layout of the big kernel: lane l holds mask words 2l (rem0) and 2l+1 (rem1).
__device__ __forceinline__ bool removed_bit(const uint32_t rem0, const uint32_t rem1, const uint32_t j) {
    const int wi = (int)(j >> 5);
    const uint32_t w0 = __shfl_sync(FULL, rem0, (wi >> 1) & 31), w1 = __shfl_sync(FULL, rem1, (wi >> 1) & 31);
    return (((wi & 1) ? w1 : w0) >> (j & 31)) & 1u;
}

__device__ __noinline__ void zero_division_check_big(const float* boxes, int box_ld, int box_vec, uint32_t* status,
                                                     const int32_t* srow, const uint16_t* ord, int n,
                                                     uint32_t rem0, uint32_t rem1, uint32_t ci, int pos, int lane) {
    const float4 bi = load_box(boxes, srow[ci], box_ld, box_vec);
    const float ai = area_f32(bi);
    bool zd = false;
    for (int base = pos + 1; base < n; base += 32) {             // warp-uniform trip count
        const int k2 = base + lane;
        const bool act = k2 < n;
        const uint32_t j = act ? ord[k2] : 0u;
        const bool gone = removed_bit(rem0, rem1, j);
        if (act && !gone) {
            const float4 bj = load_box(boxes, srow[j], box_ld, box_vec);
            float inter, uni;
            inter_union_f32(bi, ai, bj, area_f32(bj), inter, uni);
            zd |= (uni == 0.0f);
        }
    }
    if (__any_sync(FULL, zd) && lane == 0) atomicOr(status, VDET_STATUS_ZERO_DIVISION);
}

__global__ void __launch_bounds__(BIG_THREADS, 1) nms_frames_big_kernel(const NmsFramesParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NB = p.nb;            // multiple of 256
    const int W = NB >> 5;          // <= 64, multiple of 8
    constexpr int NPAD = BIG_MAX;
    // phase C (per class): srow | sorted keys per warp | order per warp.  Phases A and B use the space behind
    // srow for the x1-sorted boxes, their areas, the permutation and the sort scratch instead.
    int32_t* srow = reinterpret_cast<int32_t*>(smem_raw);                   // original index -> packed row
    uint32_t* skeys = reinterpret_cast<uint32_t*>(srow + NB);               // [BIG_WARPS][BIG_SK_LD] sorted keys
    uint16_t* sord = reinterpret_cast<uint16_t*>(skeys + BIG_WARPS * BIG_SK_LD);   // [BIG_WARPS][NPAD] order
    float4* sbox = reinterpret_cast<float4*>(srow + NB);                    // A/B: boxes in x1-sorted order
    float* sarea = reinterpret_cast<float*>(sbox + NB);
    uint16_t* sperm = reinterpret_cast<uint16_t*>(sarea + NB);              // A/B: sorted position -> original index
    uint64_t* ssort = reinterpret_cast<uint64_t*>(sperm + NB);              // A: (x1 key, index), 16 KB
    __shared__ int s_zero_union;
    uint32_t* gmask = p.gmask + (size_t)blockIdx.x * NB * W;
    uint16_t* gcnt = p.gcnt + ((size_t)blockIdx.x * BIG_WARPS + (threadIdx.x >> 5)) * NPAD;   // tie counters (cold path)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.n_classes;
    const float T = p.thresh_f32;
    uint32_t* sk = skeys + (size_t)warp * BIG_SK_LD;
    uint16_t* ord = sord + (size_t)warp * NPAD;
    uint16_t* ct = gcnt;

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
        if (n > NB) {
            if (tid == 0) atomicOr(p.status, 0x80000000u);
            continue;
        }
        // ---- A: rows, x1 sort, boxes staged in sorted order; the matrix slot is zero-filled ---------
        if (tid == 0) s_zero_union = 0;
        int npow2 = 2;
        while (npow2 < n) npow2 <<= 1;
        bool all_sane = true;
        for (int e = tid; e < NPAD; e += BIG_THREADS) {
            uint64_t key = ~0ull;
            if (e < NB) {
                int32_t row = -1;
                if (e < n) {
                    row = p.row_ids ? p.row_ids[off + e] : off + e;
                    const float4 b = load_box(p.boxes, row, p.box_ld, p.box_vec);
                    all_sane &= box_sane(b);
                    key = ((uint64_t)f32_key_asc(b.x) << 32) | (uint32_t)e;
                }
                srow[e] = row;
            }
            if (e < npow2) ssort[e] = key;
        }
        {
            uint4* z = reinterpret_cast<uint4*>(gmask);
            const int nvec = n * (W >> 2);
            const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
            for (int v = tid; v < nvec; v += BIG_THREADS) __stcg(z + v, zero4);
        }
        const bool sane = __syncthreads_and(all_sane) != 0;     // CTA-uniform: cheaper pair test, x cut-off allowed
        cta_bitonic_sort_u64(ssort, npow2, tid);
        for (int q = tid; q < NB; q += BIG_THREADS) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t e = 0;
            if (q < n) {
                e = (uint32_t)ssort[q];
                b = load_box(p.boxes, srow[e], p.box_ld, p.box_vec);
            }
            sbox[q] = b;
            sarea[q] = area_f32(b);
            sperm[q] = (uint16_t)e;
        }
        __syncthreads();
        const int Wn = (n + 31) >> 5;
        // ---- B: bit matrix by an x-sorted sweep, one warp per row of the sorted order ------------------
        {
            const bool cut = sane && (T > 0.0f);              // inter == 0 can only reach a threshold <= 0
            const bool fast = p.fast_filter != 0;
            bool any_zero = false;
            for (int i = warp; i + 1 < n; i += BIG_WARPS) {
                const float4 bi = sbox[i];
                const float ai = sarea[i];
                const uint32_t pi = sperm[i];
                const float lim = __fadd_rn(bi.z, 2.0f);      // x1_j > x2_i + 2  =>  w == 0 for every later j too
                uint32_t* rowi = gmask + (size_t)pi * W;
                const f32x2 ai2 = pk2(ai, ai), Thi2 = pk2(p.thresh_hi, p.thresh_hi), Tlo2 = pk2(p.thresh_lo, p.thresh_lo);
                // two candidates per lane and trip (j and j + 32) on packed float32 pairs (common.cuh)
                for (int j0 = i + 1; j0 < n; j0 += 64) {
                    if (cut && sbox[j0].x > lim) break;       // warp-uniform
                    const int ja = j0 + lane, jb = ja + 32;
                    const bool va = ja < n, vb = jb < n && !(cut && sbox[min(j0 + 32, n - 1)].x > lim);
                    const int jac = va ? ja : n - 1, jbc = jb < n ? jb : n - 1;
                    f32x2 inter2, uni2, nuni2;
                    inter_union_f32x2(bi, ai2, sbox[jac], sbox[jbc], pk2(sarea[jac], sarea[jbc]), inter2, uni2, nuni2);
                    float ia, ib, ua, ub;
                    upk2(inter2, ia, ib);
                    upk2(uni2, ua, ub);
                    bool supa, supb;
                    if (fast) {
                        float ha, hb, la, lb;
                        upk2(mul2(Thi2, uni2), ha, hb);
                        upk2(mul2(Tlo2, uni2), la, lb);
                        supa = ia > ha;
                        supb = ib > hb;
                        bool unc = (va && !supa && !(ia < la)) || (vb && !supb && !(ib < lb));
                        if (!sane) unc |= (va && !(ua > 1e-30f && ua < 1e30f)) || (vb && !(ub > 1e-30f && ub < 1e30f));
                        if (__any_sync(FULL, unc)) {
                            supa = iou_ge(ia, ua, T);
                            supb = iou_ge(ib, ub, T);
                        }
                    } else {
                        supa = iou_ge(ia, ua, T);
                        supb = iou_ge(ib, ub, T);
                    }
                    if (!sane) any_zero |= (va && ua == 0.0f) || (vb && ub == 0.0f);
                    if (supa && va) {
                        const uint32_t pj = sperm[ja];
                        atomicOr(rowi + (pj >> 5), 1u << (pj & 31));
                        atomicOr(gmask + (size_t)pj * W + (pi >> 5), 1u << (pi & 31));
                    }
                    if (supb && vb) {
                        const uint32_t pj = sperm[jb];
                        atomicOr(rowi + (pj >> 5), 1u << (pj & 31));
                        atomicOr(gmask + (size_t)pj * W + (pi >> 5), 1u << (pi & 31));
                    }
                }
            }
            if (__any_sync(FULL, any_zero) && lane == 0) s_zero_union = 1;
        }
        __syncthreads();      // block-scope visibility of this CTA's own global atomics
        const bool check_zero = (s_zero_union != 0);

        for (int c = c_begin + warp; c < c_end; c += BIG_WARPS) {
            const float* sc_glob = p.scores + (int64_t)c * p.score_ldc;
            auto score_key = [&](const int e) -> uint32_t {
                return f32_key_desc(__ldg(sc_glob + (int64_t)srow[e] * p.score_ldr));
            };
            // -- keys: 64 per lane through the register network (ascending key = descending score),
            //    then parked in shared memory for the rank search
            bool tie = false;
            {
                constexpr int H = BIG_NPER / 2;                   // 32 keys per lane and half
                uint32_t klo[H], khi[H];
#pragma unroll
                for (int r = 0; r < H; ++r) {                     // striped: coalesced over the frame's rows
                    const int e = r * 32 + lane;
                    klo[r] = e < n ? score_key(e) : 0xffffffffu;
                    khi[r] = e + 1024 < n ? score_key(e + 1024) : 0xffffffffu;
                }
                warp_bitonic_sort2_u32<H>(klo, khi, lane);        // position = half*1024 + lane*32 + r
#pragma unroll
                for (int r = 0; r + 1 < H; ++r) {
                    tie |= (klo[r] == klo[r + 1]) && (lane * H + r + 1 < n);
                    tie |= (khi[r] == khi[r + 1]) && (1024 + lane * H + r + 1 < n);
                }
                const uint32_t nlo = __shfl_down_sync(FULL, klo[0], 1), nhi = __shfl_down_sync(FULL, khi[0], 1);
                const uint32_t first_hi = __shfl_sync(FULL, khi[0], 0);
                tie |= (lane < 31) && (klo[H - 1] == nlo) && ((lane + 1) * H < n);
                tie |= (lane < 31) && (khi[H - 1] == nhi) && (1024 + (lane + 1) * H < n);
                tie |= (lane == 31) && (klo[H - 1] == first_hi) && (1024 < n);
#pragma unroll
                for (int r = 0; r < H; ++r) {
                    sk[lane * (H + 1) + r] = klo[r];                       // = skw(lane*32 + r)
                    sk[1024 + 32 + lane * (H + 1) + r] = khi[r];           // = skw(1024 + lane*32 + r)
                }
            }
            __syncwarp();
            const bool has_tie = __any_sync(FULL, tie);
            // -- rank of every element = lower bound of its key among the sorted keys; the element's index
            //    goes straight to that slot of the order array
            if (!has_tie) {
                // two elements per lane per trip (independent probe chains), keys one trip ahead
                uint32_t ka = lane < n ? score_key(lane) : 0xffffffffu;
                uint32_t kb = lane + 32 < n ? score_key(lane + 32) : 0xffffffffu;
                for (int base = 0; base < n; base += 64) {
                    const uint32_t key0 = ka, key1 = kb;
                    const int e0 = base + lane, e1 = e0 + 32;
                    ka = e0 + 64 < n ? score_key(e0 + 64) : 0xffffffffu;
                    kb = e1 + 64 < n ? score_key(e1 + 64) : 0xffffffffu;
                    uint32_t pos0 = 0, pos1 = 0;
#pragma unroll
                    for (int step = NPAD >> 1; step > 0; step >>= 1) {
                        const uint32_t q0 = pos0 + step - 1, q1 = pos1 + step - 1;
                        const uint32_t v0 = sk[skw((int)(q0 < (uint32_t)n ? q0 : 0u))];
                        const uint32_t v1 = sk[skw((int)(q1 < (uint32_t)n ? q1 : 0u))];
                        if (q0 < (uint32_t)n && v0 < key0) pos0 += step;
                        if (q1 < (uint32_t)n && v1 < key1) pos1 += step;
                    }
                    if (e0 < n) ord[pos0] = (uint16_t)e0;
                    if (e1 < n) ord[pos1] = (uint16_t)e1;
                }
            } else {
                for (int e = lane; e < n; e += 32) ct[e] = 0;
                __syncwarp();
                for (int base = 0; base < n; base += 32) {
                    const int e = base + lane;
                    const bool act = e < n;
                    const uint32_t key = act ? score_key(e) : 0xffffffffu;
                    uint32_t pos = 0;
                    for (int step = NPAD >> 1; step > 0; step >>= 1) {
                        const uint32_t q = pos + step - 1;
                        if (q < (uint32_t)n && sk[skw((int)q)] < key) pos += step;
                    }
                    // elements arrive in index order: the ordinal among equal keys is the running count
                    // of that key (kept at its lower-bound slot) plus the lanes below me with the same key
                    const unsigned peers = __match_any_sync(FULL, act ? key : (0xfffffff0u ^ (uint32_t)lane));
                    const int leader = __ffs(peers) - 1;
                    uint32_t old = 0;
                    if (act && lane == leader) { old = ct[pos]; ct[pos] = (uint16_t)(old + __popc(peers)); }
                    old = __shfl_sync(FULL, old, leader);
                    pos += old + __popc(peers & lanemask_lt());
                    __syncwarp();
                    if (act) ord[pos] = (uint16_t)e;
                }
            }
            __syncwarp();

            uint32_t rem0 = 0, rem1 = 0;          // lane l: mask words 2l and 2l+1 of the removed set
            int cnt = 0;
            const int64_t blk = p.frame_major ? ((int64_t)off * C + (int64_t)c * n) : ((int64_t)c * p.n_rows + off);
            int32_t* out_idx = p.keep_idx + blk;
            uint8_t* out_m = p.keep_mask ? p.keep_mask + blk : nullptr;
            const unsigned lt = lanemask_lt();
            const bool my_words = 2 * lane < W;
            // Greedy walk, 32 candidates of the score order per step.  The mask rows of the step's still-alive
            // candidates are fetched up to 8 at a time (independent 8-byte loads per lane, one L2 latency per
            // batch), then the batch is resolved in order in registers: a candidate that is still alive at its
            // turn is kept, its row is ORed into the removed set and -- through two shuffles and a ballot -- kills
            // the later candidates of the step it suppresses; a candidate killed earlier in the batch is skipped
            // (its row was fetched for nothing: bandwidth, not latency).  One row load per candidate serves both
            // the in-step resolution and the removed set.
#pragma unroll 1
            for (int g = 0; g < Wn; ++g) {
                const bool valid = (g * 32 + lane) < n;
                const uint32_t i = valid ? ord[g * 32 + lane] : 0u;
                const bool gone = removed_bit(rem0, rem1, i);      // shuffles: every lane takes part, valid or not
                unsigned alive = __ballot_sync(FULL, valid && !gone);
                const int wi = (int)(i >> 5);
                const int src_lane = (wi >> 1) & 31;               // lane holding my candidate's word of a fetched row
                const bool odd = (wi & 1) != 0;
                const uint32_t ibit = 1u << (i & 31);
                unsigned kgrp = 0;
                while (alive) {
                    uint2 r[8];
                    int ls[8];
                    unsigned t = alive;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        ls[q] = t ? (__ffs(t) - 1) : -1;
                        t &= t - 1;
                        const uint32_t ci = __shfl_sync(FULL, i, ls[q] & 31);
                        r[q] = (ls[q] >= 0 && my_words)
                                   ? __ldcg(reinterpret_cast<const uint2*>(gmask + (size_t)ci * W) + lane)
                                   : make_uint2(0u, 0u);
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (ls[q] >= 0 && ((alive >> ls[q]) & 1u)) {             // warp-uniform
                            if (check_zero) {
                                const uint32_t ci = __shfl_sync(FULL, i, ls[q]);
                                zero_division_check_big(p.boxes, p.box_ld, p.box_vec, p.status, srow, ord, n, rem0, rem1, ci,
                                                        g * 32 + ls[q], lane);
                            }
                            kgrp |= 1u << ls[q];
                            rem0 |= r[q].x;
                            rem1 |= r[q].y;
                            const uint32_t w0 = __shfl_sync(FULL, r[q].x, src_lane), w1 = __shfl_sync(FULL, r[q].y, src_lane);
                            const unsigned dead = __ballot_sync(FULL, ((odd ? w1 : w0) & ibit) != 0u);
                            alive &= ~(dead | (1u << ls[q]));
                        }
                    }
                }
                const bool mine = (kgrp >> lane) & 1u;
                if (mine) out_idx[cnt + __popc(kgrp & lt)] = srow[i];
                if (out_m && valid) out_m[i] = (uint8_t)mine;
                cnt += __popc(kgrp);
            }
#pragma unroll 4
            for (int g = cnt >> 5; g < Wn; ++g) {
                const int e = g * 32 + lane;
                if (e >= cnt && e < n) out_idx[e] = -1;
            }
            if (lane == 0) p.keep_cnt[p.frame_major ? ((int64_t)seg * C + c) : ((int64_t)c * p.n_segs + seg)] = cnt;
            __syncwarp();
        }
        __syncthreads();
    }
}

static size_t big_smem_bytes(int nb, int npad) {
    const size_t phase_c = (size_t)BIG_WARPS * (BIG_SK_LD * sizeof(uint32_t) + (size_t)npad * sizeof(uint16_t));
    const size_t phase_ab = (size_t)nb * (sizeof(float4) + sizeof(float) + sizeof(uint16_t)) + (size_t)npad * sizeof(uint64_t);
    return (size_t)nb * sizeof(int32_t) + (phase_c > phase_ab ? phase_c : phase_ab);
}
// global scratch per persistent CTA: the bit-matrix slot, then the tie counters of its warps
static size_t big_ws_bytes(int grid, int nb) {
    return (size_t)grid * ((size_t)nb * (nb / 32) * sizeof(uint32_t) + (size_t)BIG_WARPS * BIG_MAX * sizeof(uint16_t));
}

}  // namespace vdet

using namespace vdet;

extern "C" size_t vdet_nms_frames_workspace_bytes(int max_seg_len, int n_classes, int device) {
    (void)n_classes; (void)device;
    if (max_seg_len <= 1024) return 256;   // register-sort variants keep everything in shared memory
    const size_t nb = ((size_t)max_seg_len + 255) / 256 * 256;
    return big_ws_bytes(sm_count_cached(), (int)nb) + 256;   // upper bound: all SMs
}

extern "C" int vdet_nms_frames_f32(const float* boxes, int box_ld,
                                   const float* scores, int64_t score_ldr, int64_t score_ldc,
                                   const int32_t* seg_offsets, int n_segs, int max_seg_len,
                                   const int32_t* row_ids, int n_classes, double thresh,
                                   int32_t* keep_idx, int32_t* keep_cnt, uint8_t* keep_mask,
                                   int64_t n_rows, int out_layout, uint32_t* status,
                                   void* ws, size_t ws_bytes, void* stream) {
    VDET_REQUIRE(out_layout == VDET_LAYOUT_CLASS_MAJOR || out_layout == VDET_LAYOUT_FRAME_MAJOR, "nms_frames: bad out_layout");
    VDET_REQUIRE(n_segs >= 0 && n_classes >= 1 && n_rows >= 0 && max_seg_len >= 0, "nms_frames: negative size");
    VDET_REQUIRE(box_ld >= 4, "nms_frames: box_ld must be >= 4");
    VDET_REQUIRE(status != nullptr && keep_idx != nullptr && keep_cnt != nullptr, "nms_frames: null output");
    if (n_segs == 0 || n_rows == 0 && max_seg_len == 0) {
        if (n_segs > 0)
            VDET_CUDA(cudaMemsetAsync(keep_cnt, 0, sizeof(int32_t) * (size_t)n_segs * n_classes, (cudaStream_t)stream));
        return VDET_OK;
    }
    if (max_seg_len > BIG_MAX) {
        set_error("nms_frames: max_seg_len %d > %d is not supported by this build", max_seg_len, BIG_MAX);
        return VDET_ERR_UNSUPPORTED;
    }
    NmsFramesParams p;
    p.gmask = nullptr; p.gcnt = nullptr; p.npad = 0; p.cls_chunk = n_classes;
    p.frame_major = (out_layout == VDET_LAYOUT_FRAME_MAJOR) ? 1 : 0;
    {
        const float Tf = thresh_ceil_f32(thresh);
        p.fast_filter = (Tf >= 9.5367431640625e-07f && Tf <= 2.0f) ? 1 : 0;
        p.thresh_hi = (float)((double)Tf * (1.0 + 4.76837158203125e-07));
        p.thresh_lo = (float)((double)Tf * (1.0 - 4.76837158203125e-07));
    }
    p.boxes = boxes; p.box_ld = box_ld;
    p.box_vec = (box_ld == 4) && ((reinterpret_cast<uintptr_t>(boxes) & 15) == 0);
    p.scores = scores; p.score_ldr = score_ldr; p.score_ldc = score_ldc;
    p.seg_offsets = seg_offsets; p.n_segs = n_segs; p.row_ids = row_ids;
    p.n_classes = n_classes;
    p.thresh_f32 = thresh_ceil_f32(thresh);
    p.keep_idx = keep_idx; p.keep_cnt = keep_cnt; p.keep_mask = keep_mask;
    p.n_rows = n_rows; p.status = status;
    if (max_seg_len > 1024) {
        const int nb = (max_seg_len + 255) / 256 * 256;
        p.nb = nb;
        p.npad = BIG_MAX;
        p.stage = 0;
        int grid = usable_sm_count();
        plan_items(p, grid);
        if (grid > p.n_items) grid = p.n_items;
        const size_t need = big_ws_bytes(grid, nb);
        if (ws == nullptr || ws_bytes < need) {
            set_error("nms_frames: workspace of %zu bytes needed for %d-box frames", need, max_seg_len);
            return VDET_ERR_WORKSPACE;
        }
        p.gmask = (uint32_t*)ws;
        p.gcnt = (uint16_t*)((char*)ws + (size_t)grid * nb * (nb / 32) * sizeof(uint32_t));
        const size_t smem = big_smem_bytes(nb, p.npad);
        if (smem > max_dynamic_smem(nms_frames_big_kernel)) {
            set_error("nms_frames: %zu bytes of shared memory needed", smem);
            return VDET_ERR_UNSUPPORTED;
        }
        VDET_CUDA(allow_dynamic_smem(nms_frames_big_kernel, smem));
        nms_frames_big_kernel<<<grid, BIG_THREADS, smem, (cudaStream_t)stream>>>(p);
        VDET_LAUNCH_CHECK();
        return VDET_OK;
    }
    const int nb = max_seg_len <= 32 ? 32 : (max_seg_len + 31) / 32 * 32;   // shared-memory capacity
    p.nb = nb;
    p.so_words = (nb / 32) * 33;
    int nper = 1;                                                           // sort network: 32*nper >= nb
    while (32 * nper < nb) nper <<= 1;
    // Stage scores when the block is box-major and the CTA still fits >= 2 per SM.  Residency: the
    // register budget allows 4 CTAs per SM (64 registers x 256 threads; 1 for the 1024-box variant);
    // shared memory decides the rest.  When staging every class at once would cost a CTA slot, the
    // classes are staged in up to 3 chunks (multiples of the warp count) instead.
    const bool want_stage = (score_ldr != 1);
    const size_t base = nms_smem_bytes(nb, nper, 0, false);
    const size_t per_class = (size_t)(nb + 1) * sizeof(float);
    const size_t sm_smem = 228 * 1024, cta_reserved = 1024;
    const int reg_limit = (nper > 16) ? 1 : 4;
    auto fit = [&](size_t smem_cta) {                       // CTAs of that size per SM
        int k = (int)(sm_smem / (smem_cta + cta_reserved));
        return k > reg_limit ? reg_limit : k;
    };
    p.stage = (want_stage && base + n_classes * per_class <= 100 * 1024) ? 1 : 0;
    p.cls_chunk = n_classes;
    int per_sm = fit(base + (p.stage ? n_classes * per_class : 0));
    int forced = 0;
    if (const char* e = getenv("VDET_NMS_PER_SM")) forced = atoi(e);      // measurement hook
    if (p.stage && n_classes > NMS_WARPS) {
        for (int want = reg_limit; want > per_sm; --want) {
            if (forced > 0 && want > forced) continue;
            const size_t budget = sm_smem / want - cta_reserved;
            if (budget <= base) continue;
            const int chunk_max = (int)((budget - base) / per_class);
            if (chunk_max < NMS_WARPS) continue;
            const int n_pass = (n_classes + chunk_max - 1) / chunk_max;
            if (n_pass > 3) continue;
            int chunk = (n_classes + n_pass - 1) / n_pass;
            const int rounded = (chunk + NMS_WARPS - 1) / NMS_WARPS * NMS_WARPS;
            if (rounded <= chunk_max) chunk = rounded;
            p.cls_chunk = chunk;
            per_sm = want;
            break;
        }
    }
    if (forced > 0 && per_sm > forced) per_sm = forced;
    if (per_sm < 1) per_sm = 1;
    const size_t smem = base + (p.stage ? (size_t)(p.cls_chunk < n_classes ? p.cls_chunk : n_classes) * per_class : 0);
    int grid = usable_sm_count() * per_sm;
    plan_items(p, grid);
    if (grid > p.n_items) grid = p.n_items;
    cudaStream_t st = (cudaStream_t)stream;
    switch (nper) {
        case 1:  return launch_nms_frames<1>(p, smem, grid, st);
        case 2:  return launch_nms_frames<2>(p, smem, grid, st);
        case 4:  return launch_nms_frames<4>(p, smem, grid, st);
        case 8:  return launch_nms_frames<8>(p, smem, grid, st);
        case 16: return launch_nms_frames<16>(p, smem, grid, st);
        default: return launch_nms_frames<32>(p, smem, grid, st);
    }
}
