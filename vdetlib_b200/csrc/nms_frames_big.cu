// nms_frames_big.cu -- the big-frame variant of the per-frame NMS (1025..2048 boxes per frame; BASELINE config 5).
// Shares parameters and the pair arithmetic with nms_frames.cuh.
#include "nms_frames.cuh"

namespace vdet {

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
size_t nms_frames_big_ws_bytes(int grid, int nb) {
    return (size_t)grid * ((size_t)nb * (nb / 32) * sizeof(uint32_t) + (size_t)BIG_WARPS * BIG_MAX * sizeof(uint16_t));
}

int launch_nms_frames_big(const NmsFramesParams& p, int grid, cudaStream_t st) {
    const size_t smem = big_smem_bytes(p.nb, p.npad);
    if (smem > max_dynamic_smem(nms_frames_big_kernel)) {
        set_error("nms_frames: %zu bytes of shared memory needed", smem);
        return VDET_ERR_UNSUPPORTED;
    }
    VDET_CUDA(allow_dynamic_smem(nms_frames_big_kernel, smem));
    nms_frames_big_kernel<<<grid, BIG_THREADS, smem, st>>>(p);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

}  // namespace vdet
