// link_sorted.cu -- the frame-to-frame link (link.cu) with the pairs that cannot overlap in x left out.
//
// link.cu evaluates all N x M pair IoUs of a frame pair.  Two boxes whose x ranges are more than the "+1" pixel
// convention apart have inter == 0, i.e. IoU == +0 -- and the link only needs the FIRST arg-max: a zero never beats
// the running best (which starts at "box 0, IoU 0", exactly what the full scan yields when nothing overlaps).  So
//   1. sort_frames_x1_kernel sorts every frame's boxes by x1 once (CTA-wide bitonic sort in shared memory) and
//      leaves the permutation in the caller's workspace; frames with an insane box (common.cuh) are flagged;
//   2. link_frames_sorted_kernel stages frame t+1 in x1 order (boxes, areas, original indices, prefix maximum of
//      x2), takes ITS rows in frame t's x1 order -- a warp's rows are neighbours in x -- and scans only
//      [lo, hi): lo = first position whose prefix-max x2 reaches the warp's smallest x1 - 2, hi = first position
//      whose x1 exceeds the warp's largest x2 + 2.  On BASELINE's synthetic frames that is a third of the pairs.
// First-arg-max in ORIGINAL index order: the running best is one 64-bit key (IoU bits << 32 | ~original index),
// IoUs of sane boxes are >= +0 so their bit patterns order like the floats.  Same bits as link.cu for every
// input: a frame pair with an insane box, or frames longer than LINKS_MAX, take link.cu's full scan.
#include "common.cuh"

namespace vdet {

constexpr int LINKS_MAX = 2048;          // boxes per frame the sorted variant stages at once
constexpr int SORT_THREADS = 256;

// ws layout: perm int32 [n_rows + halo_cap] (sorted position -> index within the frame), flags int32 [n_segs + 1]
// (1 = every box of the frame is sane; slot n_segs = the halo), halo count int32 [1].
__global__ void __launch_bounds__(SORT_THREADS) sort_frames_x1_kernel(const float4* __restrict__ boxes,
                                                                     const int32_t* __restrict__ seg_offsets,
                                                                     int n_segs, const float4* __restrict__ halo,
                                                                     int n_halo, const int32_t* __restrict__ n_halo_dev,
                                                                     int64_t n_rows, int32_t* __restrict__ perm,
                                                                     int32_t* __restrict__ flags) {
    __shared__ uint64_t s_key[LINKS_MAX];
    const int seg = blockIdx.x;                       // n_segs = the halo
    const float4* src;
    int n;
    int64_t out_off;
    if (seg < n_segs) {
        const int off = seg_offsets[seg];
        n = seg_offsets[seg + 1] - off;
        src = boxes + off;
        out_off = off;
    } else {
        n = n_halo;
        if (n_halo_dev != nullptr) {
            const int md = *n_halo_dev;
            n = md < 0 ? 0 : (md < n_halo ? md : n_halo);
        }
        src = halo;
        out_off = n_rows;
        if (threadIdx.x == 0) flags[n_segs + 1] = n;  // the halo's box count, for the link kernel
    }
    if (n > LINKS_MAX) {                              // not handled here: the link kernel takes the full scan
        if (threadIdx.x == 0) flags[seg] = 0;
        return;
    }
    int npow2 = 2;
    while (npow2 < n) npow2 <<= 1;
    bool sane = true;
    for (int e = threadIdx.x; e < npow2; e += SORT_THREADS) {
        uint64_t key = ~0ull;
        if (e < n) {
            const float4 b = __ldg(src + e);
            sane = sane && box_sane(b);
            key = ((uint64_t)f32_key_asc(b.x) << 32) | (uint32_t)e;
        }
        s_key[e] = key;
    }
    const bool all_sane = __syncthreads_and(sane) != 0;
    for (int size = 2; size <= npow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (npow2 >> 1); t += SORT_THREADS) {
                const int i = 2 * t - (t & (stride - 1));
                const int j = i + stride;
                const bool up = (i & size) == 0;
                const uint64_t a = s_key[i], b = s_key[j];
                if ((a > b) == up) { s_key[i] = b; s_key[j] = a; }
            }
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < n; q += SORT_THREADS) perm[out_off + q] = (int32_t)(uint32_t)s_key[q];
    if (threadIdx.x == 0) flags[seg] = all_sane ? 1 : 0;
}

template <int THREADS, int ROWS>
__global__ void __launch_bounds__(THREADS) link_frames_sorted_kernel(const float4* __restrict__ boxes,
                                                                     const int32_t* __restrict__ seg_offsets,
                                                                     int n_segs, const float4* __restrict__ halo,
                                                                     int halo_row_base, int64_t n_rows,
                                                                     const int32_t* __restrict__ perm,
                                                                     const int32_t* __restrict__ flags,
                                                                     int32_t* __restrict__ succ,
                                                                     float* __restrict__ best_iou) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seg = blockIdx.x;
    const int off = seg_offsets[seg];
    const int n = seg_offsets[seg + 1] - off;
    const int row0 = blockIdx.y * (THREADS * ROWS);
    if (row0 >= n) return;                                     // whole CTA beyond this frame
    const bool last = (seg == n_segs - 1);
    const float4* nxt = last ? halo : boxes + seg_offsets[seg + 1];
    const int m = last ? flags[n_segs + 1] : (seg_offsets[seg + 2] - seg_offsets[seg + 1]);
    const int32_t* nperm = perm + (last ? n_rows : (int64_t)seg_offsets[seg + 1]);
    const int out_base = last ? halo_row_base : seg_offsets[seg + 1];
    const bool sorted_ok = flags[seg] != 0 && flags[last ? n_segs : seg + 1] != 0 && n <= LINKS_MAX && m <= LINKS_MAX;
    const int lane = threadIdx.x & 31;

    if (!sorted_ok) {
        // full scan, generic IEEE division: link.cu's fallback path, for frame pairs with an insane box
        float4* s_box = reinterpret_cast<float4*>(smem_raw);
        float* s_area = reinterpret_cast<float*>(s_box + 1024);
        for (int r = 0; r < ROWS; ++r) {
            const int i = row0 + r * THREADS + threadIdx.x;
            const float4 bi = i < n ? __ldg(boxes + off + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float ai = area_f32(bi);
            float best = -1.0f;
            int arg = -1;
            for (int j0 = 0; j0 < m; j0 += 1024) {
                const int cnt = (m - j0) < 1024 ? (m - j0) : 1024;
                __syncthreads();
                for (int e = threadIdx.x; e < cnt; e += THREADS) {
                    const float4 b = __ldg(nxt + j0 + e);
                    s_box[e] = b;
                    s_area[e] = area_f32(b);
                }
                __syncthreads();
                for (int j = 0; j < cnt; ++j) {
                    float inter, uni;
                    inter_union_f32(bi, ai, s_box[j], s_area[j], inter, uni);
                    const float v = iou_quotient(inter, uni);
                    if (uni != 0.0f && (arg < 0 || v > best)) { best = v; arg = j0 + j; }    // NaN / union 0 never win
                }
            }
            if (i < n) {
                succ[off + i] = arg < 0 ? -1 : out_base + arg;
                best_iou[off + i] = arg < 0 ? 0.0f : best;
            }
        }
        return;
    }

    // ---- stage frame t+1 in x1 order ---------------------------------------------------------------
    const int mp = (m + 31) & ~31;
    float4* s_box = reinterpret_cast<float4*>(smem_raw);                 // [mp]
    float* s_area = reinterpret_cast<float*>(s_box + mp);                // [mp]
    float* s_pmx2 = s_area + mp;                                         // [mp] prefix maximum of x2
    uint32_t* s_low = reinterpret_cast<uint32_t*>(s_pmx2 + mp);          // [mp] ~original index (low key word)
    __shared__ float s_wmax[32];
    for (int q = threadIdx.x; q < m; q += THREADS) {
        const int j = nperm[q];
        const float4 b = __ldg(nxt + j);
        s_box[q] = b;
        s_area[q] = area_f32(b);
        s_low[q] = ~(uint32_t)j;
        s_pmx2[q] = b.z;
    }
    __syncthreads();
    // inclusive prefix maximum of x2 over the sorted order: per-thread runs, then a scan of the run maxima
    {
        const int per = (m + THREADS - 1) / THREADS;
        const int q0 = threadIdx.x * per, q1 = (q0 + per < m) ? q0 + per : m;
        float run = -INFINITY;
        for (int q = q0; q < q1; ++q) { run = fmaxf(run, s_pmx2[q]); s_pmx2[q] = run; }
        float incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl = fmaxf(incl, o);
        }
        if (lane == 31) s_wmax[threadIdx.x >> 5] = incl;
        __syncthreads();
        float before = -INFINITY;                                        // maximum over all earlier threads
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before = fmaxf(before, s_wmax[w]);
        const float up = __shfl_up_sync(FULL, incl, 1);
        if (lane > 0) before = fmaxf(before, up);
        for (int q = q0; q < q1; ++q) s_pmx2[q] = fmaxf(s_pmx2[q], before);
    }
    __syncthreads();

    // ---- this thread's rows, in frame t's x1 order (a warp's 32*ROWS rows are consecutive there) ---------
    const int32_t* rperm = perm + off;
    float4 bi[ROWS];
    float ai[ROWS];
    int orig[ROWS];
    unsigned long long best[ROWS];                       // (IoU bits << 32) | ~original index of the best box so far
    float xlo = INFINITY, xhi = -INFINITY;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int q = row0 + threadIdx.x * ROWS + r;
        orig[r] = q < n ? rperm[q] : -1;
        bi[r] = q < n ? __ldg(boxes + off + orig[r]) : make_float4(0.f, 0.f, 0.f, 0.f);
        ai[r] = area_f32(bi[r]);
        best[r] = 0xffffffffull;                         // IoU +0 with original index 0: what a scan that never overlaps yields
        if (q < n) { xlo = fminf(xlo, bi[r].x); xhi = fmaxf(xhi, bi[r].z); }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        xlo = fminf(xlo, __shfl_xor_sync(FULL, xlo, d));
        xhi = fmaxf(xhi, __shfl_xor_sync(FULL, xhi, d));
    }
    // [lo, hi): positions that can overlap one of the warp's rows in x (margin 2 > the +1 convention)
    int lo = 0, hi = 0;
    if (m > 0 && xlo <= xhi) {
        const float need = __fsub_rn(xlo, 2.0f), lim = __fadd_rn(xhi, 2.0f);
        int a = 0, b = m;                                // first q with pmx2[q] >= need
        while (a < b) { const int mid = (a + b) >> 1; if (s_pmx2[mid] < need) a = mid + 1; else b = mid; }
        lo = a;
        a = lo; b = m;                                   // first q with x1[q] > lim
        while (a < b) { const int mid = (a + b) >> 1; if (s_box[mid].x <= lim) a = mid + 1; else b = mid; }
        hi = a;
    }
    if (ROWS >= 2) {
#pragma unroll 2
        for (int j = lo; j < hi; ++j) {
            const float4 bj = s_box[j];
            const float aj = s_area[j];
            const uint32_t low = s_low[j];
            const f32x2 aj2 = pk2(aj, aj);
#pragma unroll
            for (int r = 0; r + 1 < ROWS; r += 2) {
                f32x2 inter, uni, nuni;
                inter_union_f32x2(bj, aj2, bi[r], bi[r + 1], pk2(ai[r], ai[r + 1]), inter, uni, nuni);
                float v0, v1;
                upk2(div_sane2(inter, nuni), v0, v1);
                const unsigned long long k0 = ((unsigned long long)__float_as_uint(v0) << 32) | low;
                const unsigned long long k1 = ((unsigned long long)__float_as_uint(v1) << 32) | low;
                best[r] = k0 > best[r] ? k0 : best[r];
                best[r + 1] = k1 > best[r + 1] ? k1 : best[r + 1];
            }
        }
    } else {
        const f32x2 ai2 = pk2(ai[0], ai[0]);
        int j = lo;
#pragma unroll 2
        for (; j + 1 < hi; j += 2) {
            f32x2 inter, uni, nuni;
            inter_union_f32x2(bi[0], ai2, s_box[j], s_box[j + 1], pk2(s_area[j], s_area[j + 1]), inter, uni, nuni);
            float v0, v1;
            upk2(div_sane2(inter, nuni), v0, v1);
            const unsigned long long k0 = ((unsigned long long)__float_as_uint(v0) << 32) | s_low[j];
            const unsigned long long k1 = ((unsigned long long)__float_as_uint(v1) << 32) | s_low[j + 1];
            best[0] = k0 > best[0] ? k0 : best[0];
            best[0] = k1 > best[0] ? k1 : best[0];
        }
        if (j < hi) {
            float inter, uni;
            inter_union_f32(bi[0], ai[0], s_box[j], s_area[j], inter, uni);
            const unsigned long long k0 = ((unsigned long long)__float_as_uint(div_sane(inter, uni)) << 32) | s_low[j];
            best[0] = k0 > best[0] ? k0 : best[0];
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (orig[r] >= 0) {
            const int arg = (int)~(uint32_t)best[r];
            succ[off + orig[r]] = m > 0 ? out_base + arg : -1;
            best_iou[off + orig[r]] = m > 0 ? __uint_as_float((uint32_t)(best[r] >> 32)) : 0.0f;
        }
    }
}

template <int THREADS, int ROWS>
static int launch_sorted(const float* boxes, const int32_t* seg_offsets, int n_segs, int max_seg_len,
                         const float* halo_boxes, int n_halo, int halo_row_base, int64_t n_rows, const int32_t* perm,
                         const int32_t* flags, int32_t* succ, float* best_iou, cudaStream_t st) {
    constexpr int RPC = THREADS * ROWS;
    int cap = max_seg_len > n_halo ? max_seg_len : n_halo;
    if (cap > LINKS_MAX) cap = LINKS_MAX;
    const int mp = (cap + 31) & ~31;
    size_t smem = (size_t)mp * (sizeof(float4) + 3 * sizeof(float));
    const size_t fallback = 1024 * (sizeof(float4) + sizeof(float));
    if (smem < fallback) smem = fallback;
    VDET_CUDA(allow_dynamic_smem(link_frames_sorted_kernel<THREADS, ROWS>, smem));
    dim3 grid((unsigned)n_segs, (unsigned)((max_seg_len + RPC - 1) / RPC));
    link_frames_sorted_kernel<THREADS, ROWS><<<grid, THREADS, smem, st>>>(
        (const float4*)boxes, seg_offsets, n_segs, (const float4*)halo_boxes, halo_row_base, n_rows, perm, flags, succ,
        best_iou);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

size_t link_sorted_ws_bytes(int64_t n_rows, int n_segs, int n_halo) {
    return align_up((size_t)(n_rows + n_halo) * sizeof(int32_t), 256) + (size_t)(n_segs + 2) * sizeof(int32_t) + 256;
}

// The sorted link: returns VDET_OK after enqueueing both kernels.  The caller (link.cu) decides when it applies.
int launch_link_frames_sorted(const float* boxes, const int32_t* seg_offsets, int n_segs, int max_seg_len,
                              const float* halo_boxes, int n_halo, const int32_t* n_halo_dev, int halo_row_base,
                              int32_t* succ, float* best_iou, int64_t n_rows, void* ws, cudaStream_t st) {
    int32_t* perm = (int32_t*)ws;
    int32_t* flags = (int32_t*)((char*)ws + align_up((size_t)(n_rows + n_halo) * sizeof(int32_t), 256));
    sort_frames_x1_kernel<<<(unsigned)(n_segs + 1), SORT_THREADS, 0, st>>>(
        (const float4*)boxes, seg_offsets, n_segs, (const float4*)halo_boxes, halo_boxes ? n_halo : 0, n_halo_dev, n_rows,
        perm, flags);
    VDET_LAUNCH_CHECK();
    // rows per CTA as in link.cu: fewest padded row slots, then fewest CTAs per frame
    const int cand[4] = {64, 128, 256, 512};
    int best_rpc = 64;
    long best_pad = -1;
    for (int k = 0; k < 4; ++k) {
        const long pad = ((long)max_seg_len + cand[k] - 1) / cand[k] * cand[k];
        if (best_pad < 0 || pad <= best_pad) { best_pad = pad; best_rpc = cand[k]; }
    }
    switch (best_rpc) {
        case 64:  return launch_sorted<64, 1>(boxes, seg_offsets, n_segs, max_seg_len, halo_boxes, n_halo, halo_row_base, n_rows, perm, flags, succ, best_iou, st);
        case 128: return launch_sorted<64, 2>(boxes, seg_offsets, n_segs, max_seg_len, halo_boxes, n_halo, halo_row_base, n_rows, perm, flags, succ, best_iou, st);
        case 256: return launch_sorted<128, 2>(boxes, seg_offsets, n_segs, max_seg_len, halo_boxes, n_halo, halo_row_base, n_rows, perm, flags, succ, best_iou, st);
        default:  return launch_sorted<128, 4>(boxes, seg_offsets, n_segs, max_seg_len, halo_boxes, n_halo, halo_row_base, n_rows, perm, flags, succ, best_iou, st);
    }
}

}  // namespace vdet
