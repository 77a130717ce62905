// compact.cu -- the keep lists of vdet_nms_frames_f32 in the form a caller takes home.
//
// utils/nms.pyx:43-66 returns, per problem, the kept rows in descending score -- a Python list of
// K ints.  vdet_nms_frames_f32 leaves them in padded per-(frame, class) blocks of n slots (so that
// the NMS kernel never needs another problem's count).  A host caller wants exactly sum(K) entries:
// this file turns the padded blocks into ONE contiguous array in (frame, class) order plus its
// prefix offsets -- what apply_vid_nms (vdet/video_det.py:51-61) would return for every class --
// and, optionally, the same keep sets as bit masks.
//
// Both outputs may live in MAPPED PINNED HOST memory: the kernels then write the results straight
// across PCIe (posted writes, 16-byte vectors on 16-byte boundaries), so a data-dependent number of
// bytes goes home without the host knowing sum(K) in advance -- a copy-engine D2H would need that
// size on the host first, i.e. a synchronisation in the middle of the step.
#include "common.cuh"

namespace vdet {

constexpr int KO_THREADS = 1024;
constexpr int KO_ITEMS = 4;

// Exclusive prefix sum of n counts -> off[0..n] (n + 1 entries), one CTA walking the array in tiles
// of 4096 (n is frames x classes: 30,000 for BASELINE config 2, 60,000 for a config-5 video).
__global__ void __launch_bounds__(KO_THREADS) keep_offsets_kernel(const int32_t* __restrict__ cnt, int n,
                                                                  int32_t* __restrict__ off_a,
                                                                  int32_t* __restrict__ off_b) {
    __shared__ int s_warp[KO_THREADS / 32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += KO_THREADS * KO_ITEMS) {
        const int i0 = base + tid * KO_ITEMS;
        int v[KO_ITEMS];
        int s = 0;
#pragma unroll
        for (int q = 0; q < KO_ITEMS; ++q) {
            v[q] = (i0 + q < n) ? cnt[i0 + q] : 0;
            s += v[q];
        }
        int incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_warp[lane];
            int wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(FULL, wi, d);
                if (lane >= d) wi += o;
            }
            s_warp[lane] = wi - w;                 // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int carry = s_carry;
        int run = carry + s_warp[warp] + incl - s;
#pragma unroll
        for (int q = 0; q < KO_ITEMS; ++q) {
            if (i0 + q < n) {
                off_a[i0 + q] = run;
                if (off_b) off_b[i0 + q] = run;
            }
            run += v[q];
        }
        __syncthreads();
        if (tid == KO_THREADS - 1) s_carry = run;  // the last thread's running sum is the tile's inclusive total
        __syncthreads();
    }
    if (tid == 0) {
        off_a[n] = s_carry;
        if (off_b) off_b[n] = s_carry;
    }
}

constexpr int CK_THREADS = 256;
constexpr int CK_WARPS = CK_THREADS / 32;
constexpr int CK_CAP_BYTES = 32 * 1024;       // staged output bytes per chunk of blocks
constexpr int CK_MAX_WORDS = 64;              // bit-mask words per block (frames <= 2048 boxes)

// One CTA owns chunks of `bpc` consecutive (frame, class) blocks: their kept entries are gathered into
// shared memory at the positions they take in the output, then the chunk's byte range -- contiguous
// in the output -- is written with 16-byte stores aligned to the DESTINATION (the staging buffer is
// shifted by the destination's 16-byte phase).  OutT = uint16_t: index within the frame; int32_t: packed row.
template <typename OutT>
__global__ void __launch_bounds__(CK_THREADS) compact_keep_kernel(const int32_t* __restrict__ keep_idx,
                                                                  const int32_t* __restrict__ keep_cnt,
                                                                  const int32_t* __restrict__ seg_offsets,
                                                                  int n_blocks, int C, int bpc,
                                                                  const int32_t* __restrict__ off,
                                                                  OutT* __restrict__ out,
                                                                  uint32_t* __restrict__ bits, int words) {
    __shared__ __align__(16) unsigned char s_out[CK_CAP_BYTES + 16];
    __shared__ uint32_t s_bits[CK_WARPS][CK_MAX_WORDS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_chunks = (n_blocks + bpc - 1) / bpc;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int b0 = chunk * bpc;
        const int b1 = (b0 + bpc < n_blocks) ? b0 + bpc : n_blocks;
        const int o0 = off[b0], o1 = off[b1];
        unsigned char* gb = reinterpret_cast<unsigned char*>(out + o0);
        const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(gb) & 15);
        OutT* stage = reinterpret_cast<OutT*>(s_out + phase);
        for (int b = b0 + warp; b < b1; b += CK_WARPS) {
            const int s = b / C, c = b - s * C;
            const int so = seg_offsets[s];
            const int n = seg_offsets[s + 1] - so;
            const int cnt = keep_cnt[b];
            const int32_t* src = keep_idx + (int64_t)so * C + (int64_t)c * n;
            OutT* dst = stage + (off[b] - o0);
            if (bits) {
                for (int w = lane; w < words; w += 32) s_bits[warp][w] = 0u;
                __syncwarp();
            }
            for (int k = lane; k < cnt; k += 32) {
                const int row = __ldg(src + k);
                const int loc = row - so;
                dst[k] = sizeof(OutT) == 2 ? (OutT)loc : (OutT)row;
                if (bits) atomicOr(&s_bits[warp][loc >> 5], 1u << (loc & 31));
            }
            if (bits) {
                __syncwarp();
                for (int w = lane; w < words; w += 32) bits[(int64_t)b * words + w] = s_bits[warp][w];
                __syncwarp();
            }
        }
        __syncthreads();
        // aligned write-out of [o0, o1)
        const size_t nbytes = (size_t)(o1 - o0) * sizeof(OutT);
        size_t head = (16 - phase) & 15;
        if (head > nbytes) head = nbytes;
        const unsigned char* sb = s_out + phase;
        for (size_t i = (size_t)tid * sizeof(OutT); i < head; i += (size_t)CK_THREADS * sizeof(OutT))
            *reinterpret_cast<OutT*>(gb + i) = *reinterpret_cast<const OutT*>(sb + i);
        const size_t nvec = (nbytes - head) / 16;
        uint4* gv = reinterpret_cast<uint4*>(gb + head);
        const uint4* sv = reinterpret_cast<const uint4*>(sb + head);
        for (size_t v = tid; v < nvec; v += CK_THREADS) gv[v] = sv[v];
        const size_t done = head + nvec * 16;
        for (size_t i = done + (size_t)tid * sizeof(OutT); i < nbytes; i += (size_t)CK_THREADS * sizeof(OutT))
            *reinterpret_cast<OutT*>(gb + i) = *reinterpret_cast<const OutT*>(sb + i);
        __syncthreads();
    }
}

}  // namespace vdet

using namespace vdet;

extern "C" int vdet_compact_keep(const int32_t* keep_idx, const int32_t* keep_cnt, const int32_t* seg_offsets,
                                 int n_segs, int max_seg_len, int n_classes, int out_dtype,
                                 int32_t* keep_off, int32_t* keep_off_mirror, void* keep_out,
                                 uint32_t* keep_bits, void* stream) {
    VDET_REQUIRE(n_segs >= 0 && n_classes >= 1 && max_seg_len >= 0, "compact_keep: negative size");
    VDET_REQUIRE(out_dtype == VDET_KEEP_U16_LOCAL || out_dtype == VDET_KEEP_I32_ROW, "compact_keep: bad out_dtype");
    VDET_REQUIRE(keep_idx && keep_cnt && seg_offsets && keep_off && keep_out, "compact_keep: null pointer");
    const size_t esz = out_dtype == VDET_KEEP_U16_LOCAL ? 2 : 4;
    if (out_dtype == VDET_KEEP_U16_LOCAL && max_seg_len > 65536) {
        set_error("compact_keep: frames of %d boxes do not fit 16-bit local indices", max_seg_len);
        return VDET_ERR_UNSUPPORTED;
    }
    const int words = (max_seg_len + 31) / 32;
    if ((size_t)max_seg_len * esz > (size_t)CK_CAP_BYTES || (keep_bits && words > CK_MAX_WORDS)) {
        set_error("compact_keep: frames of %d boxes are not supported by this build", max_seg_len);
        return VDET_ERR_UNSUPPORTED;
    }
    VDET_REQUIRE(((uintptr_t)keep_out & (esz - 1)) == 0, "compact_keep: misaligned output");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nb64 = (int64_t)n_segs * n_classes;
    VDET_REQUIRE(nb64 < (1ll << 31), "compact_keep: too many (frame, class) blocks");
    const int n_blocks = (int)nb64;
    keep_offsets_kernel<<<1, KO_THREADS, 0, st>>>(keep_cnt, n_blocks, keep_off, keep_off_mirror);
    VDET_LAUNCH_CHECK();
    if (n_blocks == 0) return VDET_OK;
    int bpc = max_seg_len > 0 ? (int)((size_t)CK_CAP_BYTES / ((size_t)max_seg_len * esz)) : n_blocks;
    if (bpc < 1) bpc = 1;
    if (bpc > 256) bpc = 256;
    const int n_chunks = (n_blocks + bpc - 1) / bpc;
    int grid = sm_count_cached() * 4;
    if (grid > n_chunks) grid = n_chunks;
    if (out_dtype == VDET_KEEP_U16_LOCAL)
        compact_keep_kernel<uint16_t><<<grid, CK_THREADS, 0, st>>>(keep_idx, keep_cnt, seg_offsets, n_blocks, n_classes,
                                                                   bpc, keep_off, (uint16_t*)keep_out, keep_bits, words);
    else
        compact_keep_kernel<int32_t><<<grid, CK_THREADS, 0, st>>>(keep_idx, keep_cnt, seg_offsets, n_blocks, n_classes,
                                                                  bpc, keep_off, (int32_t*)keep_out, keep_bits, words);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}
