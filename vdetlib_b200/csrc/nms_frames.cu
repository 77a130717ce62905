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
#include "common.cuh"
#include "warp_sort.cuh"

namespace vdet {

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
};

template <int NPER>
__global__ void __launch_bounds__(NMS_THREADS) nms_frames_kernel(const NmsFramesParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NB = p.nb;
    const int W = NB >> 5;          // mask words per row (<= 32 in this variant)
    const int WS = W | 1;           // odd row stride: column writes of phase B are conflict free
    const int SST = NB + 1;         // odd class stride of the staged score block
    float4* sbox = reinterpret_cast<float4*>(smem_raw);
    float* sarea = reinterpret_cast<float*>(sbox + NB);
    int32_t* srow = reinterpret_cast<int32_t*>(sarea + NB);
    uint32_t* smask = reinterpret_cast<uint32_t*>(srow + NB);
    float* sscore = reinterpret_cast<float*>(smask + (size_t)NB * WS);
    __shared__ int s_zero_union;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int C = p.n_classes;
    const float T = p.thresh_f32;

    for (int seg = blockIdx.x; seg < p.n_segs; seg += gridDim.x) {
        const int off = p.seg_offsets[seg];
        const int n = p.seg_offsets[seg + 1] - off;
        if (n > NB) {   // caller's max_seg_len was wrong: refuse loudly instead of truncating
            if (tid == 0) atomicOr(p.status, 0x80000000u);
            continue;
        }
        // ---- A: stage boxes, areas, original row ids (and scores) ------------------------
        if (tid == 0) s_zero_union = 0;
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
        }
        __syncthreads();
        if (p.stage) {
            if (p.row_ids == nullptr && p.score_ldc == 1 && p.score_ldr == C) {
                // contiguous [n, C] block: flat coalesced read, transposed conflict-free write
                const float* src = p.scores + (int64_t)off * C;
                const int total = n * C;
                for (int f = tid; f < total; f += NMS_THREADS) {
                    const int r = f / C, c = f - r * C;
                    sscore[c * SST + r] = __ldg(src + f);
                }
            } else {
                const int total = n * C;
                for (int f = tid; f < total; f += NMS_THREADS) {
                    const int r = f / C, c = f - r * C;
                    sscore[c * SST + r] = __ldg(p.scores + (int64_t)srow[r] * p.score_ldr + (int64_t)c * p.score_ldc);
                }
            }
        }
        // ---- B: suppression bit matrix, original index space ------------------------------
        {
            const int Wn = (n + 31) >> 5;          // blocks actually populated by this frame
            const int ntiles = Wn * Wn;
            for (int tile = warp; tile < ntiles; tile += NMS_WARPS) {
                const int rb = tile / Wn, cb = tile - rb * Wn;
                const int i = rb * 32 + lane;
                const float4 bi = sbox[i];
                const float ai = sarea[i];
                uint32_t word = 0;
                bool zero = false;
#pragma unroll 8
                for (int jj = 0; jj < 32; ++jj) {
                    const int j = cb * 32 + jj;
                    const float4 bj = sbox[j];
                    const float aj = sarea[j];
                    float inter, uni;
                    inter_union_f32(bi, ai, bj, aj, inter, uni);
                    const float ovr = __fdiv_rn(inter, uni);
                    if (ovr >= T) word |= (1u << jj);
                    zero |= (uni == 0.0f) && (i != j) && (i < n) && (j < n);
                }
                // columns beyond the frame never suppress / are never visited
                const int valid = n - cb * 32;
                if (valid < 32) word &= (valid <= 0) ? 0u : ((1u << valid) - 1u);
                smask[i * WS + cb] = word;
                if (__any_sync(FULL, zero) && lane == 0) s_zero_union = 1;
            }
        }
        __syncthreads();
        const bool check_zero = (s_zero_union != 0);
        const int Wn = (n + 31) >> 5;

        // ---- C: per class: sort + greedy walk (one warp per class) -----------------------
        for (int c = warp; c < C; c += NMS_WARPS) {
            uint64_t key[NPER];
#pragma unroll
            for (int r = 0; r < NPER; ++r) {
                const int e = r * 32 + lane;      // striped load: conflict-free / coalesced
                key[r] = ~0ull;                   // padding sorts last
                if (e < n) {
                    const float s = p.stage
                        ? sscore[c * SST + e]
                        : __ldg(p.scores + (int64_t)srow[e] * p.score_ldr + (int64_t)c * p.score_ldc);
                    key[r] = ((uint64_t)f32_key_desc(s) << 32) | (uint32_t)e;
                }
            }
            warp_bitonic_sort<NPER>(key, lane);

            uint32_t rem = 0;        // lane w: word w of the removed set
            uint32_t kept = 0;       // lane w: word w of the kept set
            int cnt = 0;
            int32_t buf = -1;        // lane (cnt & 31) buffers the cnt-th kept row
            int32_t* out_idx = p.keep_idx + (int64_t)c * p.n_rows + off;
            for (int src = 0; src < 32; ++src) {
                if (src * NPER >= n) break;
#pragma unroll
                for (int r = 0; r < NPER; ++r) {
                    const uint32_t i = __shfl_sync(FULL, (uint32_t)key[r], src);
                    if (src * NPER + r < n) {                     // warp-uniform
                        const uint32_t w = __shfl_sync(FULL, rem, (int)(i >> 5));
                        if (!((w >> (i & 31)) & 1u)) {            // warp-uniform: i is kept
                            if (check_zero) {
                                // exact ZeroDivisionError test of nms.pyx:64: a pair (i, j) is
                                // visited iff j comes later in the order and is not yet removed
                                const int pos = src * NPER + r;
                                const float4 bi = sbox[i];
                                const float ai = sarea[i];
                                bool zd = false;
#pragma unroll
                                for (int r2 = 0; r2 < NPER; ++r2) {
                                    const uint32_t j = (uint32_t)key[r2];
                                    const uint32_t wj = __shfl_sync(FULL, rem, (int)((j >> 5) & 31));
                                    const int pos2 = lane * NPER + r2;
                                    if (pos2 > pos && pos2 < n && !((wj >> (j & 31)) & 1u)) {
                                        float inter, uni;
                                        inter_union_f32(bi, ai, sbox[j], sarea[j], inter, uni);
                                        zd |= (uni == 0.0f);
                                    }
                                }
                                if (__any_sync(FULL, zd) && lane == 0) atomicOr(p.status, VDET_STATUS_ZERO_DIVISION);
                            }
                            if (lane < Wn) rem |= smask[i * WS + lane];
                            if (lane == (int)(i >> 5)) kept |= (1u << (i & 31));
                            if (lane == (cnt & 31)) buf = srow[i];
                            ++cnt;
                            if ((cnt & 31) == 0) out_idx[cnt - 32 + lane] = buf;
                        }
                    }
                }
            }
            // flush the partial group, pad the frame's remaining slots with -1
            {
                const int done = cnt & ~31;
                if (done + lane < cnt) out_idx[done + lane] = buf;
                for (int e = cnt + lane; e < n; e += 32) out_idx[e] = -1;
                if (lane == 0) p.keep_cnt[(int64_t)c * p.n_segs + seg] = cnt;
            }
            if (p.keep_mask) {
                uint8_t* out_m = p.keep_mask + (int64_t)c * p.n_rows + off;
                for (int wi = 0; wi < Wn; ++wi) {
                    const uint32_t kw = __shfl_sync(FULL, kept, wi);
                    const int e = wi * 32 + lane;
                    if (e < n) out_m[e] = (uint8_t)((kw >> lane) & 1u);
                }
            }
        }
        __syncthreads();   // smem is reused by the next frame
    }
}

static size_t nms_smem_bytes(int nb, int n_classes, bool stage) {
    const int W = nb / 32, WS = W | 1;
    size_t b = (size_t)nb * (sizeof(float4) + sizeof(float) + sizeof(int32_t));
    b += (size_t)nb * WS * sizeof(uint32_t);
    if (stage) b += (size_t)n_classes * (nb + 1) * sizeof(float);
    return b;
}

template <int NPER>
static int launch_nms_frames(const NmsFramesParams& p, size_t smem, int grid, cudaStream_t st) {
    if (smem > max_dynamic_smem(nms_frames_kernel<NPER>)) {
        set_error("nms_frames: %zu bytes of shared memory needed", smem);
        return VDET_ERR_UNSUPPORTED;
    }
    VDET_CUDA(allow_dynamic_smem(nms_frames_kernel<NPER>, smem));
    nms_frames_kernel<NPER><<<grid, NMS_THREADS, smem, st>>>(p);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

}  // namespace vdet

using namespace vdet;

extern "C" size_t vdet_nms_frames_workspace_bytes(int max_seg_len, int n_classes, int device) {
    (void)max_seg_len; (void)n_classes; (void)device;
    return 256;   // the register-sort variants keep everything in shared memory
}

extern "C" int vdet_nms_frames_f32(const float* boxes, int box_ld,
                                   const float* scores, int64_t score_ldr, int64_t score_ldc,
                                   const int32_t* seg_offsets, int n_segs, int max_seg_len,
                                   const int32_t* row_ids, int n_classes, double thresh,
                                   int32_t* keep_idx, int32_t* keep_cnt, uint8_t* keep_mask,
                                   int64_t n_rows, uint32_t* status,
                                   void* ws, size_t ws_bytes, void* stream) {
    (void)ws; (void)ws_bytes;
    VDET_REQUIRE(n_segs >= 0 && n_classes >= 1 && n_rows >= 0 && max_seg_len >= 0, "nms_frames: negative size");
    VDET_REQUIRE(box_ld >= 4, "nms_frames: box_ld must be >= 4");
    VDET_REQUIRE(status != nullptr && keep_idx != nullptr && keep_cnt != nullptr, "nms_frames: null output");
    if (n_segs == 0 || n_rows == 0 && max_seg_len == 0) {
        if (n_segs > 0)
            VDET_CUDA(cudaMemsetAsync(keep_cnt, 0, sizeof(int32_t) * (size_t)n_segs * n_classes, (cudaStream_t)stream));
        return VDET_OK;
    }
    if (max_seg_len > 1024) {
        set_error("nms_frames: max_seg_len %d > 1024 is not supported by this build", max_seg_len);
        return VDET_ERR_UNSUPPORTED;
    }
    NmsFramesParams p;
    p.boxes = boxes; p.box_ld = box_ld;
    p.box_vec = (box_ld == 4) && ((reinterpret_cast<uintptr_t>(boxes) & 15) == 0);
    p.scores = scores; p.score_ldr = score_ldr; p.score_ldc = score_ldc;
    p.seg_offsets = seg_offsets; p.n_segs = n_segs; p.row_ids = row_ids;
    p.n_classes = n_classes;
    p.thresh_f32 = thresh_ceil_f32(thresh);
    p.keep_idx = keep_idx; p.keep_cnt = keep_cnt; p.keep_mask = keep_mask;
    p.n_rows = n_rows; p.status = status;
    const int nb = max_seg_len <= 32 ? 32 : (max_seg_len + 31) / 32 * 32;   // shared-memory capacity
    p.nb = nb;
    int nper = 1;                                                           // sort network: 32*nper >= nb
    while (32 * nper < nb) nper <<= 1;
    // Stage scores when the block is box-major and the CTA still fits >= 2 per SM.
    const bool want_stage = (score_ldr != 1);
    const size_t smem_stage = nms_smem_bytes(nb, n_classes, true);
    p.stage = (want_stage && smem_stage <= 100 * 1024) ? 1 : 0;
    const size_t smem = nms_smem_bytes(nb, n_classes, p.stage != 0);
    int per_sm = (int)((size_t)(220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int grid = sm_count_cached() * per_sm;
    if (grid > n_segs) grid = n_segs;
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
