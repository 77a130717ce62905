// nms_frames.cu -- host entry of the per-frame NMS (vdet_nms_frames_f32) and the single-array kernel variants.
// The kernel itself: nms_frames.cuh; two-array sort variants: nms_frames_split.cu; frames of 1025..2048 boxes:
// nms_frames_big.cu.
#include "nms_frames.cuh"
#include "nms_plan.h"

namespace vdet {

static_assert(NMS_PLAN_WARPS_DEFAULT * 32 == NMS_THREADS && NMS_PLAN_CTAS_DEFAULT == VDET_NMS_CTAS_PER_SM &&
              NMS_PLAN_WARPS_WIDE * 32 == NMS_THREADS_WIDE && NMS_PLAN_CTAS_WIDE == NMS_CTAS_WIDE,
              "nms_plan.h models the CTA shapes of nms_frames.cuh");

// Work items for a persistent grid of `slots` CTAs over n_segs frames (see NmsFramesParams).
static void plan_items(NmsFramesParams& p, int slots) {
    p.split_from = p.n_segs; p.nsplit = 1; p.n_items = p.n_segs;
    if (slots <= 0 || p.n_classes < 2) return;
    // Only launches that cannot fill the grid once are split (a single image with 30 classes, a short
    // clip): measured on config 2, splitting the LAST round of a multi-round launch does not pay --
    // the CTAs of a thin last round already run alone on their SMs and finish early.
    // (again with the 10-warp shape: 1000 frames over 444 slots, the 112 frames of the third round cut in three:
    // 0.288 ms against 0.278)
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

static size_t nms_smem_bytes(int nb, int sort_blocks, int n_classes, bool stage, int warps) {
    const int W = nb / 32, WS = W | 1;
    size_t b = (size_t)nb * (sizeof(float4) + sizeof(float) + sizeof(int32_t));
    b += (size_t)nb * WS * sizeof(uint32_t);
    b += (size_t)warps * sort_blocks * 33 * sizeof(uint32_t);      // per-warp key / order scratch, skewed 32-key blocks
    if (stage) b += (size_t)n_classes * (nb + 1) * sizeof(float);
    return b;
}

template <int NPER>
static int launch_nms_frames(int threads, const NmsFramesParams& p, size_t smem, int grid, cudaStream_t st) {
    if (NPER <= 8 && threads == NMS_THREADS_WIDE)          // staged only (see the plan in vdet_nms_frames_f32)
        return launch_nms_frames_t<(NPER <= 8 ? NPER : 8), 0, true, NMS_THREADS_WIDE>(p, smem, grid, st);
    return p.stage ? launch_nms_frames_t<NPER, 0, true>(p, smem, grid, st)
                   : launch_nms_frames_t<NPER, 0, false>(p, smem, grid, st);
}

}  // namespace vdet


using namespace vdet;

extern "C" size_t vdet_nms_frames_workspace_bytes(int max_seg_len, int n_classes, int device) {
    (void)n_classes; (void)device;
    if (max_seg_len <= 1024) return 256;   // register-sort variants keep everything in shared memory
    const size_t nb = ((size_t)max_seg_len + 255) / 256 * 256;
    return nms_frames_big_ws_bytes(sm_count_cached(), (int)nb) + 256;   // upper bound: all SMs
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
    if (max_seg_len > 2048) {
        set_error("nms_frames: max_seg_len %d > %d is not supported by this build", max_seg_len, 2048);
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
        p.npad = 2048;
        p.stage = 0;
        int grid = usable_sm_count();
        plan_items(p, grid);
        if (grid > p.n_items) grid = p.n_items;
        const size_t need = nms_frames_big_ws_bytes(grid, nb);
        if (ws == nullptr || ws_bytes < need) {
            set_error("nms_frames: workspace of %zu bytes needed for %d-box frames", need, max_seg_len);
            return VDET_ERR_WORKSPACE;
        }
        p.gmask = (uint32_t*)ws;
        p.gcnt = (uint16_t*)((char*)ws + (size_t)grid * nb * (nb / 32) * sizeof(uint32_t));
        return launch_nms_frames_big(p, grid, (cudaStream_t)stream);
    }
    const int nb = max_seg_len <= 32 ? 32 : (max_seg_len + 31) / 32 * 32;   // shared-memory capacity
    p.nb = nb;
    int nper = 1;                                                           // sort network: 32*nper >= nb
    while (32 * nper < nb) nper <<= 1;
    // a frame that fills at most 5/8 of the power-of-two network sorts two arrays instead (A = half the network,
    // B = an eighth of it: 300 boxes -> 256 + 64 keys), see nms_frames.cuh
    int npb = 0;
    // (measured on B200, 1000 frames x 30 classes: 300 boxes 0.326 ms against 0.351, 150 boxes 0.146 against 0.152;
    // a 3/4 split -- 190 boxes as 128 + 64, 380 as 256 + 128 -- is slower than the padded network and is not built)
    if ((nper == 8 || nper == 16) && nb <= 32 * (nper / 2 + nper / 8)) { npb = nper / 8; nper /= 2; }
    if (const char* e = getenv("VDET_NMS_NO_SPLIT")) {                      // measurement hook: the single-array network
        if (atoi(e) && npb) { nper *= 2; npb = 0; }
    }
    const int sort_blocks = npb ? nper + npb : nb / 32;
    p.so_words = sort_blocks * 33;
    // Stage scores when the block is box-major and the CTA still fits >= 2 per SM.  Residency: the
    // register budget allows 4 CTAs per SM (64 registers x 256 threads; 1 for the 1024-box variant);
    // shared memory decides the rest.  When staging every class at once would cost a CTA slot, the
    // classes are staged in up to 3 chunks (multiples of the warp count) instead.
    const bool want_stage = (score_ldr != 1);
    const size_t per_class = (size_t)(nb + 1) * sizeof(float);
    const size_t sm_smem = 228 * 1024, cta_reserved = 1024;
    int forced = 0;
    if (const char* e = getenv("VDET_NMS_PER_SM")) forced = atoi(e);      // measurement hook
    struct Plan { int stage, cls_chunk, per_sm; size_t smem; };
    auto plan_for = [&](const int warps, const int reg_limit) {
        Plan q;
        const size_t base = nms_smem_bytes(nb, sort_blocks, 0, false, warps);
        auto fit = [&](size_t smem_cta) {                   // CTAs of that size per SM
            int k = (int)(sm_smem / (smem_cta + cta_reserved));
            return k > reg_limit ? reg_limit : k;
        };
        q.stage = (want_stage && base + n_classes * per_class <= 100 * 1024) ? 1 : 0;
        q.cls_chunk = n_classes;
        q.per_sm = fit(base + (q.stage ? n_classes * per_class : 0));
        if (q.stage && n_classes > warps) {
            for (int want = reg_limit; want > q.per_sm; --want) {
                if (forced > 0 && want > forced) continue;
                const size_t budget = sm_smem / want - cta_reserved;
                if (budget <= base) continue;
                const int chunk_max = (int)((budget - base) / per_class);
                if (chunk_max < warps) continue;
                const int n_pass = (n_classes + chunk_max - 1) / chunk_max;
                if (n_pass > (VDET_NMS_CTAS_PER_SM > 4 ? 4 : 3)) continue;
                int chunk = (n_classes + n_pass - 1) / n_pass;
                const int rounded = (chunk + warps - 1) / warps * warps;
                if (rounded <= chunk_max) chunk = rounded;
                q.cls_chunk = chunk;
                q.per_sm = want;
                break;
            }
        }
        if (forced > 0 && q.per_sm > forced) q.per_sm = forced;
        if (q.per_sm < 1) q.per_sm = 1;
        q.smem = base + (q.stage ? (size_t)(q.cls_chunk < n_classes ? q.cls_chunk : n_classes) * per_class : 0);
        return q;
    };
    // CTA shape (nms_plan.h): the default, or -- when it fits with every class staged at once and the launch
    // quantises better on its 3-per-SM grid -- the wide one; frames of at most 320 boxes only (the kernels built).
    int threads = NMS_THREADS;
    Plan plan = plan_for(NMS_THREADS / 32, (nper > 16) ? 1 : VDET_NMS_CTAS_PER_SM);
    if (nper <= 8 && want_stage) {
        const Plan wide = plan_for(NMS_THREADS_WIDE / 32, NMS_CTAS_WIDE);
        const bool fits = wide.stage && wide.cls_chunk >= n_classes && wide.per_sm == NMS_CTAS_WIDE;
        bool take = fits && nms_prefer_wide(n_segs, n_classes, usable_sm_count());
        if (const char* e = getenv("VDET_NMS_THREADS")) {                 // measurement hook: 256 / 320
            const int t = atoi(e);
            if (t == NMS_THREADS) take = false;
            if (t == NMS_THREADS_WIDE) take = wide.stage != 0;
        }
        if (take) { threads = NMS_THREADS_WIDE; plan = wide; }
    }
    p.stage = plan.stage;
    p.cls_chunk = plan.cls_chunk;
    const size_t smem = plan.smem;
    int grid = usable_sm_count() * plan.per_sm;
    plan_items(p, grid);
    if (grid > p.n_items) grid = p.n_items;
    cudaStream_t st = (cudaStream_t)stream;
    if (npb) return launch_nms_frames_split(nper, npb, threads, p, smem, grid, st);
    switch (nper) {
        case 1:  return launch_nms_frames<1>(threads, p, smem, grid, st);
        case 2:  return launch_nms_frames<2>(threads, p, smem, grid, st);
        case 4:  return launch_nms_frames<4>(threads, p, smem, grid, st);
        case 8:  return launch_nms_frames<8>(threads, p, smem, grid, st);
        case 16: return launch_nms_frames<16>(threads, p, smem, grid, st);
        default: return launch_nms_frames<32>(threads, p, smem, grid, st);
    }
}
