// vid_nms.cu -- drop-in NMS entry points: nms / vid_nms / track_det_nms
// (utils/nms.pyx:17-68, :71-125, :128-189) built from the per-frame kernel of nms_frames.cu.
//
// vid_nms(dets[M,6]) in the reference is one O(M^2) loop that skips cross-frame pairs
// (nms.pyx:110-112).  Here: (1) rows are grouped by frame with a stable radix sort of the
// frame column, (2) every frame is solved by one CTA of nms_frames_kernel, (3) the kept rows
// are put into the reference's GLOBAL descending-score order (nms.pyx:80: keep is appended
// while walking one global argsort) with a stable radix sort of the score keys over rows in
// ascending row order -- ties therefore resolve to "ascending original row", the documented
// deterministic rule (the reference's own tie order is numpy-build dependent).
#include "primitives.cuh"

namespace vdet {

constexpr uint32_t KEY_SENTINEL = 0xffffffffu;

struct SegCounters { int32_t n_packed; int32_t n_segs; int32_t max_len; int32_t pad; };

__global__ void k_frame_keys(const float* __restrict__ frames, int ld, int64_t n,
                             const uint8_t* __restrict__ row_valid, uint32_t* __restrict__ keys,
                             uint32_t* __restrict__ vals, SegCounters* cnt) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) { cnt->n_packed = (int32_t)n; cnt->n_segs = 0; cnt->max_len = 0; }
    if (p >= n) return;
    const bool valid = row_valid ? (row_valid[p] != 0) : true;
    keys[p] = valid ? f32_key_asc(__ldg(frames + p * (int64_t)ld)) : KEY_SENTINEL;
    vals[p] = (uint32_t)p;
}

// head[p] = 1 where a new frame starts among the valid (non-sentinel) sorted keys.
__global__ void k_seg_heads(const uint32_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ head,
                            SegCounters* cnt) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t k = keys[p];
    const uint32_t prev = p > 0 ? keys[p - 1] : ~k;
    if (k == KEY_SENTINEL) {
        if (p == 0 || prev != KEY_SENTINEL) cnt->n_packed = (int32_t)p;   // first dropped row
        head[p] = 0;
    } else {
        head[p] = (p == 0 || k != prev) ? 1u : 0u;
    }
}

__global__ void k_seg_offsets(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ head_scan,
                              const uint32_t* __restrict__ vals, const float* __restrict__ frames, int ld,
                              int64_t n, const uint32_t* __restrict__ n_heads,
                              int32_t* __restrict__ seg_offsets, float* __restrict__ seg_frame,
                              SegCounters* cnt) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t k = keys[p];
    if (k != KEY_SENTINEL && (p == 0 || k != keys[p - 1])) {
        const uint32_t s = head_scan[p];
        seg_offsets[s] = (int32_t)p;
        if (seg_frame) seg_frame[s] = __ldg(frames + (int64_t)vals[p] * ld);
    }
    if (p == 0) {
        const uint32_t S = *n_heads;
        cnt->n_segs = (int32_t)S;
    }
}

__global__ void k_seg_finish(int32_t* __restrict__ seg_offsets, int64_t n, SegCounters* cnt) {
    // runs after k_seg_offsets: close the last segment and find the longest
    const int S = cnt->n_segs;
    if (blockIdx.x == 0 && threadIdx.x == 0) seg_offsets[S] = cnt->n_packed;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        const int32_t end = (s + 1 < S) ? seg_offsets[s + 1] : cnt->n_packed;
        atomicMax(&cnt->max_len, end - seg_offsets[s]);
    }
}

__global__ void k_set_two(int32_t* p, int32_t a, int32_t b) { p[0] = a; p[1] = b; }

__global__ void k_flag_kept(const int32_t* __restrict__ keep_idx, int64_t n, uint32_t* __restrict__ flag) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int32_t r = keep_idx[p];
    if (r >= 0) flag[r] = 1u;
}

__global__ void k_emit_pairs(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                             const float* __restrict__ scores, int ld, int64_t n,
                             uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || !flag[r]) return;
    const uint32_t o = pos[r];
    keys[o] = f32_key_desc(__ldg(scores + r * (int64_t)ld));
    vals[o] = (uint32_t)r;
}

__global__ void k_u32_to_i64(const uint32_t* __restrict__ v, int64_t n, int64_t* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) out[p] = (int64_t)v[p];
}

__global__ void k_i32_to_i64(const int32_t* __restrict__ v, const int32_t* __restrict__ cnt, int64_t n,
                             int64_t* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n && p < *cnt) out[p] = (int64_t)v[p];
}

// nms.pyx:163-183: det i is dropped when it overlaps ANY same-frame track box.
__global__ void k_track_round1(const float* __restrict__ tracks, int64_t q, int tld,
                               const float* __restrict__ dets, int64_t k, int dld, float T,
                               uint8_t* __restrict__ valid, uint32_t* status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const float* d = dets + i * dld;
    const float f = d[0];
    const float4 bi = make_float4(d[1], d[2], d[3], d[4]);
    const float ai = area_f32(bi);
    bool sup = false;
    for (int64_t j = 0; j < q; ++j) {
        const float* t = tracks + j * tld;
        if (f != t[0]) continue;
        const float4 bj = make_float4(t[1], t[2], t[3], t[4]);
        float inter, uni;
        inter_union_f32(bi, ai, bj, area_f32(bj), inter, uni);
        if (uni == 0.0f) { atomicOr(status, VDET_STATUS_ZERO_DIVISION); break; }
        if (iou_ge(inter, uni, T)) { sup = true; break; }
    }
    valid[i] = sup ? 0 : 1;
}


__global__ void k_score_keys(const float* __restrict__ scores, int ld, int64_t n, uint32_t* __restrict__ keys,
                             uint32_t* __restrict__ vals) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    keys[p] = f32_key_desc(__ldg(scores + p * (int64_t)ld));
    vals[p] = (uint32_t)p;
}

// frame keys of rows taken in the order `vals` (already sorted by score)
__global__ void k_frame_keys_perm(const float* __restrict__ frames, int ld, int64_t n,
                                  const uint8_t* __restrict__ row_valid, const uint32_t* __restrict__ vals,
                                  uint32_t* __restrict__ keys, SegCounters* cnt) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) { cnt->n_packed = (int32_t)n; cnt->n_segs = 0; cnt->max_len = 0; }
    if (p >= n) return;
    const uint32_t r = vals[p];
    const bool valid = row_valid ? (row_valid[r] != 0) : true;
    keys[p] = valid ? f32_key_asc(__ldg(frames + (int64_t)r * ld)) : KEY_SENTINEL;
}

__global__ void k_gather_boxes(const float* __restrict__ boxes, int ld, const int32_t* __restrict__ row_ids,
                               int64_t n, float4* __restrict__ out_box, float* __restrict__ out_area) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float* b = boxes + (int64_t)row_ids[p] * ld;
    const float4 v = make_float4(__ldg(b), __ldg(b + 1), __ldg(b + 2), __ldg(b + 3));
    out_box[p] = v;
    out_area[p] = area_f32(v);
}

// Any-length frames (more boxes than the bit-matrix kernels hold): one CTA per frame walks the
// frame's boxes in descending score exactly like nms.pyx:43-66 -- a kept box i is tested on
// the fly against every later box that is still alive (the CTA's threads stride over them),
// so the pairs evaluated are precisely the pairs the reference visits, no bit matrix is stored
// and the removed set (n bits) lives in shared memory.
constexpr int OTF_THREADS = 512;

__global__ void __launch_bounds__(OTF_THREADS) nms_segment_otf_kernel(const float4* __restrict__ sbox,
                                                                      const float* __restrict__ sarea,
                                                                      const int32_t* __restrict__ row_ids,
                                                                      const int32_t* __restrict__ seg_offsets,
                                                                      float T, int32_t* __restrict__ keep_idx,
                                                                      int32_t* __restrict__ keep_cnt,
                                                                      uint32_t* status) {
    extern __shared__ uint32_t s_removed[];
    const int seg = blockIdx.x;
    const int off = seg_offsets[seg];
    const int n = seg_offsets[seg + 1] - off;
    const int tid = threadIdx.x;
    for (int w = tid; w < (n + 31) / 32; w += OTF_THREADS) s_removed[w] = 0;
    __syncthreads();
    int cnt = 0;
    for (int k = 0; k < n; ++k) {
        if ((s_removed[k >> 5] >> (k & 31)) & 1u) continue;            // uniform across the CTA
        if (tid == 0) keep_idx[off + cnt] = row_ids[off + k];
        ++cnt;
        const float4 bi = sbox[off + k];
        const float ai = sarea[off + k];
        for (int q = k + 1 + tid; q < n; q += OTF_THREADS) {
            if ((s_removed[q >> 5] >> (q & 31)) & 1u) continue;
            float inter, uni;
            inter_union_f32(bi, ai, sbox[off + q], sarea[off + q], inter, uni);
            if (uni == 0.0f) atomicOr(status, VDET_STATUS_ZERO_DIVISION);
            else if (iou_ge(inter, uni, T)) atomicOr(&s_removed[q >> 5], 1u << (q & 31));
        }
        __syncthreads();
    }
    for (int e = cnt + tid; e < n; e += OTF_THREADS) keep_idx[off + e] = -1;
    if (tid == 0) keep_cnt[seg] = cnt;
}

static inline unsigned blocks_for(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// Workspace layout of the segmenting step.
struct SegWs {
    uint32_t *keys, *keys_alt, *vals_alt, *head, *radix;
    SegCounters* cnt;
    uint32_t* n_heads;
};

static bool carve_seg(WsCarver& c, int64_t n, SegWs& w) {
    w.keys = c.take<uint32_t>(n + 1);
    w.keys_alt = c.take<uint32_t>(n + 1);
    w.vals_alt = c.take<uint32_t>(n + 1);
    w.head = c.take<uint32_t>(n + 1);
    w.radix = c.take<uint32_t>(radix_scratch_elems(n) + scan_scratch_elems(n));
    w.cnt = c.take<SegCounters>(1);
    w.n_heads = c.take<uint32_t>(1);
    return c.ok();
}

static size_t seg_ws_bytes(int64_t n) {
    return 4 * (size_t)(n + 1) * 4 + (radix_scratch_elems(n) + scan_scratch_elems(n)) * 4 + 8 * 256 + 64;
}


// ---- stable order of rows by DESCENDING score (ties: ascending row), float32 or float64 ------
// float64 keys are 64 bits: an LSD pair of 32-bit sorts (low word first, then the high word).
__device__ __forceinline__ uint64_t f64_key_desc(double s) {
    s = __dadd_rn(s, 0.0);                                   // -0.0 -> +0.0
    const uint64_t b = (uint64_t)__double_as_longlong(s);
    const uint64_t asc = b ^ ((b >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
    return ~asc;
}

__global__ void k_score_keys64_lo(const double* __restrict__ scores, int ld, int64_t n, uint32_t* __restrict__ keys,
                                  uint32_t* __restrict__ vals) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    keys[p] = (uint32_t)f64_key_desc(scores[p * (int64_t)ld]);
    vals[p] = (uint32_t)p;
}

__global__ void k_score_keys64_hi(const double* __restrict__ scores, int ld, int64_t n,
                                  const uint32_t* __restrict__ vals, uint32_t* __restrict__ keys) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    keys[p] = (uint32_t)(f64_key_desc(scores[(int64_t)vals[p] * ld]) >> 32);
}

// On return `vals` (length n) holds the rows in descending-score order.  keys/keys_alt/vals_alt/radix
// are scratch as in SegWs.
static int order_by_score(const void* scores, int sld, int dtype, int64_t n, uint32_t* keys, uint32_t* vals,
                          uint32_t* keys_alt, uint32_t* vals_alt, uint32_t* radix, cudaStream_t st) {
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (dtype == VDET_DTYPE_F32) {
        k_score_keys<<<nb, 256, 0, st>>>((const float*)scores, sld, n, keys, vals);
        VDET_LAUNCH_CHECK();
        const int f = radix_sort_pairs(keys, vals, keys_alt, vals_alt, n, 0, 32, radix, st);
        if (f != 0) { if (f > 0) set_error("sort: unexpected parity"); return f < 0 ? f : VDET_ERR_INVALID; }
        return VDET_OK;
    }
    k_score_keys64_lo<<<nb, 256, 0, st>>>((const double*)scores, sld, n, keys, vals);
    VDET_LAUNCH_CHECK();
    int f = radix_sort_pairs(keys, vals, keys_alt, vals_alt, n, 0, 32, radix, st);
    if (f != 0) { if (f > 0) set_error("sort: unexpected parity"); return f < 0 ? f : VDET_ERR_INVALID; }
    k_score_keys64_hi<<<nb, 256, 0, st>>>((const double*)scores, sld, n, vals, keys);
    VDET_LAUNCH_CHECK();
    f = radix_sort_pairs(keys, vals, keys_alt, vals_alt, n, 0, 32, radix, st);
    if (f != 0) { if (f > 0) set_error("sort: unexpected parity"); return f < 0 ? f : VDET_ERR_INVALID; }
    return VDET_OK;
}

// Asynchronous part of vdet_segment_by_frame; counters stay on the device in w.cnt.
// With `scores` != nullptr the rows are first put in descending-score order (stable), so that
// after the stable sort by frame every segment is internally in descending score with ties
// by ascending row -- the order the any-length path walks.
static int segment_async(const float* frames, int ld, int64_t n, const uint8_t* row_valid,
                         int32_t* row_ids_out, int32_t* seg_offsets_out, float* seg_frame_out,
                         SegWs& w, cudaStream_t st, const void* scores = nullptr, int sld = 0,
                         int sdtype = VDET_DTYPE_F32) {
    if (scores != nullptr) {
        const int rc0 = order_by_score(scores, sld, sdtype, n, w.keys, (uint32_t*)row_ids_out, w.keys_alt, w.vals_alt,
                                       w.radix, st);
        if (rc0 != VDET_OK) return rc0;
        k_frame_keys_perm<<<blocks_for(n), 256, 0, st>>>(frames, ld, n, row_valid, (const uint32_t*)row_ids_out,
                                                         w.keys, w.cnt);
    } else {
        k_frame_keys<<<blocks_for(n), 256, 0, st>>>(frames, ld, n, row_valid, w.keys, (uint32_t*)row_ids_out, w.cnt);
    }
    VDET_LAUNCH_CHECK();
    int flip = radix_sort_pairs(w.keys, (uint32_t*)row_ids_out, w.keys_alt, w.vals_alt, n, 0, 32, w.radix, st);
    if (flip < 0) return flip;
    if (flip != 0) { set_error("segment: unexpected sort parity"); return VDET_ERR_INVALID; }
    k_seg_heads<<<blocks_for(n), 256, 0, st>>>(w.keys, n, w.head, w.cnt);
    VDET_LAUNCH_CHECK();
    // head -> exclusive scan in keys_alt (free after the sort)
    int rc = exclusive_scan_u32(w.head, w.keys_alt, n, w.n_heads, w.radix, st);
    if (rc != VDET_OK) return rc;
    k_seg_offsets<<<blocks_for(n), 256, 0, st>>>(w.keys, w.keys_alt, (const uint32_t*)row_ids_out, frames, ld, n,
                                                 w.n_heads, seg_offsets_out, seg_frame_out, w.cnt);
    VDET_LAUNCH_CHECK();
    k_seg_finish<<<64, 256, 0, st>>>(seg_offsets_out, n, w.cnt);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

}  // namespace vdet

using namespace vdet;

extern "C" size_t vdet_segment_workspace_bytes(int64_t n) { return seg_ws_bytes(n > 0 ? n : 1); }

extern "C" int vdet_segment_by_frame(const float* frames, int ld, int64_t n, const uint8_t* row_valid,
                                     const void* scores, int scores_ld, int scores_dtype,
                                     int32_t* row_ids_out, int32_t* seg_offsets_out, float* seg_frame_out,
                                     int32_t* n_segs_host, int32_t* max_seg_len_host, int64_t* n_packed_host,
                                     void* ws, size_t ws_bytes, void* stream) {
    VDET_REQUIRE(n >= 0 && n < 0x7fffffff && ld >= 1, "segment_by_frame: bad size");
    VDET_REQUIRE(scores_dtype == VDET_DTYPE_F32 || scores_dtype == VDET_DTYPE_F64, "segment_by_frame: bad dtype");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        *n_segs_host = 0; *max_seg_len_host = 0; *n_packed_host = 0;
        k_set_two<<<1, 1, 0, st>>>(seg_offsets_out, 0, 0);
        VDET_LAUNCH_CHECK();
        return VDET_OK;
    }
    WsCarver c(ws, ws_bytes);
    SegWs w;
    if (!carve_seg(c, n, w)) { set_error("segment_by_frame: workspace too small"); return VDET_ERR_WORKSPACE; }
    int rc = segment_async(frames, ld, n, row_valid, row_ids_out, seg_offsets_out, seg_frame_out, w, st, scores, scores_ld,
                           scores_dtype);
    if (rc != VDET_OK) return rc;
    SegCounters h;
    VDET_CUDA(cudaMemcpyAsync(&h, w.cnt, sizeof(h), cudaMemcpyDeviceToHost, st));
    VDET_CUDA(cudaStreamSynchronize(st));
    *n_segs_host = h.n_segs; *max_seg_len_host = h.max_len; *n_packed_host = h.n_packed;
    return VDET_OK;
}

extern "C" size_t vdet_nms_workspace_bytes(int64_t n, int device) {
    (void)device;
    if (n < 1) n = 1;
    // segmenting + row_ids/seg_offsets/keep_idx/keep_cnt/flag/pos/k2/v2/k2a/v2a + valid + status
    size_t b = seg_ws_bytes(n) + 10 * (size_t)(n + 2) * 4 + (size_t)n + radix_scratch_elems(n) * 4 +
               scan_scratch_elems(n) * 4 + 40 * 256;
    b += (size_t)(n + 2) * 20;                                   // gathered boxes + areas (any-length path)
    if (n > 1024) b += vdet_nms_frames_workspace_bytes((int)(n < 2048 ? n : 2048), 1, device);
    return b;
}

namespace {

// Shared tail of nms / vid_nms / track_det_nms.  frames == nullptr: one frame.
int64_t nms_core(const float* dets, int64_t n, int ld, const float* frames, int box_col, int score_col,
                 const uint8_t* row_valid, double thresh, int64_t* keep, uint32_t* status_host,
                 WsCarver& c, cudaStream_t st) {
    int32_t* row_ids = c.take<int32_t>(n + 1);
    int32_t* seg_offsets = c.take<int32_t>(n + 2);
    int32_t* keep_idx = c.take<int32_t>(n + 1);
    int32_t* keep_cnt = c.take<int32_t>(n + 1);
    uint32_t* flag = c.take<uint32_t>(n + 1);
    uint32_t* pos = c.take<uint32_t>(n + 1);
    uint32_t* k2 = c.take<uint32_t>(n + 1);
    uint32_t* v2 = c.take<uint32_t>(n + 1);
    uint32_t* k2a = c.take<uint32_t>(n + 1);
    uint32_t* v2a = c.take<uint32_t>(n + 1);
    uint32_t* radix = c.take<uint32_t>(radix_scratch_elems(n) + scan_scratch_elems(n));
    uint32_t* status = c.take<uint32_t>(1);
    uint32_t* total = c.take<uint32_t>(1);
    SegWs w;
    if (!carve_seg(c, n, w)) { set_error("nms: workspace too small"); return VDET_ERR_WORKSPACE; }

    VDET_CUDA(cudaMemsetAsync(status, 0, sizeof(uint32_t), st));
    int32_t n_segs = 1, max_len = (int32_t)n;
    int64_t n_packed = n;
    const int32_t* row_ids_arg = nullptr;
    if (frames != nullptr || row_valid != nullptr) {
        // frames == nullptr with a row filter: every row is frame 0 (stride-0 read of dets[0])
        const float* fr = frames ? frames : dets;
        const int fld = frames ? ld : 0;
        int rc = segment_async(fr, fld, n, row_valid, row_ids, seg_offsets, nullptr, w, st);
        if (rc != VDET_OK) return rc;
        SegCounters h;
        VDET_CUDA(cudaMemcpyAsync(&h, w.cnt, sizeof(h), cudaMemcpyDeviceToHost, st));
        VDET_CUDA(cudaStreamSynchronize(st));
        n_segs = h.n_segs; max_len = h.max_len; n_packed = h.n_packed;
        row_ids_arg = row_ids;
    } else {
        k_set_two<<<1, 1, 0, st>>>(seg_offsets, 0, (int32_t)n);
        VDET_LAUNCH_CHECK();
    }
    if (n_packed == 0 || n_segs == 0) { *status_host = 0; return 0; }

    int rc;
    if (max_len <= 2048) {
        void* fws = nullptr;
        size_t fws_bytes = 0;
        if (max_len > 1024) {
            fws_bytes = vdet_nms_frames_workspace_bytes(max_len, 1, 0);
            fws = c.take<char>(fws_bytes);
            if (!c.ok()) { set_error("nms: workspace too small"); return VDET_ERR_WORKSPACE; }
        }
        rc = vdet_nms_frames_f32(dets + box_col, ld, dets + score_col, ld, 0, seg_offsets, n_segs, max_len,
                                 row_ids_arg, 1, thresh, keep_idx, keep_cnt, nullptr, n_packed, VDET_LAYOUT_CLASS_MAJOR,
                                 status, fws, fws_bytes, st);
        if (rc != VDET_OK) return rc;
    } else {
        // any-length path: order every frame by descending score, gather, walk on the fly
        float4* gbox = c.take<float4>(n + 1);
        float* garea = c.take<float>(n + 1);
        if (!c.ok()) { set_error("nms: workspace too small"); return VDET_ERR_WORKSPACE; }
        const float* fr = frames ? frames : dets;
        const int fld = frames ? ld : 0;
        rc = segment_async(fr, fld, n, row_valid, row_ids, seg_offsets, nullptr, w, st, dets + score_col, ld);
        if (rc != VDET_OK) return rc;
        k_gather_boxes<<<blocks_for(n_packed), 256, 0, st>>>(dets + box_col, ld, row_ids, n_packed, gbox, garea);
        VDET_LAUNCH_CHECK();
        const size_t smem = ((size_t)max_len + 31) / 32 * sizeof(uint32_t);
        if (smem > max_dynamic_smem(nms_segment_otf_kernel)) {
            set_error("nms: a frame with %d boxes exceeds what this build handles", max_len);
            return VDET_ERR_UNSUPPORTED;
        }
        VDET_CUDA(allow_dynamic_smem(nms_segment_otf_kernel, smem));
        nms_segment_otf_kernel<<<n_segs, OTF_THREADS, smem, st>>>(gbox, garea, row_ids, seg_offsets,
                                                                  thresh_ceil_f32(thresh), keep_idx, keep_cnt, status);
        VDET_LAUNCH_CHECK();
    }

    uint32_t h_total = 0, h_status = 0;
    if (n_segs == 1) {
        // one frame: keep_idx[0..cnt) is already the descending-score keep list
        k_i32_to_i64<<<blocks_for(n_packed), 256, 0, st>>>(keep_idx, keep_cnt, n_packed, keep);
        VDET_LAUNCH_CHECK();
        int32_t h_cnt = 0;
        VDET_CUDA(cudaMemcpyAsync(&h_cnt, keep_cnt, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        VDET_CUDA(cudaMemcpyAsync(&h_status, status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        VDET_CUDA(cudaStreamSynchronize(st));
        h_total = (uint32_t)h_cnt;
    } else {
        VDET_CUDA(cudaMemsetAsync(flag, 0, sizeof(uint32_t) * (size_t)n, st));
        k_flag_kept<<<blocks_for(n_packed), 256, 0, st>>>(keep_idx, n_packed, flag);
        VDET_LAUNCH_CHECK();
        rc = exclusive_scan_u32(flag, pos, n, total, radix, st);
        if (rc != VDET_OK) return rc;
        k_emit_pairs<<<blocks_for(n), 256, 0, st>>>(flag, pos, dets + score_col, ld, n, k2, v2);
        VDET_LAUNCH_CHECK();
        VDET_CUDA(cudaMemcpyAsync(&h_total, total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        VDET_CUDA(cudaMemcpyAsync(&h_status, status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        VDET_CUDA(cudaStreamSynchronize(st));
        int flip = radix_sort_pairs(k2, v2, k2a, v2a, (int64_t)h_total, 0, 32, radix, st);
        if (flip < 0) return flip;
        k_u32_to_i64<<<blocks_for(h_total > 0 ? h_total : 1), 256, 0, st>>>(flip ? v2a : v2, (int64_t)h_total, keep);
        VDET_LAUNCH_CHECK();
        VDET_CUDA(cudaStreamSynchronize(st));
    }
    if (h_status & 0x80000000u) { set_error("nms: internal frame-length mismatch"); return VDET_ERR_INVALID; }
    *status_host = h_status;
    return (int64_t)h_total;
}

}  // namespace

extern "C" int64_t vdet_nms_f32(const float* dets, int64_t n, int ld, double thresh,
                                int64_t* keep, uint32_t* status_host,
                                void* ws, size_t ws_bytes, void* stream) {
    VDET_REQUIRE(n >= 0 && n < 0x7fffffff && ld >= 5, "nms: bad size");
    *status_host = 0;
    if (n == 0) return 0;
    WsCarver c(ws, ws_bytes);
    return nms_core(dets, n, ld, nullptr, 0, 4, nullptr, thresh, keep, status_host, c, (cudaStream_t)stream);
}

extern "C" int64_t vdet_vid_nms_f32(const float* dets, int64_t n, int ld, double thresh,
                                    int64_t* keep, uint32_t* status_host,
                                    void* ws, size_t ws_bytes, void* stream) {
    VDET_REQUIRE(n >= 0 && n < 0x7fffffff && ld >= 6, "vid_nms: bad size");
    *status_host = 0;
    if (n == 0) return 0;
    WsCarver c(ws, ws_bytes);
    return nms_core(dets, n, ld, dets, 1, 5, nullptr, thresh, keep, status_host, c, (cudaStream_t)stream);
}

extern "C" int64_t vdet_track_det_nms_f32(const float* tracks, int64_t q, int tracks_ld,
                                          const float* dets, int64_t k, int dets_ld, double thresh,
                                          int64_t* keep, uint32_t* status_host,
                                          void* ws, size_t ws_bytes, void* stream) {
    VDET_REQUIRE(q >= 0 && k >= 0 && k < 0x7fffffff && tracks_ld >= 5 && dets_ld >= 6, "track_det_nms: bad size");
    *status_host = 0;
    if (k == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    WsCarver c(ws, ws_bytes);
    uint8_t* valid = c.take<uint8_t>(k + 1);
    uint32_t* status1 = c.take<uint32_t>(1);
    if (!c.ok()) { set_error("track_det_nms: workspace too small"); return VDET_ERR_WORKSPACE; }
    VDET_CUDA(cudaMemsetAsync(status1, 0, sizeof(uint32_t), st));
    k_track_round1<<<blocks_for(k), 256, 0, st>>>(tracks, q, tracks_ld, dets, k, dets_ld,
                                                  thresh_ceil_f32(thresh), valid, status1);
    VDET_LAUNCH_CHECK();
    uint32_t s2 = 0;
    int64_t r = nms_core(dets, k, dets_ld, dets, 1, 5, valid, thresh, keep, &s2, c, st);
    uint32_t s1 = 0;
    VDET_CUDA(cudaMemcpy(&s1, status1, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    *status_host = s1 | s2;
    return r;
}


// ---- stable descending sort of (score, id) pairs (cross-rank keep-list merge, top_detections) -----
extern "C" size_t vdet_sort_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    return 4 * (size_t)(n + 1) * 4 + radix_scratch_elems(n) * 4 + 8 * 256;
}

namespace vdet {
template <typename T>
__global__ void k_sort_gather_t(const uint32_t* __restrict__ perm, const T* __restrict__ scores,
                                const int64_t* __restrict__ ids, int64_t n, T* __restrict__ scores_out,
                                int64_t* __restrict__ ids_out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t q = perm[p];
    scores_out[p] = scores[q];
    ids_out[p] = ids[q];
}
}  // namespace vdet

extern "C" int vdet_sort_by_score_desc(const void* scores, int dtype, const int64_t* ids, int64_t n,
                                       void* scores_out, int64_t* ids_out,
                                       void* ws, size_t ws_bytes, void* stream) {
    VDET_REQUIRE(n >= 0 && n < 0x7fffffff, "sort_by_score: bad size");
    VDET_REQUIRE(dtype == VDET_DTYPE_F32 || dtype == VDET_DTYPE_F64, "sort_by_score: bad dtype");
    if (n == 0) return VDET_OK;
    VDET_REQUIRE(scores != scores_out && ids != ids_out, "sort_by_score: in-place sort is not supported");
    cudaStream_t st = (cudaStream_t)stream;
    WsCarver c(ws, ws_bytes);
    uint32_t* k = c.take<uint32_t>(n + 1);
    uint32_t* v = c.take<uint32_t>(n + 1);
    uint32_t* ka = c.take<uint32_t>(n + 1);
    uint32_t* va = c.take<uint32_t>(n + 1);
    uint32_t* radix = c.take<uint32_t>(radix_scratch_elems(n));
    if (!c.ok()) { set_error("sort_by_score: workspace too small"); return VDET_ERR_WORKSPACE; }
    const int rc = order_by_score(scores, 1, dtype, n, k, v, ka, va, radix, st);
    if (rc != VDET_OK) return rc;
    if (dtype == VDET_DTYPE_F32)
        k_sort_gather_t<float><<<blocks_for(n), 256, 0, st>>>(v, (const float*)scores, ids, n, (float*)scores_out, ids_out);
    else
        k_sort_gather_t<double><<<blocks_for(n), 256, 0, st>>>(v, (const double*)scores, ids, n, (double*)scores_out, ids_out);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}
