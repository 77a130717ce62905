// temporal.cu -- temporal score smoothing along the frame axis (K4, K5, K6 of SURVEY 2.1).
//
//   score completion  : do_score_completion            vdet/tubelet_cls.py:284-303
//   temporal max-pool : score_proto_temporal_maxpool   vdet/tubelet_cls.py:386-414
//   temporal conv     : depthwise 1-D convolution, the build-defined stand-in for
//                       score_conv_cls (vdet/tubelet_cls.py:15-51; its Caffe net is not part
//                       of the reference)
//
// Data: [n_rows, L] score rows (one row per tubelet x class), row pitch `ld`, optional ragged
// lengths.  All three are HBM streaming kernels (read 1 element, write 1 element, <= 2w
// flops): rows are cut into tiles, a tile plus its halo is staged once in shared memory with
// coalesced loads, and every output is produced from shared memory and written coalesced.
// Arithmetic is done in the row dtype with individually rounded operations (the library is
// built with -fmad=false) so float64 rows reproduce the reference's Python-float results bit
// for bit.
#include "common.cuh"

namespace vdet {

constexpr int TP_THREADS = 256;
constexpr int TP_ITEMS = 4;
constexpr int TP_TILE = TP_THREADS * TP_ITEMS;

template <typename T> __device__ __forceinline__ T t_max(T a, T b) { return a > b ? a : b; }
template <typename T> __device__ __forceinline__ T t_mul(T a, T b);
template <> __device__ __forceinline__ float t_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double t_mul<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T t_add(T a, T b);
template <> __device__ __forceinline__ float t_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double t_add<double>(double a, double b) { return __dadd_rn(a, b); }
template <typename T> __device__ __forceinline__ T t_sub(T a, T b);
template <> __device__ __forceinline__ float t_sub<float>(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double t_sub<double>(double a, double b) { return __dsub_rn(a, b); }
template <typename T> __device__ __forceinline__ T t_div(T a, T b);
template <> __device__ __forceinline__ float t_div<float>(float a, float b) { return __fdiv_rn(a, b); }
template <> __device__ __forceinline__ double t_div<double>(double a, double b) { return __ddiv_rn(a, b); }

// ---- temporal max-pool -------------------------------------------------------------------
// out[i] = max(in[i-h .. i+h]) with out-of-range samples = pad (tubelet_cls.py:402-409).
template <typename T>
__global__ void __launch_bounds__(TP_THREADS) temporal_maxpool_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                                      int64_t L, int64_t ld,
                                                                      const int32_t* __restrict__ lengths,
                                                                      int tiles_per_row, int h, T pad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s = reinterpret_cast<T*>(smem_raw);
    const int64_t row = blockIdx.x / tiles_per_row;
    const int tile = blockIdx.x - (int)(row * tiles_per_row);
    const int64_t len = lengths ? (int64_t)lengths[row] : L;
    const int64_t t0 = (int64_t)tile * TP_TILE;
    if (t0 >= len) return;
    const T* src = in + row * ld;
    const int span = TP_TILE + 2 * h;
    for (int e = threadIdx.x; e < span; e += TP_THREADS) {
        const int64_t g = t0 - h + e;
        s[e] = (g >= 0 && g < len) ? src[g] : pad;
    }
    __syncthreads();
    T* dst = out + row * ld;
#pragma unroll
    for (int q = 0; q < TP_ITEMS; ++q) {
        const int o = q * TP_THREADS + threadIdx.x;
        if (t0 + o < len) {
            T m = s[o];
            for (int k = 1; k <= 2 * h; ++k) m = t_max(m, s[o + k]);
            dst[t0 + o] = m;
        }
    }
}

// ---- depthwise temporal convolution ------------------------------------------------------
// out[i] = sum_k taps[ch][k] * x[i + k - h], accumulated left to right from 0, separate
// multiply and add; out-of-range samples are 0 (VDET_PAD_ZERO) or the edge sample.
template <typename T>
__global__ void __launch_bounds__(TP_THREADS) temporal_conv1d_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                                     int64_t L, int64_t ld,
                                                                     const int32_t* __restrict__ lengths,
                                                                     int tiles_per_row, const T* __restrict__ taps,
                                                                     int n_channels, int w, int pad_mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s = reinterpret_cast<T*>(smem_raw);
    const int h = w / 2;
    const int span = TP_TILE + 2 * h;
    T* s_tap = s + span;
    const int64_t row = blockIdx.x / tiles_per_row;
    const int tile = blockIdx.x - (int)(row * tiles_per_row);
    const int64_t len = lengths ? (int64_t)lengths[row] : L;
    const int64_t t0 = (int64_t)tile * TP_TILE;
    if (t0 >= len) return;
    const T* src = in + row * ld;
    const int ch = (int)(row % n_channels);
    for (int e = threadIdx.x; e < w; e += TP_THREADS) s_tap[e] = taps[(int64_t)ch * w + e];
    for (int e = threadIdx.x; e < span; e += TP_THREADS) {
        int64_t g = t0 - h + e;
        T v = (T)0;
        if (pad_mode == VDET_PAD_EDGE) {
            g = g < 0 ? 0 : (g >= len ? len - 1 : g);
            v = src[g];
        } else if (g >= 0 && g < len) {
            v = src[g];
        }
        s[e] = v;
    }
    __syncthreads();
    T* dst = out + row * ld;
#pragma unroll
    for (int q = 0; q < TP_ITEMS; ++q) {
        const int o = q * TP_THREADS + threadIdx.x;
        if (t0 + o < len) {
            T acc = (T)0;
            for (int k = 0; k < w; ++k) acc = t_add(acc, t_mul(s_tap[k], s[o + k]));
            dst[t0 + o] = acc;
        }
    }
}

template <typename T> struct Vec16;
template <> struct Vec16<float> { typedef float4 type; static constexpr int V = 4; };
template <> struct Vec16<double> { typedef double2 type; static constexpr int V = 2; };

// ---- score completion --------------------------------------------------------------------
// do_score_completion (tubelet_cls.py:284-303): every maximal run [i, j) of missing scores
// (score <= -10) is rewritten from its valid neighbours l = s[i-1], r = s[j]:
//   leading run  -> r                       (:293-295)
//   trailing run -> l                       (:296-298)
//   interior     -> l + (r - l) * (k - i + 1) / (j - i + 1)     (:299-303)
// The update is in place and valid scores never change, so the rows only have to be READ once and
// the ~5 % missing elements written.  Two kernels, race-free because the first one is the only one
// that classifies elements, before anything is rewritten:
//   1. completion_scan_kernel streams the rows (coalesced 4/8-byte loads, 32 in flight per lane) and
//      writes a "missing" BITMAP, one 32-bit word per 32 elements, straight from warp ballots (lane
//      stride 1 over the row, so a ballot IS the bitmap word): 1 bit per element of side traffic;
//   2. completion_fill_kernel reads the bitmap, one thread per word.  A run belongs to the word that
//      holds its first element (missing bit set, previous bit clear): that thread finds the run's end
//      in the bitmap alone, loads the two valid neighbours and writes the closed form.  Words without
//      a run start (84 % on BASELINE's rows) cost one load.
// Traffic per element: 4 B read + 2 bits of bitmap + ~5 % x (read neighbours, write value), against the
// 8 B of a read-everything / write-everything pass.
constexpr int CS_WARPS = 8;                     // warps per CTA of the scan kernel, one 1024-element span each

template <typename T>
__global__ void __launch_bounds__(CS_WARPS * 32) completion_scan_kernel(const T* __restrict__ scores, int64_t L, int64_t ld,
                                                                        const int32_t* __restrict__ lengths,
                                                                        int words_per_row, int spans_per_row,
                                                                        int64_t n_spans, T miss_thr,
                                                                        uint32_t* __restrict__ bitmap) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t span = (int64_t)blockIdx.x * CS_WARPS + warp; span < n_spans; span += (int64_t)gridDim.x * CS_WARPS) {
        const int64_t row = span / spans_per_row;
        const int sp = (int)(span - row * spans_per_row);
        const int len = (int)(lengths ? (int64_t)lengths[row] : L);
        const int e0 = sp * 1024;                               // 32 words of 32 elements
        const T* g = scores + row * ld;
        uint32_t mine = 0;                                       // lane j keeps word j of the span
#pragma unroll
        for (int jb = 0; jb < 32; jb += 8) {
            T v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int e = e0 + (jb + q) * 32 + lane;
                v[q] = e < len ? __ldg(g + e) : (T)0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int e = e0 + (jb + q) * 32 + lane;
                const uint32_t word = __ballot_sync(FULL, e < len && v[q] <= miss_thr);
                if (lane == jb + q) mine = word;
            }
        }
        const int w = sp * 32 + lane;
        if (w < words_per_row) bitmap[row * words_per_row + w] = mine;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) completion_fill_kernel(T* __restrict__ scores, int64_t L, int64_t ld,
                                                              const int32_t* __restrict__ lengths,
                                                              int words_per_row, int64_t n_words,
                                                              const uint32_t* __restrict__ bitmap,
                                                              const T* __restrict__ bounds,
                                                              uint32_t* status) {
    const int64_t wid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wid >= n_words) return;
    const uint32_t word = __ldg(bitmap + wid);
    if (word == 0u) return;
    const int64_t row = wid / words_per_row;
    const int wi = (int)(wid - row * words_per_row);
    const uint32_t carry = wi > 0 ? (__ldg(bitmap + wid - 1) >> 31) : 0u;
    uint32_t starts = word & ~((word << 1) | carry);            // missing, and the element before it is not
    if (starts == 0u) return;
    const int len = (int)(lengths ? (int64_t)lengths[row] : L);
    const uint32_t* brow = bitmap + row * words_per_row;
    T* g = scores + row * ld;
    while (starts) {
        const int sbit = __ffs(starts) - 1;
        starts &= starts - 1;
        const int i = wi * 32 + sbit;                            // run = [i, j)
        // end of the run: first clear bit at or after i (bits beyond the row are clear)
        int j;
        {
            uint32_t rest = ~word & ~((2u << sbit) - 1u);        // clear bits above sbit in this word
            if (sbit == 31) rest = 0u;
            if (rest) {
                j = wi * 32 + __ffs(rest) - 1;
            } else {
                int w2 = wi + 1;
                uint32_t inv = 0u;
                while (w2 < words_per_row && (inv = ~__ldg(brow + w2)) == 0u) ++w2;
                j = (w2 < words_per_row) ? w2 * 32 + __ffs(inv) - 1 : len;
            }
            if (j > len) j = len;
        }
        // the valid neighbours of the run: inside the row, or -- for a row that is one frame range of a longer
        // tubelet (frame-sharded completion) -- the nearest valid score of the neighbouring shards, given as
        // bounds[row] = (left gap, left value, right gap, right value): gap = missing frames between that score
        // and this shard's first / last column, < 0 = there is none.  Positions are those of the whole tubelet.
        bool has_l = i > 0, has_r = j < len;
        T l = (T)0, r = (T)0;
        int i_eff = i, j_eff = j;
        if (has_l) l = g[i - 1];
        else if (bounds != nullptr && bounds[row * 4 + 0] >= (T)0) { has_l = true; l = bounds[row * 4 + 1]; i_eff = -(int)bounds[row * 4 + 0]; }
        if (has_r) r = g[j];
        else if (bounds != nullptr && bounds[row * 4 + 2] >= (T)0) { has_r = true; r = bounds[row * 4 + 3]; j_eff = len + (int)bounds[row * 4 + 2]; }
        if (!has_l) {
            if (!has_r) { atomicOr(status, VDET_STATUS_ALL_MISSING); continue; }     // tubelet_cls.py:295 IndexError
            for (int k = i; k < j; ++k) g[k] = r;
        } else if (!has_r) {
            for (int k = i; k < j; ++k) g[k] = l;
        } else {
            const T d = t_sub(r, l), den = (T)(j_eff - i_eff + 1);
            for (int k = i; k < j; ++k) g[k] = t_add(l, t_div(t_mul(d, (T)(k - i_eff + 1)), den));
        }
    }
}

static inline bool per_row_fits(int64_t L) { return L / 4 < 0x7fffffff; }

// ---- register-window fast path (window <= 9, vector-aligned rows) ---------------------------
// The streaming kernels above stage a tile in shared memory and synchronise; for the small
// windows the reference actually uses (3..9) that costs more than it saves.  Here every thread
// produces G*V consecutive outputs (V = 16 bytes / element) straight from registers: it loads
// its own G vectors plus HV neighbour vectors on each side with 16-byte loads (neighbour
// vectors are L1 hits, their owners load them too), evaluates the window with fully unrolled
// register indexing, and writes 16-byte vectors.  No shared memory, no barrier, 8 loads in
// flight per thread.
template <typename T, int MODE /*0 pad value, 1 edge*/>
__device__ __forceinline__ void load_window_vec(const T* __restrict__ src, const int64_t len, const int64_t e0,
                                                const T padv, T* w) {
    constexpr int V = Vec16<T>::V;
    if (e0 >= 0 && e0 + V <= len) {
        const typename Vec16<T>::type v = __ldg(reinterpret_cast<const typename Vec16<T>::type*>(src + e0));
        const T* pv = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int c = 0; c < V; ++c) w[c] = pv[c];
    } else {
#pragma unroll
        for (int c = 0; c < V; ++c) {
            const int64_t g = e0 + c;
            if (MODE == 1) w[c] = src[g < 0 ? 0 : (g >= len ? len - 1 : g)];
            else w[c] = (g >= 0 && g < len) ? src[g] : padv;
        }
    }
}

template <typename T, int H, bool CONV>
__global__ void __launch_bounds__(256) temporal_window_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                              int64_t L, int64_t ld,
                                                              const int32_t* __restrict__ lengths,
                                                              int groups_per_row, T padv,
                                                              const T* __restrict__ taps, int n_channels,
                                                              int pad_mode) {
    constexpr int V = Vec16<T>::V;
    constexpr int G = 2;                       // own vectors per thread
    constexpr int HV = (H + V - 1) / V;        // neighbour vectors per side
    constexpr int NW = (G + 2 * HV) * V;       // window elements held in registers
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = gid / groups_per_row;
    const int64_t grp = gid - row * groups_per_row;
    const int64_t len = lengths ? (int64_t)lengths[row] : L;
    const int64_t o0 = grp * (G * V);          // first output element of this thread
    if (o0 >= len) return;
    const T* src = in + row * ld;
    T win[NW];
#pragma unroll
    for (int v = 0; v < G + 2 * HV; ++v) {
        const int64_t e0 = o0 + (int64_t)(v - HV) * V;
        if (CONV && pad_mode == VDET_PAD_EDGE) load_window_vec<T, 1>(src, len, e0, padv, win + v * V);
        else load_window_vec<T, 0>(src, len, e0, padv, win + v * V);
    }
    T tap[2 * H + 1];
    if (CONV) {
        const T* tp = taps + (row % n_channels) * (2 * H + 1);
#pragma unroll
        for (int k = 0; k < 2 * H + 1; ++k) tap[k] = __ldg(tp + k);
    }
    T res[G * V];
#pragma unroll
    for (int o = 0; o < G * V; ++o) {
        const int c = HV * V + o;              // centre of the window in win[]
        if (CONV) {
            T acc = (T)0;
#pragma unroll
            for (int k = 0; k < 2 * H + 1; ++k) acc = t_add(acc, t_mul(tap[k], win[c - H + k]));
            res[o] = acc;
        } else {
            T m = win[c - H];
#pragma unroll
            for (int k = 1; k < 2 * H + 1; ++k) m = t_max(m, win[c - H + k]);
            res[o] = m;
        }
    }
    T* dst = out + row * ld + o0;
#pragma unroll
    for (int v = 0; v < G; ++v) {
        if (o0 + (v + 1) * V <= len) {
            *reinterpret_cast<typename Vec16<T>::type*>(dst + v * V) =
                *reinterpret_cast<const typename Vec16<T>::type*>(res + v * V);
        } else {
#pragma unroll
            for (int c = 0; c < V; ++c)
                if (o0 + v * V + c < len) dst[v * V + c] = res[v * V + c];
        }
    }
}

template <typename T, bool CONV>
static int run_window_fast(const void* in, void* out, int64_t n_rows, int64_t L, int64_t ld, const int32_t* lengths,
                           int h, double pad, const void* taps, int n_channels, int pad_mode, cudaStream_t st) {
    constexpr int V = Vec16<T>::V;
    const int64_t per_row = (L + 2 * V - 1) / (2 * V);
    const int64_t total = per_row * n_rows;
    const int64_t grid = (total + 255) / 256;
    if (grid > 0x7fffffff) { set_error("temporal: grid too large"); return VDET_ERR_UNSUPPORTED; }
#define VDET_WIN_LAUNCH(HH) temporal_window_kernel<T, HH, CONV><<<(unsigned)grid, 256, 0, st>>>( \
        (const T*)in, (T*)out, L, ld, lengths, (int)per_row, (T)pad, (const T*)taps, n_channels, pad_mode)
    switch (h) {
        case 1: VDET_WIN_LAUNCH(1); break;
        case 2: VDET_WIN_LAUNCH(2); break;
        case 3: VDET_WIN_LAUNCH(3); break;
        default: VDET_WIN_LAUNCH(4); break;
    }
#undef VDET_WIN_LAUNCH
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

template <typename T>
static bool window_fast_ok(const void* in, const void* out, int64_t ld, int64_t L, int window) {
    constexpr int V = Vec16<T>::V;
    return window >= 3 && window <= 9 && (ld % V) == 0 && L < 0x7fffffff && per_row_fits(L) &&
           ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0;
}

template <typename T>
static int run_maxpool(const void* in, void* out, int64_t n_rows, int64_t L, int64_t ld, const int32_t* lengths,
                       int window, double pad, cudaStream_t st) {
    const int h = window / 2;
    if (window_fast_ok<T>(in, out, ld, L, window))
        return run_window_fast<T, false>(in, out, n_rows, L, ld, lengths, h, pad, nullptr, 1, 0, st);
    const size_t smem = (size_t)(TP_TILE + 2 * h) * sizeof(T);
    if (smem > max_dynamic_smem(temporal_maxpool_kernel<T>)) { set_error("temporal_maxpool: window %d too large", window); return VDET_ERR_UNSUPPORTED; }
    VDET_CUDA(allow_dynamic_smem(temporal_maxpool_kernel<T>, smem));
    const int64_t tiles = (L + TP_TILE - 1) / TP_TILE;
    const int64_t grid = tiles * n_rows;
    if (grid > 0x7fffffff) { set_error("temporal_maxpool: grid too large"); return VDET_ERR_UNSUPPORTED; }
    temporal_maxpool_kernel<T><<<(unsigned)grid, TP_THREADS, smem, st>>>((const T*)in, (T*)out, L, ld, lengths,
                                                                         (int)tiles, h, (T)pad);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

template <typename T>
static int run_conv(const void* in, void* out, int64_t n_rows, int64_t L, int64_t ld, const int32_t* lengths,
                    const void* taps, int n_channels, int window, int pad_mode, cudaStream_t st) {
    const int h = window / 2;
    if (window_fast_ok<T>(in, out, ld, L, window) && ((uintptr_t)taps & (sizeof(T) - 1)) == 0)
        return run_window_fast<T, true>(in, out, n_rows, L, ld, lengths, h, 0.0, taps, n_channels, pad_mode, st);
    const size_t smem = (size_t)(TP_TILE + 2 * h + window) * sizeof(T);
    if (smem > max_dynamic_smem(temporal_conv1d_kernel<T>)) { set_error("temporal_conv1d: window %d too large", window); return VDET_ERR_UNSUPPORTED; }
    VDET_CUDA(allow_dynamic_smem(temporal_conv1d_kernel<T>, smem));
    const int64_t tiles = (L + TP_TILE - 1) / TP_TILE;
    const int64_t grid = tiles * n_rows;
    if (grid > 0x7fffffff) { set_error("temporal_conv1d: grid too large"); return VDET_ERR_UNSUPPORTED; }
    temporal_conv1d_kernel<T><<<(unsigned)grid, TP_THREADS, smem, st>>>((const T*)in, (T*)out, L, ld, lengths,
                                                                        (int)tiles, (const T*)taps, n_channels,
                                                                        window, pad_mode);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

static inline size_t completion_bitmap_words(int64_t n_rows, int64_t L) {
    return (size_t)(n_rows > 0 ? n_rows : 0) * (size_t)((L + 31) / 32);
}

template <typename T>
static int run_completion(void* scores, int64_t n_rows, int64_t L, int64_t ld, const int32_t* lengths,
                          double miss_thr, const void* bounds, uint32_t* status, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (L >= 0x7fffffff - 1024) { set_error("score_completion: too large"); return VDET_ERR_UNSUPPORTED; }
    const int words_per_row = (int)((L + 31) / 32);
    const int spans_per_row = (words_per_row + 31) / 32;
    const int64_t n_words = (int64_t)words_per_row * n_rows;
    const int64_t n_spans = (int64_t)spans_per_row * n_rows;
    const size_t need = (size_t)n_words * sizeof(uint32_t);
    if (ws == nullptr || ws_bytes < need) {
        set_error("score_completion: workspace of %zu bytes needed", need);
        return VDET_ERR_WORKSPACE;
    }
    uint32_t* bitmap = (uint32_t*)ws;
    int64_t grid = (n_spans + CS_WARPS - 1) / CS_WARPS;
    const int64_t cap = (int64_t)sm_count_cached() * 8;                // 8 resident CTAs per SM, persistent
    if (grid > cap) grid = cap;
    completion_scan_kernel<T><<<(unsigned)grid, CS_WARPS * 32, 0, st>>>((const T*)scores, L, ld, lengths, words_per_row,
                                                                        spans_per_row, n_spans, (T)miss_thr, bitmap);
    VDET_LAUNCH_CHECK();
    completion_fill_kernel<T><<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>((T*)scores, L, ld, lengths, words_per_row,
                                                                                 n_words, bitmap, (const T*)bounds, status);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

}  // namespace vdet

using namespace vdet;

extern "C" size_t vdet_score_completion_workspace_bytes(int64_t n_rows, int64_t L, int dtype) {
    (void)dtype;
    return completion_bitmap_words(n_rows, L) * sizeof(uint32_t) + 256;      // the "missing" bitmap
}

extern "C" int vdet_score_completion(void* scores, int dtype, int64_t n_rows, int64_t L, int64_t ld,
                                     const int32_t* lengths, double miss_thr, uint32_t* status,
                                     void* ws, size_t ws_bytes, void* stream) {
    VDET_REQUIRE(n_rows >= 0 && L >= 0 && ld >= L, "score_completion: bad shape");
    VDET_REQUIRE(dtype == VDET_DTYPE_F32 || dtype == VDET_DTYPE_F64, "score_completion: bad dtype");
    VDET_REQUIRE(status != nullptr, "score_completion: null status");
    if (n_rows == 0 || L == 0) return VDET_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == VDET_DTYPE_F32 ? run_completion<float>(scores, n_rows, L, ld, lengths, miss_thr, nullptr, status, ws, ws_bytes, st)
                                   : run_completion<double>(scores, n_rows, L, ld, lengths, miss_thr, nullptr, status, ws, ws_bytes, st);
}

extern "C" int vdet_score_completion_bounded(void* scores, int dtype, int64_t n_rows, int64_t L, int64_t ld,
                                             const int32_t* lengths, double miss_thr, const void* bounds,
                                             uint32_t* status, void* ws, size_t ws_bytes, void* stream) {
    VDET_REQUIRE(n_rows >= 0 && L >= 0 && ld >= L, "score_completion: bad shape");
    VDET_REQUIRE(dtype == VDET_DTYPE_F32 || dtype == VDET_DTYPE_F64, "score_completion: bad dtype");
    VDET_REQUIRE(status != nullptr && bounds != nullptr, "score_completion_bounded: null status / bounds");
    if (n_rows == 0 || L == 0) return VDET_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == VDET_DTYPE_F32 ? run_completion<float>(scores, n_rows, L, ld, lengths, miss_thr, bounds, status, ws, ws_bytes, st)
                                   : run_completion<double>(scores, n_rows, L, ld, lengths, miss_thr, bounds, status, ws, ws_bytes, st);
}

extern "C" int vdet_temporal_maxpool(const void* scores, void* out, int dtype, int64_t n_rows, int64_t L,
                                     int64_t ld, const int32_t* lengths, int window, double pad, void* stream) {
    VDET_REQUIRE(n_rows >= 0 && L >= 0 && ld >= L, "temporal_maxpool: bad shape");
    VDET_REQUIRE(dtype == VDET_DTYPE_F32 || dtype == VDET_DTYPE_F64, "temporal_maxpool: bad dtype");
    VDET_REQUIRE(window >= 1 && (window % 2) == 1, "Window size must be odd!");     // tubelet_cls.py:389-390
    VDET_REQUIRE(scores != out, "temporal_maxpool: out must not alias the input");
    if (n_rows == 0 || L == 0) return VDET_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == VDET_DTYPE_F32 ? run_maxpool<float>(scores, out, n_rows, L, ld, lengths, window, pad, st)
                                   : run_maxpool<double>(scores, out, n_rows, L, ld, lengths, window, pad, st);
}

extern "C" int vdet_temporal_conv1d(const void* x, void* out, int dtype, int64_t n_rows, int64_t L,
                                    int64_t ld, const int32_t* lengths, const void* taps, int n_channels,
                                    int window, int pad_mode, void* stream) {
    VDET_REQUIRE(n_rows >= 0 && L >= 0 && ld >= L && n_channels >= 1, "temporal_conv1d: bad shape");
    VDET_REQUIRE(dtype == VDET_DTYPE_F32 || dtype == VDET_DTYPE_F64, "temporal_conv1d: bad dtype");
    VDET_REQUIRE(window >= 1 && (window % 2) == 1, "temporal_conv1d: window must be odd");
    VDET_REQUIRE(pad_mode == VDET_PAD_ZERO || pad_mode == VDET_PAD_EDGE, "temporal_conv1d: bad pad mode");
    VDET_REQUIRE(x != out, "temporal_conv1d: out must not alias the input");
    if (n_rows == 0 || L == 0) return VDET_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == VDET_DTYPE_F32
               ? run_conv<float>(x, out, n_rows, L, ld, lengths, taps, n_channels, window, pad_mode, st)
               : run_conv<double>(x, out, n_rows, L, ld, lengths, taps, n_channels, window, pad_mode, st);
}
