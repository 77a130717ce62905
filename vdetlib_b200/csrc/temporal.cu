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
// Streaming formulation (one read + one write of the row, any row length): rows are cut into
// 256-element tiles, one warp each.
//   pass 1 (completion_bounds_kernel): for tiles whose first / last element is missing, the nearest
//          valid element to the LEFT / RIGHT of the tile (index + value) -- a ballot search that
//          normally ends in its first 32-element probe -- goes to a small side buffer.  Reading
//          neighbours in a separate pass keeps the in-place update of pass 2 race-free.
//   pass 2 (completion_fill_kernel): 16-byte loads, warp max-scan of "last valid index" and min-scan
//          of "next valid index" (shuffles only, no block barrier), closed form per missing
//          element, 16-byte stores by the lanes that changed something; tiles without a missing
//          element return after the load.
constexpr int CP_ITEMS = 8;                    // elements per lane
constexpr int CP_TILE = 32 * CP_ITEMS;         // one warp = one 256-element tile
constexpr int CP_WARPS = 8;                    // tiles per CTA (independent of each other)

template <typename T>
struct TileBounds { int32_t left_idx; int32_t right_idx; T left_val; T right_val; };

// Pass 1: only tiles whose first / last element is missing need a neighbour outside the tile.
// One THREAD per tile (two loads); the rare tiles that do need a neighbour walk outwards serially.
template <typename T>
__global__ void __launch_bounds__(256) completion_bounds_kernel(const T* __restrict__ scores, int64_t L, int64_t ld,
                                                                const int32_t* __restrict__ lengths,
                                                                int tiles_per_row, int64_t n_tiles, T miss_thr,
                                                                TileBounds<T>* __restrict__ bounds) {
    const int64_t tile_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tile_id >= n_tiles) return;
    const int64_t row = tile_id / tiles_per_row;
    const int tile = (int)(tile_id - row * tiles_per_row);
    const int len = (int)(lengths ? (int64_t)lengths[row] : L);
    const int t0 = tile * CP_TILE;
    if (t0 >= len) return;
    const int t1 = (t0 + CP_TILE < len) ? t0 + CP_TILE : len;
    const T* g = scores + row * ld;
    const bool need_left = g[t0] <= miss_thr, need_right = g[t1 - 1] <= miss_thr;
    if (!need_left && !need_right) return;
    TileBounds<T> b;
    b.left_idx = -1; b.right_idx = len; b.left_val = (T)0; b.right_val = (T)0;
    if (need_left) {
        for (int k = t0 - 1; k >= 0; --k) {                     // nearest valid element left of the tile
            const T v = g[k];
            if (!(v <= miss_thr)) { b.left_idx = k; b.left_val = v; break; }
        }
    }
    if (need_right) {
        for (int k = t1; k < len; ++k) {                        // nearest valid element right of it
            const T v = g[k];
            if (!(v <= miss_thr)) { b.right_idx = k; b.right_val = v; break; }
        }
    }
    bounds[tile_id] = b;
}

// Pass 2: persistent warps, one 256-element tile at a time, no block barrier.  Lanes load 8
// consecutive elements (16-byte loads) and publish them plus an 8-bit validity mask in the warp's
// shared scratch.  The tile's MISSING elements are then dealt out evenly over the lanes (ballots per
// position + __fns pick the m-th one): missing elements are ~5 % of a row, so one pass of the closed
// form serves the whole tile -- walking each lane's own 8 positions instead would make every warp
// execute the divergent body 8 times (first version: 489 instructions per tile).  Last / next valid
// index come from the 256-bit validity set, values from the scratch or the tile's boundary record.
template <typename T>
__global__ void __launch_bounds__(CP_WARPS * 32) completion_fill_kernel(T* __restrict__ scores, int64_t L, int64_t ld,
                                                                        const int32_t* __restrict__ lengths,
                                                                        int tiles_per_row, int64_t n_tiles, T miss_thr,
                                                                        int vec_ok,
                                                                        const TileBounds<T>* __restrict__ bounds,
                                                                        uint32_t* status) {
    __shared__ __align__(16) T s_val_all[CP_WARPS][CP_TILE];
    __shared__ __align__(16) uint32_t s_ok_all[CP_WARPS][8];      // 256 validity bits per tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* s_val = s_val_all[warp];
    uint32_t* s_ok = s_ok_all[warp];
    constexpr int V = 16 / sizeof(T);
    for (int64_t tile_id = (int64_t)blockIdx.x * CP_WARPS + warp; tile_id < n_tiles;
         tile_id += (int64_t)gridDim.x * CP_WARPS) {
        const int64_t row = tile_id / tiles_per_row;
        const int tile = (int)(tile_id - row * tiles_per_row);
        const int len = (int)(lengths ? (int64_t)lengths[row] : L);
        const int t0 = tile * CP_TILE;
        if (t0 >= len) continue;
        T* g = scores + row * ld;
        const int e0 = t0 + lane * CP_ITEMS;                   // this lane's 8 consecutive elements
        __align__(16) T v[CP_ITEMS];
        if (vec_ok && e0 + CP_ITEMS <= len) {
#pragma unroll
            for (int q = 0; q < CP_ITEMS; q += V)
                *reinterpret_cast<typename Vec16<T>::type*>(v + q) =
                    *reinterpret_cast<const typename Vec16<T>::type*>(g + e0 + q);
        } else {
#pragma unroll
            for (int q = 0; q < CP_ITEMS; ++q) v[q] = (e0 + q < len) ? g[e0 + q] : (T)0;
        }
        unsigned okm = 0, missm = 0;                           // bit q: element q valid / missing
#pragma unroll
        for (int q = 0; q < CP_ITEMS; ++q) {
            const bool in = e0 + q < len;
            const bool miss = v[q] <= miss_thr;
            if (in && !miss) okm |= 1u << q;
            if (in && miss) missm |= 1u << q;
        }
        if (!__any_sync(FULL, missm != 0)) continue;           // nothing to fill in this tile
        __syncwarp();                                          // previous tile's readers are done
#pragma unroll
        for (int q = 0; q < CP_ITEMS; q += V)
            *reinterpret_cast<typename Vec16<T>::type*>(s_val + lane * CP_ITEMS + q) =
                *reinterpret_cast<const typename Vec16<T>::type*>(v + q);
        reinterpret_cast<uint8_t*>(s_ok)[lane] = (uint8_t)okm;
        // where are the missing elements?  cq[q] = lanes whose element q is missing
        unsigned cq[CP_ITEMS];
        int total = 0;
#pragma unroll
        for (int q = 0; q < CP_ITEMS; ++q) {
            cq[q] = __ballot_sync(FULL, (missm >> q) & 1u);
            total += __popc(cq[q]);
        }
        __syncwarp();
        const uint4 w_lo = *reinterpret_cast<const uint4*>(s_ok), w_hi = *reinterpret_cast<const uint4*>(s_ok + 4);
        const uint32_t okw[8] = {w_lo.x, w_lo.y, w_lo.z, w_lo.w, w_hi.x, w_hi.y, w_hi.z, w_hi.w};
        for (int m = lane; m < total; m += 32) {
            // the m-th missing element (position-major order): position q, then the lane owning it
            int q = 0, rest = m;
#pragma unroll
            for (int qq = 0; qq < CP_ITEMS - 1; ++qq) {
                const int c = __popc(cq[qq]);
                if (q == qq && rest >= c) { rest -= c; q = qq + 1; }
            }
            unsigned sel = cq[0];
#pragma unroll
            for (int qq = 1; qq < CP_ITEMS; ++qq) sel = (q == qq) ? cq[qq] : sel;
            const int src = __fns(sel, 0, rest + 1);           // lane that owns it
            const int e = src * CP_ITEMS + q;                  // element index inside the tile
            const int k = t0 + e;
            // last valid element before e / first valid element after e, inside the tile
            int lv = -1, nv = 0x7fffffff;
#pragma unroll
            for (int w = 7; w >= 0; --w) {
                unsigned mword = okw[w];
                if (w == (e >> 5)) mword &= (1u << (e & 31)) - 1u;
                if (w <= (e >> 5) && lv < 0 && mword) lv = w * 32 + 31 - __clz(mword);
            }
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                unsigned mword = okw[w];
                if (w == (e >> 5)) mword &= ~((2u << (e & 31)) - 1u);
                if (w >= (e >> 5) && nv == 0x7fffffff && mword) nv = w * 32 + __ffs(mword) - 1;
            }
            int i, j;             // run = [i, j) in row coordinates
            T lft, rgt;
            TileBounds<T> b;
            b.left_idx = -1; b.right_idx = len; b.left_val = (T)0; b.right_val = (T)0;
            if (lv < 0 || nv == 0x7fffffff) b = bounds[tile_id];   // a run that touches the tile edge
            if (lv >= 0) { i = t0 + lv + 1; lft = s_val[lv]; }
            else { i = b.left_idx + 1; lft = b.left_val; }
            if (nv != 0x7fffffff) { j = t0 + nv; rgt = s_val[nv]; }
            else { j = b.right_idx; rgt = b.right_val; }
            T r;
            if (i == 0) {
                if (j >= len) { if (k == 0) atomicOr(status, VDET_STATUS_ALL_MISSING); continue; }
                r = rgt;
            } else if (j >= len) {
                r = lft;
            } else {
                r = t_add(lft, t_div(t_mul(t_sub(rgt, lft), (T)(k - i + 1)), (T)(j - i + 1)));
            }
            g[k] = r;
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

template <typename T>
static int run_completion(void* scores, int64_t n_rows, int64_t L, int64_t ld, const int32_t* lengths,
                          double miss_thr, uint32_t* status, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int64_t tiles = (L + CP_TILE - 1) / CP_TILE;
    const int64_t n_tiles = tiles * n_rows;
    if (n_tiles > 0x7fffffff || L >= 0x7fffffff) { set_error("score_completion: too large"); return VDET_ERR_UNSUPPORTED; }
    const size_t need = (size_t)n_tiles * sizeof(TileBounds<T>);
    if (ws == nullptr || ws_bytes < need) {
        set_error("score_completion: workspace of %zu bytes needed", need);
        return VDET_ERR_WORKSPACE;
    }
    TileBounds<T>* bounds = (TileBounds<T>*)ws;
    completion_bounds_kernel<T><<<(unsigned)((n_tiles + 255) / 256), 256, 0, st>>>((const T*)scores, L, ld, lengths,
                                                                              (int)tiles, n_tiles, (T)miss_thr, bounds);
    VDET_LAUNCH_CHECK();
    const int vec_ok = (((uintptr_t)scores & 15) == 0 && (ld % (16 / sizeof(T))) == 0) ? 1 : 0;
    int64_t fill_grid = (n_tiles + CP_WARPS - 1) / CP_WARPS;
    const int64_t fill_cap = (int64_t)sm_count_cached() * 8;          // 8 resident CTAs per SM
    if (fill_grid > fill_cap) fill_grid = fill_cap;
    completion_fill_kernel<T><<<(unsigned)fill_grid, CP_WARPS * 32, 0, st>>>(
        (T*)scores, L, ld, lengths, (int)tiles, n_tiles, (T)miss_thr, vec_ok, bounds, status);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

}  // namespace vdet

using namespace vdet;

extern "C" size_t vdet_score_completion_workspace_bytes(int64_t n_rows, int64_t L, int dtype) {
    const int64_t tiles = (L + CP_TILE - 1) / CP_TILE;
    const size_t rec = dtype == VDET_DTYPE_F64 ? sizeof(TileBounds<double>) : sizeof(TileBounds<float>);
    return (size_t)(tiles * (n_rows > 0 ? n_rows : 0)) * rec + 256;
}

extern "C" int vdet_score_completion(void* scores, int dtype, int64_t n_rows, int64_t L, int64_t ld,
                                     const int32_t* lengths, double miss_thr, uint32_t* status,
                                     void* ws, size_t ws_bytes, void* stream) {
    VDET_REQUIRE(n_rows >= 0 && L >= 0 && ld >= L, "score_completion: bad shape");
    VDET_REQUIRE(dtype == VDET_DTYPE_F32 || dtype == VDET_DTYPE_F64, "score_completion: bad dtype");
    VDET_REQUIRE(status != nullptr, "score_completion: null status");
    if (n_rows == 0 || L == 0) return VDET_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == VDET_DTYPE_F32 ? run_completion<float>(scores, n_rows, L, ld, lengths, miss_thr, status, ws, ws_bytes, st)
                                   : run_completion<double>(scores, n_rows, L, ld, lengths, miss_thr, status, ws, ws_bytes, st);
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
