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

// ---- score completion --------------------------------------------------------------------
// One CTA per row, the whole row resident in shared memory.  A forward max-scan gives, for
// every k, the last valid index <= k; a backward min-scan the next valid index >= k; every
// missing element then evaluates the reference's closed form for its own run [i, j):
//   leading run  -> s[j]            (tubelet_cls.py:293-295)
//   trailing run -> s[i-1]          (:296-298)
//   interior     -> l + (r - l) * (k - i + 1) / (j - i + 1)     (:299-303)
constexpr int CP_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(CP_THREADS) score_completion_kernel(T* __restrict__ scores, int64_t L, int64_t ld,
                                                                      const int32_t* __restrict__ lengths,
                                                                      T miss_thr, int cap, uint32_t* status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s = reinterpret_cast<T*>(smem_raw);
    int32_t* lastv = reinterpret_cast<int32_t*>(s + cap);
    int32_t* nextv = lastv + cap;
    __shared__ int32_t s_lv[CP_THREADS], s_fv[CP_THREADS];
    const int64_t row = blockIdx.x;
    const int len = (int)(lengths ? (int64_t)lengths[row] : L);
    if (len <= 0) return;
    T* g = scores + row * ld;
    const int tid = threadIdx.x;
    for (int k = tid; k < len; k += CP_THREADS) s[k] = g[k];
    __syncthreads();
    // contiguous chunk per thread; odd chunk length keeps the strided passes conflict-free
    const int ch = ((len + CP_THREADS - 1) / CP_THREADS) | 1;
    const int b = tid * ch < len ? tid * ch : len;
    const int e = b + ch < len ? b + ch : len;
    int lv = -1, fv = 0x7fffffff;
    for (int k = b; k < e; ++k) {
        const bool valid = !(s[k] <= miss_thr);
        if (valid) { lv = k; if (fv == 0x7fffffff) fv = k; }
    }
    s_lv[tid] = lv;
    s_fv[tid] = fv;
    __syncthreads();
    // exclusive max-scan of lv to the left, exclusive min-scan of fv to the right (256 entries)
    int carry_l = -1, carry_r = 0x7fffffff;
    for (int t = 0; t < tid; ++t) carry_l = max(carry_l, s_lv[t]);
    for (int t = tid + 1; t < CP_THREADS; ++t) carry_r = min(carry_r, s_fv[t]);
    for (int k = b; k < e; ++k) {
        if (!(s[k] <= miss_thr)) carry_l = k;
        lastv[k] = carry_l;
    }
    for (int k = e - 1; k >= b; --k) {
        if (!(s[k] <= miss_thr)) carry_r = k;
        nextv[k] = carry_r;
    }
    __syncthreads();
    for (int k = tid; k < len; k += CP_THREADS) {
        const T v = s[k];
        if (!(v <= miss_thr)) continue;        // valid (or NaN): untouched, as the reference
        const int i = lastv[k] + 1;            // run start
        const int j = nextv[k];                // first valid index after the run (or "none")
        T r;
        if (i == 0) {
            if (j >= len) { if (k == 0) atomicOr(status, VDET_STATUS_ALL_MISSING); continue; }
            r = s[j];
        } else if (j >= len) {
            r = s[i - 1];
        } else {
            const T lft = s[i - 1], rgt = s[j];
            r = t_add(lft, t_div(t_mul(t_sub(rgt, lft), (T)(k - i + 1)), (T)(j - i + 1)));
        }
        g[k] = r;
    }
}

template <typename T>
static int run_maxpool(const void* in, void* out, int64_t n_rows, int64_t L, int64_t ld, const int32_t* lengths,
                       int window, double pad, cudaStream_t st) {
    const int h = window / 2;
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
                          double miss_thr, uint32_t* status, cudaStream_t st) {
    const int cap = (int)((L + 3) / 4 * 4);
    const size_t smem = (size_t)cap * (sizeof(T) + 2 * sizeof(int32_t));
    if (smem > max_dynamic_smem(score_completion_kernel<T>)) {
        set_error("score_completion: rows of %lld elements exceed the shared-memory row limit of this build",
                  (long long)L);
        return VDET_ERR_UNSUPPORTED;
    }
    VDET_CUDA(allow_dynamic_smem(score_completion_kernel<T>, smem));
    if (n_rows > 0x7fffffff) { set_error("score_completion: too many rows"); return VDET_ERR_UNSUPPORTED; }
    score_completion_kernel<T><<<(unsigned)n_rows, CP_THREADS, smem, st>>>((T*)scores, L, ld, lengths, (T)miss_thr,
                                                                           cap, status);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}

}  // namespace vdet

using namespace vdet;

extern "C" int vdet_score_completion(void* scores, int dtype, int64_t n_rows, int64_t L, int64_t ld,
                                     const int32_t* lengths, double miss_thr, uint32_t* status, void* stream) {
    VDET_REQUIRE(n_rows >= 0 && L >= 0 && ld >= L, "score_completion: bad shape");
    VDET_REQUIRE(dtype == VDET_DTYPE_F32 || dtype == VDET_DTYPE_F64, "score_completion: bad dtype");
    VDET_REQUIRE(status != nullptr, "score_completion: null status");
    if (n_rows == 0 || L == 0) return VDET_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == VDET_DTYPE_F32 ? run_completion<float>(scores, n_rows, L, ld, lengths, miss_thr, status, st)
                                   : run_completion<double>(scores, n_rows, L, ld, lengths, miss_thr, status, st);
}

extern "C" int vdet_temporal_maxpool(const void* scores, void* out, int dtype, int64_t n_rows, int64_t L,
                                     int64_t ld, const int32_t* lengths, int window, double pad, void* stream) {
    VDET_REQUIRE(n_rows >= 0 && L >= 0 && ld >= L, "temporal_maxpool: bad shape");
    VDET_REQUIRE(dtype == VDET_DTYPE_F32 || dtype == VDET_DTYPE_F64, "temporal_maxpool: bad dtype");
    VDET_REQUIRE(window >= 1 && (window % 2) == 1, "Window size must be odd!");     // tubelet_cls.py:389-390
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
    if (n_rows == 0 || L == 0) return VDET_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == VDET_DTYPE_F32
               ? run_conv<float>(x, out, n_rows, L, ld, lengths, taps, n_channels, window, pad_mode, st)
               : run_conv<double>(x, out, n_rows, L, ld, lengths, taps, n_channels, window, pad_mode, st);
}
