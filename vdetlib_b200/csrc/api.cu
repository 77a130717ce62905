// api.cu -- error reporting and device queries of libvdet_b200.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include <emmintrin.h>
#include <cstring>

namespace vdet {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return VDET_ERR_CUDA;
}

static int attr_cached(cudaDeviceAttr attr, int* cache /* [64] */) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (cache[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, attr, dev) != cudaSuccess) return -1;
        cache[dev] = v;
    }
    return cache[dev];
}

int sm_count_cached() {
    static int cache[64];
    int v = attr_cached(cudaDevAttrMultiProcessorCount, cache);
    return v > 0 ? v : 148;
}

static int g_reserved_sms = 0;

int usable_sm_count() {
    const int n = sm_count_cached() - g_reserved_sms;
    return n > 0 ? n : 1;
}

void set_reserved_sms(int n) { g_reserved_sms = n < 0 ? 0 : n; }

int max_optin_smem_cached() {
    static int cache[64];
    int v = attr_cached(cudaDevAttrMaxSharedMemoryPerBlockOptin, cache);
    return v > 0 ? v : 227 * 1024;
}

}  // namespace vdet

extern "C" {

int vdet_abi_version(void) { return VDET_ABI_VERSION; }

const char* vdet_last_error(void) { return vdet::g_err; }

int vdet_host_copy_stream(void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return VDET_OK;
    if (dst == nullptr || src == nullptr) {
        vdet::set_error("host_copy_stream: null pointer");
        return VDET_ERR_INVALID;
    }
    unsigned char* d = static_cast<unsigned char*>(dst);
    const unsigned char* s = static_cast<const unsigned char*>(src);
    size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;     // stores must be 16-byte aligned
    if (head > bytes) head = bytes;
    memcpy(d, s, head);
    d += head; s += head; bytes -= head;
    const size_t nvec = bytes / 16;
    __m128i* dv = reinterpret_cast<__m128i*>(d);
    const __m128i* sv = reinterpret_cast<const __m128i*>(s);
    size_t i = 0;
    for (; i + 4 <= nvec; i += 4) {                                      // one 64-byte line per round
        const __m128i a = _mm_loadu_si128(sv + i), b = _mm_loadu_si128(sv + i + 1);
        const __m128i c = _mm_loadu_si128(sv + i + 2), e = _mm_loadu_si128(sv + i + 3);
        _mm_stream_si128(dv + i, a); _mm_stream_si128(dv + i + 1, b);
        _mm_stream_si128(dv + i + 2, c); _mm_stream_si128(dv + i + 3, e);
    }
    for (; i < nvec; ++i) _mm_stream_si128(dv + i, _mm_loadu_si128(sv + i));
    memcpy(d + nvec * 16, s + nvec * 16, bytes - nvec * 16);
    _mm_sfence();
    return VDET_OK;
}

int vdet_set_reserved_sms(int n) {
    vdet::set_reserved_sms(n);
    return VDET_OK;
}

int vdet_sm_count(int device) {
    int v = 0;
    cudaError_t e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return vdet::cuda_fail(e, "cudaDeviceGetAttribute", __FILE__, __LINE__);
    return v;
}

}  // extern "C"
