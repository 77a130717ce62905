// api.cu -- error reporting and device queries of libvdet_b200.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

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
