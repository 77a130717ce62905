// api.cu -- error reporting and device queries of libvdet_b200.so.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include <emmintrin.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <cstdlib>
#include <thread>
#include <vector>
#include <unistd.h>

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

bool cpu_has_avx512f();                                                          // host_copy.cpp
void copy_stream_range_512(unsigned char* d, const unsigned char* s, size_t bytes);

// VDET_HOST_COPY=sse|avx512 forces a flavour (measurement hook); default: AVX-512 full-line stores when the CPU has them
static int host_copy_flavour() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VDET_HOST_COPY");
        if (e && !strcmp(e, "sse")) v = 0;
        else if (e && !strcmp(e, "avx512")) v = cpu_has_avx512f() ? 1 : 0;
        else v = cpu_has_avx512f() ? 1 : 0;
    }
    return v;
}

// Non-temporal copy of one byte range (16-byte streaming stores, plain copies for the unaligned ends),
// fenced before returning so that the stores are globally visible when the caller hands the buffer to a DMA.
void copy_stream_range(unsigned char* d, const unsigned char* s, size_t bytes) {
    if (host_copy_flavour() == 1) { copy_stream_range_512(d, s, bytes); return; }
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
}

// ---- persistent staging threads -----------------------------------------------------------------
// The end-to-end path stages a 41 MB shard per step (~1 ms): starting and joining 7 threads per call costs a
// visible fraction of that, so the copy ranges go to a small pool of persistent workers.  A worker that runs
// out of work keeps polling for ~200 us (the next step's copy usually arrives within that) before it sleeps on
// a condition variable; the calling thread takes ranges too.  After a fork() the child starts a fresh pool.
struct CopyJob { unsigned char* d; const unsigned char* s; size_t n; };

namespace {
struct CopyPool {
    // One batch of ranges per call.  `state` = (generation << 32) | next range index: a worker claims a range with
    // a compare-and-swap on the whole word, so a claim can only succeed while ITS generation is the active one --
    // and then the batch fields it read (published before the state word) are that generation's.
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::thread> workers;
    std::atomic<const CopyJob*> jobs{nullptr};
    std::atomic<uint32_t> n_jobs{0};
    std::atomic<uint32_t> done{0};
    std::atomic<uint64_t> state{0};
    std::atomic<bool> stop{false};
    pid_t owner = 0;

    bool claim_and_copy() {                            // false: nothing left in the active batch
        for (;;) {
            const uint64_t s = state.load(std::memory_order_acquire);
            const uint32_t idx = (uint32_t)s;
            const CopyJob* js = jobs.load(std::memory_order_acquire);
            const uint32_t n = n_jobs.load(std::memory_order_acquire);
            if (idx >= n) return false;
            uint64_t expect = s;
            if (!state.compare_exchange_weak(expect, s + 1, std::memory_order_acq_rel)) continue;
            const CopyJob j = js[idx];
            copy_stream_range(j.d, j.s, j.n);
            done.fetch_add(1, std::memory_order_acq_rel);
        }
    }
    void worker() {
        uint64_t seen_gen = state.load(std::memory_order_acquire) >> 32;
        for (;;) {
            bool got = false;
            const auto t0 = std::chrono::steady_clock::now();
            while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200)) {     // poll, then sleep
                if (stop.load(std::memory_order_relaxed)) return;
                if ((state.load(std::memory_order_acquire) >> 32) != seen_gen) { got = true; break; }
                _mm_pause();
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop.load() || (state.load(std::memory_order_acquire) >> 32) != seen_gen; });
                if (stop.load()) return;
            }
            seen_gen = state.load(std::memory_order_acquire) >> 32;
            while (claim_and_copy()) {}
        }
    }
    void ensure(int n_workers) {
        if (owner != getpid()) {                       // first use (a forked child gets a fresh pool object, see below)
            workers.clear();
            owner = getpid();
        }
        while ((int)workers.size() < n_workers) workers.emplace_back([this] { worker(); });
    }
    void run(const std::vector<CopyJob>& js) {
        ensure((int)js.size() - 1);
        const uint64_t gen = (state.load(std::memory_order_acquire) >> 32) + 1;
        {
            std::lock_guard<std::mutex> lk(mu);        // the state word changes under the lock sleepers wait with
            done.store(0, std::memory_order_relaxed);
            jobs.store(js.data(), std::memory_order_release);
            n_jobs.store((uint32_t)js.size(), std::memory_order_release);
            state.store(gen << 32, std::memory_order_release);
        }
        cv.notify_all();
        while (claim_and_copy()) {}
        while (done.load(std::memory_order_acquire) < js.size()) _mm_pause();
    }
};
std::mutex g_pool_call_mu;                              // one staging copy at a time per process
CopyPool* g_pool = nullptr;
}  // namespace

void copy_pool_run(const std::vector<CopyJob>& jobs) {
    std::lock_guard<std::mutex> lk(g_pool_call_mu);
    if (g_pool == nullptr || g_pool->owner != getpid()) g_pool = new CopyPool();   // (a forked child leaks the parent's object)
    g_pool->run(jobs);
}

}  // namespace vdet

extern "C" {

int vdet_abi_version(void) { return VDET_ABI_VERSION; }

const char* vdet_last_error(void) { return vdet::g_err; }

int vdet_host_copy_stream(void* dst, const void* src, size_t bytes) {
    return vdet_host_copy_stream_mt(dst, src, bytes, 1);
}

int vdet_host_copy_stream_mt(void* dst, const void* src, size_t bytes, int n_threads) {
    if (bytes == 0) return VDET_OK;
    if (dst == nullptr || src == nullptr) {
        vdet::set_error("host_copy_stream: null pointer");
        return VDET_ERR_INVALID;
    }
    if (n_threads <= 0) {                                   // auto: one core moves ~10 GB/s, the DRAM bus several times that
        static const int env_threads = [] {                 // VDET_STAGE_THREADS: the auto value, fixed by the operator
            const char* e = std::getenv("VDET_STAGE_THREADS");
            return e ? std::atoi(e) : 0;
        }();
        const unsigned hw = std::thread::hardware_concurrency() / 2;     // physical cores, roughly
        n_threads = env_threads > 0 ? env_threads : (int)(hw == 0 ? 1 : (hw > 8 ? 8 : hw));
        if (bytes < (size_t)(4u << 20)) n_threads = 1;      // not worth waking the pool
    }
    if (n_threads > 64) n_threads = 64;
    unsigned char* d = static_cast<unsigned char*>(dst);
    const unsigned char* s = static_cast<const unsigned char*>(src);
    if (n_threads == 1 || bytes < (size_t)n_threads * 4096) {
        vdet::copy_stream_range(d, s, bytes);
        return VDET_OK;
    }
    // ranges start on 64-byte lines of the destination (a line is never shared by two threads)
    const size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;
    const size_t lines = (bytes - head) / 64;
    const size_t per = (lines + n_threads - 1) / n_threads;
    std::vector<vdet::CopyJob> jobs;
    jobs.reserve(n_threads);
    const size_t first_end = (per < lines) ? head + per * 64 : bytes;
    jobs.push_back({d, s, first_end});                                      // head + first range
    for (int t = 1; t < n_threads; ++t) {
        const size_t l0 = per * (size_t)t, l1 = (l0 + per < lines) ? l0 + per : lines;
        if (l0 >= l1) break;
        const size_t b0 = head + l0 * 64;
        const size_t b1 = (l1 == lines) ? bytes : head + l1 * 64;          // the last range takes the tail
        jobs.push_back({d + b0, s + b0, b1 - b0});
    }
    vdet::copy_pool_run(jobs);
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
