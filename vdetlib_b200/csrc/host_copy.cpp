// host_copy.cpp -- the 64-byte (AVX-512) flavour of the streaming staging copy (plain host code; see api.cu for
// the 16-byte SSE2 one and for what the staging copy is for).  One full-line non-temporal store per cache line:
// the write-combining buffer is filled by a single instruction instead of four.
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace vdet {

bool cpu_has_avx512f() {
    static const int v = __builtin_cpu_supports("avx512f");
    return v != 0;
}

__attribute__((target("avx512f"))) void copy_stream_range_512(unsigned char* d, const unsigned char* s, size_t bytes) {
    size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;         // stores must be 64-byte aligned
    if (head > bytes) head = bytes;
    memcpy(d, s, head);
    d += head; s += head; bytes -= head;
    const size_t nvec = bytes / 64;
    size_t i = 0;
    for (; i + 4 <= nvec; i += 4) {
        const __m512i a = _mm512_loadu_si512(reinterpret_cast<const void*>(s + 64 * i));
        const __m512i b = _mm512_loadu_si512(reinterpret_cast<const void*>(s + 64 * (i + 1)));
        const __m512i c = _mm512_loadu_si512(reinterpret_cast<const void*>(s + 64 * (i + 2)));
        const __m512i e = _mm512_loadu_si512(reinterpret_cast<const void*>(s + 64 * (i + 3)));
        _mm512_stream_si512(reinterpret_cast<__m512i*>(d + 64 * i), a);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(d + 64 * (i + 1)), b);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(d + 64 * (i + 2)), c);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(d + 64 * (i + 3)), e);
    }
    for (; i < nvec; ++i)
        _mm512_stream_si512(reinterpret_cast<__m512i*>(d + 64 * i), _mm512_loadu_si512(reinterpret_cast<const void*>(s + 64 * i)));
    memcpy(d + nvec * 64, s + nvec * 64, bytes - nvec * 64);
    _mm_sfence();
}

}  // namespace vdet
