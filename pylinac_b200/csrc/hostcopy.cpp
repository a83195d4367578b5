// Host-side staging copy of pageable source frames into the page-locked ring (pf.cu CopyPool).
// The destination is written once and next read by the DMA engine, never by this CPU: non-temporal stores skip the
// read-for-ownership of every destination line (3 -> 2 bytes of memory traffic per byte copied) and keep the ring out of the caches.
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace epid {

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void nt_copy_avx2(char* d, const char* s, size_t n) {
    size_t head = (64 - ((uintptr_t)d & 63)) & 63;
    if (head > n) head = n;
    memcpy(d, s, head);
    d += head; s += head; n -= head;
    const size_t blocks = n / 128;
    for (size_t i = 0; i < blocks; i++) {
        const char* sp = s + i * 128;
        char* dp = d + i * 128;
        _mm_prefetch(sp + 1024, _MM_HINT_NTA);
        _mm_prefetch(sp + 1088, _MM_HINT_NTA);
        const __m256i a = _mm256_loadu_si256((const __m256i*)(sp));
        const __m256i b = _mm256_loadu_si256((const __m256i*)(sp + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(sp + 64));
        const __m256i e = _mm256_loadu_si256((const __m256i*)(sp + 96));
        _mm256_stream_si256((__m256i*)(dp), a);
        _mm256_stream_si256((__m256i*)(dp + 32), b);
        _mm256_stream_si256((__m256i*)(dp + 64), c);
        _mm256_stream_si256((__m256i*)(dp + 96), e);
    }
    _mm_sfence();
    memcpy(d + blocks * 128, s + blocks * 128, n - blocks * 128);
}
#endif

void staging_copy(void* dst, const void* src, size_t bytes) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !(getenv("EPID_COPY_NT") && atoi(getenv("EPID_COPY_NT")) == 0);
    if (avx2 && bytes >= (1u << 16)) { nt_copy_avx2((char*)dst, (const char*)src, bytes); return; }
#endif
    memcpy(dst, src, bytes);
}

}  // namespace epid
