// Host-side streaming kernels of the reference-shaped entry points (deodr_b200_render_host / render_b_host): the
// fp64 <-> fp32 conversions, copies, zero fills and comparisons that move the caller's numpy arrays to and from the
// pinned staging buffers.  They are memory-bound; the AVX-512 versions use non-temporal stores so that a converted
// buffer costs one read + one write of DRAM traffic (no read-for-ownership of the destination) and does not evict the
// caller's working set.  Plain C++ (compiled by the host compiler, no CUDA): picked at run time with
// __builtin_cpu_supports, scalar code otherwise.  Results are bit-identical between the two (IEEE conversions).
#include "host_simd.h"

#include <cstdint>
#include <cstring>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define DEODR_HAVE_X86 1
#else
#define DEODR_HAVE_X86 0
#endif

namespace {

#if DEODR_HAVE_X86
bool has_avx512() {
    static const bool v = __builtin_cpu_supports("avx512f");
    return v;
}

__attribute__((target("avx512f"))) void f64_to_f32_avx512(float *dst, const double *src, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 63u)) { dst[i] = (float)src[i]; i++; }
    for (; i + 16 <= n; i += 16) {
        const __m256 lo = _mm512_cvtpd_ps(_mm512_loadu_pd(src + i));
        const __m256 hi = _mm512_cvtpd_ps(_mm512_loadu_pd(src + i + 8));
        const __m512 both = _mm512_castpd_ps(
            _mm512_insertf64x4(_mm512_castpd256_pd512(_mm256_castps_pd(lo)), _mm256_castps_pd(hi), 1));
        _mm512_stream_ps(dst + i, both);
    }
    for (; i < n; i++) dst[i] = (float)src[i];
    _mm_sfence();
}

__attribute__((target("avx512f"))) void f32_to_f64_avx512(double *dst, const float *src, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 63u)) { dst[i] = (double)src[i]; i++; }
    for (; i + 8 <= n; i += 8) _mm512_stream_pd(dst + i, _mm512_cvtps_pd(_mm256_loadu_ps(src + i)));
    for (; i < n; i++) dst[i] = (double)src[i];
    _mm_sfence();
}

__attribute__((target("avx512f"))) void f32_add_f64_avx512(double *dst, const float *src, size_t n) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8)
        _mm512_storeu_pd(dst + i, _mm512_add_pd(_mm512_loadu_pd(dst + i), _mm512_cvtps_pd(_mm256_loadu_ps(src + i))));
    for (; i < n; i++) dst[i] += (double)src[i];
}

__attribute__((target("avx512f"))) void copy_avx512(char *dst, const char *src, size_t bytes) {
    size_t i = 0;
    while (i < bytes && ((uintptr_t)(dst + i) & 63u)) { dst[i] = src[i]; i++; }
    for (; i + 64 <= bytes; i += 64) _mm512_stream_si512((__m512i *)(dst + i), _mm512_loadu_si512((const void *)(src + i)));
    for (; i < bytes; i++) dst[i] = src[i];
    _mm_sfence();
}

__attribute__((target("avx512f"))) void zero_avx512(char *dst, size_t bytes) {
    size_t i = 0;
    while (i < bytes && ((uintptr_t)(dst + i) & 63u)) dst[i++] = 0;
    const __m512i z = _mm512_setzero_si512();
    for (; i + 64 <= bytes; i += 64) _mm512_stream_si512((__m512i *)(dst + i), z);
    for (; i < bytes; i++) dst[i] = 0;
    _mm_sfence();
}

// 1 iff (float)user[i] has the bit pattern of mirror[i] for every i
__attribute__((target("avx512f"))) int equal_f32_avx512(const double *user, const float *mirror, size_t n) {
    size_t i = 0;
    __mmask8 diff = 0;
    for (; i + 8 <= n; i += 8) {
        const __m256i a = _mm256_castps_si256(_mm512_cvtpd_ps(_mm512_loadu_pd(user + i)));
        const __m256i b = _mm256_loadu_si256((const __m256i *)(mirror + i));
        const __m256i x = _mm256_xor_si256(a, b);
        diff |= (__mmask8)!_mm256_testz_si256(x, x);
    }
    int ok = diff == 0;
    for (; i < n; i++) {
        const float f = (float)user[i];
        ok &= (memcmp(&f, &mirror[i], sizeof(float)) == 0);
    }
    return ok;
}
#endif

}  // namespace

void deodr_simd_f64_to_f32(float *dst, const double *src, size_t n) {
#if DEODR_HAVE_X86
    if (has_avx512()) return f64_to_f32_avx512(dst, src, n);
#endif
    for (size_t i = 0; i < n; i++) dst[i] = (float)src[i];
}

void deodr_simd_f32_to_f64(double *dst, const float *src, size_t n) {
#if DEODR_HAVE_X86
    if (has_avx512()) return f32_to_f64_avx512(dst, src, n);
#endif
    for (size_t i = 0; i < n; i++) dst[i] = (double)src[i];
}

void deodr_simd_f32_add_f64(double *dst, const float *src, size_t n) {
#if DEODR_HAVE_X86
    if (has_avx512()) return f32_add_f64_avx512(dst, src, n);
#endif
    for (size_t i = 0; i < n; i++) dst[i] += (double)src[i];
}

void deodr_simd_copy(void *dst, const void *src, size_t bytes) {
#if DEODR_HAVE_X86
    if (has_avx512()) return copy_avx512((char *)dst, (const char *)src, bytes);
#endif
    memcpy(dst, src, bytes);
}

void deodr_simd_zero(void *dst, size_t bytes) {
#if DEODR_HAVE_X86
    if (has_avx512()) return zero_avx512((char *)dst, bytes);
#endif
    memset(dst, 0, bytes);
}

int deodr_simd_equal_f32(const double *user, const float *mirror, size_t n) {
#if DEODR_HAVE_X86
    if (has_avx512()) return equal_f32_avx512(user, mirror, n);
#endif
    int ok = 1;
    for (size_t i = 0; i < n; i++) {
        const float f = (float)user[i];
        ok &= (memcmp(&f, &mirror[i], sizeof(float)) == 0);
    }
    return ok;
}
