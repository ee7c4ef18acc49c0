// Host-side conversion / copy loops of the host-buffer pipeline (upmix_host.cu), compiled by g++ (not nvcc) so that
// the AVX2 bodies can use intrinsics behind a run-time CPU check.
//
// What they do: main.py hands float64 strided views of one interleaved [n][2] array (main.py:43-50); the device wants
// two planar float32 channels.  The workers of upmix_process_host_ex convert chunk by chunk into pinned staging slots
// and copy finished output pieces from pinned slots into the caller's arrays.  Both sides are bound by host memory
// bandwidth, so the stores are non-temporal (no read-for-ownership of lines that are overwritten whole: a third less
// traffic for a copy, a quarter less for the conversion).
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace {

bool have_avx2() {
    static const bool v = [] {
        const char* e = getenv("UPMIX_HOST_SIMD");
        if (e && atoi(e) == 0) return false;
        __builtin_cpu_init();
        return __builtin_cpu_supports("avx2") != 0;
    }();
    return v;
}

bool use_nt() {
    static const bool v = [] { const char* e = getenv("UPMIX_HOST_NT"); return !(e && atoi(e) == 0); }();
    return v;
}

template <class T>
void pair_scalar(const T* src, int64_t n, float* dl, float* dr) {
    for (int64_t i = 0; i < n; i++) {
        dl[i] = (float)src[2 * i];
        dr[i] = (float)src[2 * i + 1];
    }
}

__attribute__((target("avx2"))) void pair_f64_avx2(const double* src, int64_t n, float* dl, float* dr, bool nt) {
    int64_t i = 0;
    // head: until both destinations are 32-byte aligned (they are equally aligned in the staging slots)
    while (i < n && ((reinterpret_cast<uintptr_t>(dl + i) | reinterpret_cast<uintptr_t>(dr + i)) & 31)) {
        if (((reinterpret_cast<uintptr_t>(dl + i) ^ reinterpret_cast<uintptr_t>(dr + i)) & 31) != 0) { nt = false; break; }
        dl[i] = (float)src[2 * i];
        dr[i] = (float)src[2 * i + 1];
        i++;
    }
    for (; i + 8 <= n; i += 8) {
        const double* p = src + 2 * i;
        const __m128 c0 = _mm256_cvtpd_ps(_mm256_loadu_pd(p));          // L0 R0 L1 R1
        const __m128 c1 = _mm256_cvtpd_ps(_mm256_loadu_pd(p + 4));      // L2 R2 L3 R3
        const __m128 c2 = _mm256_cvtpd_ps(_mm256_loadu_pd(p + 8));
        const __m128 c3 = _mm256_cvtpd_ps(_mm256_loadu_pd(p + 12));
        const __m256 l = _mm256_set_m128(_mm_shuffle_ps(c2, c3, _MM_SHUFFLE(2, 0, 2, 0)), _mm_shuffle_ps(c0, c1, _MM_SHUFFLE(2, 0, 2, 0)));
        const __m256 r = _mm256_set_m128(_mm_shuffle_ps(c2, c3, _MM_SHUFFLE(3, 1, 3, 1)), _mm_shuffle_ps(c0, c1, _MM_SHUFFLE(3, 1, 3, 1)));
        if (nt) {
            _mm256_stream_ps(dl + i, l);
            _mm256_stream_ps(dr + i, r);
        } else {
            _mm256_storeu_ps(dl + i, l);
            _mm256_storeu_ps(dr + i, r);
        }
    }
    for (; i < n; i++) {
        dl[i] = (float)src[2 * i];
        dr[i] = (float)src[2 * i + 1];
    }
    if (nt) _mm_sfence();
}

__attribute__((target("avx2"))) void pair_f32_avx2(const float* src, int64_t n, float* dl, float* dr, bool nt) {
    int64_t i = 0;
    while (i < n && ((reinterpret_cast<uintptr_t>(dl + i) | reinterpret_cast<uintptr_t>(dr + i)) & 31)) {
        if (((reinterpret_cast<uintptr_t>(dl + i) ^ reinterpret_cast<uintptr_t>(dr + i)) & 31) != 0) { nt = false; break; }
        dl[i] = src[2 * i];
        dr[i] = src[2 * i + 1];
        i++;
    }
    for (; i + 8 <= n; i += 8) {
        const __m256 a = _mm256_loadu_ps(src + 2 * i);                  // L0 R0 L1 R1 | L2 R2 L3 R3
        const __m256 b = _mm256_loadu_ps(src + 2 * i + 8);              // L4 R4 L5 R5 | L6 R6 L7 R7
        const __m256 lo = _mm256_shuffle_ps(a, b, _MM_SHUFFLE(2, 0, 2, 0));   // L0 L1 L4 L5 | L2 L3 L6 L7
        const __m256 hi = _mm256_shuffle_ps(a, b, _MM_SHUFFLE(3, 1, 3, 1));
        const __m256 l = _mm256_castpd_ps(_mm256_permute4x64_pd(_mm256_castps_pd(lo), _MM_SHUFFLE(3, 1, 2, 0)));
        const __m256 r = _mm256_castpd_ps(_mm256_permute4x64_pd(_mm256_castps_pd(hi), _MM_SHUFFLE(3, 1, 2, 0)));
        if (nt) {
            _mm256_stream_ps(dl + i, l);
            _mm256_stream_ps(dr + i, r);
        } else {
            _mm256_storeu_ps(dl + i, l);
            _mm256_storeu_ps(dr + i, r);
        }
    }
    for (; i < n; i++) {
        dl[i] = src[2 * i];
        dr[i] = src[2 * i + 1];
    }
    if (nt) _mm_sfence();
}

__attribute__((target("avx2"))) void cvt_f64_avx2(const double* src, int64_t n, float* dst, bool nt) {
    int64_t i = 0;
    for (; i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31); i++) dst[i] = (float)src[i];
    for (; i + 8 <= n; i += 8) {
        const __m256 v = _mm256_set_m128(_mm256_cvtpd_ps(_mm256_loadu_pd(src + i + 4)), _mm256_cvtpd_ps(_mm256_loadu_pd(src + i)));
        if (nt) _mm256_stream_ps(dst + i, v);
        else _mm256_store_ps(dst + i, v);
    }
    for (; i < n; i++) dst[i] = (float)src[i];
    if (nt) _mm_sfence();
}

__attribute__((target("avx2"))) void copy_avx2_nt(float* dst, const float* src, int64_t n) {
    int64_t i = 0;
    for (; i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31); i++) dst[i] = src[i];
    for (; i + 32 <= n; i += 32) {
        const __m256 a = _mm256_loadu_ps(src + i), b = _mm256_loadu_ps(src + i + 8);
        const __m256 c = _mm256_loadu_ps(src + i + 16), d = _mm256_loadu_ps(src + i + 24);
        _mm256_stream_ps(dst + i, a);
        _mm256_stream_ps(dst + i + 8, b);
        _mm256_stream_ps(dst + i + 16, c);
        _mm256_stream_ps(dst + i + 24, d);
    }
    for (; i < n; i++) dst[i] = src[i];
    _mm_sfence();
}

}  // namespace

extern "C" {

// interleaved [n][2] -> two planar float32 channels
void upmix_host_pair_f64(const double* src, int64_t n, float* dl, float* dr) {
    if (have_avx2()) pair_f64_avx2(src, n, dl, dr, use_nt());
    else pair_scalar(src, n, dl, dr);
}
void upmix_host_pair_f32(const float* src, int64_t n, float* dl, float* dr) {
    if (have_avx2()) pair_f32_avx2(src, n, dl, dr, use_nt());
    else pair_scalar(src, n, dl, dr);
}
// one channel, element stride `stride`
void upmix_host_gather_f64(const double* src, int64_t stride, int64_t n, float* dst) {
    if (stride == 1 && have_avx2()) { cvt_f64_avx2(src, n, dst, use_nt()); return; }
    for (int64_t i = 0; i < n; i++) dst[i] = (float)src[i * stride];
}
void upmix_host_copy(float* dst, const float* src, int64_t n);
void upmix_host_gather_f32(const float* src, int64_t stride, int64_t n, float* dst) {
    if (stride == 1) { upmix_host_copy(dst, src, n); return; }
    for (int64_t i = 0; i < n; i++) dst[i] = src[i * stride];
}
// plain copy with non-temporal stores (large pieces whose destination is not read again soon)
void upmix_host_copy(float* dst, const float* src, int64_t n) {
    if (have_avx2() && use_nt() && n >= 1024) copy_avx2_nt(dst, src, n);
    else memcpy(dst, src, (size_t)n * sizeof(float));
}

}  // extern "C"
