// Host-side helpers of the kriging path (no device code): the exact symmetry test of an error covariance.
//
// The reference adds the error covariance as given, `obs_obs_cov += self.error_cov` (glomar_gridding/kriging.py:342,
// :576), and solves by LU, so an asymmetric matrix is a legal input there.  This implementation uploads only the lower
// block trapezoid of a symmetric matrix and factors by Cholesky; whether a matrix IS symmetric is decided here, entry by
// entry (NaN equal to NaN: both read as 0 on the device), in cache-sized blocks so that the transposed operand is read
// from cache.  Called through ctypes from the packing threads (the GIL is released for the duration of the call).
#include <cstdint>
#include <cstring>
#include <emmintrin.h>
#include "../../include/glomar_b200.h"

namespace {

template <typename T>
inline bool same(T a, T b) { return a == b || (a != a && b != b); }

template <typename T>
int block_rows_symmetric(const T* E, int64_t ld, int64_t r0, int64_t r1) {
    constexpr int64_t B = 32;                     // 32 x 32 blocks: 8 KB + 8 KB (double), both operands stay in L1
    for (int64_t i0 = r0; i0 < r1; i0 += B) {
        const int64_t i1 = i0 + B < r1 ? i0 + B : r1;
        for (int64_t j0 = 0; j0 < i1; j0 += B) {  // columns up to and including the diagonal block
            const int64_t j1 = j0 + B < i1 ? j0 + B : i1;
            for (int64_t i = i0; i < i1; ++i) {
                const T* row = E + i * ld;
                for (int64_t j = j0; j < j1; ++j)
                    if (!same(row[j], E[j * ld + i])) return 0;
            }
        }
    }
    return 1;
}

}  // namespace

extern "C" int gmg_host_rows_symmetric(const void* E, int is_f32, int64_t n, int64_t ld, int64_t r0, int64_t r1) {
    if (!E) return -1;
    if (n <= 0 || ld < n) return -3;
    if (r0 < 0 || r1 > n || r0 > r1) return -5;
    return is_f32 ? block_rows_symmetric(static_cast<const float*>(E), ld, r0, r1)
                  : block_rows_symmetric(static_cast<const double*>(E), ld, r0, r1);
}

// Packing of a strided block of a host matrix into a contiguous float64 staging buffer (the pinned buffers of the
// upload pipeline): rows x cols elements, row stride ld_src, float64 or float32 source (widened here).  The staging
// buffer is written once and read by the copy engine, never by this core again - streaming (non-temporal) stores keep
// it out of the cache and roughly double the per-thread rate of a cached copy (numpy.copyto: 4.8 GB/s per thread on
// the GPU box, tools/pin_experiment.py).  Called through ctypes by the packing threads (GIL released).
extern "C" int gmg_host_pack_f64(const void* src, int is_f32, int64_t rows, int64_t cols, int64_t ld_src, double* dst) {
    if (!src) return -1;
    if (rows < 0 || cols < 0 || ld_src < cols) return -3;
    if (!dst) return -6;
    for (int64_t i = 0; i < rows; ++i) {
        double* d = dst + i * cols;
        int64_t j = 0;
        if (is_f32) {
            const float* s = static_cast<const float*>(src) + i * ld_src;
            while (j < cols && (reinterpret_cast<uintptr_t>(d + j) & 15)) { d[j] = (double)s[j]; ++j; }
            for (; j + 2 <= cols; j += 2)
                _mm_stream_pd(d + j, _mm_cvtps_pd(_mm_castsi128_ps(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(s + j)))));
            for (; j < cols; ++j) d[j] = (double)s[j];
        } else {
            const double* s = static_cast<const double*>(src) + i * ld_src;
            while (j < cols && (reinterpret_cast<uintptr_t>(d + j) & 15)) { d[j] = s[j]; ++j; }
            for (; j + 8 <= cols; j += 8) {
                const __m128d a = _mm_loadu_pd(s + j), b = _mm_loadu_pd(s + j + 2);
                const __m128d c = _mm_loadu_pd(s + j + 4), e = _mm_loadu_pd(s + j + 6);
                _mm_stream_pd(d + j, a);
                _mm_stream_pd(d + j + 2, b);
                _mm_stream_pd(d + j + 4, c);
                _mm_stream_pd(d + j + 6, e);
            }
            for (; j + 2 <= cols; j += 2) _mm_stream_pd(d + j, _mm_loadu_pd(s + j));
            for (; j < cols; ++j) d[j] = s[j];
        }
    }
    _mm_sfence();
    return 0;
}

// float32 block that stays float32 (the descriptor's GMG_KRIGE_E_FLOAT32 flag): the same streaming copy, 4-byte elements
extern "C" int gmg_host_pack_f32(const float* src, int64_t rows, int64_t cols, int64_t ld_src, float* dst) {
    if (!src) return -1;
    if (rows < 0 || cols < 0 || ld_src < cols) return -3;
    if (!dst) return -5;
    for (int64_t i = 0; i < rows; ++i) {
        const float* s = src + i * ld_src;
        float* d = dst + i * cols;
        int64_t j = 0;
        while (j < cols && (reinterpret_cast<uintptr_t>(d + j) & 15)) { d[j] = s[j]; ++j; }
        for (; j + 16 <= cols; j += 16) {
            const __m128 a = _mm_loadu_ps(s + j), b = _mm_loadu_ps(s + j + 4);
            const __m128 c = _mm_loadu_ps(s + j + 8), e = _mm_loadu_ps(s + j + 12);
            _mm_stream_ps(d + j, a);
            _mm_stream_ps(d + j + 4, b);
            _mm_stream_ps(d + j + 8, c);
            _mm_stream_ps(d + j + 12, e);
        }
        for (; j + 4 <= cols; j += 4) _mm_stream_ps(d + j, _mm_loadu_ps(s + j));
        for (; j < cols; ++j) d[j] = s[j];
    }
    _mm_sfence();
    return 0;
}
