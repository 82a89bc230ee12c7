// Host-side helpers of the kriging path (no device code): the exact symmetry test of an error covariance.
//
// The reference adds the error covariance as given, `obs_obs_cov += self.error_cov` (glomar_gridding/kriging.py:342,
// :576), and solves by LU, so an asymmetric matrix is a legal input there.  This implementation uploads only the lower
// block trapezoid of a symmetric matrix and factors by Cholesky; whether a matrix IS symmetric is decided here, entry by
// entry (NaN equal to NaN: both read as 0 on the device), in cache-sized blocks so that the transposed operand is read
// from cache.  Called through ctypes from the packing threads (the GIL is released for the duration of the call).
#include <cstdint>
#include <cstring>
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
