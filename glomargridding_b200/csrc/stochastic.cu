// K9: batched Gaussian draws for the stochastic (perturbation) kriging stage.
//
// Replaces (reference, paths relative to its root):
//   glomar_gridding/stochastic.py:395-502  scipy_mv_normal_draw -> np.random.multivariate_normal (SVD factor,
//                                          legacy MT19937 stream; :460)
// with X = Z L^T, L = chol(cov) (gmg_potrf_batched_f64), Z ~ N(0, I) from a counter-based Philox4x32-10
// generator (Salmon et al., SC'11) and the Box-Muller transform.  The random stream differs from numpy's by
// construction (no reference test pins drawn values); it is reproducible from (seed, offset).
#include "gmg_common.cuh"
#include "../../include/glomar_b200.h"

namespace {

struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(U4 ctr, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        U4 n;
        n.x = hi1 ^ ctr.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ ctr.w ^ k1;
        n.w = lo0;
        ctr = n;
        k0 += W0;
        k1 += W1;
    }
    return ctr;
}

// two uniforms with 53 and 52+ random bits from four 32-bit words, then Box-Muller -> two normals
__global__ void __launch_bounds__(256) normal_kernel(double* __restrict__ out, int64_t n, uint64_t seed,
                                                     uint64_t offset) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;     // pair index
    if (2 * p >= n) return;
    const uint64_t c = offset + (uint64_t)p;
    U4 ctr = {(uint32_t)c, (uint32_t)(c >> 32), 0u, 0u};
    const U4 r = philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)r.x << 32) | r.y, b = ((uint64_t)r.z << 32) | r.w;
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);   // (0, 1]
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);           // [0, 1)
    const double rad = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    out[2 * p] = rad * co;
    if (2 * p + 1 < n) out[2 * p + 1] = rad * s;
}

__global__ void __launch_bounds__(256) tril_kernel(double* __restrict__ A, int64_t n, int64_t ld) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j < n && j > i) A[i * ld + j] = 0.0;
}

}  // namespace

extern "C" {

int gmg_normal_f64(double* out, int64_t n, uint64_t seed, uint64_t offset, gmg_stream_t stream) {
    if (!out) return -1;
    if (n < 0) return -2;
    if (n == 0) return GMG_OK;
    const int64_t pairs = (n + 1) / 2;
    normal_kernel<<<(unsigned)gmg_ceil_div(pairs, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_tril_f64(double* A, int64_t n, int64_t ld, gmg_stream_t stream) {
    if (!A) return -1;
    if (n <= 0 || n > 65535) return -2;
    if (ld < n) return -3;
    tril_kernel<<<dim3((unsigned)gmg_ceil_div(n, 256), (unsigned)n), 256, 0, (cudaStream_t)stream>>>(A, n, ld);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

}  // extern "C"
