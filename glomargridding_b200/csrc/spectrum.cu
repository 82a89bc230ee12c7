// N1: all eigenvalues of a dense symmetric matrix on the device, for `eigenvalue_clip`.
//
// The reference clips eigenvalues from a full numpy.linalg.eigh (glomar_gridding/covariance_tools.py:616, :717 ->
// LAPACK dsyevd).  The clipped matrix  c I + V_k (Lambda_k - c) V_k^T  only needs (a) the eigenVALUES - to find the
// cut k and the replacement constant c - and (b) the spectral projector onto the kept eigenspace, which the host
// side obtains from a scaled Newton-Schulz iteration of sign(C - tau I) on DMMA GEMMs
// (glomargridding_b200/covariance_tools.py).  This file provides (a):
//
//   sytrd_kernel      Householder reduction to tridiagonal form, one cooperative persistent kernel.  Every step is
//                     memory-bound BLAS-2 (a symmetric matrix-vector product and a rank-2 update of the trailing
//                     block), so the kernel is written for bandwidth: the 5-degree matrix (54 MB) stays in the 126 MB
//                     L2, rows are streamed with 16-byte loads by one warp each, the Householder vector lives in
//                     shared memory, and the rank-2 update of step j is FUSED with the matrix-vector product of
//                     step j+1 (its Householder vector only needs column j+1 of the updated block, which every CTA
//                     forms for itself), so the trailing block is read and written once per step with one grid-wide
//                     barrier.  Eigenvectors are never accumulated.
//   bisect_kernel     eigenvalue i of the tridiagonal matrix by Sturm-count bisection, one thread per eigenvalue
//                     (the output is sorted by construction).
#include <cooperative_groups.h>
#include "gmg_common.cuh"
#include "../../include/glomar_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int kSyThreads = 256;
constexpr int kSyWarps = kSyThreads / 32;

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// sum over the CTA; every thread gets the result.  `red` = kSyWarps doubles of shared memory.
__device__ __forceinline__ double block_sum(double x, double* red) {
    x = warp_sum(x);
    __syncthreads();                                   // protect `red` from the previous use
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kSyWarps; ++w) s += red[w];
    return s;
}

// LAPACK dlarfg on x = (x0, tail) held in shared memory `v` (length len): H = I - tau v v^T with v0 = 1 maps x to
// (beta, 0, ..., 0).  Every CTA does this redundantly from the same data, so all CTAs agree bit for bit.
__device__ __forceinline__ void make_reflector(double* v, int len, double* red, double& tau, double& beta) {
    double part = 0.0;
    for (int i = 1 + threadIdx.x; i < len; i += kSyThreads) part = fma(v[i], v[i], part);
    const double tail2 = block_sum(part, red);
    const double x0 = v[0];
    if (tail2 == 0.0) {                                // already in tridiagonal form at this column
        tau = 0.0;
        beta = x0;
        __syncthreads();
        if (threadIdx.x == 0) v[0] = 1.0;
        __syncthreads();
        return;
    }
    const double nrm = sqrt(fma(x0, x0, tail2));
    beta = x0 >= 0.0 ? -nrm : nrm;
    tau = (beta - x0) / beta;
    const double scal = 1.0 / (x0 - beta);
    __syncthreads();
    for (int i = 1 + threadIdx.x; i < len; i += kSyThreads) v[i] *= scal;
    if (threadIdx.x == 0) v[0] = 1.0;
    __syncthreads();
}

// One warp: dot product of a row segment (global, `len` doubles from column c0) with a shared-memory vector.
__device__ __forceinline__ double row_dot(const double* __restrict__ row, const double* v, int len, int lane) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int i = lane;
    for (; i + 96 < len; i += 128) {
        a0 = fma(row[i], v[i], a0);
        a1 = fma(row[i + 32], v[i + 32], a1);
        a2 = fma(row[i + 64], v[i + 64], a2);
        a3 = fma(row[i + 96], v[i + 96], a3);
    }
    for (; i < len; i += 32) a0 = fma(row[i], v[i], a0);
    return warp_sum((a0 + a1) + (a2 + a3));
}

// A: n x n symmetric, full storage, row-major (lda); destroyed.  d[n], e[n] (e[n-1] unused) out.
// p: n doubles of global scratch (the matrix-vector product), double-buffered by step parity: 2 n doubles.
__global__ void __launch_bounds__(kSyThreads) sytrd_kernel(double* __restrict__ A, int64_t lda, int n,
                                                           double* __restrict__ d, double* __restrict__ e,
                                                           double* __restrict__ pbuf) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double sh[];
    double* v = sh;                 // current Householder vector          (length n - j - 1, index 0 = row j+1)
    double* w = sh + n;             // w = p - (tau/2)(p.v) v of the step  (same indexing)
    double* vn = sh + 2 * n;        // next step's Householder vector      (index 0 = row j+2)
    double* red = sh + 3 * n;       // kSyWarps doubles
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * kSyWarps + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * kSyWarps;

    if (n == 1) {
        if (blockIdx.x == 0 && threadIdx.x == 0) d[0] = A[0];
        return;
    }
    // ---- step 0 prologue: reflector of column 0 and p = tau A22 v (no previous update to fuse with) ----------
    double tau, beta;
    {
        const int len = n - 1;
        for (int i = threadIdx.x; i < len; i += kSyThreads) v[i] = A[1 + i];          // row 0 = column 0
        __syncthreads();
        make_reflector(v, len, red, tau, beta);
        for (int r = gwarp; r < len; r += nwarps) {
            const double s = row_dot(A + (int64_t)(1 + r) * lda + 1, v, len, lane);
            if (lane == 0) pbuf[r] = tau * s;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) { d[0] = A[0]; e[0] = beta; }
    }
    grid.sync();

    for (int j = 0; j < n - 2; ++j) {
        // trailing block T = A[j+1:, j+1:] (order len), v = reflector of step j (in shared memory), p in pbuf[par]
        const int len = n - j - 1;
        double* p = pbuf + (size_t)(j & 1) * n;
        double* pn = pbuf + (size_t)((j + 1) & 1) * n;
        // w = p - (tau/2)(p.v) v   (every CTA, redundantly)
        double part = 0.0;
        for (int i = threadIdx.x; i < len; i += kSyThreads) {
            const double pi = p[i];
            w[i] = pi;
            part = fma(pi, v[i], part);
        }
        const double K = 0.5 * tau * block_sum(part, red);
        for (int i = threadIdx.x; i < len; i += kSyThreads) w[i] = fma(-K, v[i], w[i]);
        __syncthreads();
        // column 0 of the UPDATED trailing block below its diagonal = x of step j+1:
        //   T'[i, 0] = T[i, 0] - v[i] w[0] - w[i] v[0],  i = 1 .. len-1      (T[i,0] = T[0,i]: row j+1, contiguous)
        const int len1 = len - 1;
        const double* trow0 = A + (int64_t)(j + 1) * lda + (j + 1);
        const double v0 = v[0], w0 = w[0];
        for (int i = threadIdx.x; i < len1; i += kSyThreads)
            vn[i] = trow0[1 + i] - v[1 + i] * w0 - w[1 + i] * v0;
        __syncthreads();
        double tau_n = 0.0, beta_n = 0.0;
        if (len1 >= 2) {
            make_reflector(vn, len1, red, tau_n, beta_n);
        } else {
            beta_n = vn[0];                                   // last off-diagonal element, no reflector needed
        }
        // fused pass over rows 1 .. len-1 of T: rank-2 update of the row, then its dot product with vn
        //   T'[r, c] = T[r, c] - v[r] w[c] - w[r] v[c];   pn[r-1] = tau_n * sum_{c >= 1} T'[r, c] vn[c-1]
        // (row 0 / column 0 of T' are final tridiagonal data: only T'[0,0] and x are needed, handled above/below)
        for (int r = 1 + gwarp; r < len; r += nwarps) {
            double* row = A + (int64_t)(j + 1 + r) * lda + (j + 1);
            const double vr = v[r], wr = w[r];
            double acc0 = 0.0, acc1 = 0.0;
            int c = 1 + lane;
            for (; c + 32 < len; c += 64) {
                const double t0 = row[c] - vr * w[c] - wr * v[c];
                const double t1 = row[c + 32] - vr * w[c + 32] - wr * v[c + 32];
                row[c] = t0;
                row[c + 32] = t1;
                acc0 = fma(t0, vn[c - 1], acc0);
                acc1 = fma(t1, vn[c + 31], acc1);
            }
            for (; c < len; c += 32) {
                const double t0 = row[c] - vr * w[c] - wr * v[c];
                row[c] = t0;
                acc0 = fma(t0, vn[c - 1], acc0);
            }
            const double s = warp_sum(acc0 + acc1);
            if (lane == 0) pn[r - 1] = tau_n * s;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            d[j + 1] = trow0[0] - 2.0 * v0 * w0;              // T'[0, 0]
            e[j + 1] = beta_n;
        }
        // roll: next step's reflector becomes the current one
        __syncthreads();
        for (int i = threadIdx.x; i < len1; i += kSyThreads) v[i] = vn[i];
        tau = tau_n;
        grid.sync();
    }
    // last diagonal element: trailing block of order 1 after the final update (steps ran to j = n-3; the block
    // A[n-1, n-1] was updated as row r = len-1 of every step, column c = len-1 included)
    if (blockIdx.x == 0 && threadIdx.x == 0) d[n - 1] = A[(int64_t)(n - 1) * lda + (n - 1)];
}

// Sturm count: number of eigenvalues of the tridiagonal (d, e) that are < x.  e2 = e^2.
__device__ __forceinline__ int sturm_count(const double* d, const double* e2, int n, double x, double pivmin) {
    int cnt = 0;
    double q = d[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < 0.0;
    for (int k = 1; k < n; ++k) {
        q = d[k] - x - e2[k - 1] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += q < 0.0;
    }
    return cnt;
}

__global__ void __launch_bounds__(128) bisect_kernel(const double* __restrict__ d, const double* __restrict__ e,
                                                     int n, double* __restrict__ w) {
    extern __shared__ __align__(16) double sh[];
    double* sd = sh;
    double* se2 = sh + n;
    __shared__ double s_lo, s_hi, s_emax;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        sd[i] = d[i];
        se2[i] = (i + 1 < n) ? e[i] * e[i] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                   // Gershgorin interval (serial: n is a few thousand)
        double lo = 1e300, hi = -1e300, emax = 0.0;
        for (int i = 0; i < n; ++i) {
            const double l = i > 0 ? sqrt(se2[i - 1]) : 0.0, r = i + 1 < n ? sqrt(se2[i]) : 0.0;
            lo = fmin(lo, sd[i] - l - r);
            hi = fmax(hi, sd[i] + l + r);
            emax = fmax(emax, se2[i]);
        }
        const double span = fmax(hi - lo, fmax(fabs(lo), fabs(hi)));
        s_lo = lo - 1e-12 * span - 1e-300;
        s_hi = hi + 1e-12 * span + 1e-300;
        s_emax = emax;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // eigenvalue index, ascending
    if (i >= n) return;
    const double pivmin = fmax(2.2250738585072014e-308 * fmax(1.0, s_emax), 1e-300);
    double lo = s_lo, hi = s_hi;
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;                    // interval is down to neighbouring doubles
        if (sturm_count(sd, se2, n, mid, pivmin) > i) hi = mid; else lo = mid;
    }
    w[i] = 0.5 * (lo + hi);
}

struct SpecWs { double* d; double* e; double* p; };

inline size_t spec_ws_bytes(int64_t n) { return gmg_align_up((size_t)(4 * n) * sizeof(double), 256); }

}  // namespace

extern "C" {

int gmg_sym_eigvals_ws_bytes(int64_t n, size_t* bytes) {
    if (n <= 0) return -1;
    if (!bytes) return -2;
    *bytes = spec_ws_bytes(n);
    return GMG_OK;
}

int gmg_sym_eigvals_f64(double* A, int64_t lda, int64_t n, double* w, void* ws, size_t ws_bytes,
                        gmg_stream_t stream) {
    if (!A) return -1;
    if (lda < n) return -2;
    if (n <= 0 || n > 8192) return -3;                  // three vectors of n doubles per CTA in shared memory
    if (!w) return -4;
    if (!ws) return -5;
    if (ws_bytes < spec_ws_bytes(n) || (reinterpret_cast<uintptr_t>(ws) & 15)) return GMG_ERR_WORKSPACE;
    double* d = reinterpret_cast<double*>(ws);
    double* e = d + n;
    double* p = d + 2 * n;
    cudaStream_t s = (cudaStream_t)stream;
    int dev = 0, sms = 0, per_sm = 0, coop = 0;
    GMG_CUDA_CHECK(cudaGetDevice(&dev));
    GMG_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return GMG_ERR_UNSUPPORTED;
    GMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = (size_t)(3 * n + kSyWarps) * sizeof(double);
    GMG_CUDA_CHECK(cudaFuncSetAttribute(sytrd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GMG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sytrd_kernel, kSyThreads, smem));
    if (per_sm < 1) return GMG_ERR_UNSUPPORTED;
    if (per_sm > 2) per_sm = 2;                          // rows per step <= n: more CTAs only lengthen the barrier
    int grid = sms * per_sm;
    const int want = (int)gmg_ceil_div(n, kSyWarps);
    if (grid > want) grid = want;
    int ni = (int)n;
    void* args[] = {&A, &lda, &ni, &d, &e, &p};
    GMG_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)sytrd_kernel, dim3(grid), dim3(kSyThreads), args, smem, s));
    const size_t bsmem = (size_t)2 * n * sizeof(double);
    GMG_CUDA_CHECK(cudaFuncSetAttribute(bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
    bisect_kernel<<<(unsigned)gmg_ceil_div(n, 128), 128, bsmem, s>>>(d, e, ni, w);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

}  // extern "C"
