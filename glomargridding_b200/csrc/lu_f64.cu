// Indefinite-safe fallback of the kriging solve: LU with partial pivoting (the factorisation the reference
// actually uses - np.linalg.solve -> LAPACK dgesv, glomar_gridding/kriging.py:343, :585, :768).
//
// The fast path is the DMMA Cholesky (dense_f64.cu).  It cannot factor the matrices the reference accepts
// without complaint: the raw global ellipse covariance under "Modified_Met_Office" is indefinite (SURVEY.md
// H3), and the reference's own known-answer test feeds a variogram matrix (zero diagonal) as the covariance
// (test/test_kriging.py:78-88).  When Cholesky reports a non-positive pivot the Python layer re-gathers the
// system and calls these kernels, so results stay drop-in; they are written for correctness, not speed
// (one CTA per system for the factorisation, one thread per right-hand-side column for the substitutions).
#include "gmg_common.cuh"
#include "../../include/glomar_b200.h"

namespace {

constexpr int LU_THREADS = 1024;

// grid = (1, 1, batch).  Row-major A (n x n, ld) <- unit-lower L and upper U in place; piv[k] = row swapped
// with row k (stored as int32 in the descriptor's P workspace).  info = k + 1 for an exactly zero pivot.
__global__ void __launch_bounds__(LU_THREADS) getrf_kernel(const GmgSolveDesc* __restrict__ descs) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int n = d.n;
    double* A = d.S;
    const int64_t ld = d.lds;
    int* piv = reinterpret_cast<int*>(d.P);
    __shared__ double s_val[LU_THREADS / 32];
    __shared__ int s_idx[LU_THREADS / 32];
    __shared__ int s_p;
    __shared__ double s_pivinv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = 0; k < n; ++k) {
        // pivot search: first row of maximal |A[i][k]|, i >= k (LAPACK idamax tie rule: lowest index)
        double best = -1.0;
        int bi = k;
        for (int i = k + tid; i < n; i += LU_THREADS) {
            const double v = fabs(A[(int64_t)i * ld + k]);
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = lane < LU_THREADS / 32 ? s_val[lane] : -1.0;
            bi = lane < LU_THREADS / 32 ? s_idx[lane] : 0x7fffffff;
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) {
                s_p = bi;
                piv[k] = bi;
                const double pv = A[(int64_t)bi * ld + k];
                if (pv == 0.0 || pv != pv) atomicCAS(d.info, 0, k + 1);
                s_pivinv = 1.0 / pv;
            }
        }
        __syncthreads();
        const int p = s_p;
        if (p != k) {
            for (int j = tid; j < n; j += LU_THREADS) {
                const double t = A[(int64_t)k * ld + j];
                A[(int64_t)k * ld + j] = A[(int64_t)p * ld + j];
                A[(int64_t)p * ld + j] = t;
            }
        }
        __syncthreads();
        const double pinv = s_pivinv;
        // multipliers and rank-1 update of the trailing block: thread -> (row group, column) so that a warp
        // touches contiguous columns of one row
        const int cols = n - k - 1;
        if (cols > 0) {
            const int cgroups = (cols + 31) / 32;
            const int rows_per_pass = LU_THREADS / 32;
            for (int i = k + 1 + warp; i < n; i += rows_per_pass) {
                const double l = A[(int64_t)i * ld + k] * pinv;
                for (int cg = 0; cg < cgroups; ++cg) {
                    const int j = k + 1 + cg * 32 + lane;
                    if (j < n) A[(int64_t)i * ld + j] = fma(-l, A[(int64_t)k * ld + j], A[(int64_t)i * ld + j]);
                }
                __syncwarp();
                if (lane == 0) A[(int64_t)i * ld + k] = l;
            }
        }
        __syncthreads();
    }
}

// grid = (ceil(nrhs_max / 128), 1, batch): one right-hand-side column per thread, in place in R
__global__ void __launch_bounds__(128) getrs_kernel(const GmgSolveDesc* __restrict__ descs) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int n = d.n;
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (col >= d.nrhs) return;
    const double* A = d.S;
    const int64_t ld = d.lds, ldr = d.ldr;
    const int* piv = reinterpret_cast<const int*>(d.P);
    double* x = d.R + col;
    for (int k = 0; k < n; ++k) {                 // row interchanges
        const int p = piv[k];
        if (p != k) {
            const double t = x[(int64_t)k * ldr];
            x[(int64_t)k * ldr] = x[(int64_t)p * ldr];
            x[(int64_t)p * ldr] = t;
        }
    }
    for (int i = 1; i < n; ++i) {                 // L y = P b (unit lower)
        double s0 = 0.0, s1 = 0.0;
        int j = 0;
        for (; j + 1 < i; j += 2) {
            s0 = fma(A[(int64_t)i * ld + j], x[(int64_t)j * ldr], s0);
            s1 = fma(A[(int64_t)i * ld + j + 1], x[(int64_t)(j + 1) * ldr], s1);
        }
        if (j < i) s0 = fma(A[(int64_t)i * ld + j], x[(int64_t)j * ldr], s0);
        x[(int64_t)i * ldr] -= s0 + s1;
    }
    for (int i = n - 1; i >= 0; --i) {            // U x = y
        double s0 = 0.0, s1 = 0.0;
        int j = i + 1;
        for (; j + 1 < n; j += 2) {
            s0 = fma(A[(int64_t)i * ld + j], x[(int64_t)j * ldr], s0);
            s1 = fma(A[(int64_t)i * ld + j + 1], x[(int64_t)(j + 1) * ldr], s1);
        }
        if (j < n) s0 = fma(A[(int64_t)i * ld + j], x[(int64_t)j * ldr], s0);
        x[(int64_t)i * ldr] = (x[(int64_t)i * ldr] - (s0 + s1)) / A[(int64_t)i * ld + i];
    }
}

}  // namespace

extern "C" int gmg_gesv_batched_f64(const GmgSolveDesc* descs, int batch, int64_t nrhs_max, gmg_stream_t stream) {
    if (!descs) return -1;
    if (batch <= 0 || batch > 65535) return -2;
    if (nrhs_max <= 0) return -3;
    cudaStream_t s = (cudaStream_t)stream;
    getrf_kernel<<<dim3(1, 1, batch), LU_THREADS, 0, s>>>(descs);
    GMG_LAUNCH_CHECK();
    getrs_kernel<<<dim3((unsigned)gmg_ceil_div(nrhs_max, 128), 1, batch), 128, 0, s>>>(descs);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}
