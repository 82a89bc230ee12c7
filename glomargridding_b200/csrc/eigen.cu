// N1: small dense symmetric eigensolver for the device `eigenvalue_clip`.
//
// The reference clips eigenvalues from a FULL numpy.linalg.eigh of the covariance
// (glomar_gridding/covariance_tools.py:616, :717 -> LAPACK dsyevd).  The clipped matrix only needs the k kept
// (largest) eigenpairs - everything else is replaced by one constant - so the device path finds those by subspace
// iteration on DMMA GEMMs (glomargridding_b200/covariance_tools.py) and only the small m x m Rayleigh-Ritz problem
// (m ~ 2 k) is diagonalised here, by the cyclic two-sided Jacobi method with the round-robin parallel ordering:
// m/2 disjoint rotations per round, applied as a row phase and a column phase separated by grid-wide barriers
// (cooperative launch).  Jacobi is chosen for its accuracy (eigenvalues to high relative accuracy) and because it
// needs nothing but FP64 FMAs; the matrices are L2 resident (m = 512: 2 MB).
#include <cooperative_groups.h>
#include "gmg_common.cuh"
#include "../../include/glomar_b200.h"

namespace cg = cooperative_groups;

namespace {

__device__ __forceinline__ void round_robin_pair(int m, int r, int k, int& p, int& q) {
    // circle method: m even, rounds r = 0 .. m-2, pairs k = 0 .. m/2-1; every unordered pair meets exactly once
    if (k == 0) {
        p = m - 1;
        q = r % (m - 1);
    } else {
        p = (r + k) % (m - 1);
        q = (r - k + (m - 1)) % (m - 1);
    }
    if (p > q) { const int t = p; p = q; q = t; }
}

__global__ void __launch_bounds__(256) jacobi_eigh_kernel(double* __restrict__ A, int64_t lda, double* __restrict__ U,
                                                          int64_t ldu, int m, int max_sweeps, double tol,
                                                          double* __restrict__ cs, int* __restrict__ pq,
                                                          double* __restrict__ stat /* [0] off, [1] diag */,
                                                          int* __restrict__ sweeps_done) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double rows[];     // 4 x m doubles
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const int half = m / 2;
    for (int64_t e = tid; e < (int64_t)m * m; e += nthr) {
        const int i = (int)(e / m), j = (int)(e % m);
        U[i * ldu + j] = (i == j) ? 1.0 : 0.0;
    }
    grid.sync();
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (tid == 0) { stat[0] = 0.0; stat[1] = 0.0; }
        grid.sync();
        for (int r = 0; r < m - 1; ++r) {
            // phase 0: rotation of every pair of this round (Rutishauser's stable formulas)
            for (int64_t k = tid; k < half; k += nthr) {
                int p, q;
                round_robin_pair(m, r, (int)k, p, q);
                const double app = A[p * lda + p], aqq = A[q * lda + q], apq = A[p * lda + q];
                double c = 1.0, s = 0.0;
                if (apq != 0.0) {
                    const double tau = (aqq - app) / (2.0 * apq);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t * t);
                    s = t * c;
                }
                cs[2 * k] = c;
                cs[2 * k + 1] = s;
                pq[2 * k] = p;
                pq[2 * k + 1] = q;
                // convergence statistics of the sweep (max |off-diagonal| met, max |diagonal|)
                atomicMax(reinterpret_cast<unsigned long long*>(&stat[0]), (unsigned long long)__double_as_longlong(fabs(apq)));
                atomicMax(reinterpret_cast<unsigned long long*>(&stat[1]),
                          (unsigned long long)__double_as_longlong(fmax(fabs(app), fabs(aqq))));
            }
            grid.sync();
            // phase 1 (fused): every CTA takes pairs k = blockIdx.x, += gridDim.x.  Rows p, q of A and rows 2k, 2k+1 of
            // U are staged in shared memory (coalesced), rows p, q get the row rotation of their own pair, then ALL
            // column rotations of the round are applied inside shared memory (the scattered column pairs cost nothing
            // there), and the four rows are written back coalesced:  A' = J^T A J and U' = U J with one read and one
            // write of A and U per round.  (The first version rotated columns in global memory: 17 s for n = 2592.)
            for (int k = blockIdx.x; k < half; k += gridDim.x) {
                const int p = pq[2 * k], q = pq[2 * k + 1];
                double* rp = rows;
                double* rq = rows + m;
                double* u0 = rows + 2 * m;
                double* u1 = rows + 3 * m;
                for (int j = threadIdx.x; j < m; j += blockDim.x) {
                    const double ap = A[p * lda + j], aq = A[q * lda + j];
                    const double c = cs[2 * k], s = cs[2 * k + 1];
                    rp[j] = c * ap - s * aq;                       // row rotation of this pair
                    rq[j] = s * ap + c * aq;
                    u0[j] = U[(int64_t)(2 * k) * ldu + j];
                    u1[j] = U[(int64_t)(2 * k + 1) * ldu + j];
                }
                __syncthreads();
                for (int k2 = threadIdx.x; k2 < half; k2 += blockDim.x) {
                    const double c = cs[2 * k2], s = cs[2 * k2 + 1];
                    const int p2 = pq[2 * k2], q2 = pq[2 * k2 + 1];
                    double x, y;
                    x = rp[p2]; y = rp[q2]; rp[p2] = c * x - s * y; rp[q2] = s * x + c * y;
                    x = rq[p2]; y = rq[q2]; rq[p2] = c * x - s * y; rq[q2] = s * x + c * y;
                    x = u0[p2]; y = u0[q2]; u0[p2] = c * x - s * y; u0[q2] = s * x + c * y;
                    x = u1[p2]; y = u1[q2]; u1[p2] = c * x - s * y; u1[q2] = s * x + c * y;
                }
                __syncthreads();
                for (int j = threadIdx.x; j < m; j += blockDim.x) {
                    A[p * lda + j] = rp[j];
                    A[q * lda + j] = rq[j];
                    U[(int64_t)(2 * k) * ldu + j] = u0[j];
                    U[(int64_t)(2 * k + 1) * ldu + j] = u1[j];
                }
                __syncthreads();
            }
            grid.sync();
        }
        const double off = stat[0], dg = stat[1];
        grid.sync();
        if (off <= tol * dg) { ++sweep; break; }
    }
    if (tid == 0) *sweeps_done = sweep;
}

}  // namespace

extern "C" {

int gmg_jacobi_eigh_ws_bytes(int64_t m, size_t* bytes) {
    if (m <= 0) return -1;
    if (!bytes) return -2;
    *bytes = gmg_align_up((size_t)m * sizeof(double), 256) + gmg_align_up((size_t)m * sizeof(int), 256) + 256;
    return GMG_OK;
}

int gmg_jacobi_eigh_f64(double* A, int64_t lda, double* U, int64_t ldu, int64_t m, int max_sweeps, double tol,
                        void* ws, size_t ws_bytes, int* sweeps_done, gmg_stream_t stream) {
    if (!A) return -1;
    if (lda < m) return -2;
    if (!U) return -3;
    if (ldu < m) return -4;
    if (m < 2 || m > 6400 || (m & 1)) return -5;      // even order (the caller pads); 4 rows must fit shared memory
    if (max_sweeps < 1) return -6;
    if (!ws) return -8;
    size_t need;
    gmg_jacobi_eigh_ws_bytes(m, &need);
    if (ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 15)) return GMG_ERR_WORKSPACE;
    if (!sweeps_done) return -10;
    char* w = reinterpret_cast<char*>(ws);
    double* cs = reinterpret_cast<double*>(w);
    int* pq = reinterpret_cast<int*>(w + gmg_align_up((size_t)m * sizeof(double), 256));
    double* stat = reinterpret_cast<double*>(w + gmg_align_up((size_t)m * sizeof(double), 256) +
                                             gmg_align_up((size_t)m * sizeof(int), 256));
    int dev = 0, sms = 0, per_sm = 0, coop = 0;
    GMG_CUDA_CHECK(cudaGetDevice(&dev));
    GMG_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return GMG_ERR_UNSUPPORTED;
    GMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = (size_t)4 * m * sizeof(double);
    if (smem > 200 * 1024) return GMG_ERR_UNSUPPORTED;
    GMG_CUDA_CHECK(cudaFuncSetAttribute(jacobi_eigh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GMG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_eigh_kernel, 256, smem));
    if (per_sm < 1) return GMG_ERR_UNSUPPORTED;
    int grid = (int)(m / 2);                                       // one pair per CTA and round, if they all fit
    if (grid > sms * per_sm) grid = sms * per_sm;
    int mi = (int)m;
    void* args[] = {&A, &lda, &U, &ldu, &mi, &max_sweeps, &tol, &cs, &pq, &stat, &sweeps_done};
    GMG_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)jacobi_eigh_kernel, dim3(grid), dim3(256), args, smem,
                                               (cudaStream_t)stream));
    return GMG_OK;
}

}  // extern "C"
