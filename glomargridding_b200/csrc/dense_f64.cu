// K4 / K5: blocked FP64 Cholesky and triangular solves on DMMA tiles, batched over independent systems.
//
// Replaces (reference, paths relative to its root) the LAPACK calls of the kriging path:
//   glomar_gridding/kriging.py:343, :585, :768   np.linalg.solve(S, C_cross)   (dgesv: LU + two solves)
//   glomar_gridding/stochastic.py:159            np.linalg.inv(S)
// with  S = L L^T  (right-looking blocked Cholesky, trailing update on DMMA tiles)  and
//       S^-1 B = L^-T (L^-1 B)  (blocked substitution in which the 128 x 128 diagonal blocks are inverted
//       once, so that every step is a GEMM on DMMA tiles).
// The observation covariance S = C[idx, idx] + E must be symmetric positive definite; a non-positive pivot
// is reported through `info` exactly like LAPACK potrf (the reference's LU accepts indefinite
// matrices, the Python layer turns info > 0 into numpy.linalg.LinAlgError).
//
// One descriptor per system (device memory, uploaded by the caller); every kernel takes the descriptor
// array and the block step, so a batch of 480 months with different observation counts runs as ONE launch
// per step (blockIdx.z = system), not as 480 streams of small launches.
#include "gmg_common.cuh"
#include "dense_f64.cuh"
#include "../../include/glomar_b200.h"

namespace {

using namespace gmg;

constexpr int NB = 128;                 // block size of the factorisation = GEMM tile
constexpr int DIAG_LD = NB + 1;         // odd leading dimension: conflict-free column access
constexpr int DIAG_THREADS = 512;

static_assert(sizeof(GmgSolveDesc) == 72, "GmgSolveDesc layout is part of the C ABI");

// ---- diagonal block: potrf (unblocked, in shared memory) followed by the in-place inverse of L_kk ---------
// grid = (1, 1, batch).  Reads the lower triangle of S[k0:k0+nb, k0:k0+nb], writes L_kk back into S and
// L_kk^-1 (lower, zero above the diagonal, identity-padded to NB) into Linv + kblk * NB * NB.
__global__ void __launch_bounds__(DIAG_THREADS) potrf_diag_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    if (k0 >= d.n) return;
    const int nb = min(NB, d.n - k0);
    extern __shared__ __align__(16) double sm[];
    double* A = sm;                              // [NB][DIAG_LD]
    __shared__ int s_info;
    const int tid = threadIdx.x;
    if (tid == 0) s_info = 0;
    for (int e = tid; e < NB * NB; e += DIAG_THREADS) {
        const int i = e / NB, j = e % NB;
        double v = (i == j) ? 1.0 : 0.0;         // identity padding of a partial last block
        if (i < nb && j < nb && j <= i) v = d.S[(int64_t)(k0 + i) * d.lds + k0 + j];
        A[i * DIAG_LD + j] = v;
    }
    __syncthreads();
    // right-looking Cholesky: column j scaled by 1/sqrt(pivot), then the rank-1 update of the trailing block
    const int ti = tid % NB, tg = tid / NB;      // row ti, column group tg of 4
    for (int j = 0; j < NB; ++j) {
        const double piv = A[j * DIAG_LD + j];
        if (!(piv > 0.0)) {                      // also NaN
            if (tid == 0 && s_info == 0) s_info = k0 + j + 1;
        }
        const double rinv = 1.0 / sqrt(piv);
        __syncthreads();                         // everybody has read the pivot / column before it changes
        if (tg == 0) {
            if (ti == j) A[j * DIAG_LD + j] = piv * rinv;
            else if (ti > j) A[ti * DIAG_LD + j] *= rinv;
        }
        __syncthreads();
        if (ti > j) {
            const double lij = A[ti * DIAG_LD + j];
            for (int k = j + 1 + tg; k <= ti; k += DIAG_THREADS / NB)
                A[ti * DIAG_LD + k] = fma(-lij, A[k * DIAG_LD + j], A[ti * DIAG_LD + k]);
        }
        // next iteration's first barrier orders these updates before the next scale
        __syncthreads();
    }
    if (tid == 0 && s_info != 0) atomicCAS(d.info, 0, s_info);
    // write L_kk
    for (int e = tid; e < NB * NB; e += DIAG_THREADS) {
        const int i = e / NB, j = e % NB;
        if (i < nb && j <= i) d.S[(int64_t)(k0 + i) * d.lds + k0 + j] = A[i * DIAG_LD + j];
    }
    __syncthreads();
    // in-place inverse of the lower-triangular block, column by column from the right (LAPACK trti2 order):
    //   X[j][j] = 1 / L[j][j];   X[j+1:, j] = -X[j][j] * X[j+1:, j+1:] * L[j+1:, j]
    // Thread groups of 4 share one row (split the dot product, combine with shuffles).
    const int row = tid >> 2, part = tid & 3;    // 128 rows x 4 partial sums
    for (int j = NB - 1; j >= 0; --j) {
        const double xjj = 1.0 / A[j * DIAG_LD + j];
        double sum = 0.0;
        if (row > j) {
            for (int k = j + 1 + part; k <= row; k += 4) sum = fma(A[row * DIAG_LD + k], A[k * DIAG_LD + j], sum);
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        __syncthreads();                         // all reads of column j (still L) done
        if (part == 0) {
            if (row > j) A[row * DIAG_LD + j] = -xjj * sum;
            else if (row == j) A[j * DIAG_LD + j] = xjj;
        }
        __syncthreads();
    }
    double* X = d.Linv + (int64_t)kblk * NB * NB;
    for (int e = tid; e < NB * NB; e += DIAG_THREADS) {
        const int i = e / NB, j = e % NB;
        X[e] = (j <= i) ? A[i * DIAG_LD + j] : 0.0;
    }
}

// ---- panel: L21 = S21 L11^-T  (in place; a CTA owns 128 full rows of the panel) ---------------------------
// grid = (ceil((n_max - k0 - NB) / BM), 1, batch)
__global__ void __launch_bounds__(GEMM_THREADS) potrf_panel_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    const int m = d.n - k0 - NB;                 // rows below the diagonal block
    const int m0 = blockIdx.x * BM;
    if (m <= 0 || m0 >= m) return;
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = d.S + (int64_t)(k0 + NB) * d.lds + k0;     // S21: m x NB
    g.lda = d.lds;
    g.B = d.Linv + (int64_t)kblk * NB * NB;           // op(B) = Linv^T: stored n x k = NB x NB
    g.ldb = NB;
    g.C = nullptr; g.ldc = 2;
    g.D = const_cast<double*>(g.A); g.ldd = d.lds;
    g.m = m; g.n = NB; g.k = NB;
    g.alpha = 1.0; g.beta = 0.0;
    gemm_tile<false, true>(g, m0, 0, sm);
}

// ---- trailing update: S22 -= L21 L21^T, lower tiles only ---------------------------------------------------
// grid = (tiles, tiles, batch) with tiles = ceil((n_max - k0 - NB) / BM); tiles above the diagonal exit.
__global__ void __launch_bounds__(GEMM_THREADS) potrf_update_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    if (blockIdx.x > blockIdx.y) return;         // x = column tile, y = row tile: keep the lower triangle
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    const int m = d.n - k0 - NB;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m <= 0 || m0 >= m || n0 >= m) return;
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = d.S + (int64_t)(k0 + NB) * d.lds + k0;     // L21: m x NB
    g.lda = d.lds;
    g.B = g.A;                                        // op(B) = L21^T: stored n x k
    g.ldb = d.lds;
    g.C = d.S + (int64_t)(k0 + NB) * d.lds + (k0 + NB);
    g.ldc = d.lds;
    g.D = const_cast<double*>(g.C); g.ldd = d.lds;
    g.m = m; g.n = m; g.k = NB;
    g.alpha = -1.0; g.beta = 1.0;
    gemm_tile<false, true>(g, m0, n0, sm);
}

// ---- triangular solves with the inverted diagonal blocks ---------------------------------------------------
// Forward (L Y = B), block row k:
//   Y_k = Linv_kk B_k - (Linv_kk L[k, 0:k]) Y_{0:k}  =  [ -Lt_k | Linv_kk ] [ Y_{0:k} ; B_k ]
// where rows 0..(k+1)NB of the right-hand-side buffer ARE [Y_{0:k}; B_k] (in place).  `trsm_prep_fwd` builds
// the NB x (k+1)NB coefficient strip P_k = [ -Linv_kk L[k, 0:k] | Linv_kk ] in the workspace, `trsm_step_fwd`
// is then one GEMM whose CTAs own NB full rows of a column strip (so in place is safe).
// Backward (L^T X = Y), block row k, rows k..nblk of the buffer are [Y_k; X_{k+1:}]:
//   X_k = Linv_kk^T Y_k - Linv_kk^T L[k+1:, k]^T X_{k+1:}  =  Q_k^T [ Y_k ; X_{k+1:} ]
// with Q_k = [ Linv_kk ; -L[k+1:, k] Linv_kk ]  ((n - k NB) x NB, stored row-major: op(A) = Q_k^T).

// grid = (ceil(k0 / BN) , 1, batch): strip[:, 0:k0] = -Linv_kk L[k, 0:k0]; the Linv_kk part is copied.
__global__ void __launch_bounds__(GEMM_THREADS) trsm_prep_fwd_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    if (k0 >= d.n) return;
    const int64_t ldp = d.ldp;                   // leading dimension of the coefficient strip (>= n, even)
    const double* X = d.Linv + (int64_t)kblk * NB * NB;
    const int nb = min(NB, d.n - k0);
    if ((int)blockIdx.x * BN >= k0) {
        // last CTA: copy Linv_kk into columns k0..k0+NB of the strip
        if ((int)blockIdx.x * BN == k0 || k0 == 0) {
            for (int e = threadIdx.x; e < NB * NB; e += GEMM_THREADS) {
                const int i = e / NB, j = e % NB;
                if (k0 + j < ldp) d.P[(int64_t)i * ldp + k0 + j] = X[e];
            }
        }
        return;
    }
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = X; g.lda = NB;                         // Linv_kk: NB x NB
    g.B = d.S + (int64_t)k0 * d.lds;             // L[k, 0:k0]: nb x k0 (rows >= nb are zero-filled by the loader)
    g.ldb = d.lds;
    g.C = nullptr; g.ldc = 2;
    g.D = d.P; g.ldd = ldp;
    g.m = NB; g.n = k0; g.k = nb;
    g.alpha = -1.0; g.beta = 0.0;
    gemm_tile<false, false>(g, 0, blockIdx.x * BN, sm);
}

// grid = (ceil(nrhs_max / BN), 1, batch)
__global__ void __launch_bounds__(GEMM_THREADS) trsm_step_fwd_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    const int n0 = blockIdx.x * BN;
    if (k0 >= d.n || n0 >= d.nrhs) return;
    const int nb = min(NB, d.n - k0);
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = d.P; g.lda = d.ldp;                    // P_k: NB x (k0 + nb)
    g.B = d.R; g.ldb = d.ldr;                    // rows 0..k0+nb of the RHS buffer
    g.C = nullptr; g.ldc = 2;
    g.D = d.R + (int64_t)k0 * d.ldr; g.ldd = d.ldr;
    g.m = nb; g.n = d.nrhs; g.k = k0 + nb;
    g.alpha = 1.0; g.beta = 0.0;
    gemm_tile<false, false>(g, 0, n0, sm);
}

// grid = (ceil((n_max - k0) / BM), 1, batch): Q_k rows: [Linv_kk ; -L[k+1:, k] Linv_kk] into P (ld NB)
__global__ void __launch_bounds__(GEMM_THREADS) trsm_prep_bwd_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    if (k0 >= d.n) return;
    const double* X = d.Linv + (int64_t)kblk * NB * NB;
    const int m = d.n - k0 - NB;                 // rows below the diagonal block
    if (blockIdx.x == 0) {
        for (int e = threadIdx.x; e < NB * NB; e += GEMM_THREADS) d.P[e] = X[e];
        return;
    }
    const int m0 = (blockIdx.x - 1) * BM;
    if (m <= 0 || m0 >= m) return;
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = d.S + (int64_t)(k0 + NB) * d.lds + k0; g.lda = d.lds;   // L[k+1:, k]: m x NB
    g.B = X; g.ldb = NB;                                           // Linv_kk: NB x NB
    g.C = nullptr; g.ldc = 2;
    g.D = d.P + (int64_t)NB * NB; g.ldd = NB;
    g.m = m; g.n = NB; g.k = NB;
    g.alpha = -1.0; g.beta = 0.0;
    gemm_tile<false, false>(g, m0, 0, sm);
}

// grid = (ceil(nrhs_max / BN), 1, batch)
__global__ void __launch_bounds__(GEMM_THREADS) trsm_step_bwd_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    const int n0 = blockIdx.x * BN;
    if (k0 >= d.n || n0 >= d.nrhs) return;
    const int nb = min(NB, d.n - k0);
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = d.P; g.lda = NB;                       // op(A) = Q_k^T: stored k x m = (n - k0) x NB
    g.B = d.R + (int64_t)k0 * d.ldr; g.ldb = d.ldr;
    g.C = nullptr; g.ldc = 2;
    g.D = d.R + (int64_t)k0 * d.ldr; g.ldd = d.ldr;
    g.m = nb; g.n = d.nrhs; g.k = d.n - k0;
    g.alpha = 1.0; g.beta = 0.0;
    gemm_tile<true, false>(g, 0, n0, sm);
}

// ---- plain GEMM entry point ---------------------------------------------------------------------------------
template <bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_kernel(const TileArgs g) {
    extern __shared__ __align__(16) double sm[];
    gemm_tile<TA, TB>(g, blockIdx.y * BM, blockIdx.x * BN, sm);
}

template <typename K>
int set_smem(K kern, size_t bytes) {
    GMG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return GMG_OK;
}

constexpr size_t DIAG_SMEM = (size_t)NB * DIAG_LD * sizeof(double);

int ensure_attrs() {
    static bool done = false;    // per process; the attribute is per device function, set on the current device
    static int done_dev = -1;
    int dev = 0;
    GMG_CUDA_CHECK(cudaGetDevice(&dev));
    if (done && done_dev == dev) return GMG_OK;
    int rc;
    if ((rc = set_smem(potrf_diag_kernel, DIAG_SMEM))) return rc;
    if ((rc = set_smem(potrf_panel_kernel, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(potrf_update_kernel, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_prep_fwd_kernel, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_step_fwd_kernel, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_prep_bwd_kernel, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_step_bwd_kernel, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(gemm_kernel<false, false>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(gemm_kernel<false, true>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(gemm_kernel<true, false>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(gemm_kernel<true, true>, GEMM_SMEM_BYTES))) return rc;
    done = true;
    done_dev = dev;
    return GMG_OK;
}

}  // namespace

extern "C" {

int gmg_gemm_f64(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                 const double* B, int64_t ldb, double beta, double* C, int64_t ldc, gmg_stream_t stream) {
    if (m < 0 || m > 0x7fffffff) return -3;
    if (n < 0 || n > 0x7fffffff) return -4;
    if (k < 0 || k > 0x7fffffff) return -5;
    if (!A || (lda & 1) || (reinterpret_cast<uintptr_t>(A) & 15)) return -7;
    if (!B || (ldb & 1) || (reinterpret_cast<uintptr_t>(B) & 15)) return -9;
    if (!C) return -12;
    if (m == 0 || n == 0) return GMG_OK;
    int rc = ensure_attrs();
    if (rc) return rc;
    TileArgs g;
    g.A = A; g.B = B; g.C = C; g.D = C;
    g.m = (int)m; g.n = (int)n; g.k = (int)k;
    g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.ldd = ldc;
    g.alpha = alpha; g.beta = beta;
    dim3 grid((unsigned)gmg_ceil_div(n, BN), (unsigned)gmg_ceil_div(m, BM));
    cudaStream_t s = (cudaStream_t)stream;
    if (!transa && !transb) gemm_kernel<false, false><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(g);
    else if (!transa && transb) gemm_kernel<false, true><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(g);
    else if (transa && !transb) gemm_kernel<true, false><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(g);
    else gemm_kernel<true, true><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(g);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_solve_ws_doubles(int64_t n, int64_t* linv_doubles, int64_t* p_doubles, int64_t* ldp) {
    if (n <= 0) return -1;
    const int64_t nblk = gmg_ceil_div(n, NB);
    const int64_t ld = nblk * NB;                          // even, >= n
    if (linv_doubles) *linv_doubles = nblk * NB * NB;
    // the coefficient strip is NB x ldp (forward) or (n rounded up) x NB (backward): same size
    if (p_doubles) *p_doubles = (int64_t)NB * ld;
    if (ldp) *ldp = ld;
    return GMG_OK;
}

int gmg_potrf_batched_f64(const GmgSolveDesc* descs, int batch, int64_t n_max, gmg_stream_t stream) {
    if (!descs) return -1;
    if (batch <= 0 || batch > 65535) return -2;
    if (n_max <= 0 || n_max > 0x7fffffff) return -3;
    int rc = ensure_attrs();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int nblk = (int)gmg_ceil_div(n_max, NB);
    for (int k = 0; k < nblk; ++k) {
        potrf_diag_kernel<<<dim3(1, 1, batch), DIAG_THREADS, DIAG_SMEM, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        const int64_t rem = n_max - (int64_t)(k + 1) * NB;
        if (rem <= 0) break;
        const unsigned tiles = (unsigned)gmg_ceil_div(rem, BM);
        potrf_panel_kernel<<<dim3(tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        potrf_update_kernel<<<dim3(tiles, tiles, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
    }
    return GMG_OK;
}

int gmg_potrs_batched_f64(const GmgSolveDesc* descs, int batch, int64_t n_max, int64_t nrhs_max, gmg_stream_t stream) {
    if (!descs) return -1;
    if (batch <= 0 || batch > 65535) return -2;
    if (n_max <= 0 || n_max > 0x7fffffff) return -3;
    if (nrhs_max <= 0 || nrhs_max > 0x7fffffff) return -4;
    int rc = ensure_attrs();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int nblk = (int)gmg_ceil_div(n_max, NB);
    const unsigned rhs_tiles = (unsigned)gmg_ceil_div(nrhs_max, BN);
    for (int k = 0; k < nblk; ++k) {             // L Y = B
        trsm_prep_fwd_kernel<<<dim3((unsigned)k + 1, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        trsm_step_fwd_kernel<<<dim3(rhs_tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
    }
    for (int k = nblk - 1; k >= 0; --k) {        // L^T X = Y
        const int64_t rem = n_max - (int64_t)(k + 1) * NB;
        const unsigned tiles = 1 + (rem > 0 ? (unsigned)gmg_ceil_div(rem, BM) : 0);
        trsm_prep_bwd_kernel<<<dim3(tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        trsm_step_bwd_kernel<<<dim3(rhs_tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
    }
    return GMG_OK;
}

}  // extern "C"
