// N2 at speed: blocked LU with partial pivoting on DMMA tiles - the indefinite-safe solve of the kriging path.
//
// The reference solves every kriging system with np.linalg.solve -> LAPACK dgesv (glomar_gridding/kriging.py:343,
// :585, :768), which accepts the matrices Cholesky rejects: the raw global ellipse covariance under the DEFAULT
// displacement method is indefinite, an error covariance may be asymmetric, and the reference's own
// known-answer test feeds a variogram matrix.  lu_f64.cu handles those with one CTA per system (fine for a 3 x 3 test,
// hopeless for a month-sized or a 20 000-observation system).  This file is the same factorisation organised like the
// Cholesky path: right-looking over 128-column panels,
//   panel     one CTA per system: pivot search / row swap / scaling / rank-1 update inside the panel (the only part
//             with partial pivoting's column-by-column dependency; the panel stays L2 resident)
//   laswp     the panel's row interchanges applied to the columns left and right of it
//   blocks    L11^-1 (unit lower) and U11^-1 (upper) of the factored 128 x 128 diagonal block
//   U12       = L11^-1 A12                         one GEMM per column strip, in place
//   update    A22 -= L21 U12                       DMMA tiles (128 x 64, two CTAs per SM)
// and the solve as blocked substitutions whose steps are GEMMs against coefficient strips built from the inverted
// diagonal blocks - the forward one REUSES the Cholesky path's kernels (a unit-lower L stored in place of L), the
// backward one mirrors them for the upper factor.  Pivot choices are those of the unblocked algorithm (the panel is
// fully updated before each search), i.e. LAPACK's.
#include "gmg_common.cuh"
#include "dense_f64.cuh"
#include "../../include/glomar_b200.h"

namespace {

using namespace gmg;

constexpr int NB = 128;
constexpr int PANEL_THREADS = 1024;
constexpr int UBN = 64;

__device__ __forceinline__ int* piv_of(const GmgSolveDesc& d) {
    // pivots live behind the coefficient strip in the P workspace (see gmg_gesv_ws_doubles)
    return reinterpret_cast<int*>(d.P + (int64_t)NB * d.ldp);
}

// ---- panel factorisation with partial pivoting: columns k0 .. k0 + nb, rows k0 .. n.  grid = (1, 1, batch) -----------
__global__ void __launch_bounds__(PANEL_THREADS) lu_panel_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int n = d.n, k0 = kblk * NB;
    if (k0 >= n) return;
    const int nb = min(NB, n - k0);
    double* A = d.S;
    const int64_t ld = d.lds;
    int* piv = piv_of(d);
    __shared__ double s_val[PANEL_THREADS / 32];
    __shared__ int s_idx[PANEL_THREADS / 32];
    __shared__ int s_p;
    __shared__ double s_pivinv;
    __shared__ double s_row[NB];                  // the pivot row of the panel
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = 0; j < nb; ++j) {
        const int k = k0 + j;
        // pivot: first row of maximal |A[i][k]|, i >= k (idamax tie rule: lowest index)
        double best = -1.0;
        int bi = k;
        for (int i = k + tid; i < n; i += PANEL_THREADS) {
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
            best = lane < PANEL_THREADS / 32 ? s_val[lane] : -1.0;
            bi = lane < PANEL_THREADS / 32 ? s_idx[lane] : 0x7fffffff;
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
        // swap rows k and p inside the panel, and stage the (new) pivot row
        if (tid < nb) {
            const double a = A[(int64_t)k * ld + k0 + tid], b = A[(int64_t)p * ld + k0 + tid];
            if (p != k) {
                A[(int64_t)k * ld + k0 + tid] = b;
                A[(int64_t)p * ld + k0 + tid] = a;
            }
            s_row[tid] = (p != k) ? b : a;
        }
        __syncthreads();
        const double pinv = s_pivinv;
        // multipliers and rank-1 update of the panel columns right of j: one warp per row, lanes over columns
        const int c0 = j + 1;
        for (int i = k + 1 + warp; i < n; i += PANEL_THREADS / 32) {
            double* row = A + (int64_t)i * ld + k0;
            const double l = row[j] * pinv;
            for (int c = c0 + lane; c < nb; c += 32) row[c] = fma(-l, s_row[c], row[c]);
            __syncwarp();
            if (lane == 0) row[j] = l;
        }
        __syncthreads();
    }
}

// ---- the panel's row interchanges on the columns outside the panel.  grid = (ceil(n_max / 256), 1, batch) -------------
__global__ void __launch_bounds__(256) lu_laswp_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int n = d.n, k0 = kblk * NB;
    if (k0 >= n) return;
    const int nb = min(NB, n - k0);
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= n || (c >= k0 && c < k0 + nb)) return;
    const int* piv = piv_of(d);
    double* col = d.S + c;
    for (int j = 0; j < nb; ++j) {
        const int k = k0 + j, p = piv[k];
        if (p != k) {
            const double t = col[(int64_t)k * d.lds];
            col[(int64_t)k * d.lds] = col[(int64_t)p * d.lds];
            col[(int64_t)p * d.lds] = t;
        }
    }
}

// ---- all row interchanges on the right-hand sides.  grid = (ceil(nrhs_max / 256), 1, batch) --------------------------
__global__ void __launch_bounds__(256) lu_laswp_rhs_kernel(const GmgSolveDesc* __restrict__ descs) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= d.nrhs) return;
    const int* piv = piv_of(d);
    double* col = d.R + c;
    for (int k = 0; k < d.n; ++k) {
        const int p = piv[k];
        if (p != k) {
            const double t = col[(int64_t)k * d.ldr];
            col[(int64_t)k * d.ldr] = col[(int64_t)p * d.ldr];
            col[(int64_t)p * d.ldr] = t;
        }
    }
}

// ---- inverses of the factored diagonal block: Linv_k = L11^-1 (unit lower), Uinv_k = U11^-1 (upper), both identity
// padded to 128 x 128.  One CTA of 256 threads per system; the factored block sits in shared memory, thread c < 128
// computes column c of L11^-1 by forward substitution, thread 128 + c column c of U11^-1 by backward substitution -
// each thread only ever reads back its OWN column, so the inverses are built directly in their global destination
// (coalesced across threads, L1 resident).
__global__ void __launch_bounds__(256) lu_blocks_kernel(const GmgSolveDesc* __restrict__ descs, int kblk, int nblk_max) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int n = d.n, k0 = kblk * NB;
    if (k0 >= n) return;
    const int nb = min(NB, n - k0);
    extern __shared__ __align__(16) double sm[];
    double* B = sm;                               // [NB][NB + 1] the factored block
    constexpr int LD = NB + 1;
    for (int e = threadIdx.x; e < NB * NB; e += 256) {
        const int i = e / NB, j = e % NB;
        B[i * LD + j] = (i < nb && j < nb) ? d.S[(int64_t)(k0 + i) * d.lds + k0 + j] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    const int c = threadIdx.x & (NB - 1);
    if (threadIdx.x < NB) {
        double* x = d.Linv + (int64_t)kblk * NB * NB;                   // L^-1, column c
        for (int i = 0; i < NB; ++i) {
            double acc = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) acc = fma(-B[i * LD + k], x[k * NB + c], acc);
            x[i * NB + c] = (i < c) ? 0.0 : acc;                        // unit diagonal: no division
        }
    } else {
        double* x = d.Linv + (int64_t)(nblk_max + kblk) * NB * NB;      // U^-1, column c
        for (int i = NB - 1; i >= 0; --i) {
            double acc = (i == c) ? 1.0 : 0.0;
            for (int k = i + 1; k <= c; ++k) acc = fma(-B[i * LD + k], x[k * NB + c], acc);
            x[i * NB + c] = (i > c) ? 0.0 : acc / B[i * LD + i];
        }
    }
}

// ---- U12 = L11^-1 A12, in place (a CTA owns all rows of its 128-column strip).  grid = (ceil(m / 128), 1, batch) --------
__global__ void __launch_bounds__(GEMM_THREADS) lu_u12_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    const int m = d.n - k0 - NB;                  // columns right of the panel
    const int n0 = blockIdx.x * BN;
    if (m <= 0 || n0 >= m) return;
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = d.Linv + (int64_t)kblk * NB * NB; g.lda = NB;
    g.B = d.S + (int64_t)k0 * d.lds + k0 + NB; g.ldb = d.lds;      // A12: NB x m
    g.C = nullptr; g.ldc = 2;
    g.D = const_cast<double*>(g.B); g.ldd = d.lds;
    g.m = NB; g.n = m; g.k = NB;
    g.alpha = 1.0; g.beta = 0.0;
    gemm_tile<false, false>(g, 0, n0, sm);
}

// ---- A22 -= L21 U12.  grid = (2 ceil(m / 128), ceil(m / 128), batch), 128 x 64 tiles ---------------------------------
__global__ void __launch_bounds__(CfgSmall::THREADS, 2) lu_update_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    const int m = d.n - k0 - NB;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * UBN;
    if (m <= 0 || m0 >= m || n0 >= m) return;
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = d.S + (int64_t)(k0 + NB) * d.lds + k0; g.lda = d.lds;    // L21: m x NB
    g.B = d.S + (int64_t)k0 * d.lds + k0 + NB; g.ldb = d.lds;      // U12: NB x m
    g.C = d.S + (int64_t)(k0 + NB) * d.lds + k0 + NB; g.ldc = d.lds;
    g.D = const_cast<double*>(g.C); g.ldd = d.lds;
    g.m = m; g.n = m; g.k = NB;
    g.alpha = -1.0; g.beta = 1.0;
    gemm_tile_cfg<CfgSmall, false, false, UBN>(g, m0, n0, sm);
}

// ---- forward substitution with the unit-lower factor: coefficient strip P_k = [ -Linv_k L[k, 0:k0] | Linv_k ] and one
// GEMM per block row, exactly the Cholesky path's step (dense_f64.cu) - restated here on this file's kernels.
__global__ void __launch_bounds__(GEMM_THREADS) lu_prep_fwd_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    if (k0 >= d.n) return;
    const int64_t ldp = d.ldp;
    const double* X = d.Linv + (int64_t)kblk * NB * NB;
    const int nb = min(NB, d.n - k0);
    const int gemm_ctas = (k0 + BN - 1) / BN;
    if ((int)blockIdx.x >= gemm_ctas) {
        if ((int)blockIdx.x == gemm_ctas)
            for (int e = threadIdx.x; e < NB * NB; e += GEMM_THREADS) {
                const int i = e / NB, j = e % NB;
                if (k0 + j < ldp) d.P[(int64_t)i * ldp + k0 + j] = X[e];
            }
        return;
    }
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = X; g.lda = NB;
    g.B = d.S + (int64_t)k0 * d.lds; g.ldb = d.lds;                 // L[k, 0:k0]
    g.C = nullptr; g.ldc = 2;
    g.D = d.P; g.ldd = ldp;
    g.m = NB; g.n = k0; g.k = nb;
    g.alpha = -1.0; g.beta = 0.0;
    gemm_tile<false, false>(g, 0, blockIdx.x * BN, sm);
}

template <int BNT>
__global__ void __launch_bounds__(GEMM_THREADS) lu_step_fwd_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB, n0 = blockIdx.x * BNT;
    if (k0 >= d.n || n0 >= d.nrhs) return;
    const int nb = min(NB, d.n - k0);
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = d.P; g.lda = d.ldp;
    g.B = d.R; g.ldb = d.ldr;
    g.C = nullptr; g.ldc = 2;
    g.D = d.R + (int64_t)k0 * d.ldr; g.ldd = d.ldr;
    g.m = nb; g.n = d.nrhs; g.k = k0 + nb;
    g.alpha = 1.0; g.beta = 0.0;
    gemm_tile<false, false, BNT>(g, 0, n0, sm);
}

// ---- backward substitution with the upper factor: X_k = Uinv_k (Y_k - U[k, k+1:] X_{k+1:}) = Q_k [Y_k ; X_{k+1:}] with
// Q_k = [ Uinv_k | -Uinv_k U[k, k0 + nb : n] ]  (NB x (n - k0), in the strip workspace, leading dimension ldp).
__global__ void __launch_bounds__(GEMM_THREADS) lu_prep_bwd_kernel(const GmgSolveDesc* __restrict__ descs, int kblk,
                                                                   int nblk_max) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    if (k0 >= d.n) return;
    const int64_t ldp = d.ldp;
    const double* X = d.Linv + (int64_t)(nblk_max + kblk) * NB * NB;
    const int nb = min(NB, d.n - k0);
    const int rest = d.n - k0 - nb;               // columns right of the diagonal block
    if (blockIdx.x == 0) {
        for (int e = threadIdx.x; e < NB * NB; e += GEMM_THREADS) {
            const int i = e / NB, j = e % NB;
            if (j < nb) d.P[(int64_t)i * ldp + j] = X[e];
        }
        return;
    }
    const int n0 = (blockIdx.x - 1) * BN;
    if (rest <= 0 || n0 >= rest) return;
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = X; g.lda = NB;
    g.B = d.S + (int64_t)k0 * d.lds + k0 + nb; g.ldb = d.lds;      // U[k, k0 + nb : n]: nb x rest
    g.C = nullptr; g.ldc = 2;
    g.D = d.P + nb; g.ldd = ldp;
    g.m = NB; g.n = rest; g.k = nb;
    g.alpha = -1.0; g.beta = 0.0;
    gemm_tile<false, false>(g, 0, n0, sm);
}

template <int BNT>
__global__ void __launch_bounds__(GEMM_THREADS) lu_step_bwd_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB, n0 = blockIdx.x * BNT;
    if (k0 >= d.n || n0 >= d.nrhs) return;
    const int nb = min(NB, d.n - k0);
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = d.P; g.lda = d.ldp;                     // Q_k: nb x (n - k0)
    g.B = d.R + (int64_t)k0 * d.ldr; g.ldb = d.ldr;
    g.C = nullptr; g.ldc = 2;
    g.D = d.R + (int64_t)k0 * d.ldr; g.ldd = d.ldr;
    g.m = nb; g.n = d.nrhs; g.k = d.n - k0;
    g.alpha = 1.0; g.beta = 0.0;
    gemm_tile<false, false, BNT>(g, 0, n0, sm);
}

constexpr size_t LU_BLOCKS_SMEM = (size_t)NB * (NB + 1) * sizeof(double);

template <typename K>
int set_smem(K kern, size_t bytes) {
    GMG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return GMG_OK;
}

int lu_attrs() {
    static int done_dev = -1;
    int dev = 0;
    GMG_CUDA_CHECK(cudaGetDevice(&dev));
    if (done_dev == dev) return GMG_OK;
    int rc;
    if ((rc = set_smem(lu_blocks_kernel, LU_BLOCKS_SMEM))) return rc;
    if ((rc = set_smem(lu_u12_kernel, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(lu_update_kernel, GEMM_SMALL_SMEM_BYTES))) return rc;
    if ((rc = set_smem(lu_prep_fwd_kernel, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(lu_step_fwd_kernel<128>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(lu_step_fwd_kernel<64>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(lu_prep_bwd_kernel, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(lu_step_bwd_kernel<128>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(lu_step_bwd_kernel<64>, GEMM_SMEM_BYTES))) return rc;
    done_dev = dev;
    return GMG_OK;
}

}  // namespace

extern "C" {

int gmg_gesv_ws_doubles(int64_t n, int64_t* linv_doubles, int64_t* p_doubles, int64_t* ldp) {
    if (n <= 0) return -1;
    const int64_t nblk = gmg_ceil_div(n, NB);
    const int64_t ld = nblk * NB;
    if (linv_doubles) *linv_doubles = 2 * nblk * NB * NB;              // L^-1 and U^-1 of every diagonal block
    if (p_doubles) *p_doubles = (int64_t)NB * ld + (n + 1) / 2 + 2;     // coefficient strip + n int32 pivots
    if (ldp) *ldp = ld;
    return GMG_OK;
}

int gmg_gesv_blocked_f64(const GmgSolveDesc* descs, int batch, int64_t n_max, int64_t nrhs_max, gmg_stream_t stream) {
    if (!descs) return -1;
    if (batch <= 0 || batch > 65535) return -2;
    if (n_max <= 0 || n_max > 0x7fffffff) return -3;
    if (nrhs_max < 0 || nrhs_max > 0x7fffffff) return -4;
    int rc = lu_attrs();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int nblk = (int)gmg_ceil_div(n_max, NB);
    for (int k = 0; k < nblk; ++k) {
        lu_panel_kernel<<<dim3(1, 1, batch), PANEL_THREADS, 0, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        lu_laswp_kernel<<<dim3((unsigned)gmg_ceil_div(n_max, 256), 1, batch), 256, 0, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        lu_blocks_kernel<<<dim3(1, 1, batch), 256, LU_BLOCKS_SMEM, s>>>(descs, k, nblk);
        GMG_LAUNCH_CHECK();
        const int64_t rem = n_max - (int64_t)(k + 1) * NB;
        if (rem <= 0) break;
        const unsigned tiles = (unsigned)gmg_ceil_div(rem, BM);
        lu_u12_kernel<<<dim3(tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        lu_update_kernel<<<dim3(tiles * (BM / UBN), tiles, batch), CfgSmall::THREADS, GEMM_SMALL_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
    }
    if (nrhs_max == 0) return GMG_OK;
    int dev = 0, sms = 0;
    GMG_CUDA_CHECK(cudaGetDevice(&dev));
    GMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const bool narrow = gmg_ceil_div(gmg_ceil_div(nrhs_max, 64) * (int64_t)batch, sms) * 64.0 * 1.06 <
                        gmg_ceil_div(gmg_ceil_div(nrhs_max, BN) * (int64_t)batch, sms) * 128.0;
    const unsigned rhs_tiles = (unsigned)gmg_ceil_div(nrhs_max, narrow ? 64 : BN);
    lu_laswp_rhs_kernel<<<dim3((unsigned)gmg_ceil_div(nrhs_max, 256), 1, batch), 256, 0, s>>>(descs);
    GMG_LAUNCH_CHECK();
    for (int k = 0; k < nblk; ++k) {                                    // L y = P b
        lu_prep_fwd_kernel<<<dim3((unsigned)k + 1, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        if (narrow) lu_step_fwd_kernel<64><<<dim3(rhs_tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        else lu_step_fwd_kernel<128><<<dim3(rhs_tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
    }
    for (int k = nblk - 1; k >= 0; --k) {                               // U x = y
        const int64_t rest = n_max - (int64_t)(k + 1) * NB;
        const unsigned ctas = 1 + (rest > 0 ? (unsigned)gmg_ceil_div(rest, BN) : 0);
        lu_prep_bwd_kernel<<<dim3(ctas, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k, nblk);
        GMG_LAUNCH_CHECK();
        if (narrow) lu_step_bwd_kernel<64><<<dim3(rhs_tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        else lu_step_bwd_kernel<128><<<dim3(rhs_tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
    }
    return GMG_OK;
}

}  // extern "C"
