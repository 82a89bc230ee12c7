// K3 / K6 / K7: assembly of the observation system and application of the kriging weights.
//
// Replaces (reference, paths relative to its root):
//   glomar_gridding/kriging.py:337-342, :571-583   obs_obs_cov / obs_grid_cov gathers, "+= error_cov", Lagrange row
//   glomar_gridding/kriging.py:417, :677-679       kriging_weights @ obs
//   glomar_gridding/kriging.py:432-435, :694-701   diag(C - W @ C_cross) [- alpha]
//   glomar_gridding/kriging.py:487-491, :772-776   diag(C[:, idx] @ W_simple^T) / diag(C)
//   glomar_gridding/stochastic.py:233-243, :292-297  the same for StochasticKriging
// The reference forms M x M products to read their diagonals (2 M^2 n flops and an M^2 temporary each); here
// one pass over T = S^-1 [C_cross | 1] and C_cross produces the three column reductions everything derives from:
//   yT[m] = sum_a y[a] T[a, m]      d[m] = sum_a T[a, m] Cx[a, m]      oc[m] = sum_a o[a] Cx[a, m]
// The pass is HBM-bound (2 n M 8 bytes read); partial sums over 256-row slabs are written to a workspace and
// combined in a fixed order by the finalising kernel, so results are bit-reproducible run to run.
#include "gmg_common.cuh"
#include "../../include/glomar_b200.h"

namespace {

constexpr int kSlab = 256;            // rows per partial sum
constexpr int kColsPerBlock = 256;    // one column per thread: a warp reads 256 contiguous bytes per row

static_assert(sizeof(GmgKrigeDesc) == 104, "GmgKrigeDesc layout is part of the C ABI");

constexpr int kGatherRows = 8;        // rows of S / Cx / T per CTA of the gather kernel
constexpr int kGatherCols = 512;      // columns per CTA: two per thread

// Error-covariance entry (a, m) of system d, NaN read as 0.0 (kriging.py:120); `have` = false where a packed
// symmetric E holds nothing the Cholesky path reads.
__device__ __forceinline__ double error_cov_load(const GmgKrigeDesc& d, int64_t off) {
    // float32 error covariances (HadCRUT's are) travel and stay as float32: widened here, exactly
    return (d.ordinary & GMG_KRIGE_E_FLOAT32) ? (double)reinterpret_cast<const float*>(d.E)[off] : d.E[off];
}

__device__ __forceinline__ double error_cov_entry(const GmgKrigeDesc& d, int a, int m) {
    double e = 0.0;
    if (d.lde > 0) {
        e = error_cov_load(d, (int64_t)a * d.lde + m);
    } else {
        // block-trapezoid packed lower triangle of a SYMMETRIC E (see GmgKrigeDesc): row i of row block
        // b = i / hb holds its first L_b = min(n, hb (b + 1)) entries.  The Cholesky path only reads the
        // lower triangle and the diagonal tiles of S, all inside the trapezoid; the full-solve (LU) path,
        // recognisable by Cx != NULL, gets the mirrored entries as well.
        const int hb = (int)(-d.lde);
        int i = a, j = m;
        const int La = min(d.n, hb * (a / hb + 1));
        bool have = j < La;
        if (!have && d.Cx) { i = m; j = a; have = true; }
        if (have) {
            const int b = i / hb;
            const int Lb = min(d.n, hb * (b + 1));
            e = error_cov_load(d, (int64_t)hb * hb * ((int64_t)b * (b + 1) / 2) + (int64_t)(i - hb * b) * Lb + j);
        }
    }
    return e != e ? 0.0 : e;
}

// grid = (ceil((M + 2) / 512), ceil(n_max / 8), batch): rows a0 .. a0 + 7 of S, Cx and T for every system.
// A thread owns two adjacent columns (one 16-byte load and one 16-byte store per row and matrix when the leading
// dimensions and base pointers allow it) of eight rows, all loads issued before the first store, so that a CTA moves 64 KB instead of the 4 KB of the first
// version (one row, one column per thread: 8.4 M CTAs for 480 months, bound by the CTA launch rate at 1.8 TB/s).
// t_from > 0: only the columns >= t_from of T are written (S as always) and the grid's LAST column block is the one
// that holds them (`tail_block`, or -1 when it is one of the S blocks anyway) - gmg_krige_solve_f64, whose solve reads
// C[idx, :] itself.
__global__ void __launch_bounds__(256) gather_kernel(const double* __restrict__ C, int64_t M, int64_t ldc,
                                                     const GmgKrigeDesc* __restrict__ descs, int64_t t_from,
                                                     int tail_block) {
    const GmgKrigeDesc d = descs[blockIdx.z];
    const int a0 = blockIdx.y * kGatherRows;
    if (a0 >= d.n) return;
    const int64_t xblk = (tail_block >= 0 && blockIdx.x == gridDim.x - 1) ? tail_block : blockIdx.x;
    const int64_t m0 = 2 * (xblk * (int64_t)blockDim.x + threadIdx.x);
    const bool vec = (((ldc | d.ldt | d.lds | d.ldx) & 1) == 0) &&
                     (((reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(d.T) |
                        reinterpret_cast<uintptr_t>(d.S) | reinterpret_cast<uintptr_t>(d.Cx)) & 15) == 0);
    // S columns of this thread: C[idx[a], idx[m]] is a gather, one index per column for all rows
    // Observation cells are validated by the host (`KrigeBatch`: IndexError outside [-M, M), negatives wrapped like
    // numpy); a caller of the C entry point that skips that gets CLAMPED cells here - wrong values for those
    // observations, never an out-of-bounds read (a faulting read takes the CUDA context with it).
    const int64_t top = M - 1;
    const bool s0 = m0 < d.n, s1 = m0 + 1 < d.n;
    const int64_t j0 = s0 ? min(max(d.idx[m0], (int64_t)0), top) : 0;
    const int64_t j1 = s1 ? min(max(d.idx[m0 + 1], (int64_t)0), top) : 0;
    // Rows past the end of a ragged last block are clamped to the last row: they redo its (identical) stores, and
    // the loop bodies stay branch-free so that all loads of the kGatherRows rows are in flight together.
    const double* crow[kGatherRows];
    int arow[kGatherRows];
#pragma unroll
    for (int r = 0; r < kGatherRows; ++r) {
        arow[r] = min(a0 + r, d.n - 1);
        crow[r] = C + min(max(d.idx[arow[r]], (int64_t)0), top) * ldc;
    }
    // ---- T = [C[idx, :] | 1 | y] (and the C_cross copy of the full-solve paths) ----------------------------------
    if (m0 + 1 < t_from) {
        // nothing of T for these two columns
    } else if (m0 + 1 < M && vec) {
        double2 v[kGatherRows];
#pragma unroll
        for (int r = 0; r < kGatherRows; ++r) v[r] = *reinterpret_cast<const double2*>(crow[r] + m0);
#pragma unroll
        for (int r = 0; r < kGatherRows; ++r) {
            if (d.Cx) *reinterpret_cast<double2*>(d.Cx + (int64_t)arow[r] * d.ldx + m0) = v[r];
            *reinterpret_cast<double2*>(d.T + (int64_t)arow[r] * d.ldt + m0) = v[r];
        }
    } else {
#pragma unroll
        for (int r = 0; r < kGatherRows; ++r) {
            const int a = arow[r];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int64_t m = m0 + k;
                if (m < t_from) {
                } else if (m < M) {
                    const double v = crow[r][m];
                    if (d.Cx) d.Cx[(int64_t)a * d.ldx + m] = v;
                    d.T[(int64_t)a * d.ldt + m] = v;
                } else if (m == M) {
                    d.T[(int64_t)a * d.ldt + M] = 1.0;          // right-hand side for o = S^-1 1 (Lagrange column, :579-583)
                } else if (m == M + 1) {
                    d.T[(int64_t)a * d.ldt + M + 1] = d.y[a];   // right-hand side for S^-1 y
                }
            }
        }
    }
    // ---- S = C[idx, idx] + E (:337, :342, :571, :576): the first ceil(n / 512) column blocks ------------------
    if (s0) {
        double v0[kGatherRows], v1[kGatherRows];
#pragma unroll
        for (int r = 0; r < kGatherRows; ++r) {
            v0[r] = crow[r][j0];
            v1[r] = s1 ? crow[r][j1] : 0.0;
        }
        if (d.E) {
#pragma unroll
            for (int r = 0; r < kGatherRows; ++r) {
                v0[r] += error_cov_entry(d, arow[r], (int)m0);
                if (s1) v1[r] += error_cov_entry(d, arow[r], (int)m0 + 1);
            }
        }
#pragma unroll
        for (int r = 0; r < kGatherRows; ++r) {
            double* srow = d.S + (int64_t)arow[r] * d.lds + m0;
            if (s1 && vec) {
                *reinterpret_cast<double2*>(srow) = make_double2(v0[r], v1[r]);
            } else {
                srow[0] = v0[r];
                if (s1) srow[1] = v1[r];
            }
        }
    }
}

// Lower triangle (diagonal included) of the error covariance of system b as a plain row-major n x n matrix: what a
// Cholesky factorisation of E itself needs (the noise draws of the stochastic ensemble, stochastic.py:379-383) when E
// travelled as its packed block trapezoid.  NaN -> 0 like the gather.  grid = (ceil(n / 256), n)
__global__ void __launch_bounds__(256) error_cov_lower_kernel(const GmgKrigeDesc* __restrict__ descs, int b,
                                                              double* __restrict__ out, int64_t ldo) {
    const GmgKrigeDesc d = descs[b];
    const int a = blockIdx.y;
    const int m = blockIdx.x * 256 + threadIdx.x;
    if (a >= d.n || m >= d.n) return;
    out[(int64_t)a * ldo + m] = (m <= a && d.E) ? error_cov_entry(d, a, m) : 0.0;
}

// Forward-only variant: T holds Y = L^-1 [Cx | 1 | y] (only the FIRST triangular solve has been done).  Because
// S^-1 = L^-T L^-1, every quantity the kriging classes return is a column dot product of Y with itself or with
// its two extra columns:
//   d[m]  = Cx[:,m]^T S^-1 Cx[:,m] = sum_a Y[a,m]^2          oc[m] = 1^T S^-1 Cx[:,m] = sum_a Y[a,M]   Y[a,m]
//   yT[m] = y^T S^-1 Cx[:,m]       = sum_a Y[a,M+1] Y[a,m]
// so the second triangular solve (half of the flops of the whole path) is only needed when the caller asks for
// the weights matrix itself.  One read of Y (n M 8 bytes) instead of T and Cx.
// grid = (ceil(M / 256), ceil(n_max / kSlab), batch)
__global__ void __launch_bounds__(kColsPerBlock) reduce_fwd_kernel(int64_t M, const GmgKrigeDesc* __restrict__ descs) {
    const GmgKrigeDesc d = descs[blockIdx.z];
    const int a0 = blockIdx.y * kSlab;
    if (a0 >= d.n) return;
    const int a1 = min(d.n, a0 + kSlab);
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    __shared__ double sy[kSlab], so[kSlab];
    for (int e = threadIdx.x; e < a1 - a0; e += blockDim.x) {
        so[e] = d.T[(int64_t)(a0 + e) * d.ldt + M];
        sy[e] = d.T[(int64_t)(a0 + e) * d.ldt + M + 1];
    }
    __syncthreads();
    if (m >= M) return;
    double yT = 0.0, dd = 0.0, oc = 0.0;
    const double* t = d.T + (int64_t)a0 * d.ldt + m;
#pragma unroll 4
    for (int a = 0; a < a1 - a0; ++a) {
        const double tv = t[(int64_t)a * d.ldt];
        yT = fma(sy[a], tv, yT);
        dd = fma(tv, tv, dd);
        oc = fma(so[a], tv, oc);
    }
    const int64_t slabs = (d.n + kSlab - 1) / kSlab;
    double* p = d.partial + (int64_t)blockIdx.y * M + m;
    p[0] = yT;
    p[slabs * M] = dd;
    p[2 * slabs * M] = oc;
}

// grid = (ceil(M / 256), ceil(n_max / kSlab), batch)
__global__ void __launch_bounds__(kColsPerBlock) reduce_kernel(int64_t M, const GmgKrigeDesc* __restrict__ descs) {
    const GmgKrigeDesc d = descs[blockIdx.z];
    const int a0 = blockIdx.y * kSlab;
    if (a0 >= d.n) return;
    const int a1 = min(d.n, a0 + kSlab);
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    __shared__ double sy[kSlab], so[kSlab];
    for (int e = threadIdx.x; e < a1 - a0; e += blockDim.x) {
        sy[e] = d.y[a0 + e];
        so[e] = (d.ordinary & 1) ? d.T[(int64_t)(a0 + e) * d.ldt + M] : 0.0;
    }
    __syncthreads();
    if (m >= M) return;
    double yT = 0.0, dd = 0.0, oc = 0.0, st = 0.0;
    const double* t = d.T + (int64_t)a0 * d.ldt + m;
    const double* x = d.Cx + (int64_t)a0 * d.ldx + m;
#pragma unroll 4
    for (int a = 0; a < a1 - a0; ++a) {
        const double tv = t[(int64_t)a * d.ldt], xv = x[(int64_t)a * d.ldx];
        yT = fma(sy[a], tv, yT);
        dd = fma(tv, xv, dd);
        oc = fma(so[a], xv, oc);
        st += tv;                                // 1^T S^-1 c: equals o^T c only for a symmetric S
    }
    const int64_t slabs = (d.n + kSlab - 1) / kSlab;
    double* p = d.partial + (int64_t)blockIdx.y * M + m;
    p[0] = yT;
    p[slabs * M] = dd;
    p[2 * slabs * M] = oc;
    p[3 * slabs * M] = st;
}

// grid = (ceil(M / 256), 1, batch)
__global__ void __launch_bounds__(256) finalize_kernel(const double* __restrict__ C, int64_t M, int64_t ldc,
                                                       const GmgKrigeDesc* __restrict__ descs, int fwd_only) {
    const GmgKrigeDesc d = descs[blockIdx.z];
    if (d.n <= 0) return;
    // scalar reductions over the observations: 1^T o and o^T y (each block recomputes them; n is small
    // next to M and the order of summation is fixed)
    __shared__ double red[2][256];
    double so = 0.0, soy = 0.0;
    if (d.ordinary & 1) {
        for (int a = threadIdx.x; a < d.n; a += blockDim.x) {
            const double o = d.T[(int64_t)a * d.ldt + M];
            if (fwd_only) {                      // o = L^-1 1, T[:, M+1] = L^-1 y: 1^T S^-1 1 and 1^T S^-1 y
                so = fma(o, o, so);
                soy = fma(o, d.T[(int64_t)a * d.ldt + M + 1], soy);
            } else {                             // o = S^-1 1
                so += o;
                soy = fma(o, d.y[a], soy);
            }
        }
    }
    red[0][threadIdx.x] = so;
    red[1][threadIdx.x] = soy;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    const double sum_o = red[0][0], o_dot_y = red[1][0];
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int64_t slabs = (d.n + kSlab - 1) / kSlab;
    double yT = 0.0, dd = 0.0, oc = 0.0, st = 0.0;
    for (int64_t s = 0; s < slabs; ++s) {
        yT += d.partial[s * M + m];
        dd += d.partial[(slabs + s) * M + m];
        oc += d.partial[(2 * slabs + s) * M + m];
        if (!fwd_only) st += d.partial[(3 * slabs + s) * M + m];
    }
    const double cmm = C[m * ldc + m];
    double field = yT, rowdot = dd, alpha = 0.0, u2 = cmm - dd;
    if (d.ordinary & 1) {
        // Schur complement of the bordered system (:579-586): alpha = (1^T S^-1 c - 1) / (1^T S^-1 1).  On the
        // Cholesky path S is symmetric and 1^T S^-1 c = o^T c; the full-solve path (LU: indefinite or asymmetric S,
        // user-supplied weights) sums the solved column itself
        alpha = ((fwd_only ? oc : st) - 1.0) / sum_o;
        field = yT - alpha * o_dot_y;            // W_ext @ [y; 0]                              (:677-679)
        rowdot = dd - alpha * oc;                // rowdot(W, Cx^T)
        u2 = cmm - rowdot - alpha - alpha;       // diag(C - W_ext @ Cx_ext) - alpha            (:699-701)
    }
    d.out[m] = field;
    d.out[M + m] = u2;
    d.out[2 * M + m] = dd / cmm;                 // constraint mask uses the SIMPLE weights       (:772-776)
    d.out[3 * M + m] = alpha;
    d.out[4 * M + m] = rowdot;
}

// W[m, a] = T[a, m] - o[a] alpha[m],  W[m, n] = alpha[m]: 32 x 32 transpose tiles through shared memory
__global__ void __launch_bounds__(256) weights_kernel(const double* __restrict__ T, int64_t ldt,
                                                      const double* __restrict__ o, int64_t o_stride,
                                                      const double* __restrict__ alpha, int64_t n, int64_t M,
                                                      double* __restrict__ W, int64_t ldw) {
    __shared__ double tile[32][33];
    const int64_t m0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int64_t a = a0 + r, m = m0 + tx;
        tile[r][tx] = (a < n && m < M) ? T[a * ldt + m] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t m = m0 + r, a = a0 + tx;
        if (m < M && a < n) {
            double v = tile[tx][r];
            if (alpha) v = fma(-o[a * o_stride], alpha[m], v);
            W[m * ldw + a] = v;
        }
    }
    if (alpha && blockIdx.y == 0 && threadIdx.x < 32) {
        const int64_t m = m0 + threadIdx.x;
        if (m < M) W[m * ldw + n] = alpha[m];
    }
}

}  // namespace

extern "C" {

// dense_f64.cu (library-internal, not part of the public header)
int gmg_internal_potrf_solve_gathered(const GmgSolveDesc* descs, int batch, int64_t n_max, int64_t nrhs_max,
                                      const double* C, int64_t M, int64_t ldc, const GmgKrigeDesc* kd, cudaStream_t stream);


int gmg_krige_gather_batched_f64(const double* C, int64_t M, int64_t ldc, const GmgKrigeDesc* descs, int batch,
                                 int64_t n_max, gmg_stream_t stream) {
    if (!C) return -1;
    if (M <= 0) return -2;
    if (ldc < M) return -3;
    if (!descs) return -4;
    if (batch <= 0 || batch > 65535) return -5;
    if (n_max <= 0 || n_max > 65535 || n_max > M) return -6;
    dim3 grid((unsigned)gmg_ceil_div(M + 2, kGatherCols), (unsigned)gmg_ceil_div(n_max, kGatherRows),
              (unsigned)batch);                                               // columns M, M + 1: the extra RHS
    gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(C, M, ldc, descs, 0, -1);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_krige_solve_f64(const double* C, int64_t M, int64_t ldc, const GmgKrigeDesc* kd, const GmgSolveDesc* sd, int batch,
                        int64_t n_max, gmg_stream_t stream) {
    if (!C) return -1;
    if (M <= 0) return -2;
    if (ldc < M) return -3;
    if (!kd) return -4;
    if (!sd) return -5;
    if (batch <= 0 || batch > 65535) return -6;
    if (n_max <= 0 || n_max > 65535 || n_max > M) return -7;
    // the on-the-fly rows are fetched with 16-byte cp.async: two columns per chunk, never straddling column M
    if ((M & 1) || (ldc & 1) || (reinterpret_cast<uintptr_t>(C) & 15)) return GMG_ERR_UNSUPPORTED;
    const int64_t sblocks = gmg_ceil_div(n_max, kGatherCols), tail = M / kGatherCols;
    const bool extra = tail >= sblocks;
    dim3 grid((unsigned)(sblocks + (extra ? 1 : 0)), (unsigned)gmg_ceil_div(n_max, kGatherRows), (unsigned)batch);
    gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(C, M, ldc, kd, M, extra ? (int)tail : -1);
    GMG_LAUNCH_CHECK();
    return gmg_internal_potrf_solve_gathered(sd, batch, n_max, M + 2, C, M, ldc, kd, (cudaStream_t)stream);
}

int gmg_krige_error_cov_lower_f64(const GmgKrigeDesc* descs, int b, int64_t n, double* out, int64_t ldo,
                                  gmg_stream_t stream) {
    if (!descs) return -1;
    if (b < 0) return -2;
    if (n <= 0 || n > 65535) return -3;
    if (!out) return -4;
    if (ldo < n) return -5;
    error_cov_lower_kernel<<<dim3((unsigned)gmg_ceil_div(n, 256), (unsigned)n), 256, 0, (cudaStream_t)stream>>>(
        descs, b, out, ldo);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_krige_apply_batched_f64(const double* C, int64_t M, int64_t ldc, const GmgKrigeDesc* descs, int batch,
                                int64_t n_max, int fwd_only, gmg_stream_t stream) {
    if (!C) return -1;
    if (M <= 0) return -2;
    if (ldc < M) return -3;
    if (!descs) return -4;
    if (batch <= 0 || batch > 65535) return -5;
    if (n_max <= 0) return -6;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned cb = (unsigned)gmg_ceil_div(M, kColsPerBlock);
    const dim3 rgrid(cb, (unsigned)gmg_ceil_div(n_max, kSlab), (unsigned)batch);
    if (fwd_only) reduce_fwd_kernel<<<rgrid, kColsPerBlock, 0, s>>>(M, descs);
    else reduce_kernel<<<rgrid, kColsPerBlock, 0, s>>>(M, descs);
    GMG_LAUNCH_CHECK();
    finalize_kernel<<<dim3(cb, 1, (unsigned)batch), 256, 0, s>>>(C, M, ldc, descs, fwd_only);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_krige_weights_f64(const double* T, int64_t ldt, const double* o, int64_t o_stride, const double* alpha,
                          int64_t n, int64_t M, double* W, int64_t ldw, gmg_stream_t stream) {
    if (!T) return -1;
    if (ldt < M) return -2;
    if (alpha && !o) return -3;
    if (n <= 0) return -6;
    if (M <= 0) return -7;
    if (!W) return -8;
    if (ldw < n + (alpha ? 1 : 0)) return -9;
    dim3 grid((unsigned)gmg_ceil_div(M, 32), (unsigned)gmg_ceil_div(n, 32));
    if (grid.y > 65535) return GMG_ERR_UNSUPPORTED;
    weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(T, ldt, o, o_stride, alpha, n, M, W, ldw);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

}  // extern "C"

// ---- bordered (Lagrange-extended) inverse --------------------------------------------------------------------
// replaces `_extended_inverse`, glomar_gridding/kriging.py:815-829 (Guttman 1946 block inverse with B = 1, d = 0):
//   e = S^-1 1,  f = -(1^T e),  K = S^-1 + e e^T / f,  result = [[K, -e/f], [-e^T/f, 1/f]]
namespace {

__global__ void __launch_bounds__(256) rowsum_kernel(const double* __restrict__ A, int64_t n, int64_t ld,
                                                     double* __restrict__ e) {
    // one warp per row, fixed summation order
    const int64_t row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    double s = 0.0;
    for (int64_t j = lane; j < n; j += 32) s += A[row * ld + j];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) e[row] = s;
}

__global__ void __launch_bounds__(256) bordered_kernel(const double* __restrict__ A, int64_t n, int64_t ld,
                                                       const double* __restrict__ e, double* __restrict__ out,
                                                       int64_t ldo) {
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += 256) s += e[j];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    const double finv = 1.0 / (0.0 - red[0]);
    const int64_t i = blockIdx.y;
    const int64_t j = blockIdx.x * 256 + threadIdx.x;
    if (j > n) return;
    double v;
    if (i < n && j < n) v = A[i * ld + j] + e[i] * (finv * e[j]);
    else if (i < n) v = -(finv * e[i]);
    else if (j < n) v = -(finv * e[j]);
    else v = finv;
    out[i * ldo + j] = v;
}

}  // namespace

extern "C" int gmg_extended_inverse_f64(const double* inv, int64_t n, int64_t ld, double* ws_e, double* out,
                                        int64_t ldo, gmg_stream_t stream) {
    if (!inv) return -1;
    if (n <= 0 || n >= 65535) return -2;
    if (ld < n) return -3;
    if (!ws_e) return -4;
    if (!out) return -5;
    if (ldo < n + 1) return -6;
    cudaStream_t s = (cudaStream_t)stream;
    rowsum_kernel<<<(unsigned)gmg_ceil_div(n, 8), 256, 0, s>>>(inv, n, ld, ws_e);
    GMG_LAUNCH_CHECK();
    bordered_kernel<<<dim3((unsigned)gmg_ceil_div(n + 1, 256), (unsigned)(n + 1)), 256, 0, s>>>(inv, n, ld, ws_e, out, ldo);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}
