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
#include <dlfcn.h>
#include <cstdlib>
#include <mutex>
#include <vector>
#include "gmg_common.cuh"
#include "dense_f64.cuh"
#include "../../include/glomar_b200.h"

namespace {

using namespace gmg;

constexpr int NB = 128;                 // block size of the factorisation = GEMM tile
constexpr int DIAG_LD = NB + 1;         // odd leading dimension: conflict-free column access
constexpr int DIAG_THREADS = 512;

static_assert(sizeof(GmgSolveDesc) == 72, "GmgSolveDesc layout is part of the C ABI");

// ---- diagonal block: 128 x 128 Cholesky and the inverse of its factor, one CTA, all in shared memory ---------
// grid = (1, 1, batch).  Reads the lower triangle of S[k0:k0+nb, k0:k0+nb], writes L_kk back into S and
// L_kk^-1 (lower, zero above the diagonal, identity-padded to NB) into Linv + kblk * NB * NB.
// This kernel is the serial step of the factorisation (one CTA per system per block step), so it is blocked
// with 16 x 16 sub-blocks to cut the number of CTA-wide barriers from ~640 to ~40 (ncu r01: the unblocked
// version took 275 us per call, 70 % of it at barriers):
//   for each 16-column block:  warp 0 factors the 16 x 16 diagonal sub-block;
//                              warp 0 inverts it while warps 1..4 solve the panel rows below it (one row per thread);
//                              all 16 warps apply the rank-16 update to the trailing lower triangle (4 x 4 register tiles)
//   inverse:  X = L^-1 by block forward substitution with the inverted 16 x 16 diagonal sub-blocks.
constexpr int SB = 16;                          // sub-block
constexpr int NSB = NB / SB;                    // 8
constexpr int DIAG_G_LD = NB - SB + 1;          // staging of one block row of L X (16 x 112), odd ld
// The kernel only ever touches the LOWER triangle of the block (factor and inverse are lower triangular), so the
// triangle is stored FOLDED: row i < 64 shares a row of 129 doubles with row 127 - i - the long row from the left,
// the short one reversed from the right (i + 1 and 128 - i entries: 129 together).  66 KB instead of 132: the CTA needs
// 97 KB in all, so two of them - or one and a 92 KB GEMM CTA of the solve stream - fit on an SM (round 1 / early
// round 2: 163 KB, one CTA per SM, 480 months = 3.2 waves of 134 us per block step with the rest of the SM idle).
constexpr int DIAG_ROWS = NB / 2;
// inverted diagonal sub-blocks, rows padded to 17: in the block-row products of the inverse the lanes of a half warp read
// 16 different ROWS of one sub-block at the same column - with a leading dimension of 16 doubles that is one bank, 16-way
// (ncu r02b: 14 % of the kernel's stall samples on that one line)
constexpr int DINV_LD = SB + 1;
constexpr size_t DIAG_SMEM = (size_t)(DIAG_ROWS * DIAG_LD + NSB * SB * DINV_LD + SB * DIAG_G_LD + SB) * sizeof(double);
__device__ __forceinline__ int diag_at(int i, int j) {           // element (i, j), j <= i, of the folded lower triangle
    return i < DIAG_ROWS ? i * DIAG_LD + (NB - j) : (NB - 1 - i) * DIAG_LD + j;
}
#define AT(i, j) A[diag_at((i), (j))]

__global__ void __launch_bounds__(DIAG_THREADS, 2) potrf_diag_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0g = kblk * NB;
    if (k0g >= d.n) return;
    const int nb = min(NB, d.n - k0g);
    extern __shared__ __align__(16) double sm[];
    double* A = sm;                              // [DIAG_ROWS][DIAG_LD], folded lower triangle (see diag_at)
    double* Dinv = A + DIAG_ROWS * DIAG_LD;      // [NSB][SB][DINV_LD] inverted diagonal sub-blocks
    double* G = Dinv + NSB * SB * DINV_LD;       // [SB][DIAG_G_LD]
    double* rd = G + SB * DIAG_G_LD;             // [SB] reciprocal diagonal of the current sub-block
    __shared__ int s_info;
    __shared__ double s_piv, s_col[SB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_info = 0;
    for (int e = tid; e < NB * NB; e += DIAG_THREADS) {
        const int i = e / NB, j = e % NB;
        double v = (i == j) ? 1.0 : 0.0;         // identity padding of a partial last block
        if (i < nb && j < nb && j <= i) v = d.S[(int64_t)(k0g + i) * d.lds + k0g + j];
        if (j <= i) AT(i, j) = v;                // only the lower triangle is stored
    }
    __syncthreads();

    for (int kb = 0; kb < NSB; ++kb) {
        const int k0 = kb * SB;
        // (1) 16 x 16 diagonal sub-block, right-looking, ONE ELEMENT PER THREAD on warps 0..7 (thread 16 i + c holds
        //     A[i][c]); per column: the pivot travels through shared memory, every thread of the column scales its own
        //     entry, the column is posted, every thread right of it applies its one FMA.  Two 256-thread named barriers
        //     per column.  (Round 1 / early round 2 ran this in the registers of a single warp with shuffles: 1 430
        //     straight-line instructions executed by one warp while fifteen waited - a third of the kernel's time, ncu
        //     r02.)  Same operations on the same operands: the factor is bit-identical.
        if (tid < SB * SB) {
            const int i = tid >> 4, c = tid & (SB - 1);
            const bool low = c <= i;
            double v = low ? AT(k0 + i, k0 + c) : 0.0;
            double my_rinv = 0.0;
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                if (i == j && c == j) s_piv = v;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (c == j && i >= j) {
                    const double piv = s_piv;
                    if (!(piv > 0.0) && i == j && s_info == 0) s_info = k0g + k0 + j + 1;   // also NaN
                    // 1/sqrt(piv) without libm calls: MUFU seed, one cubic step (< 2^-64), and sqrt(piv) = piv * rinv
                    // with one Heron correction
                    double rinv;
                    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rinv) : "d"(piv));
                    const double e = fma(-(piv * rinv), rinv, 1.0);
                    rinv = fma(rinv * e, fma(e, 0.375, 0.5), rinv);
                    double sq = piv * rinv;
                    sq = fma(fma(-sq, sq, piv), 0.5 * rinv, sq);
                    v = (i == j) ? sq : v * rinv;                        // column j of L
                    if (i == j) my_rinv = rinv;
                    s_col[i] = v;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (c > j && low) v = fma(-s_col[i], s_col[c], v);      // A[i][c] -= L[i][j] L[c][j]
            }
            if (low) AT(k0 + i, k0 + c) = v;
            if (i == c) rd[i] = my_rinv;                                 // reciprocal diagonal 1 / L_ii
        }
        __syncthreads();
        // (2) warp 0 inverts the sub-block (lane j = column j of the inverse, forward substitution in registers);
        //     warps 1..4 solve the panel rows below it:  x L_kk^T = a, one row per thread
        if (warp == 0) {
            if (lane < SB) {
                double x[SB];
#pragma unroll
                for (int i = 0; i < SB; ++i) {
                    double acc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
                    for (int k = 0; k < i; ++k) acc = fma(-AT(k0 + i, k0 + k), x[k], acc);
                    x[i] = (i < lane) ? 0.0 : acc * rd[i];
                    Dinv[(kb * SB + i) * DINV_LD + lane] = x[i];
                }
            }
        } else {
            const int i = k0 + SB + (tid - 32);
            if (tid - 32 < NB && i < NB) {
                double x[SB];
#pragma unroll
                for (int j = 0; j < SB; ++j) {
                    double acc = AT(i, k0 + j);
#pragma unroll
                    for (int k = 0; k < j; ++k) acc = fma(-x[k], AT(k0 + j, k0 + k), acc);
                    x[j] = acc * rd[j];
                }
#pragma unroll
                for (int j = 0; j < SB; ++j) AT(i, k0 + j) = x[j];
            }
        }
        __syncthreads();
        // (3) rank-16 update of the trailing lower triangle in 4 x 4 register tiles
        const int m = NB - k0 - SB;              // trailing order (multiple of 16)
        const int T = m / 4;                     // tiles per side
        for (int t = tid; t < T * (T + 1) / 2; t += DIAG_THREADS) {
            int tr = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
            while (tr * (tr + 1) / 2 > t) --tr;
            while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
            const int tc = t - tr * (tr + 1) / 2;
            const int i0 = k0 + SB + 4 * tr, c0 = k0 + SB + 4 * tc;
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 4
            for (int j = 0; j < SB; ++j) {
                double ai[4], ac[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    ai[a] = AT(i0 + a, k0 + j);
                    ac[a] = AT(c0 + a, k0 + j);
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fma(ai[a], ac[b], acc[a][b]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (c0 + b <= i0 + a) AT(i0 + a, c0 + b) -= acc[a][b];
        }
        __syncthreads();
    }
    if (tid == 0 && s_info != 0) atomicCAS(d.info, 0, s_info);
    // write L_kk
    for (int e = tid; e < NB * NB; e += DIAG_THREADS) {
        const int i = e / NB, j = e % NB;
        if (i < nb && j <= i) d.S[(int64_t)(k0g + i) * d.lds + k0g + j] = AT(i, j);
    }
    __syncthreads();
    // ---- X = L^-1, in place over the lower triangle: diagonal sub-blocks first, then block rows top to bottom:
    //      X_ij = -Dinv_i (L_i,j..i-1 X_j..i-1,j)
    for (int e = tid; e < NSB * SB * SB; e += DIAG_THREADS) {
        const int kb = e / (SB * SB), i = (e / SB) % SB, j = e % SB;
        if (j <= i) AT(kb * SB + i, kb * SB + j) = Dinv[(kb * SB + i) * DINV_LD + j];
    }
    __syncthreads();
    for (int bi = 1; bi < NSB; ++bi) {
        const int i0 = bi * SB, ncols = bi * SB;
        for (int e = tid; e < SB * ncols; e += DIAG_THREADS) {
            const int r = e % SB, c = e / SB;    // lanes vary r: conflict-free rows of L, broadcast X[k][c]
            double s0 = 0.0, s1 = 0.0;
            int k = c;
            for (; k + 1 < i0; k += 2) {
                s0 = fma(AT(i0 + r, k), AT(k, c), s0);
                s1 = fma(AT(i0 + r, k + 1), AT(k + 1, c), s1);
            }
            if (k < i0) s0 = fma(AT(i0 + r, k), AT(k, c), s0);
            G[r * DIAG_G_LD + c] = s0 + s1;
        }
        __syncthreads();
        for (int e = tid; e < SB * ncols; e += DIAG_THREADS) {
            const int r = e % SB, c = e / SB;
            double s = 0.0;
            for (int q = 0; q <= r; ++q) s = fma(Dinv[(bi * SB + r) * DINV_LD + q], G[q * DIAG_G_LD + c], s);
            AT(i0 + r, c) = -s;
        }
        __syncthreads();
    }
    double* X = d.Linv + (int64_t)kblk * NB * NB;
    for (int e = tid; e < NB * NB; e += DIAG_THREADS) {
        const int i = e / NB, j = e % NB;
        X[e] = (j <= i) ? AT(i, j) : 0.0;
    }
}

// ---- panel: L21 = S21 L11^-T  (in place; a CTA owns 128 full rows of the panel) ---------------------------
// grid = (ceil((n_max - k0 - NB) / BM), 1, batch)
// (tr_mod, tr_rem): only tile rows with (global tile row) % tr_mod == tr_rem are processed - the 1-D block-cyclic
// ownership of the multi-GPU factorisation (tr_mod = 1: every row).
// `pack` (optional, batch 1): the tiles this rank computes are ALSO written, tile after tile (128 x 128 doubles each,
// leading dimension 128), into the send buffer of the panel exchange - the j-th tile row this rank owns below the
// diagonal goes to pack + j * 128 * 128 (see `panel_unpack_kernel` for the receiving side).
__global__ void __launch_bounds__(GEMM_THREADS) potrf_panel_kernel(const GmgSolveDesc* __restrict__ descs, int kblk,
                                                                   int tr_mod, int tr_rem, double* pack) {
    if ((kblk + 1 + (int)blockIdx.x) % tr_mod != tr_rem) return;
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
    if (pack) {
        const int x0 = ((tr_rem - (kblk + 1)) % tr_mod + tr_mod) % tr_mod;     // first tile row below the diagonal it owns
        const int j = ((int)blockIdx.x - x0) / tr_mod;
        g.D2 = pack + (int64_t)j * NB * NB - (int64_t)m0 * NB;                  // gemm_tile indexes rows from the panel top
        g.ldd2 = NB;
    }
    gemm_tile<false, true>(g, m0, 0, sm);
}

// Receiving side of the panel exchange (all_gather of the packed tiles): recv = [rank q][j-th owned tile][128 x 128];
// every tile row below the diagonal that ANOTHER rank owns is copied into the panel column of this rank's S.
// grid = (tile rows below the diagonal, 1, 1), 256 threads.
__global__ void __launch_bounds__(256) panel_unpack_kernel(const GmgSolveDesc* __restrict__ descs, int kblk,
                                                           const double* __restrict__ recv, int world, int rank,
                                                           int64_t cap) {
    const GmgSolveDesc d = descs[0];
    const int x = blockIdx.x;
    const int t = kblk + 1 + x;
    const int q = t % world;
    if (q == rank) return;
    const int x0 = ((q - (kblk + 1)) % world + world) % world;
    const int j = (x - x0) / world;
    const double* src = recv + ((int64_t)q * cap + j) * NB * NB;
    const int rows = min(NB, d.n - t * NB);
    double* dst = d.S + (int64_t)t * NB * d.lds + (int64_t)kblk * NB;
    for (int e = threadIdx.x; e < rows * (NB / 2); e += 256) {
        const int i = e / (NB / 2), c2 = e % (NB / 2);
        *reinterpret_cast<double2*>(dst + (int64_t)i * d.lds + 2 * c2) = *reinterpret_cast<const double2*>(src + i * NB + 2 * c2);
    }
}

// Diagonal exchange: buf = [L_kk as a full 128 x 128 block (zeros above the diagonal / outside a partial block) |
// Linv_k].  unpack = 0: the owner fills buf from its S and Linv; unpack = 1: the others write buf into theirs.
// grid = (64, 1, 1), 256 threads.
__global__ void __launch_bounds__(256) diag_xchg_kernel(const GmgSolveDesc* __restrict__ descs, int kblk,
                                                        double* __restrict__ buf, int unpack) {
    const GmgSolveDesc d = descs[0];
    const int k0 = kblk * NB;
    const int nb = min(NB, d.n - k0);
    double* X = d.Linv + (int64_t)kblk * NB * NB;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < NB * NB; e += gridDim.x * 256) {
        const int i = e / NB, j = e % NB;
        const bool inside = i < nb && j <= i;
        if (!unpack) {
            buf[e] = inside ? d.S[(int64_t)(k0 + i) * d.lds + k0 + j] : 0.0;
            buf[NB * NB + e] = X[e];
        } else {
            if (inside) d.S[(int64_t)(k0 + i) * d.lds + k0 + j] = buf[e];
            X[e] = buf[NB * NB + e];
        }
    }
}

// ---- trailing update: S22 -= L21 L21^T, lower tiles only ---------------------------------------------------
// grid = (tiles, tiles, batch) with tiles = ceil((n_max - k0 - NB) / BM); tiles above the diagonal exit.
// x_off: first column tile of this launch (the look-ahead split of the overlapped driver: column tile 0 - the next
// panel - is updated on the critical-path stream, column tiles 1.. on another one).
// Tiles are 128 x 64 on the two-CTAs-per-SM configuration (K is only 128 here: with one 512-thread CTA per SM a third of
// a CTA's life was prologue and epilogue with the tensor pipe idle; batched months factored at 10.7 TFLOP/s).
// x_off and gridDim.x count 64-column tiles.
constexpr int UBN = 64;
__global__ void __launch_bounds__(CfgSmall::THREADS, 2) potrf_update_kernel(const GmgSolveDesc* __restrict__ descs,
                                                                            int kblk, int tr_mod, int tr_rem, int x_off) {
    const int cx = (int)blockIdx.x + x_off;
    const int m0 = blockIdx.y * BM, n0 = cx * UBN;
    if (n0 >= m0 + BM) return;                   // cx = column tile, y = row tile: keep the lower triangle
    if ((kblk + 1 + (int)blockIdx.y) % tr_mod != tr_rem) return;
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    const int m = d.n - k0 - NB;
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
    gemm_tile_cfg<CfgSmall, false, true, UBN>(g, m0, n0, sm);
}


// ---- left-looking column update: S[k:, k] -= L[k:, 0:k] L[k, 0:k]^T  (batched small systems) -----------------------
// The right-looking trailing update touches every tile of the trailing matrix once per block step with K = 128: for
// hundreds of month-sized systems (4.6 GB of S in total, far beyond L2) that is a DRAM-bound read-modify-write - ncu
// r02: tensor pipe 56 % active, 5.9 long-scoreboard stalls per issue, 10.7-11.6 TFLOP/s for the whole factorisation.
// Left-looking, block column k receives ALL its updates in one GEMM with K = 128 k right before it is factored: every
// tile of S is read and written once per factorisation instead of once per step.  Same flops, different summation
// order (one long accumulation instead of k rounded subtractions), so the two variants agree to rounding, not bit
// for bit; which one runs depends only on n_max (see `use_left_looking`), never on the batch size.
// grid = (2, tile rows of rows k.., batch)
__global__ void __launch_bounds__(CfgSmall::THREADS, 2) potrf_colupdate_kernel(const GmgSolveDesc* __restrict__ descs,
                                                                               int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    const int m = d.n - k0;                       // rows of block column k, diagonal block included
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * UBN;
    if (k0 == 0 || m <= 0 || m0 >= m || n0 >= min(NB, m)) return;
    extern __shared__ __align__(16) double sm[];
    TileArgs g;
    g.A = d.S + (int64_t)k0 * d.lds;              // L[k:, 0:k0]: m x k0
    g.lda = d.lds;
    g.B = g.A;                                    // op(B) = L[k, 0:k0]^T: stored n x k, the first 128 rows of A
    g.ldb = d.lds;
    g.C = d.S + (int64_t)k0 * d.lds + k0;
    g.ldc = d.lds;
    g.D = const_cast<double*>(g.C); g.ldd = d.lds;
    g.m = m; g.n = min(NB, m); g.k = k0;
    g.alpha = -1.0; g.beta = 1.0;
    gemm_tile_cfg<CfgSmall, false, true, UBN>(g, m0, n0, sm);
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

// grid = (ceil(k0 / BNT) + 1, 1, batch): strip[:, 0:k0] = -Linv_kk L[k, 0:k0]; the LAST CTA copies the Linv_kk part.
// Cfg / BNT: CfgBig / 128, or CfgSmall / 64 (two CTAs per SM) for short strips.
template <typename Cfg, int BNT>
__global__ void __launch_bounds__(Cfg::THREADS, (Cfg::THREADS == 256 ? 2 : 1)) trsm_prep_fwd_kernel(
    const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    if (k0 >= d.n) return;
    const int64_t ldp = d.ldp;                   // leading dimension of the coefficient strip (>= n, even)
    const double* X = d.Linv + (int64_t)kblk * NB * NB;
    const int nb = min(NB, d.n - k0);
    const int gemm_ctas = (k0 + BNT - 1) / BNT;
    if ((int)blockIdx.x >= gemm_ctas) {
        if ((int)blockIdx.x == gemm_ctas) {      // copy Linv_kk into columns k0..k0+NB of the strip
            for (int e = threadIdx.x; e < NB * NB; e += Cfg::THREADS) {
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
    gemm_tile_cfg<Cfg, false, false, BNT>(g, 0, blockIdx.x * BNT, sm);
}

// grid = (ceil(nrhs_max / BNT), 1, batch)
// GATHER: the right-hand-side rows of block row k (the last nb k-rows of the GEMM) are read from the covariance matrix,
// rows idx[k0 ..], for the columns < M - the kriging path's T = [C[idx, :] | 1 | y] is then never written by a gather
// kernel, only its two extra columns are (gmg_krige_solve_f64).  Same values, same arithmetic: bit-identical.
struct RhsGather {
    const double* C;               // M x M covariance, leading dimension ldc (even, 16-byte aligned base)
    int64_t M, ldc;
    const GmgKrigeDesc* kd;        // same systems, same order as the solve descriptors
};

template <typename Cfg, int BNT, bool GATHER = false>
__global__ void __launch_bounds__(Cfg::THREADS, (Cfg::THREADS == 256 ? 2 : 1)) trsm_step_fwd_kernel(
    const GmgSolveDesc* __restrict__ descs, int kblk, RhsGather rg) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    const int n0 = blockIdx.x * BNT;
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
    if constexpr (GATHER) {
        __shared__ int64_t rowoff[NB];
        const int64_t* idx = rg.kd[blockIdx.z].idx;
        for (int r = threadIdx.x; r < nb; r += Cfg::THREADS)
            rowoff[r] = min(max(idx[k0 + r], (int64_t)0), rg.M - 1) * rg.ldc;   // clamped like the gather kernel
        __syncthreads();
        GatherArgs ga;
        ga.G = rg.C; ga.rowoff = rowoff; ga.kg0 = k0; ga.ng = (int)rg.M;
        gemm_tile_cfg<Cfg, false, false, BNT, true>(g, 0, n0, sm, &ga);
    } else {
        gemm_tile_cfg<Cfg, false, false, BNT>(g, 0, n0, sm);
    }
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

// grid = (ceil(nrhs_max / BNT), 1, batch)
template <int BNT>
__global__ void __launch_bounds__(GEMM_THREADS) trsm_step_bwd_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int k0 = kblk * NB;
    const int n0 = blockIdx.x * BNT;
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
    gemm_tile<true, false, BNT>(g, 0, n0, sm);
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

int ensure_attrs() {
    static bool done = false;    // per process; the attribute is per device function, set on the current device
    static int done_dev = -1;
    int dev = 0;
    GMG_CUDA_CHECK(cudaGetDevice(&dev));
    if (done && done_dev == dev) return GMG_OK;
    int rc;
    if ((rc = set_smem(potrf_diag_kernel, DIAG_SMEM))) return rc;
    if ((rc = set_smem(potrf_panel_kernel, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(potrf_update_kernel, GEMM_SMALL_SMEM_BYTES))) return rc;
    if ((rc = set_smem(potrf_colupdate_kernel, GEMM_SMALL_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_prep_fwd_kernel<CfgBig, 128>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_prep_fwd_kernel<CfgSmall, 64>, GEMM_SMALL_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_step_fwd_kernel<CfgBig, 128>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_step_fwd_kernel<CfgBig, 64>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_step_fwd_kernel<CfgSmall, 64>, GEMM_SMALL_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_step_fwd_kernel<CfgTall, 56>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_step_fwd_kernel<CfgBig, 128, true>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_step_fwd_kernel<CfgBig, 64, true>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_step_fwd_kernel<CfgSmall, 64, true>, GEMM_SMALL_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_prep_bwd_kernel, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_step_bwd_kernel<128>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(trsm_step_bwd_kernel<64>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(gemm_kernel<false, false>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(gemm_kernel<false, true>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(gemm_kernel<true, false>, GEMM_SMEM_BYTES))) return rc;
    if ((rc = set_smem(gemm_kernel<true, true>, GEMM_SMEM_BYTES))) return rc;
    done = true;
    done_dev = dev;
    return GMG_OK;
}


// ---- launch helpers shared by the sequential entry points and the overlapped driver ---------------------------------
enum { UPD_ALL = 0, UPD_NEXT = 1, UPD_REST = 2 };

// trailing update of block step k: all column tiles, only the next panel column (128 columns), or the rest
int launch_update(const GmgSolveDesc* descs, int k, int tr_mod, int tr_rem, unsigned tiles, int batch, int which,
                  cudaStream_t s) {
    constexpr unsigned per = BM / UBN;           // 64-column tiles per 128-column panel
    unsigned gx = tiles * per, x_off = 0;
    if (which == UPD_NEXT) gx = per;
    if (which == UPD_REST) { gx = (tiles - 1) * per; x_off = per; }
    if (gx == 0) return GMG_OK;
    potrf_update_kernel<<<dim3(gx, tiles, batch), CfgSmall::THREADS, GEMM_SMALL_SMEM_BYTES, s>>>(descs, k, tr_mod, tr_rem,
                                                                                                 (int)x_off);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

// Systems of up to this order are factored left-looking on one GPU (GMG_LEFT_MAX overrides; 0 = never).  The choice
// depends on n_max alone, so that a month solved on its own and the same month inside a batch take the same route.
bool use_left_looking(int64_t n_max, int world) {
    static int64_t v = -1;
    if (v < 0) {
        const char* e = getenv("GMG_LEFT_MAX");
        v = e ? atoll(e) : 4096;
    }
    return world == 1 && n_max <= v;
}

int launch_colupdate(const GmgSolveDesc* descs, int k, int64_t n_max, int batch, cudaStream_t s) {
    const int64_t m = n_max - (int64_t)k * NB;
    if (k == 0 || m <= 0) return GMG_OK;
    potrf_colupdate_kernel<<<dim3(BM / UBN, (unsigned)gmg_ceil_div(m, BM), batch), CfgSmall::THREADS,
                             GEMM_SMALL_SMEM_BYTES, s>>>(descs, k);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

// Column-strip width of the long-K solve steps (512-thread tile): 128, or 64 when that takes fewer "column-waves" - the
// CTAs of a step all run the same K, so a step costs ceil(strips / SMs) x width; the narrow tile re-reads the
// coefficient strip twice as often (~6 % slower per column).  One 1-degree rank of 8: 8 102 columns = 64 strips of 128
// (1 wave x 128) or 127 of 64 (1 wave x 64); the whole grid on one GPU: 507 x 128 (4 waves = 512) or 1 013 x 64
// (7 waves = 448).  Bit-identical either way.
// A third width, 56 columns on the 16 x 1 warp layout (CfgTall), is taken when it saves a further ~10 %: one rank of 8
// then runs 145 strips on 148 SMs instead of 127 - but ONLY for a solve that has the GPU to itself (`alone`).  Inside
// the overlapped factor + solve pipeline the 21 SMs the 64-wide strips leave idle are what the factorisation's
// latency-bound critical path (diagonal block, panel, look-ahead update: high-priority stream, but CTAs are not
// pre-empted) runs on: measured on 8 GPUs, 1 degree / 20 000 observations, 0.172 s with 64-wide strips and 0.214 s
// with 56-wide ones (profiles/r02_strip_n8_ab.txt).  GMG_STRIP forces a width (experiments).
int strip_width(int64_t nrhs_max, int batch, int sms, bool alone) {
    if (nrhs_max <= 0) return BN;
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("GMG_STRIP");
        forced = e ? atoi(e) : 0;
    }
    if (forced == 128 || forced == 64 || forced == 56) return forced;
    auto cost = [&](int w, double penalty) {
        return (double)gmg_ceil_div(gmg_ceil_div(nrhs_max, w) * (int64_t)batch, sms) * w * penalty;
    };
    const double c128 = cost(128, 1.0), c64 = cost(64, 1.06), c56 = cost(56, 1.14);
    if (alone && c56 < c64 && c56 < c128) return 56;
    return c64 < c128 ? 64 : 128;
}

// K up to which a forward-solve step runs on the two-CTAs-per-SM tile (bit-identical either way; GMG_SMALL_K overrides)
int small_k_limit() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GMG_SMALL_K");
        v = e ? atoi(e) : 2048;   // A/B on B200, 480 months: K limit 0 / 512 / 1024 / 2048 -> forward solve 71.8 / 69.9 / 66.6 / 64.3 ms
    }
    return v;
}

// forward-solve step of block row k: coefficient strip, then one GEMM (rg != nullptr: right-hand sides of the block row
// gathered from the covariance matrix on the fly)
int launch_fwd_step(const GmgSolveDesc* descs, int k, int batch, int64_t nrhs_max, int width, cudaStream_t s,
                    const RhsGather* rg = nullptr) {
    const int k0 = k * NB;
    const RhsGather none{nullptr, 0, 0, nullptr};
    if (k0 + NB <= small_k_limit()) {
        trsm_prep_fwd_kernel<CfgSmall, 64><<<dim3((unsigned)gmg_ceil_div(k0, 64) + 1, 1, batch), CfgSmall::THREADS,
                                             GEMM_SMALL_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        const dim3 grid((unsigned)gmg_ceil_div(nrhs_max, 64), 1, batch);
        if (rg) trsm_step_fwd_kernel<CfgSmall, 64, true><<<grid, CfgSmall::THREADS, GEMM_SMALL_SMEM_BYTES, s>>>(descs, k, *rg);
        else trsm_step_fwd_kernel<CfgSmall, 64><<<grid, CfgSmall::THREADS, GEMM_SMALL_SMEM_BYTES, s>>>(descs, k, none);
        GMG_LAUNCH_CHECK();
        return GMG_OK;
    }
    trsm_prep_fwd_kernel<CfgBig, 128><<<dim3((unsigned)k + 1, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
    GMG_LAUNCH_CHECK();
    if (width == 56 && !rg)
        trsm_step_fwd_kernel<CfgTall, 56><<<dim3((unsigned)gmg_ceil_div(nrhs_max, 56), 1, batch), GEMM_THREADS,
                                            GEMM_SMEM_BYTES, s>>>(descs, k, none);
    else if (width == 64 || width == 56) {
        const dim3 grid((unsigned)gmg_ceil_div(nrhs_max, 64), 1, batch);
        if (rg) trsm_step_fwd_kernel<CfgBig, 64, true><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k, *rg);
        else trsm_step_fwd_kernel<CfgBig, 64><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k, none);
    } else {
        const dim3 grid((unsigned)gmg_ceil_div(nrhs_max, BN), 1, batch);
        if (rg) trsm_step_fwd_kernel<CfgBig, 128, true><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k, *rg);
        else trsm_step_fwd_kernel<CfgBig, 128><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k, none);
    }
    GMG_LAUNCH_CHECK();
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
    const bool left = use_left_looking(n_max, 1);
    for (int k = 0; k < nblk; ++k) {
        if (left && (rc = launch_colupdate(descs, k, n_max, batch, s))) return rc;
        potrf_diag_kernel<<<dim3(1, 1, batch), DIAG_THREADS, DIAG_SMEM, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        const int64_t rem = n_max - (int64_t)(k + 1) * NB;
        if (rem <= 0) break;
        const unsigned tiles = (unsigned)gmg_ceil_div(rem, BM);
        potrf_panel_kernel<<<dim3(tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k, 1, 0, nullptr);
        GMG_LAUNCH_CHECK();
        if (!left && (rc = launch_update(descs, k, 1, 0, tiles, batch, UPD_ALL, s))) return rc;
    }
    return GMG_OK;
}

int gmg_potrf_step_f64(const GmgSolveDesc* descs, int batch, int64_t n_max, int kblk, int phase, int tr_mod, int tr_rem,
                       gmg_stream_t stream) {
    if (!descs) return -1;
    if (batch <= 0 || batch > 65535) return -2;
    if (n_max <= 0 || n_max > 0x7fffffff) return -3;
    if (kblk < 0 || (int64_t)kblk * NB >= n_max) return -4;
    if (phase < GMG_POTRF_DIAG || phase > GMG_POTRF_UPDATE_REST) return -5;
    if (tr_mod < 1 || tr_rem < 0 || tr_rem >= tr_mod) return -6;
    int rc = ensure_attrs();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t rem = n_max - (int64_t)(kblk + 1) * NB;
    const unsigned tiles = rem > 0 ? (unsigned)gmg_ceil_div(rem, BM) : 0;
    if (phase == GMG_POTRF_DIAG) {
        potrf_diag_kernel<<<dim3(1, 1, batch), DIAG_THREADS, DIAG_SMEM, s>>>(descs, kblk);
    } else if (tiles == 0) {
        return GMG_OK;
    } else if (phase == GMG_POTRF_PANEL) {
        potrf_panel_kernel<<<dim3(tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, kblk, tr_mod, tr_rem, nullptr);
    } else if (phase == GMG_POTRF_UPDATE) {
        return launch_update(descs, kblk, tr_mod, tr_rem, tiles, batch, UPD_ALL, s);
    } else if (phase == GMG_POTRF_UPDATE_NEXT) {         // the next panel column only
        return launch_update(descs, kblk, tr_mod, tr_rem, tiles, batch, UPD_NEXT, s);
    } else {                                             // GMG_POTRF_UPDATE_REST: everything right of it
        return launch_update(descs, kblk, tr_mod, tr_rem, tiles, batch, UPD_REST, s);
    }
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_potrs_batched_f64(const GmgSolveDesc* descs, int batch, int64_t n_max, int64_t nrhs_max, int which,
                          gmg_stream_t stream) {
    if (!descs) return -1;
    if (batch <= 0 || batch > 65535) return -2;
    if (n_max <= 0 || n_max > 0x7fffffff) return -3;
    if (nrhs_max <= 0 || nrhs_max > 0x7fffffff) return -4;
    if (which < 1 || which > 3) return -5;
    int rc = ensure_attrs();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int nblk = (int)gmg_ceil_div(n_max, NB);
    // Column strips of the right-hand sides: 128 wide, or 64 wide when the wide ones would leave more than a quarter
    // of the SMs without a CTA (a column-sharded solve on 8 GPUs: 64 strips of 128 for 148 SMs ran at 12.8 TFLOP/s).
    int dev = 0, sms = 0;
    GMG_CUDA_CHECK(cudaGetDevice(&dev));
    GMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int width = strip_width(nrhs_max, batch, sms, true);
    const bool narrow = width != BN;              // the backward steps know 128 and 64
    const unsigned rhs_tiles = (unsigned)gmg_ceil_div(nrhs_max, narrow ? 64 : BN);
    for (int k = 0; k < nblk && (which & GMG_SOLVE_FORWARD); ++k)               // L Y = B
        if ((rc = launch_fwd_step(descs, k, batch, nrhs_max, width, s))) return rc;
    for (int k = nblk - 1; k >= 0 && (which & GMG_SOLVE_BACKWARD); --k) {       // L^T X = Y
        const int64_t rem = n_max - (int64_t)(k + 1) * NB;
        const unsigned tiles = 1 + (rem > 0 ? (unsigned)gmg_ceil_div(rem, BM) : 0);
        trsm_prep_bwd_kernel<<<dim3(tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        if (narrow)
            trsm_step_bwd_kernel<64><<<dim3(rhs_tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        else
            trsm_step_bwd_kernel<128><<<dim3(rhs_tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
    }
    return GMG_OK;
}

}  // extern "C"

// =====================================================================================================================
// Overlapped factor + forward solve, one GPU or the ranks of one box (gmg_dist_* / gmg_potrf_solve_f64)
// =====================================================================================================================
// The right-looking factorisation has a latency-bound critical path (diagonal block -> panel -> update of the NEXT panel
// column) next to bulk work that nothing waits for immediately (the rest of the trailing update), and the forward
// triangular solve of block row k only needs row k of L and the inverted diagonal block k - both final as soon as
// step k's diagonal block is done.  So the driver below runs three streams:
//   F (high priority)  diagonal block, [exchange], panel, [exchange], update of column tile k+1 (look-ahead)
//   R                  update of column tiles k+2.. of step k               (waits for panel k; F's step k+1 look-ahead
//                                                                            waits for it: per-tile update order kept)
//   U (low priority)   forward-solve step k = prep + one GEMM               (waits for diagonal block k)
// Every tile receives the same updates in the same order as in gmg_potrf_batched_f64 + gmg_potrs_batched_f64, so the
// factor and the solved right-hand sides are bit-identical to the sequential path - on any number of ranks.
// Multi-GPU (one process per GPU, 1-D block-cyclic tile rows): the owner of block k broadcasts [L_kk | Linv_k]
// (256 KB), every rank forms its tile rows of the panel and packs them, ONE all_gather of the packed tiles gives
// everyone the whole panel.  The collectives are NCCL calls issued from this C loop on stream F (no Python, no
// per-step host synchronisation); NCCL is resolved at run time from the libnccl.so.2 the process already has loaded
// (PyTorch's), so the library has no link-time dependency on it.
namespace {

struct NcclUid { char internal[128]; };
struct NcclApi {
    int (*GetUniqueId)(NcclUid*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    bool ok = false;
};
constexpr int kNcclFloat64 = 8;      // ncclDouble / ncclFloat64 of nccl.h

NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy PyTorch brought in, if any
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(h, "ncclBroadcast"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllGather;
    });
    return api;
}

struct DistCtx {
    int world = 1, rank = 0, device = 0;
    void* comm = nullptr;
    cudaStream_t sF = nullptr, sR = nullptr, sU = nullptr;
    std::vector<cudaEvent_t> events;
    std::mutex mu;                       // one factorisation at a time per context
};

int ctx_streams(DistCtx* c) {
    GMG_CUDA_CHECK(cudaGetDevice(&c->device));
    int lo = 0, hi = 0;                  // lo = lowest priority (largest number), hi = highest
    GMG_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const int mid = (lo + hi) / 2;
    GMG_CUDA_CHECK(cudaStreamCreateWithPriority(&c->sF, cudaStreamNonBlocking, hi));
    GMG_CUDA_CHECK(cudaStreamCreateWithPriority(&c->sR, cudaStreamNonBlocking, mid));
    GMG_CUDA_CHECK(cudaStreamCreateWithPriority(&c->sU, cudaStreamNonBlocking, lo));
    return GMG_OK;
}

int ctx_event(DistCtx* c, size_t i, cudaEvent_t* out) {
    while (c->events.size() <= i) {
        cudaEvent_t e;
        GMG_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->events.push_back(e);
    }
    *out = c->events[i];
    return GMG_OK;
}

// process-wide context of the plain single-GPU calls, one per device
DistCtx* local_ctx() {
    static std::mutex mu;
    static DistCtx* ctxs[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!ctxs[dev]) {
        DistCtx* c = new DistCtx();
        if (ctx_streams(c) != GMG_OK) { delete c; return nullptr; }
        ctxs[dev] = c;
    }
    return ctxs[dev];
}

inline int64_t panel_cap(int64_t n, int world) {           // packed tiles per rank in the largest (first) panel
    const int64_t nblk = gmg_ceil_div(n, NB);
    return nblk > 1 ? gmg_ceil_div(nblk - 1, world) : 1;
}

#define GMG_NCCL_CHECK(expr)                                 \
    do {                                                     \
        if ((expr) != 0) return GMG_ERR_COMM;                \
    } while (0)

int factor_solve(DistCtx* c, const GmgSolveDesc* descs, int batch, int64_t n_max, int64_t nrhs_max, double* xbuf,
                 cudaStream_t caller, const RhsGather* rg = nullptr) {
    const int world = c->world, rank = c->rank;
    const int nblk = (int)gmg_ceil_div(n_max, NB);
    NcclApi& nc = nccl_api();
    double* dbuf = xbuf;                                   // [2][NB][NB]
    const int64_t cap0 = panel_cap(n_max, world);
    double* sendbuf = xbuf ? xbuf + 2 * NB * NB : nullptr; // [cap0][NB][NB]
    double* recvbuf = xbuf ? sendbuf + cap0 * NB * NB : nullptr;   // [world][cap][NB][NB], cap = this step's
    int sms = 0;
    GMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    const int width = strip_width(nrhs_max, batch, sms, false);

    size_t ne = 0;
    cudaEvent_t ev_fork, ev;
    int rc;
    if ((rc = ctx_event(c, ne++, &ev_fork))) return rc;
    GMG_CUDA_CHECK(cudaEventRecord(ev_fork, caller));
    GMG_CUDA_CHECK(cudaStreamWaitEvent(c->sF, ev_fork, 0));
    GMG_CUDA_CHECK(cudaStreamWaitEvent(c->sR, ev_fork, 0));
    GMG_CUDA_CHECK(cudaStreamWaitEvent(c->sU, ev_fork, 0));
    cudaEvent_t ev_rest_prev = nullptr;
    const bool left = use_left_looking(n_max, world);
    for (int k = 0; k < nblk; ++k) {
        const int owner = k % world;
        if (left && (rc = launch_colupdate(descs, k, n_max, batch, c->sF))) return rc;
        // ---- F: diagonal block (its column was brought up to date by the look-ahead update of step k - 1) ----------
        if (rank == owner) {
            potrf_diag_kernel<<<dim3(1, 1, batch), DIAG_THREADS, DIAG_SMEM, c->sF>>>(descs, k);
            GMG_LAUNCH_CHECK();
        }
        if (world > 1) {
            if (rank == owner) {
                diag_xchg_kernel<<<64, 256, 0, c->sF>>>(descs, k, dbuf, 0);
                GMG_LAUNCH_CHECK();
            }
            GMG_NCCL_CHECK(nc.Broadcast(dbuf, dbuf, (size_t)2 * NB * NB, kNcclFloat64, owner, c->comm, c->sF));
            if (rank != owner) {
                diag_xchg_kernel<<<64, 256, 0, c->sF>>>(descs, k, dbuf, 1);
                GMG_LAUNCH_CHECK();
            }
        }
        // ---- U: forward-solve step k ---------------------------------------------------------------------------------
        if (nrhs_max > 0) {
            if ((rc = ctx_event(c, ne++, &ev))) return rc;
            GMG_CUDA_CHECK(cudaEventRecord(ev, c->sF));
            GMG_CUDA_CHECK(cudaStreamWaitEvent(c->sU, ev, 0));
            if ((rc = launch_fwd_step(descs, k, batch, nrhs_max, width, c->sU, rg))) return rc;
        }
        const int64_t rem = n_max - (int64_t)(k + 1) * NB;
        if (rem <= 0) break;
        const unsigned tiles = (unsigned)gmg_ceil_div(rem, BM);
        // ---- F: panel (own tile rows), exchange ------------------------------------------------------------------------
        potrf_panel_kernel<<<dim3(tiles, 1, batch), GEMM_THREADS, GEMM_SMEM_BYTES, c->sF>>>(
            descs, k, world, rank, world > 1 ? sendbuf : nullptr);
        GMG_LAUNCH_CHECK();
        if (world > 1) {
            const int64_t cap = gmg_ceil_div((int64_t)tiles, world);
            GMG_NCCL_CHECK(nc.AllGather(sendbuf, recvbuf, (size_t)cap * NB * NB, kNcclFloat64, c->comm, c->sF));
            panel_unpack_kernel<<<tiles, 256, 0, c->sF>>>(descs, k, recvbuf, world, rank, cap);
            GMG_LAUNCH_CHECK();
        }
        if (left) continue;                     // left-looking: block column k + 1 is brought up to date at the top of its step
        // ---- R: trailing update of column tiles k + 2 .. ------------------------------------------------------------
        cudaEvent_t ev_rest = nullptr;
        if (tiles > 1) {
            if ((rc = ctx_event(c, ne++, &ev))) return rc;
            GMG_CUDA_CHECK(cudaEventRecord(ev, c->sF));                        // the whole panel is in S
            GMG_CUDA_CHECK(cudaStreamWaitEvent(c->sR, ev, 0));
            if ((rc = launch_update(descs, k, world, rank, tiles, batch, UPD_REST, c->sR))) return rc;
            if ((rc = ctx_event(c, ne++, &ev_rest))) return rc;
            GMG_CUDA_CHECK(cudaEventRecord(ev_rest, c->sR));
        }
        // ---- F: look-ahead - column tile k + 1, after every earlier update of it (rest of step k - 1) ------------------
        if (ev_rest_prev) GMG_CUDA_CHECK(cudaStreamWaitEvent(c->sF, ev_rest_prev, 0));
        if ((rc = launch_update(descs, k, world, rank, tiles, batch, UPD_NEXT, c->sF))) return rc;
        ev_rest_prev = ev_rest;
    }
    // join: the caller's stream continues when all three are done
    for (cudaStream_t s : {c->sF, c->sR, c->sU}) {
        if ((rc = ctx_event(c, ne++, &ev))) return rc;
        GMG_CUDA_CHECK(cudaEventRecord(ev, s));
        GMG_CUDA_CHECK(cudaStreamWaitEvent(caller, ev, 0));
    }
    return GMG_OK;
}

}  // namespace

extern "C" {

int gmg_dist_unique_id(void* id128) {
    if (!id128) return -1;
    NcclApi& nc = nccl_api();
    if (!nc.ok) return GMG_ERR_COMM;
    NcclUid uid;
    if (nc.GetUniqueId(&uid) != 0) return GMG_ERR_COMM;
    memcpy(id128, &uid, sizeof(uid));
    return GMG_OK;
}

int gmg_dist_init(const void* id128, int world, int rank, void** ctx) {
    if (world < 1 || world > 64) return -2;
    if (rank < 0 || rank >= world) return -3;
    if (!ctx) return -4;
    if (world > 1 && !id128) return -1;
    DistCtx* c = new DistCtx();
    c->world = world;
    c->rank = rank;
    int rc = ctx_streams(c);
    if (rc) { delete c; return rc; }
    if (world > 1) {
        NcclApi& nc = nccl_api();
        NcclUid uid;
        memcpy(&uid, id128, sizeof(uid));
        if (!nc.ok || nc.CommInitRank(&c->comm, world, uid, rank) != 0) { delete c; return GMG_ERR_COMM; }
    }
    *ctx = c;
    return GMG_OK;
}

int gmg_dist_destroy(void* ctx) {
    if (!ctx) return -1;
    DistCtx* c = static_cast<DistCtx*>(ctx);
    if (c->comm) nccl_api().CommDestroy(c->comm);
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    for (cudaStream_t s : {c->sF, c->sR, c->sU})
        if (s) cudaStreamDestroy(s);
    delete c;
    return GMG_OK;
}

int gmg_dist_ws_doubles(int64_t n, int world, int64_t* doubles) {
    if (n <= 0) return -1;
    if (world < 1) return -2;
    if (!doubles) return -3;
    *doubles = world > 1 ? (int64_t)NB * NB * (2 + panel_cap(n, world) * (1 + world)) : 0;
    return GMG_OK;
}

// The overlapped factor + forward-solve pipeline of gmg_potrf_solve_f64 on this GPU, with the right-hand-side rows of
// every block row read from C[idx, :] on the fly (used by gmg_krige_solve_f64, kriging.cu).
int gmg_internal_potrf_solve_gathered(const GmgSolveDesc* descs, int batch, int64_t n_max, int64_t nrhs_max,
                                      const double* C, int64_t M, int64_t ldc, const GmgKrigeDesc* kd, cudaStream_t stream) {
    DistCtx* c = local_ctx();
    if (!c) return GMG_ERR_CUDA;
    int rc = ensure_attrs();
    if (rc) return rc;
    const RhsGather rg{C, M, ldc, kd};
    std::lock_guard<std::mutex> lock(c->mu);
    return factor_solve(c, descs, batch, n_max, nrhs_max, nullptr, stream, &rg);
}

int gmg_potrf_solve_f64(void* ctx, const GmgSolveDesc* descs, int batch, int64_t n_max, int64_t nrhs_max,
                        double* xbuf, int64_t xbuf_doubles, gmg_stream_t stream) {
    if (!descs) return -2;
    if (batch <= 0 || batch > 65535) return -3;
    if (n_max <= 0 || n_max > 0x7fffffff) return -4;
    if (nrhs_max < 0 || nrhs_max > 0x7fffffff) return -5;
    DistCtx* c = ctx ? static_cast<DistCtx*>(ctx) : local_ctx();
    if (!c) return GMG_ERR_CUDA;
    if (c->world > 1) {
        if (batch != 1) return GMG_ERR_UNSUPPORTED;
        int64_t need = 0;
        gmg_dist_ws_doubles(n_max, c->world, &need);
        if (!xbuf || xbuf_doubles < need || (reinterpret_cast<uintptr_t>(xbuf) & 15)) return GMG_ERR_WORKSPACE;
    }
    int rc = ensure_attrs();
    if (rc) return rc;
    int dev = 0;
    GMG_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev != c->device) return GMG_ERR_UNSUPPORTED;     // a context belongs to the device it was created on
    std::lock_guard<std::mutex> lock(c->mu);
    return factor_solve(c, descs, batch, n_max, nrhs_max, xbuf, (cudaStream_t)stream);
}

}  // extern "C"
