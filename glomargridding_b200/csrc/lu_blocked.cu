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
#include <algorithm>
#include <cstdlib>

#include "gmg_common.cuh"
#include "dense_f64.cuh"
#include "../../include/glomar_b200.h"

namespace {

using namespace gmg;

constexpr int NB = 128;
constexpr int UBN = 64;


// ---- panel factorisation with partial pivoting: columns k0 .. k0 + nb, rows k0 .. n ------------------------------------
// The one part of LU with a column-by-column dependency.  grid = (G, 1, systems): the rows below the diagonal are split
// over G co-resident CTAs per system (cooperative launch), which meet at a per-system barrier once per column.  The
// panel is worked in sub-panels of SW = 16 columns:
//   * a CTA holds its rows of the sub-panel in shared memory, column-major (one thread per row, conflict free);
//   * per column: local pivot candidate -> every CTA publishes (|value|, row, the candidate ROW's 16 entries) and the
//     owner of row k publishes row k -> ONE barrier -> everyone picks the winner (largest magnitude, lowest row: the
//     unblocked algorithm's / LAPACK's choice), the owners of rows k and p exchange them from the published copies,
//     and every CTA scales and rank-1 updates its own rows.  Publications are double buffered by column parity, which
//     is what lets one barrier per column suffice;
//   * per sub-panel: rows are written back, CTA 0 applies the sixteen interchanges to the panel's other columns as ONE
//     gather (the net permutation of the <= 32 touched rows, all loads before all stores), a barrier, then every CTA
//     forms U12 = L11^-1 A12 for the remaining columns of the panel redundantly in shared memory and updates its own
//     rows with it (CTA 0 writes U12 back after the next barrier, when nobody reads the original any more).
// Global memory written by another CTA is always read with ld.cg (L1 is not coherent).
constexpr int SW = 16;
constexpr int COOP_THREADS = 512;
constexpr int COOP_WARPS = COOP_THREADS / 32;
constexpr int COOP_MAX_G = 296;
constexpr int SLOT = 2 + SW;                                  // |value|, row, the candidate row
constexpr int COOP_DOUBLES = 16 + 2 * COOP_MAX_G * SLOT + 2 * SW;   // barrier counter, slots x 2, row k x 2
constexpr int COOP_MAX_ROWS = 1536;                           // rows of one CTA (shared memory)
constexpr int NOROW = 0x7fffffff;

__device__ __forceinline__ double* coop_of(const GmgSolveDesc& d) { return d.P + (int64_t)NB * d.ldp; }
__device__ __forceinline__ int* piv_of(const GmgSolveDesc& d) {
    // pivots live behind the coefficient strip and the panel's exchange area in the P workspace (gmg_gesv_ws_doubles)
    return reinterpret_cast<int*>(d.P + (int64_t)NB * d.ldp + COOP_DOUBLES);
}
__host__ __device__ constexpr int coop_stride(int rows_cap) { return rows_cap | 1; }
static size_t coop_smem(int rows_cap) { return sizeof(double) * ((size_t)SW * coop_stride(rows_cap) + (size_t)SW * NB + SW * SW + SW) + 512; }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// all CTAs of one system; `target` = arrivals expected so far in this launch (the counter is zeroed between launches)
__device__ __forceinline__ void system_barrier(unsigned* cnt, unsigned target, int G) {
    __syncthreads();
    if (G > 1 && threadIdx.x == 0) {
        // release: the CTA's writes before the barrier above are ordered before the arrival; acquire on the way out
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
        while (ld_acquire_u32(cnt) < target) {}
    }
    __syncthreads();
}

__global__ void __launch_bounds__(COOP_THREADS, 2)
lu_panel_kernel(const GmgSolveDesc* __restrict__ descs, int kblk, int G, int rows_cap, int rows_target) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int n = d.n, k0 = kblk * NB;
    if (k0 >= n) return;
    const int nb = min(NB, n - k0), m = n - k0;
    // CTAs this system really uses: one per rows_target rows, and enough for its rows to fit the shared memory
    const int Gs = min(G, max(max(1, (m + rows_target - 1) / rows_target), (m + rows_cap - 1) / rows_cap));
    const int g = blockIdx.x;
    if (g >= Gs) return;
    const int RL = (m + Gs - 1) / Gs;
    const int r0 = k0 + g * RL, r1 = min(n, r0 + RL);
    const int nloc = max(0, r1 - r0);
    const int RLP = coop_stride(rows_cap);
    double* A = d.S;
    const int64_t ld = d.lds;
    int* piv = piv_of(d);
    double* ws = coop_of(d);
    unsigned* cnt = reinterpret_cast<unsigned*>(ws);
    double* slots = ws + 16;
    double* rowk = slots + 2 * COOP_MAX_G * SLOT;

    extern __shared__ __align__(16) double sm[];
    double* s_P = sm;                              // [SW][RLP]  my rows of the sub-panel, column-major
    double* s_U = s_P + (size_t)SW * RLP;          // [SW][NB]   U12 of the sub-panel
    double* s_L11 = s_U + SW * NB;                 // [SW][SW]
    double* s_prow = s_L11 + SW * SW;              // [SW]       the pivot row
    __shared__ double s_wval[COOP_WARPS];
    __shared__ int s_widx[COOP_WARPS];
    __shared__ int s_piv[SW], s_src[2 * SW], s_dst[2 * SW];
    __shared__ int s_p;
    __shared__ double s_pinv;
    __shared__ double s_rowk[SW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned bar = 0;
    int pend_row = -1, pend_col = 0, pend_rest = 0;                 // U12 waiting to be written back (CTA 0)

    for (int j0 = 0; j0 < nb; j0 += SW) {
        const int w = min(SW, nb - j0), kk0 = k0 + j0;
        for (int e = tid; e < nloc * SW; e += COOP_THREADS) {
            const int i = e / SW, c = e % SW;
            s_P[c * RLP + i] = c < w ? __ldcg(A + (int64_t)(r0 + i) * ld + kk0 + c) : 0.0;
        }
        __syncthreads();
        for (int j = 0; j < w; ++j) {
            const int k = kk0 + j, par = j & 1;
            // candidate among my rows: first row of maximal magnitude
            double best = -1.0;
            int bi = NOROW;
            for (int i = max(k, r0) - r0 + tid; i < nloc; i += COOP_THREADS) {
                const double v = fabs(s_P[j * RLP + i]);
                if (v > best) { best = v; bi = r0 + i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) { s_wval[warp] = best; s_widx[warp] = bi; }
            __syncthreads();
            if (warp == 0) {
                best = lane < COOP_WARPS ? s_wval[lane] : -1.0;
                bi = lane < COOP_WARPS ? s_widx[lane] : NOROW;
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                double* slot = slots + (size_t)(par * COOP_MAX_G + g) * SLOT;
                if (lane == 0) { slot[0] = best; slot[1] = (double)bi; }
                if (lane < SW && bi != NOROW) slot[2 + lane] = s_P[lane * RLP + (bi - r0)];
                if (k >= r0 && k < r1 && lane < SW) rowk[par * SW + lane] = s_P[lane * RLP + (k - r0)];
            }
            system_barrier(cnt, (unsigned)Gs * (++bar), Gs);
            if (pend_row >= 0) {                                        // (CTA 0) the previous sub-panel's U12
                for (int e = tid; e < SW * pend_rest; e += COOP_THREADS) {
                    const int t = e / pend_rest, c = e % pend_rest;
                    A[(int64_t)(pend_row + t) * ld + pend_col + c] = s_U[t * NB + c];
                }
                pend_row = -1;
            }
            if (warp == 1 && lane < SW) s_rowk[lane] = __ldcg(rowk + par * SW + lane);
            if (warp == 0) {
                double bv = -1.0;
                int bp = NOROW, bg = 0;
                for (int q = lane; q < Gs; q += 32) {
                    const double* sl = slots + (size_t)(par * COOP_MAX_G + q) * SLOT;
                    const double v = __ldcg(sl);
                    const int ix = (int)__ldcg(sl + 1);
                    if (v > bv || (v == bv && ix < bp)) { bv = v; bp = ix; bg = q; }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bp, o);
                    const int og = __shfl_xor_sync(0xffffffffu, bg, o);
                    if (ov > bv || (ov == bv && oi < bp)) { bv = ov; bp = oi; bg = og; }
                }
                // (a column of NaNs has no candidate: keep row k, and report it)
                const double* src = bp == NOROW ? rowk + par * SW : slots + (size_t)(par * COOP_MAX_G + bg) * SLOT + 2;
                if (bp == NOROW) bp = k;
                double pr = 0.0;
                if (lane < SW) { pr = __ldcg(src + lane); s_prow[lane] = pr; }
                const double pv = __shfl_sync(0xffffffffu, pr, j);
                if (lane == 0) {
                    s_p = bp;
                    s_piv[j] = bp;
                    s_pinv = 1.0 / pv;
                    if (g == 0) {
                        piv[k] = bp;
                        if (pv == 0.0 || pv != pv) atomicCAS(d.info, 0, k + 1);
                    }
                }
            }
            __syncthreads();
            const int p = s_p;
            if (p != k) {                                                // rows k and p trade places (published copies)
                if (p >= r0 && p < r1 && tid < SW) s_P[tid * RLP + (p - r0)] = s_rowk[tid];
                if (k >= r0 && k < r1 && tid >= 32 && tid < 32 + SW) s_P[(tid - 32) * RLP + (k - r0)] = s_prow[tid - 32];
            }
            __syncthreads();
            const double pinv = s_pinv;
            for (int i = max(k + 1, r0) - r0 + tid; i < nloc; i += COOP_THREADS) {
                const double l = s_P[j * RLP + i] * pinv;
                s_P[j * RLP + i] = l;
                for (int c = j + 1; c < w; ++c) s_P[c * RLP + i] = fma(-l, s_prow[c], s_P[c * RLP + i]);
            }
        }
        __syncthreads();
        // my rows of the finished sub-panel go back (rows above kk0 are earlier sub-panels' U rows: not mine to touch)
        for (int e = tid; e < nloc * SW; e += COOP_THREADS) {
            const int i = e / SW, c = e % SW;
            if (r0 + i >= kk0 && c < w) A[(int64_t)(r0 + i) * ld + kk0 + c] = s_P[c * RLP + i];
        }
        if (g == 0) {
            // the w interchanges on the panel's other columns: where does the final content of each touched row come from
            if (tid < 2 * w) {
                const int q = tid < w ? kk0 + tid : s_piv[tid - w];
                int cur = q;
                for (int t = w - 1; t >= 0; --t) {
                    const int kt = kk0 + t, pt = s_piv[t];
                    if (cur == kt) cur = pt;
                    else if (cur == pt) cur = kt;
                }
                s_dst[tid] = q;
                s_src[tid] = cur;
            }
            __syncthreads();
            const int nother = nb - w, total = 2 * w * nother;
            constexpr int PER = (2 * SW * NB + COOP_THREADS - 1) / COOP_THREADS;
            double v[PER];
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int e = tid + u * COOP_THREADS;
                if (e < total) {
                    const int qi = e / nother, cc = e % nother;
                    v[u] = __ldcg(A + (int64_t)s_src[qi] * ld + k0 + (cc < j0 ? cc : cc + w));
                }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int e = tid + u * COOP_THREADS;
                if (e < total) {
                    const int qi = e / nother, cc = e % nother;
                    if (s_src[qi] != s_dst[qi]) A[(int64_t)s_dst[qi] * ld + k0 + (cc < j0 ? cc : cc + w)] = v[u];
                }
            }
        }
        const int rest = nb - j0 - w;
        if (rest <= 0) break;
        system_barrier(cnt, (unsigned)Gs * (++bar), Gs);
        // U12 = L11^-1 A12 (here w == SW), by everyone
        for (int e = tid; e < SW * SW; e += COOP_THREADS) s_L11[e] = __ldcg(A + (int64_t)(kk0 + e / SW) * ld + kk0 + e % SW);
        for (int e = tid; e < SW * rest; e += COOP_THREADS) {
            const int t = e / rest, c = e % rest;
            s_U[t * NB + c] = __ldcg(A + (int64_t)(kk0 + t) * ld + kk0 + SW + c);
        }
        __syncthreads();
        if (tid < rest) {
            double x[SW];
#pragma unroll
            for (int t = 0; t < SW; ++t) {
                double acc = s_U[t * NB + tid];
#pragma unroll
                for (int q = 0; q < t; ++q) acc = fma(-s_L11[t * SW + q], x[q], acc);
                x[t] = acc;
            }
#pragma unroll
            for (int t = 0; t < SW; ++t) s_U[t * NB + tid] = x[t];
        }
        __syncthreads();
        if (g == 0) { pend_row = kk0; pend_col = kk0 + SW; pend_rest = rest; }
        // my rows below the sub-panel: A[i][c] -= sum_t L[i][t] U12[t][c]; a warp per row, lanes over the columns
        for (int i = max(kk0 + SW, r0) - r0 + warp; i < nloc; i += COOP_WARPS) {
            const double lv = lane < SW ? s_P[lane * RLP + i] : 0.0;
            double l[SW];
#pragma unroll
            for (int t = 0; t < SW; ++t) l[t] = __shfl_sync(0xffffffffu, lv, t);
            double* row = A + (int64_t)(r0 + i) * ld + kk0 + SW;
            double a[NB / 32];
#pragma unroll
            for (int u = 0; u < NB / 32; ++u) a[u] = lane + 32 * u < rest ? __ldcg(row + lane + 32 * u) : 0.0;
#pragma unroll
            for (int u = 0; u < NB / 32; ++u) {
                const int c = lane + 32 * u;
                if (c < rest) {
                    double acc = a[u];
#pragma unroll
                    for (int t = 0; t < SW; ++t) acc = fma(-l[t], s_U[t * NB + c], acc);
                    row[c] = acc;
                }
            }
        }
        __syncthreads();
    }
}

// zero the panel barrier's counter (before the first panel; lu_laswp_kernel re-zeroes it after every panel)
__global__ void lu_init_kernel(const GmgSolveDesc* __restrict__ descs, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) *reinterpret_cast<unsigned*>(coop_of(descs[b])) = 0u;
}

// ---- the panel's row interchanges on the columns outside the panel.  grid = (ceil(n_max / 256), 1, batch) -------------
__global__ void __launch_bounds__(256) lu_laswp_kernel(const GmgSolveDesc* __restrict__ descs, int kblk) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int n = d.n, k0 = kblk * NB;
    if (k0 >= n) return;
    const int nb = min(NB, n - k0);
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c == 0) *reinterpret_cast<unsigned*>(coop_of(d)) = 0u;           // the panel barrier's counter, for the next panel
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
// padded to 128 x 128.  grid = (2, 1, batch): CTA 0 inverts L11, CTA 1 U11, each by BLOCKED substitution over 16-row
// groups - the contribution of the rows already solved is a (16 x 16 I) x (16 I x 128) product spread over all 512
// threads (four outputs each), the 16 x 16 diagonal group a short per-column substitution - so the dependent chain is
// 8 x (<= 112 + 16 x 8) steps instead of 128 x 64.  The factored block sits in shared memory; the inverse is built in
// its global destination and read back through L1 (same CTA, after a barrier).
constexpr int IG = 16;                            // rows per group
__global__ void __launch_bounds__(512) lu_blocks_kernel(const GmgSolveDesc* __restrict__ descs, int kblk, int nblk_max) {
    const GmgSolveDesc d = descs[blockIdx.z];
    const int n = d.n, k0 = kblk * NB;
    if (k0 >= n) return;
    const int nb = min(NB, n - k0);
    extern __shared__ __align__(16) double sm[];
    double* B = sm;                               // [NB][NB + 1] the factored block
    constexpr int LD = NB + 1;
    double* T = B + NB * LD;                      // [IG][NB] right-hand side of the current group
    for (int e = threadIdx.x; e < NB * NB; e += 512) {
        const int i = e / NB, j = e % NB;
        B[i * LD + j] = (i < nb && j < nb) ? d.S[(int64_t)(k0 + i) * d.lds + k0 + j] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    const bool lower = blockIdx.x == 0;
    double* X = d.Linv + (int64_t)((lower ? 0 : nblk_max) + kblk) * NB * NB;
    const int c = threadIdx.x & (NB - 1), rq = (threadIdx.x >> 7) * 4;       // my column, my four rows of the group
    for (int g = 0; g < NB / IG; ++g) {
        const int i0 = lower ? g * IG : NB - IG - g * IG;
        const int ka = lower ? 0 : i0 + IG, kb = lower ? i0 : NB;           // rows already solved
        double acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = (i0 + rq + u == c) ? 1.0 : 0.0;
        for (int k = ka; k < kb; ++k) {
            const double xv = X[k * NB + c];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fma(-B[(i0 + rq + u) * LD + k], xv, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) T[(rq + u) * NB + c] = acc[u];
        __syncthreads();
        if (threadIdx.x < NB) {
            double x[IG];
            if (lower) {
#pragma unroll
                for (int r = 0; r < IG; ++r) {
                    double a = T[r * NB + c];
#pragma unroll
                    for (int q = 0; q < r; ++q) a = fma(-B[(i0 + r) * LD + i0 + q], x[q], a);
                    x[r] = a;                                                  // unit diagonal: no division
                }
            } else {
#pragma unroll
                for (int r = IG - 1; r >= 0; --r) {
                    double a = T[r * NB + c];
#pragma unroll
                    for (int q = r + 1; q < IG; ++q) a = fma(-B[(i0 + r) * LD + i0 + q], x[q], a);
                    x[r] = a / B[(i0 + r) * LD + i0 + r];
                }
            }
#pragma unroll
            for (int r = 0; r < IG; ++r) {
                const int i = i0 + r;
                X[i * NB + c] = (lower ? i < c : i > c) ? 0.0 : x[r];
            }
        }
        __syncthreads();
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

constexpr size_t LU_BLOCKS_SMEM = ((size_t)NB * (NB + 1) + (size_t)IG * NB) * sizeof(double);

template <typename K>
int set_smem(K kern, size_t bytes) {
    GMG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return GMG_OK;
}

// The panel of the largest remaining order m over `batch` systems: as many co-resident CTAs per system as the device
// holds (two per SM while their rows fit in half the shared memory), at least enough for the rows to fit at all, and
// no more than one per `rows_target` rows; systems that do not fit together go in successive launches.
int panel_rows_target() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GMG_LU_ROWS");
        v = e ? atoi(e) : 128;
        if (v < 32) v = 32;
    }
    return v;
}

int launch_panel(const GmgSolveDesc* descs, int batch, int kblk, int m, int sms, cudaStream_t s) {
    const int target = panel_rows_target();
    const int g_need = (int)gmg_ceil_div(m, COOP_MAX_ROWS), g_want = (int)gmg_ceil_div(m, target);
    int G = 1, rows_cap = m, chunk = batch;
    size_t smem = 0;
    for (int occ_try = 2; occ_try >= 1; --occ_try) {
        const int cap = std::min(occ_try * sms, COOP_MAX_G);
        if (g_need > cap) return -3;
        chunk = std::min(batch, std::max(1, cap / std::max(g_need, 1)));
        G = std::max(g_need, std::min(cap / chunk, g_want));
        rows_cap = (int)gmg_ceil_div(m, G);
        smem = coop_smem(rows_cap);
        int occ = 0;
        GMG_CUDA_CHECK(cudaFuncSetAttribute(lu_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GMG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lu_panel_kernel, COOP_THREADS, smem));
        if (occ >= occ_try) break;
        if (occ_try == 1) return GMG_ERR_CUDA;
    }
    int rows_target = target;
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        const GmgSolveDesc* dd = descs + b0;
        const int nb = std::min(chunk, batch - b0);
        if (G == 1) {
            lu_panel_kernel<<<dim3(1, 1, nb), COOP_THREADS, smem, s>>>(dd, kblk, G, rows_cap, rows_target);
            GMG_LAUNCH_CHECK();
        } else {
            void* args[] = {(void*)&dd, (void*)&kblk, (void*)&G, (void*)&rows_cap, (void*)&rows_target};
            GMG_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)lu_panel_kernel, dim3(G, 1, nb), dim3(COOP_THREADS), args, smem, s));
        }
    }
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
    // strip + panel exchange + pivots, rounded to 16 bytes: workspaces carved one after the other out of one allocation
    // stay 16-byte aligned (the GEMM tiles read the strip with 16-byte cp.async)
    if (p_doubles) *p_doubles = ((int64_t)NB * ld + COOP_DOUBLES + (n + 1) / 2 + 2 + 1) / 2 * 2;
    if (ldp) *ldp = ld;
    return GMG_OK;
}

int64_t gmg_gesv_pivots_offset(int64_t ldp) { return (int64_t)NB * ldp + COOP_DOUBLES; }

int gmg_gesv_blocked_f64(const GmgSolveDesc* descs, int batch, int64_t n_max, int64_t nrhs_max, gmg_stream_t stream) {
    if (!descs) return -1;
    if (batch <= 0 || batch > 65535) return -2;
    if (n_max <= 0 || n_max > 0x7fffffff) return -3;
    if (nrhs_max < 0 || nrhs_max > 0x7fffffff) return -4;
    int rc = lu_attrs();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int nblk = (int)gmg_ceil_div(n_max, NB);
    int dev = 0, sms = 0;
    GMG_CUDA_CHECK(cudaGetDevice(&dev));
    GMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    lu_init_kernel<<<(unsigned)gmg_ceil_div(batch, 256), 256, 0, s>>>(descs, batch);
    GMG_LAUNCH_CHECK();
    for (int k = 0; k < nblk; ++k) {
        if ((rc = launch_panel(descs, batch, k, (int)(n_max - (int64_t)k * NB), sms, s))) return rc;
        lu_laswp_kernel<<<dim3((unsigned)gmg_ceil_div(n_max, 256), 1, batch), 256, 0, s>>>(descs, k);
        GMG_LAUNCH_CHECK();
        lu_blocks_kernel<<<dim3(2, 1, batch), 512, LU_BLOCKS_SMEM, s>>>(descs, k, nblk);
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
