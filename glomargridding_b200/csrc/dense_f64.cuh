// FP64 dense building blocks for the kriging path on sm_100a: a DMMA (mma.sync m8n8k4 f64) GEMM tile
// mainloop with a 3-stage cp.async pipeline, shared by the GEMM entry point, the blocked Cholesky and the
// blocked triangular solves (dense_f64.cu, kriging.cu).
//
// Why mma.sync and not tcgen05: tcgen05.mma has no f64 kind on sm_100a (ptxas rejects it); the FP64 tensor
// path is the DMMA.8x8x4 instruction that mma.sync.m8n8k4.f64 lowers to (measured 37.1 TFLOP/s on B200,
// profiles/r01_microbench_fp64.json, against 34.2 TFLOP/s for plain DFMA).
//
// All matrices are ROW-MAJOR with leading dimensions in elements.  Pointers and leading dimensions must be
// multiples of 2 doubles (16 B) because tiles are fetched with 16-byte cp.async.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace gmg {

constexpr int BM = 128;            // CTA tile rows
constexpr int BN = 128;            // CTA tile columns
#ifndef GMG_BK
#define GMG_BK 32
#endif
constexpr int BK = GMG_BK;         // k-slab per pipeline stage (8 DMMA k-steps; 31.5 vs 30.6 TFLOP/s for BK = 16)
#ifndef GMG_STAGES
#define GMG_STAGES 3
#endif
constexpr int STAGES = GMG_STAGES;
constexpr int GEMM_THREADS = 512;  // 16 warps as 4 (m) x 4 (n), warp tile 32 x 32 = 4 x 4 DMMA tiles
constexpr int LDK = BK + 4;        // [row][k] operand tile: 36 doubles (72 words = 8 mod 32) -> the 8 x 4 fragment is conflict-free
constexpr int LDN = BN + 4;        // [k][col] operand tile: 132 doubles -> the 4 x 8 fragment is conflict-free
constexpr int TILE_DOUBLES = (BM * LDK > BK * LDN) ? BM * LDK : BK * LDN;   // 4608
constexpr size_t GEMM_SMEM_BYTES = (size_t)STAGES * 2 * TILE_DOUBLES * sizeof(double);   // 221 184

struct TileArgs {
    const double* A;   // op(A) is m x k:  TA ? stored k x m : stored m x k
    const double* B;   // op(B) is k x n:  TB ? stored n x k : stored k x n
    const double* C;   // m x n, read only when beta != 0 (may alias D)
    double* D;         // m x n
    int m, n, k;
    int64_t lda, ldb, ldc, ldd;
    double alpha, beta;
    double* D2 = nullptr;   // optional second copy of the result tile (D2[row * ldd2 + col], same row / col indices as D):
    int64_t ldd2 = 0;       // the panel kernel of the distributed factorisation packs its tiles for the exchange with it
};

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Tile configurations.  BIG: 512 threads = 16 warps as 4 (m) x 4 (n), k-slab 32, 3 stages, 221 KB of shared memory -
// one CTA per SM, the shape for long-K GEMMs (31.5 TFLOP/s at 4096^3).  SMALL: 256 threads = 8 warps as 4 x 2 on a
// 128 x 64 tile, k-slab 16, 3 stages, 92 KB - TWO CTAs per SM, so that one CTA's prologue (first slabs in flight, no
// math yet) and epilogue (tile read-modify-write) overlap the other's main loop: the shape for the short-K GEMMs of
// batched small systems (K = 128 trailing updates, the first triangular-solve steps), where a lone CTA per SM spends a
// third of its life outside the main loop.  Every output element is accumulated over k in the same order in both, so
// results are bit-identical.
struct CfgBig {
    static constexpr int THREADS = GEMM_THREADS, WARPS_N = 4, KB = BK, NST = STAGES;
};
struct CfgSmall {
    static constexpr int THREADS = 256, WARPS_N = 2, KB = 16, NST = 3;
};
// TALL: the 512 threads as 16 (m) x 1 (n) warps, each warp one 8-row DMMA tile deep and the whole strip wide - the only
// layout that admits strip widths that are not multiples of 32 (56 = 7 x 8).  A column-sharded solve whose strips do
// not fill the SMs picks its width from {128, 64, 56}: one 1-degree rank of 8 has 8 102 columns = 127 strips of 64 on
// 148 SMs, or 145 strips of 56.  Costs more shared-memory traffic per DMMA (8 fragment loads per 7) than the 4 x 4 layout.
struct CfgTall {
    static constexpr int THREADS = GEMM_THREADS, WARPS_N = 1, KB = BK, NST = STAGES;
};
template <typename Cfg, int BNT>
struct TileGeom {
    static constexpr int LDK_ = Cfg::KB + 4;                 // [row][k] tile: KB + 4 doubles per row (8 mod 32 words)
    static constexpr int LDN_A = BM + 4;                     // [k][col] tile of A^T
    static constexpr int LDN_B = BNT + 4;                    // [k][col] tile of B
    static constexpr int A_DOUBLES = (BM * LDK_ > Cfg::KB * LDN_A) ? BM * LDK_ : Cfg::KB * LDN_A;
    static constexpr int B_DOUBLES = (Cfg::KB * LDN_B > BNT * LDK_) ? Cfg::KB * LDN_B : BNT * LDK_;
    static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
    static constexpr size_t SMEM_BYTES = (size_t)Cfg::NST * STAGE_DOUBLES * sizeof(double);
};
constexpr size_t GEMM_SMALL_SMEM_BYTES = TileGeom<CfgSmall, 64>::SMEM_BYTES;      // 92 160

// Fetch one [rows][KB] k-slab of a matrix stored rows x k (row-major) into a [row][k] tile (ld = KB + 4); NR = rows.
template <typename Cfg, int NR>
__device__ __forceinline__ void load_rowk_c(double* s, const double* G, int64_t ld, int row0, int rows_total, int k0,
                                            int k_total, int tid) {
    constexpr int KB = Cfg::KB, LD = KB + 4;
#pragma unroll
    for (int c = tid; c < NR * (KB / 2); c += Cfg::THREADS) {
        const int r = c / (KB / 2), c2 = c % (KB / 2);
        const int gr = row0 + r, gk = k0 + 2 * c2;
        int bytes = (gr < rows_total) ? (k_total - gk) * 8 : 0;
        bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const double* src = bytes ? G + (int64_t)gr * ld + gk : G;
        cp_async16_zfill(s + r * LD + 2 * c2, src, bytes);
    }
}

// Fetch one [KB][cols] k-slab of a matrix stored k x cols (row-major) into a [k][col] tile (ld = NC + 4); NC = columns.
template <typename Cfg, int NC>
__device__ __forceinline__ void load_kcol_c(double* s, const double* G, int64_t ld, int col0, int cols_total, int k0,
                                            int k_total, int tid) {
    constexpr int KB = Cfg::KB, LD = NC + 4;
#pragma unroll
    for (int c = tid; c < KB * (NC / 2); c += Cfg::THREADS) {
        const int r = c / (NC / 2), c2 = c % (NC / 2);
        const int gk = k0 + r, gc = col0 + 2 * c2;
        int bytes = (gk < k_total) ? (cols_total - gc) * 8 : 0;
        bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const double* src = bytes ? G + (int64_t)gk * ld + gc : G;
        cp_async16_zfill(s + r * LD + 2 * c2, src, bytes);
    }
}

// Rows of op(B) (TB = false) that are GATHERED from another matrix instead of read from B: k-rows kg0 .. k - 1, columns
// < ng, come from G[rowoff[kk - kg0] + col] (rowoff = row index x leading dimension, in SHARED memory).  The kriging
// path uses it to read the right-hand sides C[idx, :] of block row k straight from the covariance matrix inside the
// forward-solve step, instead of from a copy a gather kernel would have to write first (n x M doubles per system).
struct GatherArgs {
    const double* G;
    const int64_t* rowoff;   // shared memory, NB entries
    int kg0, ng;
};

template <typename Cfg, int NC>
__device__ __forceinline__ void load_kcol_gather_c(double* s, const double* B, int64_t ld, int col0, int cols_total,
                                                   int k0, int k_total, int tid, const GatherArgs& ga) {
    constexpr int KB = Cfg::KB, LD = NC + 4;
#pragma unroll
    for (int c = tid; c < KB * (NC / 2); c += Cfg::THREADS) {
        const int r = c / (NC / 2), c2 = c % (NC / 2);
        const int gk = k0 + r, gc = col0 + 2 * c2;
        int bytes = (gk < k_total) ? (cols_total - gc) * 8 : 0;
        bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const double* src = B;
        if (bytes) src = (gk >= ga.kg0 && gc < ga.ng) ? ga.G + ga.rowoff[gk - ga.kg0] + gc : B + (int64_t)gk * ld + gc;
        cp_async16_zfill(s + r * LD + 2 * c2, src, bytes);
    }
}

// The DMMAs of one k-slab for the first V row fragments of a warp (V is a compile-time count: no predicates - ptxas turns
// `if (i < valid) dmma` into PREDICATED DMMAs, which still take their turn on the tensor pipe: profiles/
// r02b_ragged_fragments_ab.txt).  Fragment i of warp row wm holds tile rows i * RSTEP + wm * 8 .. + 7.
template <int V, int MI, int NJ, int KB, bool TA, bool TB, int LDA, int LDB, int LDK_, int RSTEP, int WN>
__device__ __forceinline__ void mma_slab(double (&acc)[MI][NJ][2], const double* __restrict__ sA,
                                         const double* __restrict__ sB, int wm, int wn, int grp, int t4) {
#pragma unroll
    for (int k4 = 0; k4 < KB / 4; ++k4) {
        double a[V], b[NJ];
#pragma unroll
        for (int i = 0; i < V; ++i)
            a[i] = TA ? sA[(k4 * 4 + t4) * LDA + i * RSTEP + wm * 8 + grp]
                      : sA[(i * RSTEP + wm * 8 + grp) * LDK_ + k4 * 4 + t4];
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            b[j] = TB ? sB[(wn * WN + j * 8 + grp) * LDK_ + k4 * 4 + t4]
                      : sB[(k4 * 4 + t4) * LDB + wn * WN + j * 8 + grp];
#pragma unroll
        for (int i = 0; i < V; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// D[m0.., n0..] = alpha op(A) op(B) + beta C for one BM x BNT tile.  `smem` = TileGeom<Cfg, BNT>::SMEM_BYTES, 16 B
// aligned.  BNT = 128 or 64 for CfgBig (the narrow tile halves the work per CTA and is used by the triangular-solve
// steps when the right-hand sides give fewer 128-column strips than the GPU has SMs), 64 for CfgSmall.
// Every thread of the Cfg::THREADS-thread CTA must call this (it contains __syncthreads).
// In-place note: a CTA reads all of its A / B slabs before the epilogue writes D, so D may alias rows of A
// or B that only THIS tile reads (used by the panel solve and the triangular solves).
template <typename Cfg, bool TA, bool TB, int BNT, bool GATHER = false>
__device__ void gemm_tile_cfg(const TileArgs& g, int m0, int n0, double* smem, const GatherArgs* ga = nullptr) {
    static_assert(!GATHER || !TB, "gathered rows are rows of a k x n operand");
    using G_ = TileGeom<Cfg, BNT>;
    constexpr int KB = Cfg::KB, NST = Cfg::NST;
    constexpr int WN = BNT / Cfg::WARPS_N;    // columns per warp
    constexpr int NJ = WN / 8;                // 8-column DMMA tiles per warp along n
    constexpr int WARPS_M = Cfg::THREADS / 32 / Cfg::WARPS_N;
    constexpr int MI = BM / (WARPS_M * 8);    // 8-row DMMA tiles per warp along m (4 for the 4 x 4 and 4 x 2 layouts, 1 for 16 x 1)
    constexpr int LDK_ = G_::LDK_, LDA = G_::LDN_A, LDB = G_::LDN_B;
    static_assert(WN * Cfg::WARPS_N == BNT && NJ * 8 == WN && BNT % 2 == 0, "strip width: a multiple of 8 columns per warp");
    static_assert(MI * WARPS_M * 8 == BM, "warps along m must cover the 128 tile rows");
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
    const int grp = lane >> 2, t4 = lane & 3;

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // Rows of the tile are dealt to the warps in 8-row DMMA fragments ROUND ROBIN: fragment i of warp row wm holds tile rows
    // i * RSTEP + wm * 8 .. + 7 (RSTEP = 8 x warps along m = 32 for the 4 x 4 and 4 x 2 layouts).  A tile on the ragged
    // bottom edge of a matrix (the last block row of a system whose order is not a multiple of 128: half a tile on average,
    // and the forward-solve step with the longest K) then has the same number of valid fragments in every warp, and each
    // warp issues only those.  Skipped fragments feed rows the epilogue never stores; every stored element sees the same
    // operations in the same order as with contiguous bands, so results are bit-identical.
    constexpr int RSTEP = WARPS_M * 8;
    static_assert(MI == 4 || MI == 1, "fragment-count dispatch below is written for 4 or 1 row fragments per warp");
    const int rows_left = g.m - m0 - wm * 8;
    const int mi_valid = rows_left <= 0 ? 0 : (rows_left > (MI - 1) * RSTEP ? MI : (rows_left + RSTEP - 1) / RSTEP);

    const int nk = (g.k + KB - 1) / KB;
    auto load_stage = [&](int kt) {
        double* sA = smem + (size_t)(kt % NST) * G_::STAGE_DOUBLES;
        double* sB = sA + G_::A_DOUBLES;
        if (TA) load_kcol_c<Cfg, BM>(sA, g.A, g.lda, m0, g.m, kt * KB, g.k, tid);
        else    load_rowk_c<Cfg, BM>(sA, g.A, g.lda, m0, g.m, kt * KB, g.k, tid);
        if (TB) load_rowk_c<Cfg, BNT>(sB, g.B, g.ldb, n0, g.n, kt * KB, g.k, tid);
        else if (GATHER && (kt + 1) * KB > ga->kg0) load_kcol_gather_c<Cfg, BNT>(sB, g.B, g.ldb, n0, g.n, kt * KB, g.k, tid, *ga);
        else    load_kcol_c<Cfg, BNT>(sB, g.B, g.ldb, n0, g.n, kt * KB, g.k, tid);
    };
#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
        if (s < nk) load_stage(s);
        cp_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_wait<NST - 2>();
        __syncthreads();                     // slab kt has landed; slab kt-1's buffer is free for the prefetch
        if (kt + NST - 1 < nk) load_stage(kt + NST - 1);
        cp_commit();
        const double* sA = smem + (size_t)(kt % NST) * G_::STAGE_DOUBLES;
        const double* sB = sA + G_::A_DOUBLES;
#define GMG_SLAB(V) mma_slab<V, MI, NJ, KB, TA, TB, LDA, LDB, LDK_, RSTEP, WN>(acc, sA, sB, wm, wn, grp, t4)
        if (mi_valid == MI) {
            GMG_SLAB(MI);
        } else if constexpr (MI == 4) {          // ragged bottom edge: the same count of valid fragments in every warp (+- 1)
            if (mi_valid == 3) GMG_SLAB(3);
            else if (mi_valid == 2) GMG_SLAB(2);
            else if (mi_valid == 1) GMG_SLAB(1);
        }
#undef GMG_SLAB
    }
    cp_wait<0>();
    __syncthreads();                         // all warps done with shared memory (callers may reuse it)

    const bool vec_ok = ((g.ldd & 1) == 0) && ((reinterpret_cast<uintptr_t>(g.D) & 15) == 0) &&
                        ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int row = m0 + i * RSTEP + wm * 8 + grp;
        if (row >= g.m) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int col = n0 + wn * WN + j * 8 + 2 * t4;
            if (col >= g.n) continue;
            double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
            double* d = g.D + (int64_t)row * g.ldd + col;
            if (col + 1 < g.n && vec_ok) {
                if (g.beta != 0.0) {
                    const double2 c = *reinterpret_cast<const double2*>(g.C + (int64_t)row * g.ldc + col);
                    v0 = fma(g.beta, c.x, v0);
                    v1 = fma(g.beta, c.y, v1);
                }
                *reinterpret_cast<double2*>(d) = make_double2(v0, v1);
            } else {
                if (g.beta != 0.0) v0 = fma(g.beta, g.C[(int64_t)row * g.ldc + col], v0);
                d[0] = v0;
                if (col + 1 < g.n) {
                    if (g.beta != 0.0) v1 = fma(g.beta, g.C[(int64_t)row * g.ldc + col + 1], v1);
                    d[1] = v1;
                }
            }
            if (g.D2) {
                double* d2 = g.D2 + (int64_t)row * g.ldd2 + col;
                d2[0] = v0;
                if (col + 1 < g.n) d2[1] = v1;
            }
        }
    }
}

// the 512-thread tile (what every kernel used so far)
template <bool TA, bool TB, int BNT = BN>
__device__ __forceinline__ void gemm_tile(const TileArgs& g, int m0, int n0, double* smem) {
    gemm_tile_cfg<CfgBig, TA, TB, BNT>(g, m0, n0, smem);
}

}  // namespace gmg
