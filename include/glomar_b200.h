/* libglomar_b200.so - C ABI of the B200-native GloMarGridding hot path.
 *
 * The reference (jtsiddons/GloMarGridding, `glomar_gridding` 1.0.0) is pure Python and has no
 * FFI layer: its boundary for this path is the Python class API
 * (`EllipseCovarianceBuilder`, `SimpleKriging`, `OrdinaryKriging`, `StochasticKriging`,
 * `scipy_mv_normal_draw`).  Those classes are kept by the `glomargridding_b200` Python package
 * and call the entry points below through `ctypes`; INTEGRATION.md shows the binding a
 * maintainer of the reference would add.  Each entry point cites the reference code it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *  - All pointers are DEVICE pointers on the current CUDA device unless named `h_*`.
 *  - The library never allocates or frees persistent device memory: buffers and workspaces are
 *    caller-owned (`*_ws_bytes` queries give the sizes).
 *  - Every call only enqueues work on `stream` (a `cudaStream_t` passed as void*; NULL = legacy
 *    default stream) and does not synchronise, except where documented.
 *  - Return value: 0 ok; negative = 1-based index of a bad argument; positive = failure class
 *    (1 CUDA error, 2 matrix not positive definite, 3 workspace too small, 4 unsupported,
 *    5 singular, 6 NCCL unavailable / collective failed).  `gmg_strerror` / `gmg_last_cuda_error` give detail.  No exceptions, no exit().
 *  - Matrices are row-major (C order, numpy default) with an explicit leading dimension in
 *    elements.
 */
#ifndef GLOMAR_B200_H
#define GLOMAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gmg_stream_t; /* cudaStream_t */

/* delta_x_method (reference glomar_gridding/types.py:46, ellipse/covariance.py:182-203) */
#define GMG_DISP_MET_OFFICE 0
#define GMG_DISP_MODIFIED_MET_OFFICE 1

/* Matern evaluation strategy */
#define GMG_MATERN_AUTO 0   /* closed form for nu in {0.5, 1.5, 2.5}, table otherwise          */
#define GMG_MATERN_TABLE 1  /* piecewise-polynomial table of x^nu K_nu(x) e^x built on device  */
#define GMG_MATERN_DIRECT 2 /* Temme / CF2 Bessel-K per pair (slow, used to validate the table) */

int gmg_version(void);
const char* gmg_strerror(int code);
int gmg_last_cuda_error(void);

/* ---------------------------------------------------------------------------------------------
 * Ellipse covariance (Paciorek-Schervish non-stationary Matern)
 * replaces glomar_gridding/ellipse/covariance.py:205-468 and the helpers at :512-654
 * ------------------------------------------------------------------------------------------- */

/* Bytes of workspace needed by gmg_ellipse_prepare_f64 / gmg_ellipse_cov_f64 for N points. */
int gmg_ellipse_ws_bytes(int64_t N, size_t* bytes);

/* Per-point pre-terms: Sigma_i = R(theta) diag(Lx^2, Ly^2) R(theta)^T, sqrt(det Sigma_i), cos(lat),
 * and the Matern table for `nu`, packed into `ws`.
 * replaces `_sigma_rot_func_multi` (:512-533), `_det_22_multi` (:543-547) and the pre-terms of
 * `_calculate_covariance` (:218-224).  lat/lon are RADIANS as produced by numpy.deg2rad on the host
 * (:171-172) so that the +-pi longitude-wrap ties are decided on bit-identical inputs.
 * `disp_method` selects the displacement function of `_get_disp_fn` (:182-203): it only changes the
 * per-point mean-cos-latitude term, so it is fixed here and not passed to the build calls. */
int gmg_ellipse_prepare_f64(const double* Lx, const double* Ly, const double* theta,
                            const double* stdev, const double* lat_rad, const double* lon_rad,
                            int64_t N, double nu, int disp_method, void* ws, size_t ws_bytes,
                            gmg_stream_t stream);

/* Rows [row0, row0 + nrows) of the symmetric N x N covariance, written to `C_rows`
 * (nrows x N, leading dimension `ld` elements, FP64 or - if store_f32 - FP32 rounded on store).
 * Diagonal entries are stdev^2.  Pairs whose haversine distance exceeds `max_dist_km` are 0
 * (max_dist_km < 0 disables the cut-off).
 * replaces `calculate_covariance_array` / `_batched` / `_loop` (:250-403), `c_ij_anisotropic_array`
 * (:405-468), the displacement helpers (:591-654), `_haversine_multi` (:572-588), the mirror
 * (:277, :401) and the diagonal add (:246).  Row blocks are how the build shards across GPUs. */
int gmg_ellipse_cov_f64(const void* ws, int64_t N, int64_t row0, int64_t nrows, double nu,
                        double max_dist_km, int matern_mode, void* C_rows, int64_t ld,
                        int store_f32, gmg_stream_t stream);

/* Multi-GPU build in which every pair of points is evaluated ONCE across the box: rank `rank` of `world`
 * (one process per GPU) owns the tile-aligned row block [h_block_row0[rank], h_block_row0[rank + 1]) and
 * computes its diagonal block plus the blocks at cyclic distance 1 .. world/2 to its right; each of those tiles
 * is stored into its own rows and, transposed, straight into the owning peer's row block over NVLink
 * (`h_block_rows[g]` = device pointer of rank g's row block, mapped into this process with gmg_ipc_open_f64).
 * Both `h_*` arrays are HOST arrays.  The caller synchronises the ranks after the launch (every rank's block is
 * complete when ALL ranks have finished).  Same replaced reference code as gmg_ellipse_cov_f64; the reference has
 * no multi-device path. */
int gmg_ellipse_cov_p2p_f64(const void* ws, int64_t N, double nu, double max_dist_km, int matern_mode,
                            int world, int rank, const int64_t* h_block_row0, void* const* h_block_rows,
                            int64_t ld, int store_f32, gmg_stream_t stream);

/* CUDA IPC plumbing for the call above.  An IPC handle names a whole cudaMalloc block, which a caching allocator
 * cannot promise, so the peer-visible row blocks are allocated by gmg_ipc_alloc_f64 (the one exception to "the
 * library never allocates") and freed by gmg_ipc_free_f64; gmg_ipc_open_f64 / gmg_ipc_close_f64 map / unmap a
 * peer's block (64-byte cudaIpcMemHandle_t obtained in the peer process) in this process. */
int gmg_ipc_alloc_f64(size_t bytes, void** dev_ptr, void* handle64_out); /* cudaMalloc + cudaIpcGetMemHandle */
int gmg_ipc_free_f64(void* dev_ptr);
int gmg_ipc_open_f64(const void* handle64, void** dev_ptr);
int gmg_ipc_close_f64(void* dev_ptr);

/* Covariance of arbitrary index pairs (i[k], j[k]) - no diagonal term, no max_dist.
 * replaces `EllipseCovarianceBuilder.c_ij_anisotropic_array(i_s, j_s)` (:405-468). */
int gmg_ellipse_pairs_f64(const void* ws, int64_t N, const int64_t* i, const int64_t* j,
                          int64_t npairs, double nu, int matern_mode, double* out,
                          gmg_stream_t stream);

/* out[k] = K_nu(x[k]) (scaled != 0: e^x K_nu(x)), Temme/CF2 on device.
 * replaces the `scipy.special.kv` call sites (ellipse/covariance.py:346, :466). */
int gmg_bessel_kv_f64(double nu, const double* x, int64_t n, int scaled, double* out,
                      gmg_stream_t stream);

/* out[k] = x^nu K_nu(x) / (Gamma(nu) 2^(nu-1)) (the Matern correlation at scaled distance x)
 * evaluated through the table in `ws` (matern_mode TABLE) or directly (DIRECT / closed forms). */
int gmg_matern_f64(const void* ws, double nu, int matern_mode, const double* x, int64_t n,
                   double* out, gmg_stream_t stream);

/* Haversine (asin form) distance matrix rows on a sphere of the given radius (km), at most 65535 rows per call.
 * replaces `distances.haversine_distance_from_frame` (distances.py:191-218: sklearn haversine_distances * radius). */
int gmg_haversine_f64(const double* lat_rad, const double* lon_rad, int64_t N, int64_t row0,
                      int64_t nrows, double radius, double* D_rows, int64_t ld, gmg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Dense FP64 linear algebra of the kriging path (DMMA = mma.sync m8n8k4 f64 tiles)
 * replaces the LAPACK/BLAS calls of glomar_gridding/kriging.py:343,585,768 (np.linalg.solve),
 * stochastic.py:159 (np.linalg.inv) and the `@` products at kriging.py:378,634, stochastic.py:210-216
 * All matrices row-major; pointers and leading dimensions must be multiples of 2 doubles (16 B).
 * ------------------------------------------------------------------------------------------- */

/* C (m x n) = alpha op(A) op(B) + beta C;  op(X) = X or X^T (trans != 0); A stored m x k (k x m if transa),
 * B stored k x n (n x k if transb). */
int gmg_gemm_f64(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                 const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
                 int64_t ldc, gmg_stream_t stream);

/* One symmetric positive definite system per descriptor (array in DEVICE memory, built by the caller). */
typedef struct GmgSolveDesc {
    double* S;      /* n x n, lower triangle read; overwritten by the Cholesky factor L (lower)     */
    int64_t lds;
    double* Linv;   /* workspace: ceil(n/128) inverted 128 x 128 diagonal blocks of L               */
    double* R;      /* n x nrhs right-hand sides, overwritten by S^-1 R                             */
    int64_t ldr;
    double* P;      /* workspace: coefficient strip, `p_doubles` of gmg_solve_ws_doubles            */
    int64_t ldp;    /* leading dimension of the forward strip, `ldp` of gmg_solve_ws_doubles        */
    int* info;      /* 0, or the order of the first non-positive pivot (LAPACK potrf convention)    */
    int32_t n;
    int32_t nrhs;
} GmgSolveDesc;

/* Workspace sizes (in doubles) for one system of order n. */
int gmg_solve_ws_doubles(int64_t n, int64_t* linv_doubles, int64_t* p_doubles, int64_t* ldp);

/* S = L L^T for `batch` systems (n_max = largest order in the batch): blocked right-looking
 * Cholesky, trailing update on DMMA tiles.  *info must be zeroed by the caller. */
int gmg_potrf_batched_f64(const GmgSolveDesc* descs, int batch, int64_t n_max, gmg_stream_t stream);

/* One phase of block step `kblk` of the same factorisation, for callers that interleave communication:
 * GMG_POTRF_DIAG factors and inverts the 128 x 128 diagonal block, GMG_POTRF_PANEL forms L21 = S21 L11^-T,
 * GMG_POTRF_UPDATE applies S22 -= L21 L21^T; PANEL / UPDATE only touch tile rows r (in units of 128 rows from the
 * top of the matrix) with r % tr_mod == tr_rem - the 1-D block-cyclic ownership of the multi-GPU factorisation
 * (glomargridding_b200/distributed.py: the panel is exchanged between the phases with one NCCL all_reduce). */
#define GMG_POTRF_DIAG 0
#define GMG_POTRF_PANEL 1
#define GMG_POTRF_UPDATE 2
#define GMG_POTRF_UPDATE_NEXT 3 /* only the first column tile of the trailing matrix (what the next panel needs) */
#define GMG_POTRF_UPDATE_REST 4 /* the other column tiles: NEXT + REST == UPDATE, the look-ahead split             */
int gmg_potrf_step_f64(const GmgSolveDesc* descs, int batch, int64_t n_max, int kblk, int phase,
                       int tr_mod, int tr_rem, gmg_stream_t stream);

/* Triangular solves with the factored systems, each block step one GEMM (diagonal blocks were inverted
 * by gmg_potrf_batched_f64):  which = GMG_SOLVE_FORWARD: R <- L^-1 R;  GMG_SOLVE_BACKWARD: R <- L^-T R;
 * GMG_SOLVE_BOTH: R <- S^-1 R. */
#define GMG_SOLVE_FORWARD 1
#define GMG_SOLVE_BACKWARD 2
#define GMG_SOLVE_BOTH 3
int gmg_potrs_batched_f64(const GmgSolveDesc* descs, int batch, int64_t n_max, int64_t nrhs_max,
                          int which, gmg_stream_t stream);

/* Factorisation AND forward solve as one overlapped pipeline: S = L L^T and R <- L^-1 R (nrhs_max = 0: factor only)
 * with exactly the arithmetic of gmg_potrf_batched_f64 followed by gmg_potrs_batched_f64(GMG_SOLVE_FORWARD) - results
 * are bit-identical - but on three internal streams: the latency-bound critical path of the factorisation (diagonal
 * block, panel, update of the next panel column), the bulk of the trailing update, and the forward-solve step of
 * every block row as soon as its diagonal block exists.  The call only enqueues: the work starts after everything
 * already in `stream`, and `stream` continues after all of it.
 *
 * Multi-GPU (the SURVEY 8b "gmg_dist_*" handles; replaces the LU inside np.linalg.solve, kriging.py:585, for a system
 * that one GPU factors too slowly): one process per GPU, every rank holds a replica of S and ITS OWN columns of the
 * right-hand sides.  Tile rows (128 rows) are owned block-cyclically, rank = tile row % world; per block step the owner
 * broadcasts the factored + inverted diagonal block, every rank forms and packs its tile rows of the panel and one
 * all-gather hands everyone the whole panel; all ranks end up with the complete factor.  The collectives are NCCL
 * calls issued by this library on its own communicator: gmg_dist_unique_id (one rank) -> the caller distributes the
 * 128 bytes -> gmg_dist_init on every rank (collective).  NCCL is taken from the libnccl.so.2 already loaded in the
 * process; GMG_ERR_COMM (6) if there is none.  ctx = NULL: single GPU, no communication.
 *   xbuf: device workspace of gmg_dist_ws_doubles(n_max, world) doubles (0 for one GPU), 16-byte aligned. */
int gmg_dist_unique_id(void* id128);
int gmg_dist_init(const void* id128, int world, int rank, void** ctx);
int gmg_dist_destroy(void* ctx);
int gmg_dist_ws_doubles(int64_t n, int world, int64_t* doubles);
int gmg_potrf_solve_f64(void* ctx, const GmgSolveDesc* descs, int batch, int64_t n_max, int64_t nrhs_max,
                        double* xbuf, int64_t xbuf_doubles, gmg_stream_t stream);

/* Indefinite-safe fallback: R <- S^-1 R by LU with partial pivoting (what the reference's
 * np.linalg.solve does, kriging.py:343, :585, :768).  S is the FULL n x n matrix (both triangles) and is
 * overwritten by its LU factors; pivots go to the descriptor's P workspace; *info = k + 1 for an exactly
 * zero pivot.  One CTA per system: meant for the matrices Cholesky rejects, not for speed. */
int gmg_gesv_batched_f64(const GmgSolveDesc* descs, int batch, int64_t nrhs_max, gmg_stream_t stream);

/* The same solve as a BLOCKED LU with partial pivoting on DMMA tiles (right-looking over 128-column panels: panel
 * factorisation with LAPACK's pivot choices, row interchanges, U12 = L11^-1 A12 and A22 -= L21 U12 as GEMMs, then two
 * blocked substitutions whose steps are GEMMs) - what np.linalg.solve costs at speed for the systems Cholesky rejects
 * (indefinite covariance of the default displacement method, asymmetric error covariance).  Workspaces per system, sized
 * for the LARGEST system of the batch by gmg_gesv_ws_doubles(n_max, ...): `Linv` (L^-1 and U^-1 of every diagonal
 * block), `P` (coefficient strip, leading dimension `ldp`, then the panel's exchange area, then the int32 pivots - row
 * interchanged with row k at step k, 0-based, at double offset gmg_gesv_pivots_offset(ldp) of P).  nrhs_max = 0: factor
 * only.  The panel splits its rows over co-resident CTAs (cooperative launch) that meet once per column.
 * S, R, Linv and P must be 16-byte aligned and lds, ldr even (the sizes gmg_gesv_ws_doubles returns are even, so
 * workspaces carved back to back out of one aligned allocation qualify). */
int gmg_gesv_ws_doubles(int64_t n, int64_t* linv_doubles, int64_t* p_doubles, int64_t* ldp);
int64_t gmg_gesv_pivots_offset(int64_t ldp);
int gmg_gesv_blocked_f64(const GmgSolveDesc* descs, int batch, int64_t n_max, int64_t nrhs_max, gmg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Kriging assembly and application
 * ------------------------------------------------------------------------------------------- */

/* One kriging problem (descriptor array in DEVICE memory).  T (n x (M + 2), leading dimension
 * ldt >= M + 2) starts as the right-hand sides [C_cross | 1 | y] and is overwritten by the solves:
 * after the forward solve it holds Y = L^-1 [C_cross | 1 | y], after both S^-1 [C_cross | 1 | y]
 * (the "transposed weights", column M = o = S^-1 1). */
typedef struct GmgKrigeDesc {
    const int64_t* idx;  /* n observation cells                                                    */
    const double* E;     /* n x n error covariance or NULL (float32 values when `ordinary` has
                          * GMG_KRIGE_E_FLOAT32 set: same layouts, 4-byte elements, widened on the device)  */
    int64_t lde;         /* > 0: row-major, leading dimension lde.  < 0: E is SYMMETRIC and packed as a block
                          * trapezoid with row blocks of hb = -lde rows: row i of block b = i / hb stores its first
                          * min(n, hb (b + 1)) entries, blocks one after the other (half the bytes to upload)    */
    const double* y;     /* n observations                                                         */
    double* S;           /* n x n  out: C[idx, idx] + E                                            */
    int64_t lds;
    double* Cx;          /* n x M  out: C[idx, :]; may be NULL when only the forward-solve apply is used */
    int64_t ldx;
    double* T;           /* n x (M + 2)  out: [Cx | 1 | y] = right-hand sides                      */
    int64_t ldt;
    double* partial;     /* workspace: 4 * ceil(n/256) * M doubles                                 */
    double* out;         /* 5 x M: field, u2 (variance before clamping), mask, alpha, rowdot       */
    int32_t n;
    int32_t ordinary;    /* bit 0: 0 simple kriging, 1 ordinary kriging (Lagrange-extended system);
                          * bit 1 (GMG_KRIGE_E_FLOAT32): E points at float32 values                  */
} GmgKrigeDesc;
#define GMG_KRIGE_E_FLOAT32 2

/* S = C[idx, idx] + E, Cx = C[idx, :], T = [Cx | 1 | y]   (reference kriging.py:337-342, :571-583).
 * NaN entries of E are read as 0.0 (kriging.py:120).  With even leading dimensions and 16-byte aligned base
 * pointers the rows move as 16-byte vectors; anything else takes the scalar path of the same kernel. */
int gmg_krige_gather_batched_f64(const double* C, int64_t M, int64_t ldc, const GmgKrigeDesc* descs,
                                 int batch, int64_t n_max, gmg_stream_t stream);

/* gather + Cholesky + forward triangular solve of a batch of kriging systems as ONE call on one GPU: S = C[idx, idx] + E
 * is assembled, and the overlapped factor + forward-solve pipeline of gmg_potrf_solve_f64 runs with the right-hand
 * sides [C[idx, :] | 1 | y] of every block row read from C ITSELF inside the solve step (only the two extra columns
 * of T are written beforehand) - the n x M copy of C[idx, :] that gmg_krige_gather_batched_f64 writes, and the solve
 * then reads back, never exists.  Results (S = L L^T, T = L^-1 [Cx | 1 | y], *info) are bit-identical to
 * gmg_krige_gather_batched_f64 + gmg_potrf_solve_f64.  kd[b] and sd[b] describe the same system: sd[b].S = kd[b].S,
 * sd[b].R = kd[b].T, sd[b].nrhs = M + 2; kd[b].Cx is not written.  Needs an even M and ldc and a 16-byte aligned C
 * (GMG_ERR_UNSUPPORTED otherwise: use the two separate calls).  replaces kriging.py:337-343, :571-585. */
int gmg_krige_solve_f64(const double* C, int64_t M, int64_t ldc, const GmgKrigeDesc* kd, const GmgSolveDesc* sd, int batch,
                        int64_t n_max, gmg_stream_t stream);

/* field = W y, variance = diag(C) - rowdot(W, Cx^T) [- 2 alpha], constraint mask = rowdot(W_simple,
 * Cx^T) / diag(C), alpha = Lagrange multiplier column, without forming W (M x n) or the reference's
 * M x M temporaries (kriging.py:417, :434-435, :487-491, :677-679, :699-701, :772-776).  Ordinary
 * kriging uses the Schur complement of the bordered system: alpha = (o^T Cx - 1) / (1^T o),
 * W = (S^-1 Cx - o alpha^T)^T.
 * fwd_only != 0: T holds Y = L^-1 [Cx | 1 | y] (forward solve only); since S^-1 = L^-T L^-1 every
 * result is a column dot product of Y (d = sum Y^2, ...) - the second triangular solve is skipped.
 * fwd_only == 0: T holds S^-1 [Cx | 1 | y] and Cx must be present (LU fallback, user-supplied weights). */
/* out (n x n, row-major, leading dimension ldo) <- the lower triangle incl. diagonal of the error covariance of system
 * `b` of a descriptor array (zeros above it), whether it was uploaded whole or as its packed block trapezoid; NaN -> 0.
 * What the Cholesky factor of E itself is computed from (noise draws, glomar_gridding/stochastic.py:379-383). */
int gmg_krige_error_cov_lower_f64(const GmgKrigeDesc* descs, int b, int64_t n, double* out, int64_t ldo,
                                  gmg_stream_t stream);
int gmg_krige_apply_batched_f64(const double* C, int64_t M, int64_t ldc, const GmgKrigeDesc* descs,
                                int batch, int64_t n_max, int fwd_only, gmg_stream_t stream);

/* Materialise the reference's `kriging_weights`: W (M x n, or M x (n+1) with the Lagrange column when
 * alpha != NULL) from T (n x M) :  W[m, a] = T[a, m] - o[a] alpha[m],  W[m, n] = alpha[m]. */
int gmg_krige_weights_f64(const double* T, int64_t ldt, const double* o, int64_t o_stride,
                          const double* alpha, int64_t n, int64_t M, double* W, int64_t ldw,
                          gmg_stream_t stream);

/* Inverse of [[S, 1], [1^T, 0]] from S^-1 (n x n) by the bordered-matrix identity; `ws_e` = n doubles.
 * replaces `_extended_inverse`, glomar_gridding/kriging.py:815-829. */
int gmg_extended_inverse_f64(const double* inv, int64_t n, int64_t ld, double* ws_e, double* out,
                             int64_t ldo, gmg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stochastic draws   replaces glomar_gridding/stochastic.py:395-502 (np.random.multivariate_normal)
 * A draw is X = Z L^T with L from gmg_potrf_batched_f64 (upper triangle cleared by gmg_tril_f64),
 * Z from gmg_normal_f64 and the product by gmg_gemm_f64.
 * ------------------------------------------------------------------------------------------- */

/* out[0..n) ~ N(0, 1): Philox4x32-10 counter (offset + pair index) keyed by `seed`, Box-Muller. */
int gmg_normal_f64(double* out, int64_t n, uint64_t seed, uint64_t offset, gmg_stream_t stream);

/* Zero the strict upper triangle of the row-major n x n matrix A. */
int gmg_tril_f64(double* A, int64_t n, int64_t ld, gmg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stationary variogram models on a distance matrix (the reference workflow's other covariance route)
 * replaces glomar_gridding/variogram.py:96-138, :177-218, :257-289, :387-476 (the four `.fit`) and
 * :479-504 (`variogram_to_covariance`, fused when to_covariance != 0: out = variance - variogram).
 * matern_method: 0 "sklearn", 1 "gstat", 2 "karspeck".  fill_diag: set the diagonal to the nugget
 * (MaternVariogram.fit does, :470-475).  Scalar parameters only.
 * ------------------------------------------------------------------------------------------- */
#define GMG_VARIOGRAM_SPHERICAL 0
#define GMG_VARIOGRAM_GAUSSIAN 1
#define GMG_VARIOGRAM_EXPONENTIAL 2
#define GMG_VARIOGRAM_MATERN 3
int gmg_variogram_f64(int model, const double* dist, int64_t rows, int64_t cols, int64_t ld,
                      double psill, double nugget, double range, double nu, int matern_method,
                      int fill_diag, int to_covariance, double variance, double* out, int64_t ldo,
                      gmg_stream_t stream);

/* The stationary route FUSED: haversine distance matrix (distances.py:191-218: sklearn's asin form x radius) ->
 * variogram model (.fit) -> variance - variogram (variogram.py:479-504) in one pass over the N x N output, which is
 * written once and never read; lat/lon in radians, cells in the order of the frame's rows.  Bit-identical to
 * gmg_haversine_f64 followed by gmg_variogram_f64 (same arithmetic; only the upper triangle is evaluated, the
 * distance being symmetric bit for bit).  to_covariance = 0 returns the variogram itself. */
int gmg_stationary_cov_f64(int model, const double* lat_rad, const double* lon_rad, int64_t N, double radius,
                           double psill, double nugget, double range, double nu, int matern_method, int fill_diag,
                           int to_covariance, double variance, double* out, int64_t ldo, gmg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Small symmetric eigensolver (cyclic two-sided Jacobi, cooperative launch) for the device
 * `eigenvalue_clip`: the reference takes a full numpy.linalg.eigh of the covariance
 * (glomar_gridding/covariance_tools.py:616, :717); here only the m x m Rayleigh-Ritz matrix of a
 * subspace iteration is diagonalised.  A (m x m symmetric, m even) is overwritten - its diagonal
 * holds the eigenvalues (unsorted); U (m x m) receives the eigenvectors as columns.
 * ------------------------------------------------------------------------------------------- */
int gmg_jacobi_eigh_ws_bytes(int64_t m, size_t* bytes);
int gmg_jacobi_eigh_f64(double* A, int64_t lda, double* U, int64_t ldu, int64_t m, int max_sweeps,
                        double tol, void* ws, size_t ws_bytes, int* sweeps_done /* device */,
                        gmg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * N x N bookkeeping of the covariance path, arithmetic in the matrix' own precision (is_f32).
 *   gmg_cov2cor     utils.cov_2_cor, glomar_gridding/utils.py:637-671 (EllipseCovarianceBuilder.calculate_cor,
 *                   ellipse/covariance.py:470-473): cor = cov / outer(sd, sd), sd = sqrt(diag(cov)) (written to
 *                   sd_ws, n elements), zeros kept.  stat (device, 2 x u64): [0] = number of diagonal entries
 *                   != 1, [1] = bit pattern of the largest |diag - 1| as a double - the caller resets the diagonal
 *                   or raises like the reference (:655-668).
 *   gmg_cor2cov     utils.cor_2_cov, utils.py:611-634: cov = cor * outer(sd, sd), zeros kept.
 *   gmg_uncompress  EllipseCovarianceBuilder.uncompress_cov, ellipse/covariance.py:475-509: out (M x M) gets
 *                   cov[inv[r], inv[c]] where both points are unmasked (inv >= 0), diag_fill on the rest of the
 *                   diagonal, fill elsewhere.
 * ------------------------------------------------------------------------------------------- */
int gmg_cov2cor(const void* cov, int64_t ld, int64_t n, int is_f32, void* sd_ws, void* cor, int64_t ldo,
                unsigned long long* stat, gmg_stream_t stream);
int gmg_cor2cov(const void* cor, int64_t ld, int64_t n, int is_f32, const void* sd, void* cov, int64_t ldo,
                gmg_stream_t stream);
int gmg_uncompress(const void* cov, int64_t ld, const int* inv, int64_t M, int is_f32, double fill,
                   double diag_fill, void* out, int64_t ldo, gmg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * All eigenvalues of a dense symmetric matrix (ascending) - the spectrum `explained_variance_clip`
 * needs to find its cut (glomar_gridding/covariance_tools.py:616, :717 take them from a full
 * numpy.linalg.eigh).  Householder tridiagonalisation in one cooperative kernel (rank-2 update fused
 * with the next matrix-vector product), then Sturm-count bisection.  A (n x n, full symmetric
 * storage, row-major) is DESTROYED; w receives n eigenvalues.  n <= 8192.
 * ------------------------------------------------------------------------------------------- */
int gmg_sym_eigvals_ws_bytes(int64_t n, size_t* bytes);
int gmg_sym_eigvals_f64(double* A, int64_t lda, int64_t n, double* w, void* ws, size_t ws_bytes,
                        gmg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Elementwise geometry helpers of the ellipse covariance path (the module-level functions next to the
 * reference's builder; its tests call them directly, test/test_matrix.py:74-140).  Device pointers,
 * n elements, radians.
 *   gmg_sigma_rot_f64          _sigma_rot_func_multi, glomar_gridding/ellipse/covariance.py:512-533:
 *                              sigmas (n x 4, 16-byte aligned) = [c2 Lx2 + s2 Ly2, cs (Lx2 - Ly2), same, s2 Lx2 + c2 Ly2]
 *   gmg_det22_f64              _det_22_single / _det_22_multi, :536-547: m0 m3 - m1 m2 of n flattened 2 x 2 matrices
 *   gmg_pair_haversine_f64     _haversine_single / _haversine_multi, :550-588: asin-form haversine of n point pairs
 *   gmg_pair_displacement_f64  _mo_disp_* (modified = 0) / _mod_mo_disp_* (modified = 1), :591-654: (dy, dx) in km,
 *                              longitude difference wrapped into [-pi, pi] with strict comparisons
 * ------------------------------------------------------------------------------------------- */
int gmg_sigma_rot_f64(const double* Lx, const double* Ly, const double* theta, int64_t n, double* sigmas,
                      gmg_stream_t stream);
int gmg_det22_f64(const double* mats, int64_t n, double* dets, gmg_stream_t stream);
int gmg_pair_haversine_f64(const double* lat0, const double* lon0, const double* lat1, const double* lon1,
                           int64_t n, double radius_km, double* dist, gmg_stream_t stream);
int gmg_pair_displacement_f64(const double* lat0, const double* lon0, const double* lat1, const double* lon1,
                              int64_t n, int modified, double radius_km, double* dy, double* dx,
                              gmg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Observation -> grid index ("map_to_grid index production"): idx[k] (+)= stride * argmin_i |grid[i] -
 * values[k]| - the nearest grid coordinate of every value, FIRST minimum on ties (utils.find_nearest,
 * glomar_gridding/utils.py:178-214, as grid.map_to_grid calls it per coordinate, grid.py:28-124).  Calling it for
 * latitude with stride = n_lon, accumulate = 0 and for longitude with stride = 1, accumulate = 1 yields the row-major
 * flat index of np.ravel_multi_index - the `grid_idx` the kriging classes take as `idx`.
 * ------------------------------------------------------------------------------------------- */
int gmg_nearest_index_f64(const double* grid, int64_t g, const double* values, int64_t n, int64_t* idx,
                          int64_t stride, int accumulate, gmg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Ellipse-parameter fit: the maximum-likelihood estimate of (Lx, Ly, theta[, sigma]) per grid cell that
 * produces the inputs of the covariance builder.  Replaces EllipseModel.negative_log_likelihood / .fit
 * (glomar_gridding/ellipse/model.py:266-384, :402-500: scipy Nelder-Mead with bounds on the Gaussian likelihood of the
 * Fisher-transformed correlations), cov_ij_anisotropic / cov_ij_isotropic (:582-686) and the per-cell Python loop of
 * EllipseBuilder.compute_params with its training-set selection (ellipse/estimate.py:441-513, :515-715).
 *
 * Problems are given in CSR form: problem b owns training pairs offs[b] .. offs[b + 1] of x0 (east-west displacement,
 * or the distance for the 1-parameter isotropic form), x1 (north-south displacement; NULL for isotropic) and y
 * (observed correlations); `total` = offs[n_problems].  n_params = 1 (R), 2 (Lx, Ly) or 3 (Lx, Ly, theta); unless
 * unit_sigma the likelihood's sigma is fitted as one more parameter.  h_guess / h_lo / h_hi: HOST arrays of
 * n_params [+ 1] initial values and bounds shared by all problems.  xatol / fatol / maxiter / maxfun as scipy's
 * options (maxiter, maxfun <= 0: N * 200).  scratch: 2 * total doubles.  Outputs per problem: x (n_params [+ 1]),
 * fun, nit, nfev, status (0 converged, 1 function-evaluation budget, 2 iteration budget - scipy's codes).
 * gmg_ellipse_nll_f64 evaluates the negative log-likelihood of given parameters (params: n_problems x N, device).
 * gmg_ellipse_train_f64 builds the training pairs of `cells` from the unmasked cells' coordinates (degrees) and their
 * correlation matrix: call with fill = 0 to get counts[n_cells], scan them into offs, call with fill = 1.
 * disp_method: -1 none (degrees), 0 Met_Office, 1 Modified_Met_Office.
 * ------------------------------------------------------------------------------------------- */
int gmg_ellipse_fit_f64(const int64_t* offs, int64_t n_problems, int64_t total, const double* x0, const double* x1,
                        const double* y, int n_params, int unit_sigma, int arctanh, double nu, const double* h_guess,
                        const double* h_lo, const double* h_hi, double xatol, double fatol, int maxiter, int maxfun,
                        double* scratch, double* x_out, double* fun_out, int* nit_out, int* nfev_out, int* status_out,
                        gmg_stream_t stream);
int gmg_ellipse_nll_f64(const int64_t* offs, int64_t n_problems, int64_t total, const double* x0, const double* x1,
                        const double* y, int n_params, int unit_sigma, int arctanh, double nu, const double* params,
                        double* scratch, double* fun_out, gmg_stream_t stream);
int gmg_ellipse_train_f64(const double* lat_deg, const double* lon_deg, int64_t N, const double* cor, int64_t ldc,
                          const int64_t* cells, int64_t n_cells, int anisotropic, int physical_distance,
                          int disp_method, int physical_distance_selection, double min_distance, double max_distance,
                          int fill, int64_t* counts, const int64_t* offs, double* x0, double* x1, double* y,
                          gmg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * HOST helper (host pointers, no device work): rows r0 .. r1 of the n x n host matrix E (leading dimension ld in
 * elements, float64 or float32) against columns r0 .. r1, for the columns up to and including the diagonal block -
 * returns 1 if every mirrored pair of entries is equal (NaN equal to NaN), 0 if not, negative for a bad argument.
 * The reference adds the error covariance as given and solves by LU (glomar_gridding/kriging.py:342-343, :576-585);
 * this library uploads only the lower block trapezoid of a SYMMETRIC matrix and factors by Cholesky, so the packing
 * threads decide symmetry exactly with this routine (calling it for every row block covers the whole matrix).
 * ------------------------------------------------------------------------------------------- */
int gmg_host_rows_symmetric(const void* E, int is_f32, int64_t n, int64_t ld, int64_t r0, int64_t r1);
/* HOST helper: rows x cols block of a host matrix (row stride ld_src elements, float64 or float32) -> contiguous
 * float64 at dst (a pinned staging buffer of the upload pipeline) with streaming stores; float32 is widened. */
int gmg_host_pack_f64(const void* src, int is_f32, int64_t rows, int64_t cols, int64_t ld_src, double* dst);
/* The same for a float32 block that STAYS float32 (GMG_KRIGE_E_FLOAT32: half the staging writes and half the bytes over
 * PCIe for HadCRUT-style float32 error covariances). */
int gmg_host_pack_f32(const float* src, int64_t rows, int64_t cols, int64_t ld_src, float* dst);

/* ---------------------------------------------------------------------------------------------
 * Measurement utility (no reference counterpart; SURVEY.md section 7: "FP64 peak is unmeasured - a DFMA-saturation
 * and a DMMA microbenchmark must be run first; roofline fractions are against those").  Runs a register-only
 * DFMA (what = GMG_MICROBENCH_DFMA) or mma.sync m8n8k4 f64 (GMG_MICROBENCH_DMMA) kernel that fills every SM, `reps`
 * launches back to back plus three isolated ones, SYNCHRONISES the stream, and reports TFLOP/s: `tflops_best` from the
 * fastest single launch (burst clocks), `tflops_sustained` (may be NULL) over the back-to-back window.
 * `d_scratch`: at least 8 bytes of device memory.  bench.py calls this next to the timed kernels.
 * ------------------------------------------------------------------------------------------- */
#define GMG_MICROBENCH_DFMA 0
#define GMG_MICROBENCH_DMMA 1
int gmg_microbench_fp64(int what, int iters, int reps, double* d_scratch, double* tflops_best,
                        double* tflops_sustained, gmg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GLOMAR_B200_H */
