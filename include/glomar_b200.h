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
 *    5 singular).  `gmg_strerror` / `gmg_last_cuda_error` give detail.  No exceptions, no exit().
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

/* Haversine (asin form) distance matrix rows in km, R = 6371.0.
 * replaces `distances.haversine_distance_from_frame` (distances.py:191-218). */
int gmg_haversine_f64(const double* lat_rad, const double* lon_rad, int64_t N, int64_t row0,
                      int64_t nrows, double* D_rows, int64_t ld, gmg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GLOMAR_B200_H */
