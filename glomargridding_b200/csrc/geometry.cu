// Elementwise geometry helpers of the ellipse covariance path as stand-alone entry points - the module-level
// functions the reference keeps next to its builder and that its tests call directly (test/test_matrix.py:74-140).
// The tile kernel (ellipse_cov.cu) fuses the same arithmetic into the pair loop; these kernels exist so that the
// helpers are drop-ins too and so that the fused code can be checked piece by piece.
//
// Replaces (reference, paths relative to its root):
//   glomar_gridding/ellipse/covariance.py:512-533  _sigma_rot_func_multi   Sigma = R(theta) diag(Lx^2, Ly^2) R(theta)^T
//   glomar_gridding/ellipse/covariance.py:536-547  _det_22_single / _det_22_multi
//   glomar_gridding/ellipse/covariance.py:550-588  _haversine_single / _haversine_multi  (pairs of points, radians)
//   glomar_gridding/ellipse/covariance.py:591-654  _mo_disp_* / _mod_mo_disp_*           (pairs of points, radians)
// All four are HBM-bound streams (a few loads, a few stores per element); one thread per element, grid-stride.
#include "gmg_common.cuh"
#include "../../include/glomar_b200.h"

namespace {

constexpr double kPi = 3.141592653589793;        // np.pi
constexpr double kTwoPi = 6.283185307179586;     // 2 * np.pi (ellipse/covariance.py:37)

inline unsigned geometry_grid(int64_t n) {
    const int64_t blocks = gmg_ceil_div(n, 256);
    return (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
}

__global__ void __launch_bounds__(256) sigma_rot_kernel(const double* __restrict__ Lx, const double* __restrict__ Ly,
                                                        const double* __restrict__ theta, int64_t n,
                                                        double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double st, ct;
        sincos(theta[i], &st, &ct);
        // separate roundings like the numpy expression (:519-531): no FMA contraction
        const double c2 = __dmul_rn(ct, ct), s2 = __dmul_rn(st, st), cs = __dmul_rn(ct, st);
        const double lx2 = __dmul_rn(Lx[i], Lx[i]), ly2 = __dmul_rn(Ly[i], Ly[i]);
        const double off = __dmul_rn(cs, __dsub_rn(lx2, ly2));
        double2* o = reinterpret_cast<double2*>(out + 4 * i);
        o[0] = make_double2(__dadd_rn(__dmul_rn(c2, lx2), __dmul_rn(s2, ly2)), off);
        o[1] = make_double2(off, __dadd_rn(__dmul_rn(s2, lx2), __dmul_rn(c2, ly2)));
    }
}

__global__ void __launch_bounds__(256) det22_kernel(const double* __restrict__ mats, int64_t n,
                                                    double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double2* m = reinterpret_cast<const double2*>(mats + 4 * i);
        const double2 a = m[0], b = m[1];
        out[i] = __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x));      // :547
    }
}

__global__ void __launch_bounds__(256) pair_haversine_kernel(const double* __restrict__ lat0,
                                                             const double* __restrict__ lon0,
                                                             const double* __restrict__ lat1,
                                                             const double* __restrict__ lon1, int64_t n, double radius,
                                                             double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double dlon = lon0[i] - lon1[i], dlat = lat0[i] - lat1[i];   // :578-579
        const double sa = sin(0.5 * dlat), so = sin(0.5 * dlon);
        const double a = __dadd_rn(__dmul_rn(sa, sa), __dmul_rn(__dmul_rn(cos(lat0[i]), cos(lat1[i])), __dmul_rn(so, so)));
        out[i] = radius * (2.0 * asin(sqrt(a)));                           // :585-588
    }
}

__global__ void __launch_bounds__(256) pair_disp_kernel(const double* __restrict__ lat0, const double* __restrict__ lon0,
                                                        const double* __restrict__ lat1, const double* __restrict__ lon1,
                                                        int64_t n, int modified, double radius, double* __restrict__ dy,
                                                        double* __restrict__ dx) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double y = lat0[i] - lat1[i];
        double x = lon0[i] - lon1[i];
        if (x > kPi) x -= kTwoPi;                                          // strict comparisons, :633-634
        if (x < -kPi) x += kTwoPi;
        if (modified) x = __dmul_rn(x, 0.5 * (cos(lat0[i]) + cos(lat1[i])));   // :636-637
        dy[i] = radius * y;                                                // :639
        dx[i] = radius * x;
    }
}

}  // namespace

extern "C" {

int gmg_sigma_rot_f64(const double* Lx, const double* Ly, const double* theta, int64_t n, double* sigmas,
                      gmg_stream_t stream) {
    if (!Lx) return -1;
    if (!Ly) return -2;
    if (!theta) return -3;
    if (n < 0) return -4;
    if (!sigmas || (reinterpret_cast<uintptr_t>(sigmas) & 15)) return -5;
    if (n == 0) return GMG_OK;
    sigma_rot_kernel<<<geometry_grid(n), 256, 0, (cudaStream_t)stream>>>(Lx, Ly, theta, n, sigmas);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_det22_f64(const double* mats, int64_t n, double* dets, gmg_stream_t stream) {
    if (!mats || (reinterpret_cast<uintptr_t>(mats) & 15)) return -1;
    if (n < 0) return -2;
    if (!dets) return -3;
    if (n == 0) return GMG_OK;
    det22_kernel<<<geometry_grid(n), 256, 0, (cudaStream_t)stream>>>(mats, n, dets);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_pair_haversine_f64(const double* lat0, const double* lon0, const double* lat1, const double* lon1, int64_t n,
                           double radius_km, double* dist, gmg_stream_t stream) {
    if (!lat0) return -1;
    if (!lon0) return -2;
    if (!lat1) return -3;
    if (!lon1) return -4;
    if (n < 0) return -5;
    if (!dist) return -7;
    if (n == 0) return GMG_OK;
    pair_haversine_kernel<<<geometry_grid(n), 256, 0, (cudaStream_t)stream>>>(lat0, lon0, lat1, lon1, n, radius_km, dist);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_pair_displacement_f64(const double* lat0, const double* lon0, const double* lat1, const double* lon1,
                              int64_t n, int modified, double radius_km, double* dy, double* dx, gmg_stream_t stream) {
    if (!lat0) return -1;
    if (!lon0) return -2;
    if (!lat1) return -3;
    if (!lon1) return -4;
    if (n < 0) return -5;
    if (modified != 0 && modified != 1) return -6;
    if (!dy) return -8;
    if (!dx) return -9;
    if (n == 0) return GMG_OK;
    pair_disp_kernel<<<geometry_grid(n), 256, 0, (cudaStream_t)stream>>>(lat0, lon0, lat1, lon1, n, modified, radius_km,
                                                                         dy, dx);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

}  // extern "C"

// ---- observation -> grid index (map_to_grid / find_nearest) ---------------------------------------------------------
// replaces `utils.find_nearest` (glomar_gridding/utils.py:178-214: `int(np.argmin(np.abs(array - value)))` per value,
// FIRST minimum on ties, NaN differences win like numpy's argmin) as used by `grid.map_to_grid`
// (glomar_gridding/grid.py:28-124) once per coordinate, and the row-major `np.ravel_multi_index` that follows.
namespace {

__global__ void __launch_bounds__(256) nearest_index_kernel(const double* __restrict__ grid, int64_t g,
                                                            const double* __restrict__ values, int64_t n,
                                                            int64_t* __restrict__ idx, int64_t stride, int accumulate) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double v = values[k];
    int64_t best = 0;
    double bd = fabs(grid[0] - v);
    if (bd == bd) {                                   // a NaN difference is numpy's argmin at its first occurrence
        for (int64_t i = 1; i < g; ++i) {
            const double d = fabs(grid[i] - v);
            if (d != d) { best = i; break; }
            if (d < bd) { bd = d; best = i; }
        }
    }
    idx[k] = (accumulate ? idx[k] : 0) + best * stride;
}

}  // namespace

extern "C" int gmg_nearest_index_f64(const double* grid, int64_t g, const double* values, int64_t n, int64_t* idx,
                                     int64_t stride, int accumulate, gmg_stream_t stream) {
    if (!grid) return -1;
    if (g <= 0) return -2;
    if (!values) return -3;
    if (n < 0) return -4;
    if (!idx) return -5;
    if (n == 0) return GMG_OK;
    nearest_index_kernel<<<(unsigned)gmg_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(grid, g, values, n, idx,
                                                                                            stride, accumulate);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}
