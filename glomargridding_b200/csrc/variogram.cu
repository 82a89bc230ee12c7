// N3: stationary variogram models evaluated on a distance matrix - the other covariance route of
// the reference's workflow (notebook cells 6-7: grid_to_distance_matrix -> MaternVariogram.fit -> covariance).
//
// Replaces (reference, paths relative to its root):
//   glomar_gridding/variogram.py:96-138   SphericalVariogram.fit
//   glomar_gridding/variogram.py:177-218  GaussianVariogram.fit
//   glomar_gridding/variogram.py:257-289  ExponentialVariogram.fit
//   glomar_gridding/variogram.py:387-476  MaternVariogram._left/_middle/_right/fit (scipy.special.kv, gamma)
//   glomar_gridding/variogram.py:479-504  variogram_to_covariance (variance - variogram, fused when asked)
// Elementwise over N^2 entries: HBM bound (8 B read + 8 B written per entry) except for general-nu Matern, which
// evaluates K_nu per entry with the Temme / CF2 routine of bessel_k.cuh.
#include "gmg_common.cuh"
#include "bessel_k.cuh"
#include "../../include/glomar_b200.h"

namespace {

__global__ void __launch_bounds__(256) variogram_kernel(int model, const double* __restrict__ dist, int64_t rows,
                                                        int64_t cols, int64_t ld, double psill, double nugget,
                                                        double range, double x_scale, GmgBesselConsts bc,
                                                        int fill_diag, int to_cov, double variance,
                                                        double* __restrict__ out, int64_t ldo) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= cols || i >= rows) return;
    const double d = dist[i * ld + j];
    double v;
    switch (model) {
        case GMG_VARIOGRAM_SPHERICAL: {           // :122-138
            const double h = d / range;
            v = 0.5 * psill * (3.0 * h - h * h * h) + nugget;
            if (d >= range) v = nugget + psill;
            break;
        }
        case GMG_VARIOGRAM_GAUSSIAN:              // :203-215
            v = psill * (1.0 - exp(-((d * d) / (range * range)))) + nugget;
            break;
        case GMG_VARIOGRAM_EXPONENTIAL:           // :283-286
            v = psill * (1.0 - exp(-(d / range))) + nugget;
            break;
        default: {                                // Matern :457-469: left * middle * right = x^nu K_nu(x) / (Gamma(nu) 2^(nu-1))
            const double x = x_scale * (d / range);
            double cor;
            if (bc.nu == 0.5) cor = exp(-x);
            else if (bc.nu == 1.5) cor = (1.0 + x) * exp(-x);
            else if (bc.nu == 2.5) cor = fma(x, fma(x, 1.0 / 3.0, 1.0), 1.0) * exp(-x);
            else cor = gmg_matern_scaled(bc, x) * exp(-x);
            if (!(x > 0.0)) cor = NAN;            // K_nu(0) = inf -> `_kv` maps it to NaN (:29-30)
            v = psill * (1.0 - cor) + nugget;
            if (fill_diag && i == j) v = nugget;  // np.fill_diagonal(out, nugget) (:470-475)
            break;
        }
    }
    out[i * ldo + j] = to_cov ? variance - v : v;
}

}  // namespace

extern "C" int gmg_variogram_f64(int model, const double* dist, int64_t rows, int64_t cols, int64_t ld, double psill,
                                 double nugget, double range, double nu, int matern_method, int fill_diag,
                                 int to_covariance, double variance, double* out, int64_t ldo, gmg_stream_t stream) {
    if (model < GMG_VARIOGRAM_SPHERICAL || model > GMG_VARIOGRAM_MATERN) return -1;
    if (!dist) return -2;
    if (rows <= 0 || rows > 65535 * 32768LL) return -3;
    if (cols <= 0) return -4;
    if (ld < cols) return -5;
    if (model == GMG_VARIOGRAM_MATERN && !(nu > 0.0)) return -9;
    if (matern_method < 0 || matern_method > 2) return -10;
    if (!out) return -14;
    if (ldo < cols) return -15;
    const GmgBesselConsts bc = gmg_make_bessel_consts(model == GMG_VARIOGRAM_MATERN ? nu : 0.5);
    // "sklearn": sqrt(2 nu) d / range;  "gstat": d / range;  "karspeck": 2 sqrt(nu) d / range   (:390-426)
    const double x_scale = matern_method == 0 ? sqrt(2.0 * nu) : (matern_method == 1 ? 1.0 : 2.0 * sqrt(nu));
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
        dim3 grid((unsigned)gmg_ceil_div(cols, 256), (unsigned)nr);
        // the diagonal test inside the kernel uses the row index relative to the first row of the matrix
        variogram_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(model, dist + r0 * ld - 0, nr, cols, ld, psill, nugget,
                                                                  range, x_scale, bc, r0 == 0 ? fill_diag : 0,
                                                                  to_covariance, variance, out + r0 * ldo, ldo);
        GMG_LAUNCH_CHECK();
        if (fill_diag && r0 > 0) return GMG_ERR_UNSUPPORTED;   // diagonal fill only for matrices of < 65536 rows per call
    }
    return GMG_OK;
}
