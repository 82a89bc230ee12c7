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

struct VarioParams {
    int model;
    double psill, nugget, range, x_scale;
    GmgBesselConsts bc;
};

__device__ __forceinline__ double variogram_value(const VarioParams& q, double d, bool fill_nugget) {
    double v;
    switch (q.model) {
        case GMG_VARIOGRAM_SPHERICAL: {           // :122-138
            const double h = d / q.range;
            v = 0.5 * q.psill * (3.0 * h - h * h * h) + q.nugget;
            if (d >= q.range) v = q.nugget + q.psill;
            break;
        }
        case GMG_VARIOGRAM_GAUSSIAN:              // :203-215
            v = q.psill * (1.0 - exp(-((d * d) / (q.range * q.range)))) + q.nugget;
            break;
        case GMG_VARIOGRAM_EXPONENTIAL:           // :283-286
            v = q.psill * (1.0 - exp(-(d / q.range))) + q.nugget;
            break;
        default: {                                // Matern :457-469: left * middle * right = x^nu K_nu(x) / (Gamma(nu) 2^(nu-1))
            const double x = q.x_scale * (d / q.range);
            double cor;
            if (q.bc.nu == 0.5) cor = exp(-x);
            else if (q.bc.nu == 1.5) cor = (1.0 + x) * exp(-x);
            else if (q.bc.nu == 2.5) cor = fma(x, fma(x, 1.0 / 3.0, 1.0), 1.0) * exp(-x);
            else cor = gmg_matern_scaled(q.bc, x) * exp(-x);
            if (!(x > 0.0)) cor = NAN;            // K_nu(0) = inf -> `_kv` maps it to NaN (:29-30)
            v = q.psill * (1.0 - cor) + q.nugget;
            if (fill_nugget) v = q.nugget;        // np.fill_diagonal(out, nugget) (:470-475)
            break;
        }
    }
    return v;
}

__global__ void __launch_bounds__(256) variogram_kernel(VarioParams q, const double* __restrict__ dist, int64_t rows,
                                                        int64_t cols, int64_t ld, int fill_diag, int to_cov,
                                                        double variance, double* __restrict__ out, int64_t ldo) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= cols || i >= rows) return;
    const double v = variogram_value(q, dist[i * ld + j], fill_diag && i == j);
    out[i * ldo + j] = to_cov ? variance - v : v;
}

// The whole stationary route in one pass ("fused"): haversine distance between cells i and j, the
// variogram model, variance - variogram - N^2 written once, no distance matrix.  Every step is the arithmetic of the
// three separate kernels (haversine_kernel in ellipse_cov.cu, variogram_kernel above), so the result is bit-identical
// to theirs; the distance and hence the covariance are symmetric bit for bit (sin(-x) = -sin x, products commute), so
// only tiles on or above the diagonal are evaluated and each is stored twice, the mirror through shared memory so
// that both stores are coalesced.  32 x 32 tiles, 32 x 8 threads.  grid = (T, T), tiles below the diagonal exit.
constexpr int ST = 32;
__global__ void __launch_bounds__(256) stationary_cov_kernel(VarioParams q, const double* __restrict__ lat,
                                                             const double* __restrict__ lon, int64_t N, double radius,
                                                             int fill_diag, int to_cov, double variance,
                                                             double* __restrict__ out, int64_t ldo) {
    if (blockIdx.y > blockIdx.x) return;
    __shared__ double tile[ST][ST + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t i0 = (int64_t)blockIdx.y * ST, j0 = (int64_t)blockIdx.x * ST;
    const int64_t j = j0 + tx;
    const double latj = j < N ? lat[j] : 0.0, lonj = j < N ? lon[j] : 0.0;
    const double cj = cos(latj);
#pragma unroll
    for (int u = 0; u < ST / 8; ++u) {
        const int64_t i = i0 + ty + 8 * u;
        if (i < N && j < N) {
            const double sl = sin(0.5 * (lat[i] - latj)), so = sin(0.5 * (lon[i] - lonj));
            const double h = sl * sl + cos(lat[i]) * cj * so * so;
            const double d = (2.0 * asin(sqrt(h))) * radius;
            const double v = variogram_value(q, d, fill_diag && i == j);
            const double c = to_cov ? variance - v : v;
            out[i * ldo + j] = c;
            tile[ty + 8 * u][tx] = c;
        }
    }
    if (blockIdx.y == blockIdx.x) return;
    __syncthreads();
    const int64_t im = i0 + tx;                    // mirror: out[j][i], lanes along i
#pragma unroll
    for (int u = 0; u < ST / 8; ++u) {
        const int64_t jm = j0 + ty + 8 * u;
        if (im < N && jm < N) out[jm * ldo + im] = tile[tx][ty + 8 * u];
    }
}

VarioParams make_params(int model, double psill, double nugget, double range, double nu, int matern_method) {
    VarioParams q;
    q.model = model;
    q.psill = psill;
    q.nugget = nugget;
    q.range = range;
    q.bc = gmg_make_bessel_consts(model == GMG_VARIOGRAM_MATERN ? nu : 0.5);
    // "sklearn": sqrt(2 nu) d / range;  "gstat": d / range;  "karspeck": 2 sqrt(nu) d / range   (:390-426)
    q.x_scale = matern_method == 0 ? sqrt(2.0 * nu) : (matern_method == 1 ? 1.0 : 2.0 * sqrt(nu));
    return q;
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
    const VarioParams q = make_params(model, psill, nugget, range, nu, matern_method);
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
        dim3 grid((unsigned)gmg_ceil_div(cols, 256), (unsigned)nr);
        // the diagonal test inside the kernel uses the row index relative to the first row of the matrix
        variogram_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q, dist + r0 * ld, nr, cols, ld, r0 == 0 ? fill_diag : 0,
                                                                  to_covariance, variance, out + r0 * ldo, ldo);
        GMG_LAUNCH_CHECK();
        if (fill_diag && r0 > 0) return GMG_ERR_UNSUPPORTED;   // diagonal fill only for matrices of < 65536 rows per call
    }
    return GMG_OK;
}

extern "C" int gmg_stationary_cov_f64(int model, const double* lat_rad, const double* lon_rad, int64_t N, double radius,
                                      double psill, double nugget, double range, double nu, int matern_method,
                                      int fill_diag, int to_covariance, double variance, double* out, int64_t ldo,
                                      gmg_stream_t stream) {
    if (model < GMG_VARIOGRAM_SPHERICAL || model > GMG_VARIOGRAM_MATERN) return -1;
    if (!lat_rad) return -2;
    if (!lon_rad) return -3;
    if (N <= 0 || N > 65535LL * ST) return -4;
    if (model == GMG_VARIOGRAM_MATERN && !(nu > 0.0)) return -9;
    if (matern_method < 0 || matern_method > 2) return -10;
    if (!out) return -14;
    if (ldo < N) return -15;
    const VarioParams q = make_params(model, psill, nugget, range, nu, matern_method);
    const unsigned T = (unsigned)gmg_ceil_div(N, ST);
    stationary_cov_kernel<<<dim3(T, T), 256, 0, (cudaStream_t)stream>>>(q, lat_rad, lon_rad, N, radius, fill_diag,
                                                                        to_covariance, variance, out, ldo);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}
