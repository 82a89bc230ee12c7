// Host build of glomargridding_b200/csrc/bessel_k.cuh for CPU-side tests (no GPU in the build
// container): the identical __host__ __device__ code is compiled with g++ and compared to scipy.
#include "../glomargridding_b200/csrc/bessel_k.cuh"
extern "C" {
void host_bessel_kve(double nu, const double* x, long n, double* out) {
    GmgBesselConsts c = gmg_make_bessel_consts(nu);
    for (long i = 0; i < n; ++i) out[i] = gmg_bessel_kve(c, x[i]);
}
void host_matern_scaled(double nu, const double* x, long n, double* out) {
    GmgBesselConsts c = gmg_make_bessel_consts(nu);
    for (long i = 0; i < n; ++i) out[i] = gmg_matern_scaled(c, x[i]);
}
}
