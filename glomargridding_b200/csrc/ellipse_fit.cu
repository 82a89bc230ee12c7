// N4: maximum-likelihood fit of the local ellipse (Matern) parameters - the step that PRODUCES the Lx / Ly / theta /
// stdev fields the covariance builder consumes - for every grid cell at once.
//
// Replaces (reference, paths relative to its root):
//   glomar_gridding/ellipse/model.py:582-686    cov_ij_anisotropic / cov_ij_isotropic (Matern correlation of a displacement)
//   glomar_gridding/ellipse/model.py:266-384    EllipseModel.negative_log_likelihood (Gaussian likelihood of the
//                                               Fisher-transformed correlations, 0.999999 clamp)
//   glomar_gridding/ellipse/model.py:402-500    EllipseModel.fit: scipy.optimize.minimize(method="Nelder-Mead", bounds=..)
//   glomar_gridding/ellipse/estimate.py:515-715 the Python loop of compute_params over the unmasked cells (58 s in the
//                                               reference's notebook)
// The optimiser is a restatement of the Nelder-Mead simplex method as scipy runs it (scipy/optimize/_optimize.py,
// `_minimize_neldermead`, scipy 1.16 pinned by the reference): initial simplex x0 (1 + 0.05 e_k), reflection /
// expansion / contraction / shrink with rho = 1, chi = 2, psi = sigma = 1/2, clipping to the bounds, the
// reflect-into-the-interior rule for an initial simplex beyond an upper bound, termination on xatol AND fatol, N * 200
// iterations or function evaluations (an evaluation refused by the budget aborts the iteration exactly like scipy's
// _MaxFuncCallError does).  Every decision compares the same quantities in the same order, so that the simplex follows
// the same path and stops at the same vertex as the reference - the fitted parameters agree to rounding, not merely
// to the termination tolerance.
//
// One CTA (128 threads) per cell: the likelihood is a sum over the cell's ~10^2..10^4 training pairs, evaluated by all
// threads (strided, fixed order, tree reduction in shared memory: bit-reproducible run to run); the simplex logic is
// executed redundantly by every thread on CTA-uniform values.
#include "gmg_common.cuh"
#include "bessel_k.cuh"
#include "../../include/glomar_b200.h"

namespace {

constexpr int kFitThreads = 128;
constexpr int kMaxP = 4;                       // Lx, Ly, theta, sigma
constexpr double kThr = 0.999999;              // model.py:330
constexpr double kNormLogC = 0.9189385332046727;   // log(sqrt(2 pi)), scipy.stats._continuous_distns._norm_pdf_logC

struct FitArgs {
    const int64_t* offs;     // [B + 1] CSR offsets of the training pairs of each problem
    const double* x0;        // anisotropic: east-west displacement; isotropic: distance
    const double* x1;        // anisotropic: north-south displacement; isotropic: unused (may be NULL)
    const double* y;         // observed correlations
    double* z;               // scratch [total]: arctanh of the clamped observed correlations
    double* yll;             // scratch [total]: model correlations of the current evaluation
    int n_params;            // 1 isotropic, 2 anisotropic, 3 anisotropic + rotation
    int unit_sigma;          // 1: the likelihood's sigma is 1, else it is the last fitted parameter
    int arctanh;             // Fisher transform (model.py:323-372); 0: plain Gaussian likelihood (:320-322)
    GmgBesselConsts bc;
    double guess[kMaxP], lo[kMaxP], hi[kMaxP];
    double xatol, fatol;
    int maxiter, maxfun;     // <= 0: scipy's default N * 200
};

__device__ __forceinline__ double block_sum(double v, double* red) {
    const int t = threadIdx.x;
    red[t] = v;
    __syncthreads();
    for (int s = kFitThreads / 2; s > 0; s >>= 1) {
        if (t < s) red[t] += red[t + s];
        __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
}

// max |v| with numpy semantics: NaN if any element is NaN
__device__ __forceinline__ double block_absmax(double v, bool nan, double* red) {
    const int t = threadIdx.x;
    red[t] = nan ? NAN : v;
    __syncthreads();
    for (int s = kFitThreads / 2; s > 0; s >>= 1) {
        if (t < s) {
            const double a = red[t], b = red[t + s];
            red[t] = (a != a || b != b) ? NAN : fmax(a, b);
        }
        __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
}

// Matern correlation of one displacement for parameters p (model.py:582-686, distances.py:41-58, :328-353, :356-474)
__device__ __forceinline__ double model_corr(const FitArgs& a, const double* p, double dx, double dy) {
    double tau;
    if (a.n_params == 1) {
        tau = fabs(dx) / p[0];                                   // :683  |delta| / R
    } else {
        double s00, s01, s10, s11;
        const double lx2 = p[0] * p[0], ly2 = p[1] * p[1];       // Lx ** 2.0, Ly ** 2.0
        if (a.n_params == 3) {
            double sn, cs;
            sincos(p[2], &sn, &cs);
            // R diag(Lx^2, Ly^2) R^T with R = [[c, -s], [s, c]], products in the order numpy's 2 x 2 matmul forms them
            const double r00 = cs * lx2, r01 = -sn * ly2, r10 = sn * lx2, r11 = cs * ly2;
            s00 = __dadd_rn(__dmul_rn(r00, cs), __dmul_rn(r01, -sn));
            s01 = __dadd_rn(__dmul_rn(r00, sn), __dmul_rn(r01, cs));
            s10 = __dadd_rn(__dmul_rn(r10, cs), __dmul_rn(r11, -sn));
            s11 = __dadd_rn(__dmul_rn(r10, sn), __dmul_rn(r11, cs));
        } else {
            s00 = lx2; s01 = 0.0; s10 = 0.0; s11 = ly2;
        }
        const double det = __dsub_rn(__dmul_rn(s00, s11), __dmul_rn(s01, s10));
        const double i00 = s11 / det, i01 = -s01 / det, i10 = -s10 / det, i11 = s00 / det;
        tau = sqrt(__dadd_rn(__dmul_rn(dx, __dadd_rn(__dmul_rn(dx, i00), __dmul_rn(dy, i01))),
                             __dmul_rn(dy, __dadd_rn(__dmul_rn(dx, i10), __dmul_rn(dy, i11)))));
    }
    const double inner = 2.0 * tau * sqrt(a.bc.nu);              // :627, :684
    return gmg_matern_scaled(a.bc, inner) * exp(-inner);         // first * inner^v * kv(v, inner), stdev = 1
}

// EllipseModel.negative_log_likelihood(X, y, params) for one problem; CTA-collective (model.py:266-384)
__device__ double nll(const FitArgs& a, int64_t i0, int64_t i1, const double* p, double* red) {
    const double sigma = a.unit_sigma ? 1.0 : p[a.n_params];
    double mx = 0.0;
    bool nan = false;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += kFitThreads) {
        const double v = model_corr(a, p, a.x0[i], a.x1 ? a.x1[i] : 0.0);
        a.yll[i] = v;
        if (v != v) nan = true;
        mx = fmax(mx, fabs(v));
    }
    double acc = 0.0;
    const double logs = log(sigma);
    if (!a.arctanh) {
        __syncthreads();
        for (int64_t i = i0 + threadIdx.x; i < i1; i += kFitThreads) {
            const double t = (a.y[i] - a.yll[i]) / sigma;
            acc += __dsub_rn(__dsub_rn(-__dmul_rn(t, t) / 2.0, kNormLogC), logs);
        }
        return -1.0 * block_sum(acc, red);
    }
    const double max_yll = block_absmax(mx, nan, red);            // also orders the yll stores before the reads below
    const bool clamp = max_yll >= 1.0;                            // :349 (the TEST is against 1, the clamp to the threshold)
    for (int64_t i = i0 + threadIdx.x; i < i1; i += kFitThreads) {
        double v = a.yll[i];
        if (clamp && fabs(v) > kThr) v = copysign(kThr, v);
        const double t = (a.z[i] - atanh(v)) / sigma;
        acc += __dsub_rn(__dsub_rn(-__dmul_rn(t, t) / 2.0, kNormLogC), logs);   // scipy norm.logpdf: _norm_logpdf((x - loc) / scale) - log(scale)
    }
    return -1.0 * block_sum(acc, red);
}

__device__ __forceinline__ double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// stable insertion sort of the simplex by function value, NaN last (numpy argsort of <= 16 elements)
__device__ __forceinline__ void sort_simplex(double sim[kMaxP + 1][kMaxP], double* fsim, int N) {
    for (int i = 1; i <= N; ++i) {
        const double f = fsim[i];
        double x[kMaxP];
        for (int k = 0; k < N; ++k) x[k] = sim[i][k];
        int j = i - 1;
        // numpy's order: a < b, with NaN sorted to the end
        while (j >= 0 && ((f < fsim[j]) || (fsim[j] != fsim[j] && f == f))) {
            fsim[j + 1] = fsim[j];
            for (int k = 0; k < N; ++k) sim[j + 1][k] = sim[j][k];
            --j;
        }
        fsim[j + 1] = f;
        for (int k = 0; k < N; ++k) sim[j + 1][k] = x[k];
    }
}

__global__ void __launch_bounds__(kFitThreads) ellipse_fit_kernel(const FitArgs a, int only_nll, const double* params_in,
                                                                  double* x_out, double* fun_out, int* nit_out,
                                                                  int* nfev_out, int* status_out) {
    __shared__ double red[kFitThreads];
    const int b = blockIdx.x;
    const int64_t i0 = a.offs[b], i1 = a.offs[b + 1];
    const int N = a.n_params + (a.unit_sigma ? 0 : 1);
    // ---- observed correlations: clamp (in place in the reference, :339-346) and Fisher-transform once -------------
    if (a.arctanh) {
        double mx = 0.0;
        bool nan = false;
        for (int64_t i = i0 + threadIdx.x; i < i1; i += kFitThreads) {
            const double v = a.y[i];
            if (v != v) nan = true;
            mx = fmax(mx, fabs(v));
        }
        const bool clamp = block_absmax(mx, nan, red) >= kThr;
        for (int64_t i = i0 + threadIdx.x; i < i1; i += kFitThreads) {
            double v = a.y[i];
            if (clamp && fabs(v) > kThr) v = copysign(kThr, v);
            a.z[i] = atanh(v);
        }
        __syncthreads();
    }
    if (only_nll) {                              // likelihood of given parameters (parity tests, EllipseModel API)
        double p[kMaxP];
        for (int k = 0; k < N; ++k) p[k] = params_in[(int64_t)b * N + k];
        const double f = nll(a, i0, i1, p, red);
        if (threadIdx.x == 0) fun_out[b] = f;
        return;
    }
    // ---- scipy _minimize_neldermead ------------------------------------------------------------------------------------
    const double rho = 1.0, chi = 2.0, psi = 0.5, sg = 0.5;
    const int maxiter = a.maxiter > 0 ? a.maxiter : N * 200;
    const int maxfun = a.maxfun > 0 ? a.maxfun : N * 200;
    double sim[kMaxP + 1][kMaxP], fsim[kMaxP + 1];
    double x0[kMaxP];
    for (int k = 0; k < N; ++k) x0[k] = clipd(a.guess[k], a.lo[k], a.hi[k]);
    for (int j = 0; j <= N; ++j) {
        for (int k = 0; k < N; ++k) sim[j][k] = x0[k];
        if (j > 0) sim[j][j - 1] = (x0[j - 1] != 0.0) ? (1.0 + 0.05) * x0[j - 1] : 0.00025;
        for (int k = 0; k < N; ++k) {
            double v = sim[j][k];
            if (v > a.hi[k]) v = __dsub_rn(2.0 * a.hi[k], v);      // reflect into the interior
            sim[j][k] = clipd(v, a.lo[k], a.hi[k]);
        }
        fsim[j] = INFINITY;
    }
    int fcalls = 0;
    bool refused = false;                        // scipy's _MaxFuncCallError
    auto func = [&](const double* p) -> double {
        if (fcalls >= maxfun) { refused = true; return 0.0; }
        ++fcalls;
        return nll(a, i0, i1, p, red);
    };
    for (int j = 0; j <= N && !refused; ++j) {
        const double f = func(sim[j]);
        if (!refused) fsim[j] = f;
    }
    sort_simplex(sim, fsim, N);
    int iterations = 1;
    while (fcalls < maxfun && iterations < maxiter) {
        refused = false;
        double dxm = 0.0, dfm = 0.0;
        bool nanx = false, nanf = false;
        for (int j = 1; j <= N; ++j) {
            for (int k = 0; k < N; ++k) {
                const double d = fabs(sim[j][k] - sim[0][k]);
                if (d != d) nanx = true;
                dxm = fmax(dxm, d);
            }
            const double d = fabs(fsim[0] - fsim[j]);
            if (d != d) nanf = true;
            dfm = fmax(dfm, d);
        }
        if (!nanx && !nanf && dxm <= a.xatol && dfm <= a.fatol) break;
        double xbar[kMaxP], xr[kMaxP], xt[kMaxP];
        for (int k = 0; k < N; ++k) {
            double s = sim[0][k];
            for (int j = 1; j < N; ++j) s += sim[j][k];          // np.add.reduce(sim[:-1], 0)
            xbar[k] = s / N;
            xr[k] = clipd(__dsub_rn(__dmul_rn(1.0 + rho, xbar[k]), __dmul_rn(rho, sim[N][k])), a.lo[k], a.hi[k]);
        }
        do {                                     // one iteration; `break` = the exception path of scipy's try block
            const double fxr = func(xr);
            if (refused) break;
            bool doshrink = false;
            if (fxr < fsim[0]) {
                for (int k = 0; k < N; ++k)
                    xt[k] = clipd(__dsub_rn(__dmul_rn(1.0 + rho * chi, xbar[k]), __dmul_rn(rho * chi, sim[N][k])), a.lo[k], a.hi[k]);
                const double fxe = func(xt);
                if (refused) break;
                if (fxe < fxr) {
                    for (int k = 0; k < N; ++k) sim[N][k] = xt[k];
                    fsim[N] = fxe;
                } else {
                    for (int k = 0; k < N; ++k) sim[N][k] = xr[k];
                    fsim[N] = fxr;
                }
            } else if (fxr < fsim[N - 1]) {
                for (int k = 0; k < N; ++k) sim[N][k] = xr[k];
                fsim[N] = fxr;
            } else if (fxr < fsim[N]) {          // contraction
                for (int k = 0; k < N; ++k)
                    xt[k] = clipd(__dsub_rn(__dmul_rn(1.0 + psi * rho, xbar[k]), __dmul_rn(psi * rho, sim[N][k])), a.lo[k], a.hi[k]);
                const double fxc = func(xt);
                if (refused) break;
                if (fxc <= fxr) {
                    for (int k = 0; k < N; ++k) sim[N][k] = xt[k];
                    fsim[N] = fxc;
                } else {
                    doshrink = true;
                }
            } else {                             // inside contraction
                for (int k = 0; k < N; ++k)
                    xt[k] = clipd(__dadd_rn(__dmul_rn(1.0 - psi, xbar[k]), __dmul_rn(psi, sim[N][k])), a.lo[k], a.hi[k]);
                const double fxcc = func(xt);
                if (refused) break;
                if (fxcc < fsim[N]) {
                    for (int k = 0; k < N; ++k) sim[N][k] = xt[k];
                    fsim[N] = fxcc;
                } else {
                    doshrink = true;
                }
            }
            if (doshrink) {
                for (int j = 1; j <= N; ++j) {
                    for (int k = 0; k < N; ++k)
                        sim[j][k] = clipd(__dadd_rn(sim[0][k], __dmul_rn(sg, sim[j][k] - sim[0][k])), a.lo[k], a.hi[k]);
                    const double f = func(sim[j]);
                    if (refused) break;
                    fsim[j] = f;
                }
                if (refused) break;
            }
            ++iterations;
        } while (false);
        sort_simplex(sim, fsim, N);
    }
    if (threadIdx.x == 0) {
        double fmin_ = fsim[0];
        for (int j = 1; j <= N; ++j) fmin_ = fmin(fmin_, fsim[j]);
        for (int k = 0; k < N; ++k) x_out[(int64_t)b * N + k] = sim[0][k];
        fun_out[b] = fmin_;
        nit_out[b] = iterations;
        nfev_out[b] = fcalls;
        status_out[b] = fcalls >= maxfun ? 1 : (iterations >= maxiter ? 2 : 0);
    }
}


// ---- training data of every cell (EllipseBuilder._get_train_data, estimate.py:441-513) --------------------------------
// For cell c the reference takes row c of the correlation matrix and every other unmasked cell j whose distance to c
// lies in [min_distance, max_distance] (and is not 0): the distance is either the haversine distance in km
// (`physical_distance_selection` with a displacement method) or the Euclidean norm of the wrapped degree differences;
// the regressors are the displacements of distances.displacements (:477-560: wrap at +-180 degrees, radians, mean
// cos-latitude scaling for "Modified_Met_Office") - times the Earth radius for the physical-distance forms - or the
// distance itself for the isotropic forms.  Element-wise work on (cell, j) pairs: one CTA per cell, first a count,
// then - after an exclusive scan of the counts by the caller - an ORDER-PRESERVING fill (j ascending, like np.where).
struct TrainArgs {
    const double* lat;       // [N] degrees, unmasked cells in row-major order
    const double* lon;
    const double* cor;       // [N x ldc] correlation matrix
    int64_t N, ldc;
    const int64_t* cells;    // [n_cells] cells to produce (indices into the unmasked cells)
    int anisotropic, physical_distance, disp_method /* -1 none, 0 Met_Office, 1 Modified_Met_Office */, pd_selection;
    double min_d, max_d;
};

constexpr double kDegToRad = 0.017453292519943295;     // np.deg2rad / np.radians multiply by pi / 180

__device__ __forceinline__ bool train_entry(const TrainArgs& a, int64_t c, int64_t j, double& x0, double& x1) {
    const double latc = a.lat[c], lonc = a.lon[c], latj = a.lat[j], lonj = a.lon[j];
    // displacements(lats_all, lons_all, lat_c, lon_c): (all) - (cell), wrap with strict comparisons
    double dy_deg = latj - latc, dx_deg = lonj - lonc;
    if (dx_deg > 180.0) dx_deg -= 360.0;
    if (dx_deg < -180.0) dx_deg += 360.0;
    double dy = dy_deg, dx = dx_deg;
    if (a.disp_method >= 0) {
        dy = dy_deg * kDegToRad;
        dx = dx_deg * kDegToRad;
        if (a.disp_method == 1) dx = dx * (0.5 * (cos(latj * kDegToRad) + cos(latc * kDegToRad)));
    }
    const bool by_degrees = a.disp_method < 0 || !a.pd_selection;
    double dist;
    if (by_degrees) {
        dist = sqrt(__dadd_rn(__dmul_rn(dy_deg, dy_deg), __dmul_rn(dx_deg, dx_deg)));     // np.linalg.norm(axis=1)
    } else {
        const double s0 = sin(0.5 * (latc * kDegToRad - latj * kDegToRad));
        const double s1 = sin(0.5 * (lonc * kDegToRad - lonj * kDegToRad));
        dist = 2.0 * asin(sqrt(s0 * s0 + cos(latc * kDegToRad) * cos(latj * kDegToRad) * s1 * s1)) * 6371.0;
    }
    if (!(dist <= a.max_d && dist >= a.min_d && dist != 0.0)) return false;
    if (a.anisotropic) {
        const bool scale = by_degrees ? (a.physical_distance != 0) : true;
        x0 = scale ? dx * 6371.0 : dx;
        x1 = scale ? dy * 6371.0 : dy;
    } else if (by_degrees && a.physical_distance) {
        const double s0 = sin(0.5 * (latc * kDegToRad - latj * kDegToRad));
        const double s1 = sin(0.5 * (lonc * kDegToRad - lonj * kDegToRad));
        x0 = 2.0 * asin(sqrt(s0 * s0 + cos(latc * kDegToRad) * cos(latj * kDegToRad) * s1 * s1)) * 6371.0;
        x1 = 0.0;
    } else {
        x0 = dist;
        x1 = 0.0;
    }
    return true;
}

__global__ void __launch_bounds__(256) ellipse_train_kernel(const TrainArgs a, int fill, int64_t* counts,
                                                            const int64_t* offs, double* x0, double* x1, double* y) {
    __shared__ int scan[256];
    __shared__ int64_t base;
    const int64_t c = a.cells[blockIdx.x];
    if (threadIdx.x == 0) base = fill ? offs[blockIdx.x] : 0;
    __syncthreads();
    int64_t total = 0;
    for (int64_t j0 = 0; j0 < a.N; j0 += 256) {
        const int64_t j = j0 + threadIdx.x;
        double v0 = 0.0, v1 = 0.0;
        const int ok = (j < a.N) ? (int)train_entry(a, c, j, v0, v1) : 0;
        scan[threadIdx.x] = ok;
        __syncthreads();
        for (int s = 1; s < 256; s <<= 1) {                  // inclusive Hillis-Steele scan
            const int v = (int)threadIdx.x >= s ? scan[threadIdx.x - s] : 0;
            __syncthreads();
            scan[threadIdx.x] += v;
            __syncthreads();
        }
        if (fill && ok) {
            const int64_t pos = base + total + scan[threadIdx.x] - 1;
            x0[pos] = v0;
            if (x1) x1[pos] = v1;
            y[pos] = a.cor[c * a.ldc + j];
        }
        total += scan[255];
        __syncthreads();
    }
    if (!fill && threadIdx.x == 0) counts[blockIdx.x] = total;
}

int fill_args(FitArgs& a, const int64_t* offs, const double* x0, const double* x1, const double* y, double* scratch,
              int64_t total, int n_params, int unit_sigma, int arctanh, double nu, const double* guess,
              const double* lo, const double* hi, double xatol, double fatol, int maxiter, int maxfun) {
    a.offs = offs; a.x0 = x0; a.x1 = x1; a.y = y;
    a.z = scratch; a.yll = scratch + total;
    a.n_params = n_params; a.unit_sigma = unit_sigma; a.arctanh = arctanh;
    a.bc = gmg_make_bessel_consts(nu);
    const int N = n_params + (unit_sigma ? 0 : 1);
    for (int k = 0; k < kMaxP; ++k) {
        a.guess[k] = k < N && guess ? guess[k] : 0.0;
        a.lo[k] = k < N && lo ? lo[k] : -INFINITY;
        a.hi[k] = k < N && hi ? hi[k] : INFINITY;
        if (a.lo[k] > a.hi[k]) return -1;
    }
    a.xatol = xatol; a.fatol = fatol; a.maxiter = maxiter; a.maxfun = maxfun;
    return GMG_OK;
}

}  // namespace

extern "C" {

int gmg_ellipse_fit_f64(const int64_t* offs, int64_t n_problems, int64_t total, const double* x0, const double* x1,
                        const double* y, int n_params, int unit_sigma, int arctanh, double nu, const double* h_guess,
                        const double* h_lo, const double* h_hi, double xatol, double fatol, int maxiter, int maxfun,
                        double* scratch, double* x_out, double* fun_out, int* nit_out, int* nfev_out, int* status_out,
                        gmg_stream_t stream) {
    if (!offs) return -1;
    if (n_problems <= 0 || n_problems > 0x7fffffff) return -2;
    if (total < 0) return -3;
    if (!x0) return -4;
    if (n_params < 1 || n_params > 3) return -7;
    if (n_params > 1 && !x1) return -5;
    if (!y) return -6;
    if (!(nu > 0.0)) return -10;
    if (!h_guess) return -11;
    if (!scratch) return -18;
    if (!x_out || !fun_out || !nit_out || !nfev_out || !status_out) return -19;
    FitArgs a;
    if (fill_args(a, offs, x0, x1, y, scratch, total, n_params, unit_sigma, arctanh, nu, h_guess, h_lo, h_hi, xatol,
                  fatol, maxiter, maxfun))
        return -12;                                   // a lower bound above its upper bound (scipy raises ValueError)
    ellipse_fit_kernel<<<(unsigned)n_problems, kFitThreads, 0, (cudaStream_t)stream>>>(a, 0, nullptr, x_out, fun_out,
                                                                                         nit_out, nfev_out, status_out);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_ellipse_train_f64(const double* lat_deg, const double* lon_deg, int64_t N, const double* cor, int64_t ldc,
                          const int64_t* cells, int64_t n_cells, int anisotropic, int physical_distance,
                          int disp_method, int physical_distance_selection, double min_distance, double max_distance,
                          int fill, int64_t* counts, const int64_t* offs, double* x0, double* x1, double* y,
                          gmg_stream_t stream) {
    if (!lat_deg) return -1;
    if (!lon_deg) return -2;
    if (N <= 0) return -3;
    if (!cor || ldc < N) return -4;
    if (!cells) return -6;
    if (n_cells <= 0 || n_cells > 0x7fffffff) return -7;
    if (disp_method < -1 || disp_method > 1) return -10;
    if (physical_distance && disp_method < 0) return -10;      // estimate.py:451-454 raises ValueError
    if (!fill && !counts) return -15;
    if (fill && (!offs || !x0 || !y)) return -16;
    TrainArgs a;
    a.lat = lat_deg; a.lon = lon_deg; a.cor = cor; a.N = N; a.ldc = ldc; a.cells = cells;
    a.anisotropic = anisotropic; a.physical_distance = physical_distance; a.disp_method = disp_method;
    a.pd_selection = physical_distance_selection; a.min_d = min_distance; a.max_d = max_distance;
    ellipse_train_kernel<<<(unsigned)n_cells, 256, 0, (cudaStream_t)stream>>>(a, fill, counts, offs, x0, x1, y);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_ellipse_nll_f64(const int64_t* offs, int64_t n_problems, int64_t total, const double* x0, const double* x1,
                        const double* y, int n_params, int unit_sigma, int arctanh, double nu, const double* params,
                        double* scratch, double* fun_out, gmg_stream_t stream) {
    if (!offs) return -1;
    if (n_problems <= 0 || n_problems > 0x7fffffff) return -2;
    if (total < 0) return -3;
    if (!x0) return -4;
    if (n_params < 1 || n_params > 3) return -7;
    if (n_params > 1 && !x1) return -5;
    if (!y) return -6;
    if (!(nu > 0.0)) return -10;
    if (!params) return -11;
    if (!scratch) return -12;
    if (!fun_out) return -13;
    FitArgs a;
    fill_args(a, offs, x0, x1, y, scratch, total, n_params, unit_sigma, arctanh, nu, nullptr, nullptr, nullptr, 0.0, 0.0,
              0, 0);
    ellipse_fit_kernel<<<(unsigned)n_problems, kFitThreads, 0, (cudaStream_t)stream>>>(a, 1, params, nullptr, fun_out,
                                                                                         nullptr, nullptr, nullptr);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

}  // extern "C"
