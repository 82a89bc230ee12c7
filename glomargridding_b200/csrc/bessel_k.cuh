// Modified Bessel function of the second kind K_nu(x), real nu >= 0, x > 0, in FP64.
//
// Replaces the reference's `scipy.special.kv` call sites (ellipse/covariance.py:346,466), which
// dispatch to AMOS `zbesk` in scipy (not vendored under /root/reference; scipy 1.16.1 pinned in
// uv.lock:1716-1717).  This is NOT a port of AMOS: it is Temme's method (series for x <= 2,
// Steed's continued fraction CF2 for x > 2, upward recurrence in the order), restated from the
// published algorithm (N. M. Temme, J. Comput. Phys. 19 (1975) 324; Numerical Recipes §6.7).
//
// Everything that depends only on nu (the fractional order mu, the Gamma-function terms of
// Temme's series, the Matern normaliser) is computed ONCE on the host in `gmg_make_bessel_consts`
// and passed to kernels by value - nu is a scalar parameter of a covariance build, never data.
//
// The functions are __host__ __device__ so that tests can exercise the identical code on the CPU
// (tests/test_bessel_host.py) where the build container has no GPU.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define GMG_HD __host__ __device__ __forceinline__
#else
#define GMG_HD inline
#endif

struct GmgBesselConsts {
    double nu;        // order
    double mu;        // nu - nl, in [-0.5, 0.5]
    int nl;           // number of upward recurrence steps, round(nu)
    double gam1;      // (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu)
    double gam2;      // (1/Gamma(1-mu) + 1/Gamma(1+mu)) / 2
    double gampl;     // 1/Gamma(1+mu)
    double gammi;     // 1/Gamma(1-mu)
    double pimu_over_sin;  // pi*mu / sin(pi*mu)  (1 at mu = 0)
    double matern_norm;    // 1 / (Gamma(nu) * 2^(nu-1))   (reference ellipse/covariance.py:223)
    double two_sqrt_nu;    // 2 sqrt(nu)                   (reference ellipse/covariance.py:224)
};

// Host-only: fill the nu-dependent constants.  long double keeps the (1/G(1-mu) - 1/G(1+mu))/(2mu)
// difference accurate for moderately small |mu|; for |mu| < 1e-3 the Taylor series of 1/Gamma(1+z)
// (Abramowitz & Stegun 6.1.34) is used instead.
inline GmgBesselConsts gmg_make_bessel_consts(double nu) {
    GmgBesselConsts c;
    c.nu = nu;
    c.nl = (int)floor(nu + 0.5);
    c.mu = nu - c.nl;
    const long double mu = (long double)c.mu;
    const long double gp = 1.0L / tgammal(1.0L + mu);
    const long double gm = 1.0L / tgammal(1.0L - mu);
    c.gampl = (double)gp;
    c.gammi = (double)gm;
    c.gam2 = (double)(0.5L * (gm + gp));
    if (fabsl(mu) < 1e-3L) {
        // 1/Gamma(1+z) = 1 + a1 z + a2 z^2 + ...; gam1 = -(a1 + a3 mu^2 + a5 mu^4 + ...)
        const long double a1 = 0.5772156649015328606065121L;
        const long double a3 = -0.0420026350340952355290040L;
        const long double a5 = -0.0421977345555443367482083L;
        const long double m2 = mu * mu;
        c.gam1 = (double)(-(a1 + m2 * (a3 + m2 * a5)));
        const long double pm = 3.14159265358979323846264338L * mu;
        c.pimu_over_sin = (double)(1.0L / (1.0L - pm * pm / 6.0L + pm * pm * pm * pm / 120.0L));
    } else {
        c.gam1 = (double)((gm - gp) / (2.0L * mu));
        const long double pm = 3.14159265358979323846264338L * mu;
        c.pimu_over_sin = (double)(pm / sinl(pm));
    }
    c.matern_norm = 1.0 / (tgamma(nu) * exp2(nu - 1.0));
    c.two_sqrt_nu = 2.0 * sqrt(nu);
    return c;
}

// e^x * K_nu(x) (exponentially scaled, like scipy.special.kve) for x > 0.
// Returns +inf at x == 0 and NaN for x < 0 or NaN input, as kv does.
GMG_HD double gmg_bessel_kve(const GmgBesselConsts& c, double x) {
    if (!(x > 0.0)) return (x == 0.0) ? INFINITY : NAN;
    const double mu = c.mu;
    const double mu2 = mu * mu;
    const double xi = 1.0 / x;
    const double xi2 = 2.0 * xi;
    double rkmu, rk1;
    if (x <= 2.0) {
        // Temme's series for K_mu and K_{mu+1}
        const double b = 0.5 * x;
        double d = -log(b);
        double e = mu * d;
        const double fact2 = (fabs(e) < 1e-8) ? 1.0 + e * e * (1.0 / 6.0) : sinh(e) / e;
        double ff = c.pimu_over_sin * (c.gam1 * cosh(e) + c.gam2 * fact2 * d);
        double sum = ff;
        e = exp(e);
        double p = 0.5 * e / c.gampl;
        double q = 0.5 / (e * c.gammi);
        double cc = 1.0;
        d = b * b;
        double sum1 = p;
        for (int i = 1; i < 500; ++i) {
            const double fi = (double)i;
            ff = (fi * ff + p + q) / (fi * fi - mu2);
            cc *= d / fi;
            p /= (fi - mu);
            q /= (fi + mu);
            const double del = cc * ff;
            sum += del;
            sum1 += cc * (p - fi * ff);
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        const double ex = exp(x);
        rkmu = sum * ex;
        rk1 = sum1 * xi2 * ex;
    } else {
        // Steed's algorithm for the continued fraction CF2
        double b = 2.0 * (1.0 + x);
        double d = 1.0 / b;
        double h = d, delh = d;
        double q1 = 0.0, q2 = 1.0;
        const double a1 = 0.25 - mu2;
        double q = a1, cc = a1;
        double a = -a1;
        double s = 1.0 + q * delh;
        for (int i = 2; i < 2000; ++i) {
            a -= 2.0 * (i - 1);
            cc = -a * cc / (double)i;
            const double qnew = (q1 - b * q2) / a;
            q1 = q2;
            q2 = qnew;
            q += cc * qnew;
            b += 2.0;
            d = 1.0 / (b + a * d);
            delh = (b * d - 1.0) * delh;
            h += delh;
            const double dels = q * delh;
            s += dels;
            if (fabs(dels) < fabs(s) * 1e-17) break;
        }
        h = a1 * h;
        rkmu = sqrt(1.5707963267948966 * xi) / s;   // sqrt(pi / (2x)) / s, e^-x left out (scaled)
        rk1 = rkmu * (mu + x + 0.5 - h) * xi;
    }
    // upward recurrence K_{m+1} = K_{m-1} + (2m/x) K_m, nl steps from (K_mu, K_{mu+1})
    if (c.nl == 0) return rkmu;
    double km = rkmu, kp = rk1;
    for (int i = 1; i < c.nl; ++i) {
        const double kn = (mu + i) * xi2 * kp + km;
        km = kp;
        kp = kn;
    }
    return kp;
}

// K_nu(x) itself (underflows to 0 near x ~ 745 like scipy's kv).
GMG_HD double gmg_bessel_kv(const GmgBesselConsts& c, double x) {
    return gmg_bessel_kve(c, x) * exp(-x);
}

// Scaled Matern factor g(x) = x^nu K_nu(x) e^x / (Gamma(nu) 2^(nu-1)); the Matern correlation is
// g(x) e^-x (reference ellipse/covariance.py:447-466 with the stdev / determinant prefactor removed).
GMG_HD double gmg_matern_scaled(const GmgBesselConsts& c, double x) {
    return c.matern_norm * pow(x, c.nu) * gmg_bessel_kve(c, x);
}
