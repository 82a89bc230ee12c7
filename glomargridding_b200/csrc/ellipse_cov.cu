// K1/K2: Paciorek-Schervish non-stationary Matern covariance tiles for sm_100a.
//
// Replaces (reference, paths relative to its root):
//   glomar_gridding/ellipse/covariance.py:205-248  _calculate_covariance (pre-terms, diagonal)
//   glomar_gridding/ellipse/covariance.py:250-403  the three build strategies + mirror
//   glomar_gridding/ellipse/covariance.py:405-468  c_ij_anisotropic_array (the pair formula)
//   glomar_gridding/ellipse/covariance.py:512-654  Sigma rotation, 2x2 det, haversine, displacements
//
// Design: a persistent grid (2 CTAs per SM) loops over 64 x 64 tiles of C in FP64.
// Per-point terms (lat, lon, cos lat / 2, Sigma / 2, stdev * det^(1/4), stdev^2) are packed once per
// build by `ellipse_prepare_kernel` (64 B per point, record-major and field-major copies) and fetched
// with cp.async into a double-buffered shared-memory stage for the tile's 64 row points and 64 column
// points while the previous tile is computed.  Each thread owns two adjacent columns (held in
// registers) and 8 rows (broadcast from shared memory), so a warp stores 512 contiguous bytes per row
// with 16-byte vector stores.  Inside the square diagonal block of a row shard only the upper-triangle
// tiles are computed; the lower ones are written as the transpose of their mirror tile through a padded
// shared-memory tile, so the FP64 work is halved and both stores stay coalesced.
// The formula is sign-symmetric under i <-> j (IEEE add/mul are commutative and the wrap is odd),
// so the mirrored entries are bit-identical to evaluating them in the reference's i<j orientation.
//
// The Matern factor x^nu K_nu(x) / (Gamma(nu) 2^(nu-1)) is evaluated as g(x) e^-x with
//   nu = 0.5 / 1.5 / 2.5 : g = 1, 1 + x, 1 + x + x^2/3              (closed forms)
//   otherwise            : g from a piecewise degree-5 polynomial table in the mantissa of x
//                          (32 linear sub-intervals per octave, 2^-6 <= x < 2^9), built on the
//                          device from the Temme/CF2 Bessel-K in bessel_k.cuh for the build's nu;
//                          lanes outside the table range fall back to the direct evaluation.
#include <cstring>
#include <type_traits>
#include "gmg_common.cuh"
#include "bessel_k.cuh"
#include "../../include/glomar_b200.h"

#ifndef K1_EXPERIMENT
#define K1_EXPERIMENT 0     // != 0: kernel-experiment builds that drop one component of the pair formula
#endif

namespace {

constexpr double kEarthRadiusKm = 6371.0;           // reference constants.py:18
constexpr double kPi = 3.141592653589793;           // np.pi
constexpr double kTwoPi = 6.283185307179586;        // 2 * np.pi, reference ellipse/covariance.py:37

constexpr int kTile = 64;
constexpr int kMaxMT = 2;                           // largest macro tile: 2 x 2 tiles
constexpr int kPadTail = kMaxMT * kTile;            // padding records after the (128-padded) points
constexpr int kThreads = 256;

// Matern table geometry
// Table geometry (A/B on B200, nu = 2.2, gpurun_out/r02_k1_ab*.log): 16 sub-intervals per octave from 2^-12 with
// degree 7: 10.8-11.1 ms; degree 6: 10.54-10.74; 32 sub-intervals from 2^-6 with degree 5 (shipped): 10.45 ms and
// 3e-14 relative error.  Arguments below 2^-6 = 0.016 (a few pairs next to the poles) take the cold path.
#ifndef K1_TABEMIN
#define K1_TABEMIN -6
#endif
#ifndef K1_TABSUBLOG2
#define K1_TABSUBLOG2 5
#endif
constexpr int kTabEmin = K1_TABEMIN;
constexpr int kTabEmax = 9;                         // exclusive: the table ends at x = 512 < 700
constexpr int kTabSubLog2 = K1_TABSUBLOG2;
constexpr int kTabSub = 1 << kTabSubLog2;                   // linear sub-intervals of the mantissa per octave
constexpr int kTabNint = (kTabEmax - kTabEmin) * kTabSub;   // 336
#ifndef K1_TABDEG
#define K1_TABDEG 5
#endif
#ifndef K1_TSHIFT
#define K1_TSHIFT 1
#endif
constexpr int kTabDeg = K1_TABDEG;                          // 7: 5e-14 relative; 6: 2e-13 (one LDS + one DFMA less per pair)
constexpr int kTabCoef = kTabDeg + 1;
constexpr int kTabDoubles = (kTabNint * kTabCoef + 1) / 2 * 2;   // 2688 (21.5 KB) at degree 7, layout [coef][interval]
static_assert(kTabDeg >= 4 && kTabDeg <= 7, "matern_table_kernel samples kTabCoef <= 8 Chebyshev nodes");

enum { MODE_NU05 = 0, MODE_NU15 = 1, MODE_NU25 = 2, MODE_TABLE = 3, MODE_DIRECT = 4 };

// packed per-point record, 8 doubles
enum { P_LAT = 0, P_LON = 1, P_HCOS = 2, P_H00 = 3, P_H01 = 4, P_H11 = 5, P_W = 6, P_VAR = 7, P_N = 8 };

__global__ void __launch_bounds__(256) ellipse_prepare_kernel(
    const double* __restrict__ Lx, const double* __restrict__ Ly, const double* __restrict__ theta,
    const double* __restrict__ stdev, const double* __restrict__ lat, const double* __restrict__ lon,
    int64_t N, int64_t npad, int disp_method, double x_scale, double* __restrict__ aos, double* __restrict__ soa) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= npad + kPadTail) return;
    double rec[P_N];
    if (i < N) {
        // reference `_sigma_rot_func_multi`, ellipse/covariance.py:512-533
        double st, ct;
        sincos(theta[i], &st, &ct);
        const double c2 = ct * ct, s2 = st * st, cs = ct * st;
        const double lx2 = Lx[i] * Lx[i], ly2 = Ly[i] * Ly[i];
        const double s00 = __dadd_rn(__dmul_rn(c2, lx2), __dmul_rn(s2, ly2));
        const double s01 = __dmul_rn(cs, lx2 - ly2);
        const double s11 = __dadd_rn(__dmul_rn(s2, lx2), __dmul_rn(c2, ly2));
        const double det = __dsub_rn(__dmul_rn(s00, s11), __dmul_rn(s01, s01));   // :543-547
        const double sd = stdev[i];
        rec[P_LAT] = lat[i];
        rec[P_LON] = lon[i];
        // 0.5*(cos a + cos b) == 0.5 cos a + 0.5 cos b exactly; "Met_Office" has no cos-latitude factor
        // (0.5 + 0.5 == 1)
        rec[P_HCOS] = disp_method == GMG_DISP_MODIFIED_MET_OFFICE ? 0.5 * cos(lat[i]) : 0.5;
        // 0.5*(S_i + S_j) == 0.5 S_i + 0.5 S_j exactly.  The constant of the Matern argument, inner = x_scale *
        // sqrt(Q / det) with x_scale = 2 sqrt(nu) R (:224, :457, :639), is folded into the records: with S' = S /
        // x_scale^2 the quadratic form and determinant become Q / x_scale^2 and det / x_scale^4, so that
        // sqrt(Q' / det') = inner and rsqrt(det') = x_scale^2 rsqrt(det) - the latter absorbed by w' = w / x_scale.
        // One FP64 multiply less per pair; the extra rounding is half an ulp per record entry.
        const double lam = x_scale * x_scale;
        rec[P_H00] = (0.5 * s00) / lam;
        rec[P_H01] = (0.5 * s01) / lam;
        rec[P_H11] = (0.5 * s11) / lam;
        rec[P_W] = sd * sqrt(sqrt(det)) / x_scale;   // stdev_i det_i^(1/4): sqrt(sqrt_det_i sqrt_det_j) = w_i w_j
        rec[P_VAR] = sd * sd;                   // diagonal, :246
    } else {
        // padding points of partial tiles: benign values (det = 1 > 0, q = 0 -> cold path), never stored
#pragma unroll
        for (int f = 0; f < P_N; ++f) rec[f] = 1.0;
    }
#pragma unroll
    for (int f = 0; f < P_N; ++f) {
        aos[i * P_N + f] = rec[f];
        if (i < npad) soa[f * npad + i] = rec[f];
    }
}

// One block per table interval: sample g at Chebyshev nodes, transform to monomial coefficients in
// t = (m - c_k) * 2 kTabSub in [-1, 1] where m in [1, 2) is the mantissa of x and c_k the sub-interval centre.
__global__ void __launch_bounds__(32) matern_table_kernel(GmgBesselConsts bc, double* __restrict__ table) {
    __shared__ double f[kTabCoef];
    __shared__ double cheb[kTabCoef];
    const int interval = blockIdx.x;
    const int e = kTabEmin + interval / kTabSub;
    const int k = interval % kTabSub;
    const double centre = ldexp(1.0 + (k + 0.5) / kTabSub, e);
    const double halfw = ldexp(0.5 / kTabSub, e);
    const int t = threadIdx.x;
    if (t < kTabCoef) {
        const double node = cospi((t + 0.5) / kTabCoef);
        f[t] = gmg_matern_scaled(bc, centre + halfw * node);
    }
    __syncthreads();
    if (t < kTabCoef) {
        double acc = 0.0;
        for (int j = 0; j < kTabCoef; ++j) acc += f[j] * cospi(t * (j + 0.5) / kTabCoef);
        cheb[t] = acc * (t == 0 ? 1.0 : 2.0) / kTabCoef;
    }
    __syncthreads();
    if (t == 0) {
        // Chebyshev -> monomial via T_{n+1} = 2 t T_n - T_{n-1}
        double mono[kTabCoef], tn[kTabCoef], tp[kTabCoef], tmp[kTabCoef];
        for (int j = 0; j < kTabCoef; ++j) { mono[j] = 0.0; tn[j] = 0.0; tp[j] = 0.0; }
        tp[0] = 1.0;                   // T_0
        tn[1] = 1.0;                   // T_1
        mono[0] += cheb[0];
        mono[1] += cheb[1];
        for (int n = 2; n < kTabCoef; ++n) {
            for (int j = 0; j < kTabCoef; ++j) tmp[j] = (j > 0 ? 2.0 * tn[j - 1] : 0.0) - tp[j];
            for (int j = 0; j < kTabCoef; ++j) { tp[j] = tn[j]; tn[j] = tmp[j]; mono[j] += cheb[n] * tmp[j]; }
        }
        for (int j = 0; j < kTabCoef; ++j) table[j * kTabNint + interval] = mono[j];   // [coef][interval]
    }
}

struct MaternCtx {
    GmgBesselConsts bc;
    const double* table;   // shared-memory (tile kernel) or global (pairs kernel) pointer, [coef][interval]
};

// Horner (degree kTabDeg); the table is coefficient-major, [coef][interval]: neighbouring lanes of a warp fall into the
// same or adjacent intervals (x varies smoothly along a row of the tile), so each of the eight 8-byte loads touches
// one or two 128-byte lines - one shared-memory wavefront - where a 64-byte-record layout read with LDS.128 needed
// ~6 per load (ncu r01: 39 shared wavefronts per warp-pair, the LSU pipe 69 % busy, against 14 for the closed forms).
__device__ __forceinline__ double table_horner(const double* c, double t) {
    double p = fma(c[kTabDeg * kTabNint], t, c[(kTabDeg - 1) * kTabNint]);
#pragma unroll
    for (int k = kTabDeg - 2; k >= 0; --k) p = fma(p, t, c[k * kTabNint]);
    return p;
}

template <int MODE>
__device__ __forceinline__ double matern_scaled_eval(const MaternCtx& m, double x) {
    if (MODE == MODE_NU05) return 1.0;
    if (MODE == MODE_NU15) return 1.0 + x;
    if (MODE == MODE_NU25) return fma(x, fma(x, 1.0 / 3.0, 1.0), 1.0);
    if (MODE == MODE_TABLE) {
        const int hi = __double2hiint(x);
        const int lo = __double2loint(x);
        const int idx = ((hi >> 20) - (1023 + kTabEmin)) * kTabSub + ((hi >> (20 - kTabSubLog2)) & (kTabSub - 1));
        if ((unsigned)idx < (unsigned)kTabNint) {     // also rejects x <= 0, inf, NaN (sign/exponent bits)
            const double mant = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
            const int cbits = (hi & ((kTabSub - 1) << (20 - kTabSubLog2))) | (1 << (19 - kTabSubLog2));
            const double centre = __hiloint2double(0x3FF00000 | cbits, 0);
            const double t = (mant - centre) * (double)(2 * kTabSub);
            return table_horner(m.table + idx, t);
        }
    }
    return gmg_matern_scaled(m.bc, x);
}

struct PointRegs { double lat, lon, hcos, h00, h01, h11, w; };

__device__ __forceinline__ PointRegs load_point(const double* p) {     // p: 8 consecutive doubles (AoS)
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *reinterpret_cast<const double2*>(p + 2);
    const double2 c = *reinterpret_cast<const double2*>(p + 4);
    PointRegs r;
    r.lat = a.x; r.lon = a.y; r.hcos = b.x; r.h00 = b.y; r.h01 = c.x; r.h11 = c.y; r.w = p[6];
    return r;
}

// ---- straight-line FP64 helpers for the hot loop ------------------------------------------------------
// The kernel is bound by the FP64 pipe (64 lanes / SM: one warp-wide FP64 instruction per 2 cycles per SM
// sub-partition), so the loop body is written to minimise FP64 instructions and to contain no branches.
// The CUDA libm sqrt / rsqrt / exp carry special-case branches and slow-path calls (ncu r01: 222 issued
// instructions per pair for 72 FP64 ones, "wait" stalls dominant); arguments here are positive and normal,
// so a cubically convergent refinement of the MUFU.RSQ64H seed and a table-driven exp suffice.  Anything
// unusual (x >= 700, x outside the Matern table, NaN from a non-positive determinant) is flagged and
// recomputed on a cold path with the libm functions.

// 2^(-j/128), j = 0..127, correctly rounded (mpmath).  Every CTA stages kExpRep = 16 interleaved copies in
// shared memory, entry j of copy c at [j * 16 + c]: a 64-bit shared load is served per half-warp, and lane l
// reads copy l & 15, i.e. always its own bank pair - the data-dependent lookup is conflict-free
// (ncu r01b: 4.8 wavefronts per lookup with a single copy).
constexpr int kExpRep = 16;
constexpr int kExpTab = 128;
__device__ const double kExp2Neg[kExpTab] = {
    1.0, 0.9945994234836332, 0.9892280131939755, 0.9838856116165879,
    0.9785720620877001, 0.9732872087896166, 0.9680308967461472, 0.9628029718180625,
    0.9576032806985737, 0.9524316709088371, 0.9472879907934828, 0.9421720895161673,
    0.93708381705515, 0.9320230241988945, 0.9269895625416927, 0.921983284479313,
    0.9170040432046712, 0.9120516927035267, 0.9071260877501994, 0.902227083903312,
    0.8973545375015536, 0.8925083056594675, 0.8876882462632606, 0.8828942179666364,
    0.8781260801866497, 0.8733836930995845, 0.8686669176368531, 0.8639756154809188,
    0.859309649061239, 0.8546688815502315, 0.8500531768592617, 0.8454623996346526,
    0.8408964152537145, 0.8363550898207983, 0.8318382901633682, 0.8273458838280972,
    0.8228777390769825, 0.8184337248834822, 0.8140137109286739, 0.8096175675974319,
    0.8052451659746271, 0.8008963778413467, 0.7965710756711335, 0.7922691326262469,
    0.7879904225539432, 0.7837348199827765, 0.7795022001189185, 0.7752924388425,
    0.7711054127039704, 0.766940998920478, 0.7627990753722692, 0.7586795205991074,
    0.7545822137967114, 0.7505070348132128, 0.7464538641456324, 0.7424225829363762,
    0.7384130729697497, 0.7344252166684909, 0.7304588970903235, 0.7265139979245263,
    0.7225904034885233, 0.7186879987244912, 0.714806669195985, 0.7109463010845828,
    0.7071067811865476, 0.7032879969095077, 0.6994898362691556, 0.6957121878859631,
    0.691954940981916, 0.6882179853772651, 0.6845012114872953, 0.6808045103191124,
    0.6771277734684463, 0.6734708931164729, 0.6698337620266515, 0.6662162735415808,
    0.6626183215798707, 0.659039800633032, 0.6554806057623822, 0.6519406325959679,
    0.6484197773255048, 0.6449179367033329, 0.6414350080393891, 0.637970889198196,
    0.6345254785958666, 0.6310986751971254, 0.6276903785123455, 0.6243004885946024,
    0.620928906036742, 0.6175755319684667, 0.614240268053435, 0.6109230164863788,
    0.6076236799902345, 0.6043421618132908, 0.6010783657263515, 0.5978321960199137,
    0.5946035575013605, 0.5913923554921705, 0.5881984958251406, 0.5850218848416251,
    0.5818624293887887, 0.5787200368168756, 0.5755946149764913, 0.5724860722159021,
    0.5693943173783458, 0.5663192597993596, 0.5632608093041209, 0.5602188762048034,
    0.5571933712979462, 0.5541842058618394, 0.5511912916539204, 0.5482145409081884,
    0.5452538663326288, 0.5423091811066546, 0.5393803988785599, 0.5364674337629878,
    0.5335702003384118, 0.530688613644631, 0.5278225891802786, 0.5249720429003436,
    0.5221368912137069, 0.5193170509806894, 0.5165124395106142, 0.5137229745593819,
    0.5109485743270583, 0.5081891574554765, 0.5054446430258502, 0.5027149505564014};

// FP64 literals of the hot loop live in constant memory: a DFMA can take a c[bank][offset] operand directly,
// whereas a 64-bit immediate costs two extra MOV issue slots per use once registers are short (ncu r01c: the
// loop was issue-bound, 53 non-FP64 instructions per pair for 45 FP64 ones).
struct HotConsts {
    double magic, inv_ln2_128, ln2_128, c4, c3, half, one, c38, pi, third, tab_scale, two, three;
};
__constant__ HotConsts kHot = {6755399441055744.0,        // 1.5 * 2^52
                               184.6649652337873,        // 128 / ln 2
                               0.0054152123481245725,      // ln 2 / 128, correctly rounded
                               4.16666666666666666667e-02, 1.66666666666666666667e-01, 0.5, 1.0, 0.375,
                               3.141592653589793, 1.0 / 3.0, (double)(2 * kTabSub), 2.0, 3.0};

// 1/sqrt(a), a > 0 normal: MUFU.RSQ64H seed and one cubically convergent step
//   e = 1 - a r^2,  r <- r + r e (1/2 + 3/8 e)      (error ~ 5/16 e^3)       5 FP64 instructions
// The quadratic step r <- r + r e / 2 (4 instructions) is NOT enough: measured on B200 it leaves 1.2e-10 relative
// on covariance entries (the seed is good to ~2^-18 only, and the error of rsqrt(Q) is amplified by the Matern
// argument), against 5e-14 with the cubic step.
__device__ __forceinline__ double rsqrt_pos(double a) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    const double e = fma(-(a * r), r, kHot.one);
    return fma(r * e, fma(e, kHot.c38, kHot.half), r);
}

// exp(-x) for 0 <= x < 700 (result normal):  k = rint(128 x / ln 2),  exp(-x) = 2^-(k>>7) T[k&127] exp(s),
// s = k ln2/128 - x in [-ln2/256, ln2/256].  The FMA forms k * c - x exactly and rounds once, so the only
// reduction error is k * |c - ln2/128| <= 9e4 * 6e-19 = 5e-14 absolute at x = 700 (1e-14 at x = 100): no
// Cody-Waite second constant is needed at the 1e-10 parity target.  Degree-4 Taylor (truncation s^5/120 <
// 1.2e-15).  8 FP64 instructions + one shared-memory load.
template <int STRIDE>
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ sexp) {
    const double t = fma(x, kHot.inv_ln2_128, kHot.magic);
    const int k = __double2loint(t);
    const double kf = t - kHot.magic;
    const double s = fma(kf, kHot.ln2_128, -x);
#ifdef K1_EXPDEG3                                                // degree-3 Taylor: truncation s^4/24 <= 2.2e-12
    double p = fma(kHot.c3, s, kHot.half);
    p = fma(p, s, kHot.one);
#else
    double p = fma(kHot.c4, s, kHot.c3);                          // 1/4!, 1/3!
    p = fma(p, s, kHot.half);
    p = fma(p, s, kHot.one);
#endif
    const double ts = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(sexp) +
                                                       (k & (kExpTab - 1)) * (STRIDE * (int)sizeof(double)));
    p = fma(ts * s, p, ts);                                       // T (1 + s p)
    // 2^-(k >> 7): subtract (k >> 7) << 20 from the high word = (k & ~127) * 2^13, one LOP3 + one IMAD
    return __hiloint2double(__double2hiint(p) - (k & ~(kExpTab - 1)) * 8192, __double2loint(p));
}

// g(x) of the fast path; `slow` is set when the lane must be recomputed by `cov_pair_slow`.
template <int MODE>
__device__ __forceinline__ double matern_scaled_fast(const MaternCtx& m, double x, bool& slow) {
    const int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    if (MODE != MODE_TABLE) {
        // !(0 <= x < 700) on the high word in the integer pipe (700.0 = 0x4085E000_00000000); also catches NaN
        slow = (unsigned)hi >= 0x4085E000u;
        if (MODE == MODE_NU05) return kHot.one;
        if (MODE == MODE_NU15) return kHot.one + x;
        if (MODE == MODE_NU25) return fma(x, fma(x, kHot.third, kHot.one), kHot.one);
        slow = true;                                              // MODE_DIRECT (validation): always the cold path
        return 0.0;
    }
    // Integer instructions run at half rate on sm_100 (profiles/r01_microbench_xu.txt: a LOP3 costs 2 cycles of
    // its pipe per warp), so the index arithmetic is kept to 6 of them.  With a positive x the interval index
    // (exponent - emin) * 16 + top 4 mantissa bits is simply (hi >> 16) - 16 * (1023 + emin); one unsigned
    // compare rejects everything outside [2^emin, 2^emax) - which includes x >= 700 (emax = 9), x <= 0 and NaN.
    static_assert(kTabEmax <= 9, "the range check below relies on the table ending at or below 512 < 700");
    const unsigned u = (unsigned)((hi >> (20 - kTabSubLog2)) - kTabSub * (1023 + kTabEmin));
    slow = u >= (unsigned)kTabNint;
    const unsigned idx = min(u, (unsigned)(kTabNint - 1));
#if K1_TSHIFT
    // t = 2 frac(kTabSub mant) - 1: the mantissa bits below the sub-interval bits, shifted up past them, ARE the
    // fraction - d = 1.f in [1, 2) is put together with one funnel shift, one LOP3 and one shift (no zero low word to
    // materialise, no second constant double), and t = 2 d - 3
    const unsigned hs = __funnelshift_l((unsigned)lo, (unsigned)hi, kTabSubLog2);
    const double d = __hiloint2double((int)((hs & 0x000FFFFFu) | 0x3FF00000u), (int)((unsigned)lo << kTabSubLog2));
    const double t = fma(d, kHot.two, -kHot.three);
#else
    static_assert(kTabSubLog2 == 4, "the constant below is written for 16 sub-intervals");
    const double mant = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    // t = (mant - centre) * 32 = mant * 32 - (33 + 2 sub): the subtrahend is the integer-valued double
    // 1.(2 sub + 1) x 2^5, whose high word is 0x40400000 | sub << 16 | 1 << 15 - one LOP3 on hi
    const double c32 = __hiloint2double((hi & 0x000F0000) | 0x40408000, 0);
    const double t = fma(mant, kHot.tab_scale, -c32);
#endif
    return table_horner(m.table + idx, t);
}

// longitude wrap, reference ellipse/covariance.py:633-634:  dx > pi -> dx - 2 pi,  dx < -pi -> dx + 2 pi
// (strict).  |dx| <= 2 pi, so this is dx - copysign(2 pi, dx) when |dx| > pi: one DSETP, one DADD and three
// integer instructions; the subtraction of +0.0 in the other case is exact.
__device__ __forceinline__ double wrap_lon(double dx) {
    const bool gt = fabs(dx) > kHot.pi;                                                // |dx| > np.pi, strict
    const unsigned hi = (unsigned)__double2hiint(dx);
    const unsigned whi = gt ? (0x401921FBu | (hi & 0x80000000u)) : 0u;                 // copysign(2 pi, dx)
    const unsigned wlo = gt ? 0x54442D18u : 0u;
    return dx - __hiloint2double((int)whi, (int)wlo);
}

// Geometry shared by the fast and slow paths: reference `c_ij_anisotropic_array`,
// ellipse/covariance.py:441-463, with
//   sqrt(sqrt_det_i sqrt_det_j / det) = w_i w_j rsqrt(det)   and   inner = 2 sqrt(nu) R sqrt(Q) rsqrt(det)
// where Q is the quadratic form of the displacements in RADIANS (the Earth radius of :639 and 2 sqrt(nu) are
// folded into the per-point records by the prepare kernel).  For "Met_Office" the prepare kernel stores hcos = 0.5, so (a.hcos + b.hcos) == 1.0 exactly and
// the mean-cos-latitude factor of "Modified_Met_Office" (:636-637) drops out without a second code path.
struct PairGeom { double det, q, dx, dy; };

template <bool GEOM_F32>
__device__ __forceinline__ PairGeom pair_geom(const PointRegs& a, const PointRegs& b) {
    double dy, dx;
    if (GEOM_F32) {
        // precision=float32: the reference holds radians as float32 and subtracts / compares / wraps in
        // float32 (np.pi and TWO_PI are weak scalars), which decides the 180-degree ties differently
        // from a float64 subtract of the same values.  Inputs here are exactly float32-representable.
        float d = __fsub_rn((float)a.lon, (float)b.lon);
        if (d > (float)kPi) d = __fsub_rn(d, (float)kTwoPi);
        if (d < -(float)kPi) d = __fadd_rn(d, (float)kTwoPi);
        dx = (double)d;
        dy = (double)__fsub_rn((float)a.lat, (float)b.lat);
    } else {
        dy = a.lat - b.lat;                                 // :631 orientation (i) - (j)
#if K1_EXPERIMENT == 6      // cost of the longitude wrap
        dx = a.lon - b.lon;
#else
        dx = wrap_lon(a.lon - b.lon);                       // :632-634 exact IEEE subtract of host radians
#endif
    }
    dx *= (a.hcos + b.hcos);                                // :636-637, after the wrap
    const double s0 = a.h00 + b.h00, s1 = a.h01 + b.h01, s3 = a.h11 + b.h11;   // :451
    PairGeom g;
    g.det = fma(s0, s3, -(s1 * s1));                        // :452
    g.q = dx * fma(dx, s3, -(dy * s1)) + dy * fma(dy, s0, -(dx * s1));         // :459-461
    g.dx = dx;
    g.dy = dy;
    return g;
}

template <int MODE, bool GEOM_F32, int EXPSTRIDE>
__device__ __forceinline__ double cov_pair(const PointRegs& a, const PointRegs& b, const MaternCtx& m,
                                           const double* __restrict__ sexp, bool& slow) {
    const PairGeom g = pair_geom<GEOM_F32>(a, b);
#if K1_EXPERIMENT == 5      // cost of the two reciprocal square roots (kernel-experiment builds, never shipped)
    const double r = g.det * 1e-12;
    const double x = g.q * (g.q * 1e-3) * r;
#else
    const double r = rsqrt_pos(g.det);
    const double x = g.q * rsqrt_pos(g.q) * r;                                  // :457 inner (x_scale is in the records)
#endif
#if K1_EXPERIMENT == 7      // cost of the Matern factor
    slow = false;
    const double gm = 1.0;
#else
    const double gm = matern_scaled_fast<MODE>(m, x, slow);
#endif
#if K1_EXPERIMENT == 4      // cost of exp(-x)
    return ((a.w * b.w) * r) * gm * (1.0 - x * 1e-9);
#else
    return ((a.w * b.w) * r) * gm * exp_neg<EXPSTRIDE>(x, sexp);                           // :447-455, :465-466
#endif
}

// Cold path for flagged lanes: libm sqrt / exp, Temme-CF2 Bessel-K; any x (incl. underflow to 0 and NaN).
// Takes pointers (to the 64 B point records and to the context) so that the hot loop never spills its
// registers to the stack to make the call.
template <int MODE, bool GEOM_F32>
__device__ __forceinline__ double cov_pair_slow(const double* pa, const double* pb, int sb, const MaternCtx* m) {
    const PointRegs a = load_point(pa);
    PointRegs b;                                     // second point: field f at pb[f * sb]
    b.lat = pb[P_LAT * sb]; b.lon = pb[P_LON * sb]; b.hcos = pb[P_HCOS * sb]; b.h00 = pb[P_H00 * sb];
    b.h01 = pb[P_H01 * sb]; b.h11 = pb[P_H11 * sb]; b.w = pb[P_W * sb];
    const PairGeom g = pair_geom<GEOM_F32>(a, b);
    const double x = sqrt(g.q / g.det);
    return (a.w * b.w) * rsqrt(g.det) * matern_scaled_eval<MODE>(*m, x) * exp(-x);
}

// `_haversine_multi`, ellipse/covariance.py:572-588 (only when max_dist is set, :256-266)
__device__ __forceinline__ bool beyond_max_dist(const double* pa, const double* pb, int sb, double max_dist) {
    const double lat0 = pa[P_LAT], lat1 = pb[P_LAT * sb];
    const double sl = sin(0.5 * (lat0 - lat1)), so = sin(0.5 * (pa[P_LON] - pb[P_LON * sb]));
    const double h = sl * sl + cos(lat0) * cos(lat1) * so * so;
    return kEarthRadiusKm * (2.0 * asin(sqrt(h))) > max_dist;
}

constexpr int kMaxPeers = 8;
constexpr int kMaxSeg = 5;

// dynamic shared-memory layout of the tile kernel, in doubles from the start of the extern array.  MT = number of
// 64-point tiles along each edge of the MACRO tile a CTA stages at once (1 or 2, see the kernel).
constexpr int kWLd = kTile + 2;                            // leading dimension of a warp's 8 x 64 mirror strip
template <int MT>
struct SmLayout {
    static constexpr int kMR = MT * kTile;                 // points along a macro-tile edge
    static constexpr int kRow = 0;                         // [2][kMR][P_N]  records (LDS.128 broadcast)
    static constexpr int kCol = kRow + 2 * P_N * kMR;      // [2][P_N][kMR]  field-major (conflict-free LDS.128)
    static constexpr int kExp = kCol + 2 * P_N * kMR;      // [kExpTab][kExpRep]
    static constexpr int kCtx = kExp + kExpTab * kExpRep;  // MaternCtx for the cold path (16 doubles reserved)
    static constexpr int kW = kCtx + 16;                   // [8 warps][8 rows][kWLd]  warp-private mirror strips
    static constexpr int kTab = kW + 8 * 8 * kWLd;         // [kTabCoef][kTabNint] (MODE_TABLE only)
};
static_assert(sizeof(MaternCtx) <= 16 * sizeof(double), "MaternCtx must fit its shared-memory slot");

// Cold paths of the tile kernel.  They take OFFSETS into the dynamic shared memory, not pointers: a generic
// pointer to shared memory costs an S2UR SR_CgaCtaId + address arithmetic where it is formed, which the compiler
// otherwise re-materialises inside the hot loop (ncu r01: 5.6 % of the stall samples).
template <int MODE, bool GEOM_F32>
__device__ __noinline__ double tile_pair_slow(int row_off, int col_off, int col_stride, int ctx_off) {
    extern __shared__ __align__(16) double smem[];
    const MaternCtx* m = reinterpret_cast<const MaternCtx*>(smem + ctx_off);
    return cov_pair_slow<MODE, GEOM_F32>(smem + row_off, smem + col_off, col_stride, m);
}

__device__ __noinline__ bool tile_beyond_max_dist(int row_off, int col_off, int col_stride, double max_dist) {
    extern __shared__ __align__(16) double smem[];
    return beyond_max_dist(smem + row_off, smem + col_off, col_stride, max_dist);
}

struct EllipseArgs {
    const double* aos;      // [npad + kTile][P_N] point records (rows of a tile: one 16 B cp.async per thread)
    const double* soa;      // [P_N][npad]         the same, field-major (columns of a tile)
    const double* table;
    void* out;
    int64_t N, npad, row0, nrows, ld;
    int tiles_r, tiles_c, tile_b0, symmetric;
    int64_t n_tri, n_tiles;
    // macro-tile enumeration of the MT = 2 kernel (128 x 128 blocks of 2 x 2 tiles)
    int mt_r, mt_c, mt_b0, m_tri;
    double max_dist;
    int vec2;
    GmgBesselConsts bc;
    // peer-to-peer symmetric mode (gmg_ellipse_cov_p2p_f64): rank `rank` of `world` owns rows
    // [blk_row0[rank], blk_row0[rank + 1]) and computes every pair of the matrix exactly once together with its
    // peers; the transposed tile is stored straight into the owner's row block over NVLink (peer_out).
    int world, rank;
    int nseg;                       // work segments after the n_tri diagonal-block tiles
    int64_t seg_id0[kMaxSeg + 1];   // first tile id of each segment (seg_id0[nseg] = n_tiles)
    int seg_blk[kMaxSeg];           // column block of the segment
    int seg_parity[kMaxSeg];        // -1: every tile; 0 / 1: only tiles with (ta_global + tb) % 2 == parity
    int64_t blk_row0[kMaxPeers + 1];
    void* peer_out[kMaxPeers];
};

template <typename OutT> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

template <typename OutT>
__device__ __forceinline__ void store_pair(OutT* p, double v0, double v1, bool second, bool vec2) {
    if (vec2 && second) {
        typename Vec2<OutT>::type v;
        v.x = (OutT)v0;
        v.y = (OutT)v1;
        *reinterpret_cast<typename Vec2<OutT>::type*>(p) = v;
    } else {
        p[0] = (OutT)v0;
        if (second) p[1] = (OutT)v1;
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- bulk-copy (TMA) + mbarrier helpers for the barrier-free record pipeline (K1_PIPE) ------------------------
#ifndef K1_PIPE
#define K1_PIPE 1
#endif
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "K1_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra K1_DONE;\n\t"
        "bra K1_WAIT;\n\t"
        "K1_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct TileId { int ta, tb; bool mirror; int peer; bool skip; };   // peer: block that receives the mirrored tile

// Tile `id` of the launch -> (row tile, column tile).  In symmetric mode (row0 aligned to the tile size) the
// first n_tri ids enumerate the upper triangle of the tiles_r x tiles_r block on the diagonal - rows r and
// tiles_r-1-r are folded into one strip of tiles_r+1 tiles, so the decode is one integer division - and
// those tiles are also written transposed; the remaining ids are the columns left and right of that block.
__device__ __forceinline__ TileId decode_tile(const EllipseArgs& a, int64_t id) {
    TileId t;
    t.mirror = false;
    t.peer = a.rank;
    t.skip = false;
    if (a.world > 1 && id >= a.n_tri) {
        // off-diagonal block segment: all tiles_r x (column tiles of the block), each mirrored into the peer
        int sgm = 0;
        while (sgm + 1 < a.nseg && id >= a.seg_id0[sgm + 1]) ++sgm;
        const int blk = a.seg_blk[sgm];
        const int ct0 = (int)(a.blk_row0[blk] / kTile);
        const int cts = (int)((a.blk_row0[blk + 1] - a.blk_row0[blk] + kTile - 1) / kTile);
        // Row tile fastest: CTAs that run concurrently (consecutive ids) then share the column tile, i.e. they
        // mirror into the SAME 64-row band of the peer's block.  Remote stores that are scattered over the whole
        // peer block run at ~40 GB/s (measured: every tile touches 16 different 2 MB pages of a mapping whose
        // translations are not cached like local ones); banded, they reach NVLink speed.
        const int64_t rest = id - a.seg_id0[sgm];
        if (a.seg_parity[sgm] < 0) {
            const int c = (int)(rest / a.tiles_r);
            t.ta = (int)(rest - (int64_t)c * a.tiles_r);
            t.tb = ct0 + c;
        } else {
            // half block: only tiles with (global row tile + column tile) of the given parity; per column tile
            // there are ceil(tiles_r / 2) candidates, the surplus one of an odd count is skipped
            const int per = (a.tiles_r + 1) / 2;
            const int c = (int)(rest / per);
            const int k = (int)(rest - (int64_t)c * per);
            t.tb = ct0 + c;
            t.ta = 2 * k + ((a.seg_parity[sgm] - (a.tile_b0 + t.tb)) & 1);
            t.skip = t.ta >= a.tiles_r;
        }
        (void)cts;
        t.mirror = true;
        t.peer = blk;
        return t;
    }
    if (a.symmetric) {
        if (id < a.n_tri) {
            const int T = a.tiles_r, w = T + 1;
            const int pr = (int)(id / w), k = (int)(id - (int64_t)pr * w);
            int r, off;
            if (k < T - pr) { r = pr; off = k; } else { r = T - 1 - pr; off = k - (T - pr); }
            t.ta = r;
            t.tb = a.tile_b0 + r + off;
            t.mirror = off > 0;
        } else {
            const int64_t rest = id - a.n_tri;
            const int wide = a.tiles_c - a.tiles_r;
            t.ta = (int)(rest / wide);
            const int c = (int)(rest - (int64_t)t.ta * wide);
            t.tb = c < a.tile_b0 ? c : c + a.tiles_r;
        }
    } else {
        t.ta = (int)(id / a.tiles_c);
        t.tb = (int)(id - (int64_t)t.ta * a.tiles_c);
    }
    return t;
}

#ifndef K1_MINB
#define K1_MINB 2
#endif
#ifndef K1_UNROLL
#define K1_UNROLL 1
#endif
constexpr int kRowsPerWarp = kTile / 8;

// What a CTA works on between two barriers: a macro tile of MT x MT 64-point tiles.
struct MacroId {
    int ta, tb;        // macro row (relative to the row block) and macro column (global), in units of MT tiles
    bool mirror;       // the tiles are also written transposed (inside the square block / into a peer's block)
    bool tri_diag;     // macro tile on the diagonal of the symmetric block: sub-tiles below the diagonal are skipped
    bool skip;
    int peer;
};

template <int MT>
__device__ __forceinline__ MacroId decode_macro(const EllipseArgs& a, int id) {
    MacroId t;
    t.mirror = false;
    t.tri_diag = false;
    t.skip = false;
    t.peer = a.rank;
    if (MT == 1) {
        const TileId s = decode_tile(a, id);
        t.ta = s.ta; t.tb = s.tb; t.mirror = s.mirror; t.skip = s.skip; t.peer = s.peer;
        return t;
    }
    // 32-bit arithmetic throughout (the host rejects launches with 2^31 or more tiles)
    if (a.symmetric) {
        if (id < a.m_tri) {
            // upper triangle of the mt_r x mt_r macro block on the diagonal; rows r and T-1-r share a strip of T+1
            const unsigned T = (unsigned)a.mt_r, w = T + 1;
            const unsigned pr = (unsigned)id / w, k = (unsigned)id - pr * w;
            unsigned r, off;
            if (k < T - pr) { r = pr; off = k; } else { r = T - 1 - pr; off = k - (T - pr); }
            t.ta = (int)r;
            t.tb = a.mt_b0 + (int)(r + off);
            t.mirror = true;
            t.tri_diag = off == 0;
        } else {
            const unsigned rest = (unsigned)(id - a.m_tri), wide = (unsigned)(a.mt_c - a.mt_r);
            const unsigned r = rest / wide, c = rest - r * wide;
            t.ta = (int)r;
            t.tb = (int)c < a.mt_b0 ? (int)c : (int)c + a.mt_r;
        }
    } else {
        const unsigned r = (unsigned)id / (unsigned)a.mt_c;
        t.ta = (int)r;
        t.tb = id - (int)r * a.mt_c;
    }
    return t;
}

// Persistent kernel: K1_MINB CTAs per SM, each looping over macro tiles id = blockIdx.x, += gridDim.x.
//
// A macro tile is MT x MT tiles of 64 x 64 entries (MT = 2 wherever the row block is 128-aligned; MT = 1 keeps the
// per-tile enumeration for 64-aligned blocks and for the peer-to-peer mode).  Per macro tile there is ONE barrier,
// ONE tile-id decode and ONE cp.async stage of the 2 x MT x 4 KB of per-point records of the NEXT macro tile into
// the other half of a double buffer (ncu r01, per-tile version: 21 % of the stall samples sat in the per-tile
// preamble and 16 % in the mirrored stores, with the FP64 pipe idle; the row loop itself ran the pipe at 80 %).
// The exp and Matern tables are staged once per CTA.
//
// Mirrored (transposed) stores are warp-private: a warp owns 8 rows x 64 columns of the tile, parks them in its own
// 8 x 66 strip of shared memory with one STS.128 per row, and writes the transposed 64 x 8 block as 64 B row pieces
// (four lanes x 16 B per mirrored row, two full 32 B sectors) - no CTA barrier, 5 instructions per store.
template <int MODE, bool MAXDIST, typename OutT, int MT>
__global__ void __launch_bounds__(kThreads, K1_MINB) ellipse_cov_tile_kernel(const EllipseArgs a) {
    using SL = SmLayout<MT>;
    constexpr int kMR = SL::kMR;
    constexpr bool GEOM_F32 = sizeof(OutT) == 4;
    using V2 = typename Vec2<OutT>::type;
    extern __shared__ __align__(16) double smem[];
    double* sexp = smem + SL::kExp;
    double* stab = smem + SL::kTab;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int jl = 2 * lane;                                        // first of this thread's two columns
    const int N = (int)a.N, row0 = (int)a.row0;                     // N < 2^31 - 256 (checked by the host)
    const int row_end = row0 + (int)a.nrows;
    const int n_ids = (int)a.n_tiles;
    OutT* out = reinterpret_cast<OutT*>(a.out);
    double* Ww = smem + SL::kW + warp * (8 * kWLd);                 // this warp's mirror strip

    auto stage = [&](int buf, const MacroId& t) {                  // 2 x MT x 16 B per thread
        const double* rsrc = a.aos + (a.row0 + (int64_t)t.ta * kMR) * P_N;
        const double* csrc = a.soa + (int64_t)t.tb * kMR;
        double* rdst = smem + SL::kRow + buf * (P_N * kMR);
        double* cdst = smem + SL::kCol + buf * (P_N * kMR);
#pragma unroll
        for (int h = 0; h < MT; ++h) {
            const int e = tid + h * kThreads;                       // 16-byte chunk, 4 * kMR of them per side
            cp_async16(rdst + 2 * e, rsrc + 2 * e);
            const int f = e / (kMR / 2), part = e % (kMR / 2);      // field f, points 2 part, 2 part + 1
            cp_async16(cdst + f * kMR + 2 * part, csrc + f * a.npad + 2 * part);
        }
    };

    // next macro tile of this CTA at or after `from` (ids advance by gridDim.x; tiles that a peer computes are skipped)
    auto next_tile = [&](int from, MacroId& tt) -> int {
        for (int i = from; i < n_ids; i += gridDim.x) {
            tt = decode_macro<MT>(a, i);
            if (!tt.skip) return i;
        }
        return n_ids;
    };
    // ---- record pipeline -------------------------------------------------------------------------------------
    // K1_PIPE: no CTA barrier in the steady state.  The records of macro tile i live in buffer i & 1 and are brought
    // in by bulk copies (cp.async.bulk, one 8 KB row block + 8 x 1 KB column fields) that complete on an mbarrier;
    // every warp walks the CTA's tile sequence on its own, waits on that mbarrier, and counts itself out of the
    // buffer when it is done with the tile - the LAST warp to leave refills the buffer with tile i + 2.  Warps drift
    // apart by up to a macro tile, so one warp's mirrored stores and sub-tile preambles overlap the other warps' FP64
    // work (ncu r01: with a CTA barrier per macro tile 6-8 % of the stall samples sat at the barrier).
    __shared__ __align__(8) unsigned long long full_bar[2];
    __shared__ int done_cnt[2];
    constexpr unsigned kStageBytes = 2u * P_N * kMR * sizeof(double);
    auto issue = [&](int buf, const MacroId& t) {                  // one thread
        const unsigned bar = smem_u32(&full_bar[buf]);
        const double* rsrc = a.aos + (a.row0 + (int64_t)t.ta * kMR) * P_N;
        const double* csrc = a.soa + (int64_t)t.tb * kMR;
        const unsigned rdst = smem_u32(smem + SL::kRow + buf * (P_N * kMR));
        const unsigned cdst = smem_u32(smem + SL::kCol + buf * (P_N * kMR));
        mbar_arrive_expect_tx(bar, kStageBytes);
        bulk_g2s(rdst, rsrc, P_N * kMR * (unsigned)sizeof(double), bar);
#pragma unroll
        for (int f = 0; f < P_N; ++f)
            bulk_g2s(cdst + f * kMR * (unsigned)sizeof(double), csrc + f * a.npad, kMR * (unsigned)sizeof(double), bar);
    };
    int id;
    {
        MacroId t0;
        id = next_tile(blockIdx.x, t0);
        if (id >= n_ids) return;
#if K1_PIPE
        if (tid == 0) {
            mbar_init(smem_u32(&full_bar[0]), 1);
            mbar_init(smem_u32(&full_bar[1]), 1);
            done_cnt[0] = done_cnt[1] = 0;
            fence_proxy_async();                                    // the barriers are visible to the copy engine
            issue(0, t0);
            MacroId t1;
            if (next_tile(id + gridDim.x, t1) < n_ids) issue(1, t1);
        }
#else
        stage(0, t0);
        cp_async_commit();
#endif
    }

    for (int e = tid; e < kExpTab * kExpRep; e += kThreads) sexp[e] = kExp2Neg[e / kExpRep];
    MaternCtx m;
    m.bc = a.bc;
    m.table = stab;
    if (MODE == MODE_TABLE) {
        for (int e = tid; e < kTabDoubles / 2; e += kThreads)
            reinterpret_cast<double2*>(stab)[e] = reinterpret_cast<const double2*>(a.table)[e];
    }
    if (tid == 0) *reinterpret_cast<MaternCtx*>(smem + SL::kCtx) = m;
    const double* sexp_lane = sexp + (lane & (kExpRep - 1));
#if K1_PIPE
    __syncthreads();                                  // tables, context and mbarriers are set up
#endif

    int buf = 0;
    for (int it = 0;; ++it) {
        int id_next;
#if K1_PIPE
        {
            MacroId tn;
            id_next = next_tile(id + gridDim.x, tn);
        }
        mbar_wait(smem_u32(&full_bar[buf]), (unsigned)(it >> 1) & 1u);   // this macro tile's records have landed
#else
        cp_async_wait<0>();                           // this macro tile's records have landed ...
        __syncthreads();                              // ... for every thread, and everybody left the previous one
        // The descriptor of the NEXT macro tile is only needed to issue its loads: it is decoded here and dropped
        // again, and this tile's own descriptor is re-decoded below (a 32-bit division) instead of being carried in
        // registers across the row loops (ncu s3a: the table variant spilled both descriptors and stalled on the
        // reloads at the head of every sub-tile).
        {
            MacroId tn;
            id_next = next_tile(id + gridDim.x, tn);
            if (id_next < n_ids) stage(buf ^ 1, tn);      // in flight while this macro tile is computed
            cp_async_commit();
        }
#endif
        const bool has_next = id_next < n_ids;
        const MacroId t = decode_macro<MT>(a, id);

        const int I0 = row0 + t.ta * kMR;             // global index of the macro tile's first row point
        const int J0 = t.tb * kMR;                    // global index of the macro tile's first column point
        // destination block of mirrored tiles: this rank's own rows, or - in the peer-to-peer mode - the row block
        // of the rank that owns those rows (stores go over NVLink into its memory)
        OutT* mbase = out;
        int mrow0 = row0, mlim = row_end;
        if (MT == 1 && a.world > 1 && t.peer != a.rank) {        // macro tiles (MT = 2) never leave the rank
            mbase = reinterpret_cast<OutT*>(a.peer_out[t.peer]);
            mrow0 = (int)a.blk_row0[t.peer];
            mlim = (int)a.blk_row0[t.peer + 1];
        }
        mlim = min(mlim, N);

        // `interior` = every sub-tile of the macro tile is a full 64 x 64 tile off the matrix diagonal (all but the
        // O(N / 128) macro tiles on the diagonal and on the ragged edge): its sub-tiles skip the classification
        // below, and vertically adjacent sub-tiles share the column registers (ncu s3a: the per-sub-tile preamble -
        // flag reloads, 7 LDS.128, pointer arithmetic - held 7 - 14 % of the stall samples with the FP64 pipe idle).
        const bool interior = !t.tri_diag && a.vec2 && (I0 + kMR <= row_end) && (J0 + kMR <= N) &&
                              (J0 - I0 >= kMR || I0 - J0 >= kMR) && (!t.mirror || J0 + kMR <= mlim);
        PointRegs c0, c1;
#pragma unroll 1
        for (int s = 0; s < MT * MT; ++s) {
            const int sb = s / MT, sa = s % MT;       // sub-tile (row, column) inside the macro tile, column-major
            const int i0 = I0 + sa * kTile, j0 = J0 + sb * kTile;
            if (!interior) {
                if (t.tri_diag && sa > sb) continue;      // below the diagonal: produced by mirroring
                if (i0 >= row_end || j0 >= N) continue;   // outside the row block / matrix (ragged edge)
            }
            const bool mirror = t.mirror && !(t.tri_diag && sa == sb);

            if (!interior || sa == 0) {
                const double* cp = smem + SL::kCol + buf * (P_N * kMR) + sb * kTile;
                auto ld2 = [&](int f, double& x0, double& x1) {
                    const double2 v = *reinterpret_cast<const double2*>(cp + f * kMR + jl);
                    x0 = v.x;
                    x1 = v.y;
                };
                ld2(P_LAT, c0.lat, c1.lat);   ld2(P_LON, c0.lon, c1.lon);   ld2(P_HCOS, c0.hcos, c1.hcos);
                ld2(P_H00, c0.h00, c1.h00);   ld2(P_H01, c0.h01, c1.h01);   ld2(P_H11, c0.h11, c1.h11);
                ld2(P_W, c0.w, c1.w);
            }
            // CTA-uniform tile classification, so that the per-row epilogue of the common case (a full tile off the
            // diagonal) is one vector store: `diag` = the tile touches the matrix diagonal, `full` = all 64 x 64
            // entries are inside the row block / matrix and the output admits 16 B (8 B for float) vector stores.
            const int dlt = j0 - i0;                  // column index - row index at the tile origin
            const bool diag = !interior && dlt > -kTile && dlt < kTile;
            const bool full = interior || ((i0 + kTile <= row_end) && (j0 + kTile <= N) && a.vec2);
            const int dl = diag ? dlt : (1 << 20);    // il == jl + dl  <=>  global row == global column
            const int rows_ok = min(kTile, row_end - i0);
            const int cols_ok = min(kTile, N - j0);
            OutT* const orow_start = out + (int64_t)(i0 - row0 + warp * kRowsPerWarp) * a.ld + (j0 + jl);

            // Row loop.  FAST = the common tile (full, off the diagonal, no max_dist): the epilogue is one vector
            // store and there is no per-row branch besides the cold-path check.  The next row's record is fetched
            // from shared memory while this row is computed (ncu r01b: 12 % of the stall samples sat on its first use).
            const int row_off0 = SL::kRow + buf * (P_N * kMR) + (sa * kTile + warp * kRowsPerWarp) * P_N;
            const int col_off = SL::kCol + buf * (P_N * kMR) + sb * kTile + jl;
            // SLOWALL = the cold re-run of the loop in which flagged lanes take the libm / Temme-CF2 path; it is taken
            // after the hot loop when any lane of the thread was flagged (x >= 700, x outside the Matern table, NaN -
            // in practice only the zero-distance lanes of diagonal tiles).  Keeping the call out of the hot loop
            // matters: with a call inside it the table variant spilled and ran 12 % slower (profiles/r01_ellipse_summary.md).
            bool slow_any = false;
            auto rows = [&](auto fast_tag, auto mirror_tag, auto slowall_tag) {
                constexpr bool FAST = decltype(fast_tag)::value, MIRROR = decltype(mirror_tag)::value;
                constexpr bool SLOWALL = decltype(slowall_tag)::value;
                OutT* orow = orow_start;
                PointRegs pnext = load_point(smem + row_off0);
                constexpr int kRowUnroll = K1_UNROLL;
#pragma unroll kRowUnroll
                for (int r = 0; r < kRowsPerWarp; ++r) {
                    const int il = warp * kRowsPerWarp + r;
                    const int row_off = row_off0 + r * P_N;
                    const PointRegs p = pnext;
                    if (r + 1 < kRowsPerWarp) pnext = load_point(smem + row_off + P_N);
                    bool slow0 = false, slow1 = false;
                    double v0, v1;
                    if (SLOWALL) {
                        // same fast evaluation again (bit-identical), then only the flagged lanes take the cold path:
                        // a value never depends on which tile or row block it was computed in
                        v0 = cov_pair<MODE, GEOM_F32, kExpRep>(p, c0, m, sexp_lane, slow0);
                        v1 = cov_pair<MODE, GEOM_F32, kExpRep>(p, c1, m, sexp_lane, slow1);
                        if (!(slow0 | slow1)) { orow += a.ld; continue; }
                        if (slow0) v0 = tile_pair_slow<MODE, GEOM_F32>(row_off, col_off, kMR, SL::kCtx);
                        if (slow1) v1 = tile_pair_slow<MODE, GEOM_F32>(row_off, col_off + 1, kMR, SL::kCtx);
                    } else {
#if K1_EXPERIMENT == 2      // stores only: no pair arithmetic (kernel-experiment builds, never shipped)
                        v0 = p.lat + c0.lat;
                        v1 = p.lat + c1.lat;
#else
                        v0 = cov_pair<MODE, GEOM_F32, kExpRep>(p, c0, m, sexp_lane, slow0);
                        v1 = cov_pair<MODE, GEOM_F32, kExpRep>(p, c1, m, sexp_lane, slow1);
#endif
                        slow_any |= slow0 | slow1;
                    }
                    if (MAXDIST) {                        // :256-266 pairs beyond max_dist stay 0
                        if (tile_beyond_max_dist(row_off, col_off, kMR, a.max_dist)) v0 = 0.0;
                        if (tile_beyond_max_dist(row_off, col_off + 1, kMR, a.max_dist)) v1 = 0.0;
                    }
                    if (!FAST && diag) {                  // diagonal = stdev^2, :246
                        const double var = smem[row_off + P_VAR];
                        if (il == jl + dl) v0 = var;
                        if (il == jl + 1 + dl) v1 = var;
                    }
#if K1_EXPERIMENT == 1      // arithmetic only: the store is kept reachable but never taken
                    if (v0 == 123.456789 && v1 == 987.654321)
#else
                    if (FAST || full)
#endif
                    {
                        V2 v;
                        v.x = (OutT)v0;
                        v.y = (OutT)v1;
                        *reinterpret_cast<V2*>(orow) = v;
                    } else if (il < rows_ok && jl < cols_ok) {
                        store_pair<OutT>(orow, v0, v1, jl + 1 < cols_ok, a.vec2 != 0);
                    }
                    orow += a.ld;
                    if (MIRROR) *reinterpret_cast<double2*>(Ww + r * kWLd + jl) = make_double2(v0, v1);
                }
            };
            using T_ = std::true_type;
            using F_ = std::false_type;
            if (interior) {
                if (mirror) rows(T_{}, T_{}, F_{}); else rows(T_{}, F_{}, F_{});
            } else {
                if (mirror) rows(F_{}, T_{}, F_{}); else rows(F_{}, F_{}, F_{});
            }
            if (slow_any) {                               // rare: redo this thread's 16 pairs on the cold path
                if (mirror) rows(F_{}, T_{}, T_{}); else rows(F_{}, F_{}, T_{});
            }
            if (mirror) {
                // ---- transposed block of this warp: rows j0 + 0..63 of C, columns i0 + 8 warp + 0..7 ------------
                // lane = 4 k + q stores entries 2q, 2q + 1 of mirrored row 8 it + k; the strip is read with a
                // leading dimension of 66 doubles, so a half warp hits 16 different 8-byte banks
                __syncwarp();
                const int k = lane >> 2, q = lane & 3;
                const double* wsrc = Ww + (2 * q) * kWLd + k;
                const int c0g = i0 + warp * kRowsPerWarp + 2 * q;     // global column of the first entry
                OutT* dst = mbase + (int64_t)(j0 + k - mrow0) * a.ld + c0g;
                const int64_t step = 8 * a.ld;
                if (interior || (full && j0 + kTile <= mlim)) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        V2 v;
                        v.x = (OutT)wsrc[8 * it];
                        v.y = (OutT)wsrc[kWLd + 8 * it];
                        *reinterpret_cast<V2*>(dst) = v;
                        dst += step;
                    }
                } else {
#pragma unroll 1
                    for (int it = 0; it < 8; ++it) {
                        if (j0 + 8 * it + k < mlim) {
                            if (c0g < row_end) dst[0] = (OutT)wsrc[8 * it];
                            if (c0g + 1 < row_end) dst[1] = (OutT)wsrc[kWLd + 8 * it];
                        }
                        dst += step;
                    }
                }
                __syncwarp();                             // strip free for the next sub-tile
            }
        }
#if K1_PIPE
        // leave the buffer; the last warp out refills it with the tile after next
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();                    // this warp's reads of the buffer are done
            if (atomicAdd(&done_cnt[buf], 1) == kThreads / 32 - 1) {
                done_cnt[buf] = 0;
                __threadfence_block();
                MacroId t2;
                if (has_next && next_tile(id_next + gridDim.x, t2) < n_ids) {
                    fence_proxy_async();              // generic-proxy reads before the async-proxy overwrite
                    issue(buf, t2);
                }
            }
        }
        __syncwarp();
#endif
        if (!has_next) break;
        id = id_next;
        buf ^= 1;
    }
#if !K1_PIPE
    cp_async_wait<0>();
#endif
}

template <int MODE>
__device__ __noinline__ double pairs_slow(const double* pa, const double* pb, const MaternCtx* m) {
    return cov_pair_slow<MODE, false>(pa, pb, 1, m);
}

template <int MODE>
__global__ void __launch_bounds__(256) ellipse_pairs_kernel(const double* __restrict__ params,
                                                            const double* __restrict__ table,
                                                            GmgBesselConsts bc, const int64_t* __restrict__ is,
                                                            const int64_t* __restrict__ js, int64_t n,
                                                            double* __restrict__ out) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    MaternCtx m;
    m.bc = bc;
    m.table = table;
    const double* pa = params + is[k] * P_N;
    const double* pb = params + js[k] * P_N;
    bool slow;
    double v = cov_pair<MODE, false, 1>(load_point(pa), load_point(pb), m, kExp2Neg, slow);   // as the tile kernel
    if (slow) v = pairs_slow<MODE>(pa, pb, &m);
    out[k] = v;
}

__global__ void __launch_bounds__(256) bessel_kv_kernel(GmgBesselConsts bc, const double* __restrict__ x,
                                                        int64_t n, int scaled, double* __restrict__ out) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    out[k] = scaled ? gmg_bessel_kve(bc, x[k]) : gmg_bessel_kv(bc, x[k]);
}

template <int MODE>
__global__ void __launch_bounds__(256) matern_kernel(GmgBesselConsts bc, const double* __restrict__ table,
                                                     const double* __restrict__ x, int64_t n,
                                                     double* __restrict__ out) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    MaternCtx m;
    m.bc = bc;
    m.table = table;
    out[k] = matern_scaled_eval<MODE>(m, x[k]) * exp(-x[k]);
}

__global__ void __launch_bounds__(256) haversine_kernel(const double* __restrict__ lat,
                                                        const double* __restrict__ lon, int64_t N,
                                                        int64_t row0, int64_t nrows, double radius,
                                                        double* __restrict__ D, int64_t ld) {
    // reference distances.py:191-218 (sklearn haversine_distances * 6371): asin form
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t i = row0 + blockIdx.y;
    if (j >= N || blockIdx.y >= nrows) return;
    const double sl = sin(0.5 * (lat[i] - lat[j])), so = sin(0.5 * (lon[i] - lon[j]));
    const double h = sl * sl + cos(lat[i]) * cos(lat[j]) * so * so;
    D[(int64_t)blockIdx.y * ld + j] = (2.0 * asin(sqrt(h))) * radius;    // sklearn: 2 asin(sqrt(h)), then "* radius"
}

int resolve_mode(double nu, int matern_mode) {
    if (matern_mode == GMG_MATERN_DIRECT) return MODE_DIRECT;
    if (matern_mode == GMG_MATERN_TABLE) return MODE_TABLE;
    if (nu == 0.5) return MODE_NU05;
    if (nu == 1.5) return MODE_NU15;
    if (nu == 2.5) return MODE_NU25;
    return (nu <= 10.0) ? MODE_TABLE : MODE_DIRECT;   // table accuracy degrades for large nu
}

// workspace layout: [Matern table, padded to 256 B][records (npad + 128) x 8 doubles][fields 8 x npad doubles]
constexpr size_t kTableBytes = ((size_t)kTabDoubles * sizeof(double) + 255) / 256 * 256;
inline int64_t pad_points(int64_t N) { return gmg_ceil_div(N, kPadTail) * kPadTail; }
inline size_t aos_bytes(int64_t N) { return gmg_align_up((size_t)(pad_points(N) + kPadTail) * P_N * sizeof(double), 256); }
inline size_t soa_bytes(int64_t N) { return gmg_align_up((size_t)pad_points(N) * P_N * sizeof(double), 256); }
inline const double* ws_table(const void* ws) { return reinterpret_cast<const double*>(ws); }
inline const double* ws_aos(const void* ws) {
    return reinterpret_cast<const double*>(reinterpret_cast<const char*>(ws) + kTableBytes);
}
inline const double* ws_soa(const void* ws, int64_t N) {
    return reinterpret_cast<const double*>(reinterpret_cast<const char*>(ws) + kTableBytes + aos_bytes(N));
}

template <int MODE, bool MAXDIST, typename OutT, int MT>
int launch_tiles(const EllipseArgs& a, int64_t n_ctas, cudaStream_t s) {
    size_t smem = (size_t)SmLayout<MT>::kTab * sizeof(double);
    if (MODE == MODE_TABLE) smem += (size_t)kTabDoubles * sizeof(double);
    auto kern = ellipse_cov_tile_kernel<MODE, MAXDIST, OutT, MT>;
    GMG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent grid: K1_MINB resident CTAs per SM (cached per device)
    static int sm_count[64] = {0};
    int dev = 0;
    GMG_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return GMG_ERR_UNSUPPORTED;
    if (sm_count[dev] == 0)
        GMG_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
    const int64_t grid = n_ctas < (int64_t)K1_MINB * sm_count[dev] ? n_ctas : (int64_t)K1_MINB * sm_count[dev];
    kern<<<(unsigned)grid, kThreads, smem, s>>>(a);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

// Chooses the enumeration: macro tiles of 2 x 2 tiles wherever the row block allows it (single-rank launches whose
// block starts on a 128-row boundary, or that do not use symmetry at all); per-tile otherwise.
template <int MODE, bool MAXDIST, typename OutT>
int launch_tiles_mt(EllipseArgs a, cudaStream_t s) {
    if (a.N >= (int64_t)2147483647 - 4 * kPadTail) return GMG_ERR_UNSUPPORTED;
#ifdef K1_FORCE_MT1         // kernel-experiment builds: per-tile enumeration everywhere
    const bool macro = false;
#else
    const bool macro = a.world == 1 && (!a.symmetric || a.row0 % (2 * kTile) == 0);
#endif
    if (macro) {
        a.mt_r = (a.tiles_r + 1) / 2;
        a.mt_c = (a.tiles_c + 1) / 2;
        a.mt_b0 = a.tile_b0 / 2;
        const int64_t m_tri = (int64_t)a.mt_r * (a.mt_r + 1) / 2;
        const int64_t n = a.symmetric ? m_tri + (int64_t)a.mt_r * (a.mt_c - a.mt_r) : (int64_t)a.mt_r * a.mt_c;
        if (n >= 2147483647) return GMG_ERR_UNSUPPORTED;
        a.m_tri = (int)m_tri;
        a.n_tiles = n;
        return launch_tiles<MODE, MAXDIST, OutT, 2>(a, n, s);
    }
    if (a.n_tiles >= 2147483647) return GMG_ERR_UNSUPPORTED;
    return launch_tiles<MODE, MAXDIST, OutT, 1>(a, a.n_tiles, s);
}

template <int MODE>
int launch_tiles_mode(const EllipseArgs& a, int64_t n, int f32, cudaStream_t s) {
    (void)n;
    const bool md = a.max_dist >= 0.0;
    if (f32) return md ? launch_tiles_mt<MODE, true, float>(a, s) : launch_tiles_mt<MODE, false, float>(a, s);
    return md ? launch_tiles_mt<MODE, true, double>(a, s) : launch_tiles_mt<MODE, false, double>(a, s);
}

template <int MODE>
int launch_pairs(const double* params, const double* table, const GmgBesselConsts& bc, const int64_t* i,
                 const int64_t* j, int64_t n, double* out, cudaStream_t s) {
    ellipse_pairs_kernel<MODE><<<(unsigned)gmg_ceil_div(n, 256), 256, 0, s>>>(params, table, bc, i, j, n, out);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int launch_for_nu(const EllipseArgs& a, double nu, int matern_mode, int store_f32, cudaStream_t s) {
    switch (resolve_mode(nu, matern_mode)) {
        case MODE_NU05: return launch_tiles_mode<MODE_NU05>(a, a.n_tiles, store_f32, s);
        case MODE_NU15: return launch_tiles_mode<MODE_NU15>(a, a.n_tiles, store_f32, s);
        case MODE_NU25: return launch_tiles_mode<MODE_NU25>(a, a.n_tiles, store_f32, s);
        case MODE_TABLE: return launch_tiles_mode<MODE_TABLE>(a, a.n_tiles, store_f32, s);
        default: return launch_tiles_mode<MODE_DIRECT>(a, a.n_tiles, store_f32, s);
    }
}

}  // namespace

extern "C" {

int gmg_ellipse_ws_bytes(int64_t N, size_t* bytes) {
    if (N <= 0) return -1;
    if (!bytes) return -2;
    *bytes = kTableBytes + aos_bytes(N) + soa_bytes(N);
    return GMG_OK;
}

int gmg_ellipse_prepare_f64(const double* Lx, const double* Ly, const double* theta, const double* stdev,
                            const double* lat_rad, const double* lon_rad, int64_t N, double nu, int disp_method,
                            void* ws, size_t ws_bytes, gmg_stream_t stream) {
    if (!Lx) return -1;
    if (!Ly) return -2;
    if (!theta) return -3;
    if (!stdev) return -4;
    if (!lat_rad) return -5;
    if (!lon_rad) return -6;
    if (N <= 0) return -7;
    if (!(nu > 0.0)) return -8;
    if (disp_method != GMG_DISP_MET_OFFICE && disp_method != GMG_DISP_MODIFIED_MET_OFFICE) return -9;
    if (!ws) return -10;
    size_t need;
    gmg_ellipse_ws_bytes(N, &need);
    if (ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 15)) return GMG_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    double* table = const_cast<double*>(ws_table(ws));
    const int64_t npad = pad_points(N);
    const GmgBesselConsts bc = gmg_make_bessel_consts(nu);
    ellipse_prepare_kernel<<<(unsigned)gmg_ceil_div(npad + kPadTail, 256), 256, 0, s>>>(
        Lx, Ly, theta, stdev, lat_rad, lon_rad, N, npad, disp_method, bc.two_sqrt_nu * kEarthRadiusKm,
        const_cast<double*>(ws_aos(ws)), const_cast<double*>(ws_soa(ws, N)));
    GMG_LAUNCH_CHECK();
    matern_table_kernel<<<kTabNint, 32, 0, s>>>(bc, table);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_ellipse_cov_f64(const void* ws, int64_t N, int64_t row0, int64_t nrows, double nu, double max_dist_km,
                        int matern_mode, void* C_rows, int64_t ld, int store_f32, gmg_stream_t stream) {
    if (!ws) return -1;
    if (N <= 0) return -2;
    if (row0 < 0 || row0 >= N) return -3;
    if (nrows <= 0 || row0 + nrows > N) return -4;
    if (!(nu > 0.0)) return -5;
    if (matern_mode < 0 || matern_mode > 2) return -7;
    if (!C_rows) return -8;
    if (ld < N) return -9;
    EllipseArgs a;
    a.aos = ws_aos(ws);
    a.soa = ws_soa(ws, N);
    a.npad = pad_points(N);
    a.table = ws_table(ws);
    a.out = C_rows;
    a.N = N; a.row0 = row0; a.nrows = nrows; a.ld = ld;
    a.tiles_r = (int)gmg_ceil_div(nrows, kTile);
    a.tiles_c = (int)gmg_ceil_div(N, kTile);
    a.symmetric = (row0 % kTile == 0) ? 1 : 0;
    a.tile_b0 = (int)(row0 / kTile);
    a.n_tri = (int64_t)a.tiles_r * (a.tiles_r + 1) / 2;
    a.max_dist = (max_dist_km >= 0.0) ? max_dist_km : -1.0;
    const size_t elt = store_f32 ? sizeof(float) : sizeof(double);
    a.vec2 = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(C_rows) % (2 * elt)) == 0);
    a.bc = gmg_make_bessel_consts(nu);
    a.world = 1; a.rank = 0; a.nseg = 0;
    a.mt_r = a.mt_c = a.mt_b0 = a.m_tri = 0;
    const int64_t n_ctas = a.symmetric ? a.n_tri + (int64_t)a.tiles_r * (a.tiles_c - a.tiles_r)
                                       : (int64_t)a.tiles_r * a.tiles_c;
    a.n_tiles = n_ctas;
    return launch_for_nu(a, nu, matern_mode, store_f32, (cudaStream_t)stream);
}

int gmg_ellipse_cov_p2p_f64(const void* ws, int64_t N, double nu, double max_dist_km, int matern_mode, int world,
                            int rank, const int64_t* h_block_row0, void* const* h_block_rows, int64_t ld,
                            int store_f32, gmg_stream_t stream) {
    if (!ws) return -1;
    if (N <= 0) return -2;
    if (!(nu > 0.0)) return -3;
    if (matern_mode < 0 || matern_mode > 2) return -5;
    if (world < 2 || world > kMaxPeers) return -6;
    if (rank < 0 || rank >= world) return -7;
    if (!h_block_row0 || h_block_row0[0] != 0 || h_block_row0[world] != N) return -8;
    if (!h_block_rows) return -9;
    if (ld < N) return -10;
    EllipseArgs a;
    for (int g = 0; g < world; ++g) {
        if (h_block_row0[g + 1] <= h_block_row0[g] || h_block_row0[g] % kTile != 0) return -8;   // tile-aligned blocks
        if (!h_block_rows[g]) return -9;
        a.blk_row0[g] = h_block_row0[g];
        a.peer_out[g] = h_block_rows[g];
    }
    a.blk_row0[world] = N;
    a.aos = ws_aos(ws);
    a.soa = ws_soa(ws, N);
    a.npad = pad_points(N);
    a.table = ws_table(ws);
    a.out = h_block_rows[rank];
    a.N = N; a.row0 = h_block_row0[rank]; a.nrows = h_block_row0[rank + 1] - h_block_row0[rank]; a.ld = ld;
    a.tiles_r = (int)gmg_ceil_div(a.nrows, kTile);
    a.tiles_c = (int)gmg_ceil_div(N, kTile);
    a.symmetric = 1;
    a.tile_b0 = (int)(a.row0 / kTile);
    a.n_tri = (int64_t)a.tiles_r * (a.tiles_r + 1) / 2;
    a.max_dist = (max_dist_km >= 0.0) ? max_dist_km : -1.0;
    const size_t elt = store_f32 ? sizeof(float) : sizeof(double);
    a.vec2 = (ld % 2 == 0);
    for (int g = 0; g < world; ++g) a.vec2 = a.vec2 && ((reinterpret_cast<uintptr_t>(h_block_rows[g]) % (2 * elt)) == 0);
    a.bc = gmg_make_bessel_consts(nu);
    a.world = world; a.rank = rank;
    a.mt_r = a.mt_c = a.mt_b0 = a.m_tri = 0;
    // Every unordered pair of row blocks {r, c} is computed by exactly one of the two ranks: rank r takes the
    // blocks at cyclic distance 1 .. (world-1)/2 to its right; at distance world/2 (even world) the two ranks
    // split the block's tiles by the parity of (row tile + column tile), rank min(r, c) taking the even ones.
    a.nseg = 0;
    int64_t id0 = a.n_tri;
    for (int dlt = 1; dlt <= world / 2; ++dlt) {
        const int c = (rank + dlt) % world;
        const bool half = (world % 2 == 0) && (dlt == world / 2);
        if (a.nseg >= kMaxSeg) return GMG_ERR_UNSUPPORTED;
        a.seg_id0[a.nseg] = id0;
        a.seg_blk[a.nseg] = c;
        a.seg_parity[a.nseg] = half ? (rank < c ? 0 : 1) : -1;
        const int64_t cts = gmg_ceil_div(h_block_row0[c + 1] - h_block_row0[c], kTile);
        id0 += (half ? (int64_t)(a.tiles_r + 1) / 2 : (int64_t)a.tiles_r) * cts;
        ++a.nseg;
    }
    a.seg_id0[a.nseg] = id0;
    a.n_tiles = id0;
    return launch_for_nu(a, nu, matern_mode, store_f32, (cudaStream_t)stream);
}

// ---- CUDA IPC plumbing for the peer-to-peer mode (one process per GPU) -------------------------------------------
// Torch's caching allocator sub-allocates cudaMalloc blocks, and an IPC handle always names a whole block, so the
// row blocks that peers write into are allocated here (the one exception to "the library never allocates").
int gmg_ipc_alloc_f64(size_t bytes, void** dev_ptr, void* handle64_out) {
    if (bytes == 0) return -1;
    if (!dev_ptr) return -2;
    if (!handle64_out) return -3;
    GMG_CUDA_CHECK(cudaMalloc(dev_ptr, bytes));
    cudaIpcMemHandle_t h;
    GMG_CUDA_CHECK(cudaIpcGetMemHandle(&h, *dev_ptr));
    memcpy(handle64_out, &h, sizeof(h));
    return GMG_OK;
}

int gmg_ipc_free_f64(void* dev_ptr) {
    if (!dev_ptr) return -1;
    GMG_CUDA_CHECK(cudaFree(dev_ptr));
    return GMG_OK;
}

int gmg_ipc_open_f64(const void* handle64, void** dev_ptr) {
    if (!handle64) return -1;
    if (!dev_ptr) return -2;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    GMG_CUDA_CHECK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GMG_OK;
}

int gmg_ipc_close_f64(void* dev_ptr) {
    if (!dev_ptr) return -1;
    GMG_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
    return GMG_OK;
}

int gmg_ellipse_pairs_f64(const void* ws, int64_t N, const int64_t* i, const int64_t* j, int64_t npairs,
                          double nu, int matern_mode, double* out, gmg_stream_t stream) {
    if (!ws) return -1;
    if (N <= 0) return -2;
    if (!i) return -3;
    if (!j) return -4;
    if (npairs < 0) return -5;
    if (!(nu > 0.0)) return -6;
    if (matern_mode < 0 || matern_mode > 2) return -7;
    if (!out) return -8;
    if (npairs == 0) return GMG_OK;
    const double* params = ws_aos(ws);
    const double* table = ws_table(ws);
    const GmgBesselConsts bc = gmg_make_bessel_consts(nu);
    cudaStream_t s = (cudaStream_t)stream;
    switch (resolve_mode(nu, matern_mode)) {
        case MODE_NU05: return launch_pairs<MODE_NU05>(params, table, bc, i, j, npairs, out, s);
        case MODE_NU15: return launch_pairs<MODE_NU15>(params, table, bc, i, j, npairs, out, s);
        case MODE_NU25: return launch_pairs<MODE_NU25>(params, table, bc, i, j, npairs, out, s);
        case MODE_TABLE: return launch_pairs<MODE_TABLE>(params, table, bc, i, j, npairs, out, s);
        default: return launch_pairs<MODE_DIRECT>(params, table, bc, i, j, npairs, out, s);
    }
}

int gmg_bessel_kv_f64(double nu, const double* x, int64_t n, int scaled, double* out, gmg_stream_t stream) {
    if (!(nu >= 0.0)) return -1;
    if (!x) return -2;
    if (n < 0) return -3;
    if (!out) return -5;
    if (n == 0) return GMG_OK;
    bessel_kv_kernel<<<(unsigned)gmg_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
        gmg_make_bessel_consts(nu), x, n, scaled, out);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_matern_f64(const void* ws, double nu, int matern_mode, const double* x, int64_t n, double* out,
                   gmg_stream_t stream) {
    if (!(nu > 0.0)) return -2;
    if (matern_mode < 0 || matern_mode > 2) return -3;
    if (!x) return -4;
    if (n < 0) return -5;
    if (!out) return -6;
    if (n == 0) return GMG_OK;
    const int mode = resolve_mode(nu, matern_mode);
    if (mode == MODE_TABLE && !ws) return -1;
    const double* table = ws ? ws_table(ws) : nullptr;   // the table sits at the start of any ellipse workspace
    const GmgBesselConsts bc = gmg_make_bessel_consts(nu);
    const unsigned blocks = (unsigned)gmg_ceil_div(n, 256);
    cudaStream_t s = (cudaStream_t)stream;
    switch (mode) {
        case MODE_NU05: matern_kernel<MODE_NU05><<<blocks, 256, 0, s>>>(bc, table, x, n, out); break;
        case MODE_NU15: matern_kernel<MODE_NU15><<<blocks, 256, 0, s>>>(bc, table, x, n, out); break;
        case MODE_NU25: matern_kernel<MODE_NU25><<<blocks, 256, 0, s>>>(bc, table, x, n, out); break;
        case MODE_TABLE: matern_kernel<MODE_TABLE><<<blocks, 256, 0, s>>>(bc, table, x, n, out); break;
        default: matern_kernel<MODE_DIRECT><<<blocks, 256, 0, s>>>(bc, table, x, n, out); break;
    }
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

int gmg_haversine_f64(const double* lat_rad, const double* lon_rad, int64_t N, int64_t row0, int64_t nrows,
                      double radius, double* D_rows, int64_t ld, gmg_stream_t stream) {
    if (!lat_rad) return -1;
    if (!lon_rad) return -2;
    if (N <= 0) return -3;
    if (row0 < 0 || row0 >= N) return -4;
    if (nrows <= 0 || row0 + nrows > N || nrows > 65535) return -5;
    if (!D_rows) return -7;
    if (ld < N) return -8;
    dim3 grid((unsigned)gmg_ceil_div(N, 256), (unsigned)nrows);
    haversine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(lat_rad, lon_rad, N, row0, nrows, radius, D_rows, ld);
    GMG_LAUNCH_CHECK();
    return GMG_OK;
}

}  // extern "C"
