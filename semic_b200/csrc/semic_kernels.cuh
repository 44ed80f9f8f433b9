// semic_kernels.cuh -- the SEMIC point-day physics and the persistent multi-day kernels (sm_100a).
//
// Included twice, both translation units compiled with -fmad=false:
//   SEMIC_STRICT=1 (semic_kernels_strict.cu): the reference's operation order, IEEE division, no fused operations; only
//     the libm calls exp/acos/pow/tanh differ from a gfortran build, by the 1-2 ulp of libdevice vs glibc;
//   SEMIC_STRICT=0 (semic_kernels_fast.cu, the benchmarked arithmetic): reciprocal multiplies, table-driven exp, one-
//     polynomial acos, every fused operation written as fma() so that all instantiations round alike.
//
// What it replaces: surface_energy_and_mass_balance (surface_physics.f90:138-153), i.e. n_ksub x energy_balance
// (:158-214) + mass_balance (:232-374) with the elementals (:378-560), fused per grid point so that every point-day
// costs 9 forcing loads and (optionally) 8 output stores instead of ~47 array sweeps per sub-step.
//
// Kernel shape (k_run_lean, k_run): persistent CTAs, one per SM.  Every WARP owns tiles of 32 consecutive points and
// walks each tile through ALL days with the prognostic state (tsurf, hsnow, hice, alb, alb_snow, qmr_res) in registers.
// Each warp feeds itself: one elected lane issues a 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) of the next day's
// forcing rows of the tile into the warp's private kStages-deep shared-memory ring, completion on an mbarrier per stage;
// lanes read their column with conflict-free LDS where first needed, the warp re-arms the stage after __syncwarp, and the
// daily diagnostics are stored coalesced.  No producer warp, no cross-warp coupling.  No tensor cores: nothing here is a
// contraction.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "semic_internal.h"

#ifndef SEMIC_STRICT
#error "define SEMIC_STRICT to 0 or 1 before including semic_kernels.cuh"
#endif

namespace semic {
namespace SEMIC_MODE_NS {

#define DEVI __device__ __forceinline__

// constants, surface_physics.f90:14-25
constexpr double c_pi    = 3.141592653589793238462643;
constexpr double c_t0    = 273.15;
constexpr double c_sigm  = 5.67e-8;
constexpr double c_eps   = 0.62197;
constexpr double c_cls   = 2.83e6;
constexpr double c_clv   = 2.5e6;
constexpr double c_rhow  = 1000.0;
constexpr double c_hsmax = 5.0;

// ------------------------------------------------------------------------------------------------
// arithmetic helpers
// ------------------------------------------------------------------------------------------------
// Literal doubles of the hot loop live in the constant bank so that they are direct c[3][..]
// operands of DFMA/DMUL; as immediates ptxas re-materialises each one with two UMOVs per use.
__device__ __constant__ double kc[26] = {
    6755399441055744.0,           //  0  1.5*2^52
    1.4426950408889634,           //  1  log2(e)
    -6.93147180369123816490e-01,  //  2  -ln2 hi
    -1.90821492927058770002e-10,  //  3  -ln2 lo
    2.5022322536502990e-08,       //  4  exp polynomial, degree 11 .. 2
    2.7630903488173108e-07,       //  5
    2.7557514545882439e-06,       //  6
    2.4801491039099165e-05,       //  7
    1.9841269589115497e-04,       //  8
    1.3888888945916380e-03,       //  9
    8.3333333334550432e-03,       // 10
    4.1666666666519754e-02,       // 11
    1.6666666666666477e-01,       // 12
    5.0000000000000122e-01,       // 13
    273.15,                       // 14  t0
    611.2 * 0.62197,              // 15  611.2*eps
    611.2 * (0.62197 - 1.0),      // 16  611.2*(eps-1)
    5.67e-8,                      // 17  sigm
    22.46, 272.62,                // 18 19  ei_sat a, b
    17.62, 243.12,                // 20 21  ew_sat a, b
    2.83e6, 2.5e6,                // 22 23  cls, clv
    -22.46, -17.62                // 24 25  -a of ei_sat, ew_sat
};
// dmax1/dmin1 as compare-select (3 instructions; IEEE fmax/fmin costs 6 on sm_100a, there is no DMNMX).
// Identical to fmax/fmin on ordered operands; with a NaN operand the second argument is returned.
DEVI double dmax1(double a, double b) { return a > b ? a : b; }
DEVI double dmin1(double a, double b) { return a < b ? a : b; }
// 2^(j/2048), j = 0..2047: filled once per device by the host (exp2l, rounded to double), staged into shared memory by
// the kernels that use fexp_t
constexpr int kExpTab = 2048;
#if !SEMIC_STRICT
__device__ double g_exp2_tab[kExpTab];
#endif
// constants of the table-driven exp: 2048/ln2, 1.5*2^52, -(ln2/2048) hi (low 22 mantissa bits zero) and lo, Taylor 1/2, 1/6
__device__ __constant__ double kt[6] = {2954.639443740597, 6755399441055744.0, -0.0003384507715509244, -2.0686139481490644e-13,
                                        0.5, 0.16666666666666666};
#if SEMIC_STRICT
DEVI double fdiv(double n, double d) { return n / d; }    // CUDA double division is IEEE round-to-nearest
DEVI double fexp(double x) { return exp(x); }
#else
// reciprocal from the MUFU.RCP64H seed (it reads the upper 32 bits only: relative error e ~ 2^-20) refined by
// r0*(1 + e + e^2), which leaves e^3 ~ 2^-60, below the rounding of the result; 3 DFMA.  Operands here are normal and
// far from 0/inf (temperatures, pressures, positive parameters).
DEVI double frcp(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}
DEVI double fdiv(double n, double d) { return n * frcp(d); }
// n/d correctly rounded from the correctly rounded reciprocal rd = RN(1/d) (Markstein): q0 = RN(n*rd),
// q = RN(q0 + RN(n - q0*d)*rd); 3 FP64 operations, no special cases (finite, normal operands)
DEVI double fdiv_rn(double n, double d, double rd)
{
    const double q0 = n * rd;
    return fma(fma(-q0, d, n), rd, q0);
}
// exp for |x| < 700 without special-case handling: Cody-Waite reduction by ln2 and the
// degree-11 minimax polynomial (max rel. error 4e-18 before rounding).
DEVI double fexp(double x)
{
    const double magic = kc[0];                           // 1.5 * 2^52
    double t = fma(x, kc[1], magic);
    const int k = __double2loint(t);
    const double kf = t - magic;
    double r = fma(kf, kc[2], x);
    r = fma(kf, kc[3], r);
    // Estrin evaluation: dependent depth 5 instead of 11 (the sub-step loop is latency bound, not issue bound)
    const double r2 = r * r;
    const double p01 = r + 1.0;                       // c0 + c1 r          (c0 = c1 = 1)
    const double p23 = fma(kc[12], r, kc[13]);        // c2 + c3 r
    const double p45 = fma(kc[10], r, kc[11]);        // c4 + c5 r
    const double p67 = fma(kc[8], r, kc[9]);          // c6 + c7 r
    const double p89 = fma(kc[6], r, kc[7]);          // c8 + c9 r
    const double pab = fma(kc[4], r, kc[5]);          // c10 + c11 r
    const double r4 = r2 * r2;
    const double q0 = fma(p23, r2, p01);              // c0..c3
    const double q1 = fma(p67, r2, p45);              // c4..c7
    const double q2 = fma(pab, r2, p89);              // c8..c11
    const double r8 = r4 * r4;
    const double h0 = fma(q1, r4, q0);                // c0..c7
    double p = fma(q2, r8, h0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
// Table-driven exp for the hot loop: exp(x) = 2^k * 2^(j/2048) * p(r), n = rint(x*2048/ln2) = 2048k + j,
// |r| <= ln2/4096, p = degree-3 Taylor (truncation r^4/24 <= 3.4e-17).  8 FP64 operations; the table lives in shared
// memory (LDS with a lane-varying index).  |x| < 700, no specials.
DEVI double fexp_t(double x, const double *tab)
{
    const double t = fma(x, kt[0], kt[1]);
    const int n = __double2loint(t);
    const double nf = t - kt[1];
    double r = fma(nf, kt[2], x);
    r = fma(nf, kt[3], r);
    const double r2 = r * r;
    const double p23 = fma(kt[5], r, kt[4]);
    const double em1 = fma(p23, r2, r);                           // e^r - 1
    const double tj = tab[n & (kExpTab - 1)];
    const double p = fma(em1, tj, tj);
    // scale by 2^(n >> 11): one shift and one multiply-add on the high word (the compiler's own form is shift, mask, add)
    int hi;
    asm("{\n.reg .s32 k;\nshr.s32 k, %1, 11;\nmad.lo.s32 %0, k, 1048576, %2;\n}" : "=r"(hi) : "r"(n), "r"(__double2hiint(p)));
    return __hiloint2double(hi, __double2loint(p));
}
// asin(s) = s + s*z*R(z), z = s*s <= 1/4: degree-11 interpolant of R at the Chebyshev nodes of [0, 1/4]
// (relative error of asin 5.6e-17 before rounding); coefficients highest degree first
__device__ __constant__ double ka[14] = {
    0.028169218060881414, -0.010749050339697808, 0.01603551434914882, 0.0078029494773533175,
    0.011875494382636922, 0.013929652902326633, 0.017355259955786323, 0.02237204763174451,
    0.03038194736709848, 0.044642857103423646, 0.07500000000020764, 0.1666666666666665,
    3.141592653589793238462643, 1.5707963267948966};   // 12, 13: pi, pi/2
// sqrt(z) for a normal z > 0: MUFU.RSQ64H seed, two coupled Newton steps, one residual correction
DEVI double fsqrt_pos(double z)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(z));
    double g = z * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    return fma(fma(-g, g, z), h, g);
}
// acos(x) for |x| < 1 (~1 ulp).  One polynomial serves both ranges, so a warp does not diverge on |x|:
// |x| <= 1/2: acos = pi/2 -+ asin(|x|);  |x| > 1/2: acos(|x|) = 2 asin(sqrt((1-|x|)/2)), acos(-|x|) = pi - acos(|x|).
DEVI double facos(double x)
{
    const double ax = fabs(x);
    const bool big = ax > 0.5;
    const bool neg = x < 0.0;
    double z, sv;
    if (big) { z = fma(ax, -0.5, 0.5); sv = fsqrt_pos(z); }       // (1-|x|)/2 is exact for |x| >= 1/2
    else     { z = x * x; sv = ax; }
    double p = ka[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) p = fma(p, z, ka[i]);
    const double r = fma(sv * z, p, sv);                            // asin(sv)
    const double r2 = big ? r + r : r;
    const double base = big ? (neg ? ka[12] : 0.0) : ka[13];
    return (neg != big) ? base + r2 : base - r2;
}
#endif

// ------------------------------------------------------------------------------------------------
// elementals (also exported one by one through the C-ABI for the known-answer tests)
// ------------------------------------------------------------------------------------------------

// sensible_heat_flux, surface_physics.f90:378-389
DEVI double sensible_heat_flux(double ts, double ta, double wind, double rhoa, double csh, double cap)
{
    return (((csh * cap) * rhoa) * wind) * (ts - ta);
}

// ei_sat / ew_sat, surface_physics.f90:546-560, selected by ts < t0 as latent_heat_flux does (:415)
DEVI double esat(double ts, bool ice)
{
#if SEMIC_STRICT
    const double a = ice ? 22.46 : 17.62;
    const double b = ice ? 272.62 : 243.12;
    return 611.2 * fexp(fdiv(a * (ts - c_t0), (b + ts) - c_t0));
#else
    const double a = ice ? kc[18] : kc[20];
    const double b = ice ? kc[19] : kc[21];
    const double x = ts - kc[14];
    return 611.2 * fexp((a * x) * frcp(b + x));
#endif
}

// latent_heat_flux, surface_physics.f90:393-430; returns q = subl (ice) or evap (water) in kg/(s m2)
#if SEMIC_STRICT
DEVI double latent_q(double ts, double wind_clh_rho, double shum, double sp, bool ice)
{
    const double es = esat(ts, ice);
    const double shum_sat = fdiv(es * c_eps, (es * (c_eps - 1.0)) + sp);
    return wind_clh_rho * (shum_sat - shum);
}
#else
// Fast form.  With em = exp(-a x/(b+x)) = 611.2/es:  es*eps/(es*(eps-1)+sp) = 611.2*eps / (611.2*(eps-1) + sp*em), so
// q = A/(611.2*(eps-1) + sp*em) - Bq with the day constants A = wind_clh_rho*611.2*eps, Bq = wind_clh_rho*shum.
// ALLG: the caller guarantees ts <= t0 (the point is gated, :201): at ts == t0 the argument a*x/(b+x) is 0 for either
// pair (a,b), so the ice constants serve both branches bit for bit and the two selects disappear.
template <bool ALLG>
DEVI double latent_q(double ts, double A, double Bq, double sp, bool ice, const double *tab)
{
    const double na = (ALLG || ice) ? kc[24] : kc[25];
    const double b = (ALLG || ice) ? kc[19] : kc[21];
    const double x = ts - kc[14];
    const double em = fexp_t((na * x) * frcp(b + x), tab);
    return fma(frcp(fma(sp, em, kc[16])), A, -Bq);
}
// warm = (raw >= t0) as ONE signed 64-bit compare of the bit patterns, feeding both selects: tsurf of the next sub-step
// (clamped at t0 when gated) and the latent heat it will use.  Written in PTX so that the two selects share the predicate
// (left to itself the compiler turns the first into an integer minimum and re-issues the compare for the second).
DEVI void clamp_and_latent(const double raw, const long long t0_bits, const double t0, const bool gate, const double cls_,
                           const double clv_, double &ts, double &lat)
{
    asm("{\n"
        ".reg .pred p, g, c;\n"
        "setp.ge.s64 p, %2, %3;\n"
        "setp.ne.u32 g, %4, 0;\n"
        "and.pred c, p, g;\n"
        "selp.f64 %0, %5, %6, c;\n"
        "selp.f64 %1, %7, %8, p;\n"
        "}\n"
        : "=d"(ts), "=d"(lat)
        : "l"(__double_as_longlong(raw)), "l"(t0_bits), "r"((unsigned)gate), "d"(t0), "d"(raw), "d"(clv_), "d"(cls_));
}
// ... and the pair (-a, b) of ew_sat / ei_sat for a warp that has lanes above the melting point
DEVI void clamp_and_latent_ab(const double raw, const long long t0_bits, const double t0, const bool gate, const double cls_,
                              const double clv_, const double nai, const double naw, const double bi, const double bw,
                              double &ts, double &lat, double &na, double &b)
{
    asm("{\n"
        ".reg .pred p, g, c;\n"
        "setp.ge.s64 p, %4, %5;\n"
        "setp.ne.u32 g, %6, 0;\n"
        "and.pred c, p, g;\n"
        "selp.f64 %0, %7, %8, c;\n"
        "selp.f64 %1, %9, %10, p;\n"
        "selp.f64 %2, %11, %12, p;\n"
        "selp.f64 %3, %13, %14, p;\n"
        "}\n"
        : "=d"(ts), "=d"(lat), "=d"(na), "=d"(b)
        : "l"(__double_as_longlong(raw)), "l"(t0_bits), "r"((unsigned)gate), "d"(t0), "d"(raw), "d"(clv_), "d"(cls_),
          "d"(naw), "d"(nai), "d"(bw), "d"(bi));
}
// the same with the pair (-a, b) of ei_sat / ew_sat already selected by the caller
DEVI double latent_q_ab(double ts, double A, double Bq, double sp, double na, double b, const double *tab)
{
    const double x = ts - kc[14];
    const double em = fexp_t((na * x) * frcp(b + x), tab);
    return fma(frcp(fma(sp, em, kc[16])), A, -Bq);
}
#endif

// longwave_upward, surface_physics.f90:434-438
DEVI double longwave_upward(double ts, double sigma)
{
#if SEMIC_STRICT
    return (((sigma * ts) * ts) * ts) * ts;
#else
    const double t2 = ts * ts;
    return sigma * (t2 * t2);
#endif
}

// diurnal_cycle, surface_physics.f90:444-472.  Returns the branch id (0 below only, 1 interior, 2 above only).
DEVI int diurnal_cycle(double amp, double tmean, double &above, double &below)
{
    if (tmean + amp < 0.0) {                        // :459
        below = tmean;
        above = 0.0;
        return 0;
    }
    above = tmean;
    below = 0.0;
    if (fabs(tmean) < amp) {                        // :465
        // Reference operation order (unfused) in BOTH arithmetic modes: for tmean -> -amp the numerator of
        // `above` cancels to O(eps^1.5), so only the reference's own rounding sequence comes close to its
        // digits.  Divisions are IEEE in strict mode and ~1 ulp in fast mode.
        const double r = fdiv(tmean, amp);
        double tmp1 = 0.0, tmp2 = 0.0;              // :452-453
        if (fabs(r) < 1.0) {                        // :454 (can differ from :465 only by the rounding of the quotient)
            tmp1 = acos(r);                         // :455
            tmp2 = sqrt(__dsub_rn(1.0, fdiv(__dmul_rn(tmean, tmean), __dmul_rn(amp, amp))));           // :456
        }
        const double at2 = __dmul_rn(amp, tmp2);
        above = fdiv(__dadd_rn(__dadd_rn(__dmul_rn(-tmean, tmp1), at2), __dmul_rn(c_pi, tmean)),
                     __dsub_rn(c_pi, tmp1));                                                           // :467
        below = fdiv(__dsub_rn(__dmul_rn(tmean, tmp1), at2), tmp1);                                    // :469
        return 1;
    }
    return 2;
}

#if !SEMIC_STRICT
// The same for the fast arithmetic mode.  `above` leaves clamped at 0 (the reference clamps it two lines later, :254).
// The interior branch (taken on the days around melt, ~1/4 of all) is ill-conditioned towards tmean -> -amp: with
// x = pi - acos(r) -> 0 the numerator of `above` cancels to amp*x^3/2, and an error e in tmp2 = sqrt(1 - tmean^2/amp^2)
// reaches `above` as amp*e/x.  The reference's tmp2 is a chain of correctly rounded operations, so it is reproduced BIT FOR
// BIT here: the two quotients are the reciprocal products corrected by one Newton step on the exact remainder
// (q = q0 + (n - q0*d)*(1/d): the correctly rounded quotient; 2 FMA each instead of the ~20-instruction IEEE division
// routine), the products and sums are the reference's own unfused sequence, sqrt is IEEE.  What is left against the
// reference is acos (facos vs glibc, <= 1-2 ulp of a number in (0, pi)), the same as in strict mode.
DEVI int diurnal_cycle_fast(const double amp, const double inv_amp, const double inv_amp2, const double tmean,
                            double &above, double &below)
{
    if (tmean + amp < 0.0) {                        // :459
        below = tmean;
        above = 0.0;
        return 0;
    }
    above = tmean;
    below = 0.0;
    if (fabs(tmean) < amp) {                        // :465
        const double r0 = tmean * inv_amp;
        const double r = fma(fma(-r0, amp, tmean), inv_amp, r0);                                        // tmean/amp, :454-455
        double tmp1 = 0.0, tmp2 = 0.0;                                                                  // :452-453
        if (fabs(r) < 1.0) {                                                                            // :454
            tmp1 = facos(r);                                                                            // :455
            const double t2 = __dmul_rn(tmean, tmean), a2 = __dmul_rn(amp, amp);
            const double q0 = t2 * inv_amp2;
            const double q = fma(fma(-q0, a2, t2), inv_amp2, q0);                                       // tmean**2/amp**2
            tmp2 = sqrt(__dsub_rn(1.0, q));                                                             // :456
        }
        const double at2 = __dmul_rn(amp, tmp2);
        above = __dadd_rn(__dadd_rn(__dmul_rn(-tmean, tmp1), at2), __dmul_rn(ka[12], tmean)) * frcp(__dsub_rn(ka[12], tmp1));   // :467
        above = dmax1(above, 0.0);
        below = __dsub_rn(__dmul_rn(tmean, tmp1), at2) * frcp(tmp1);                                   // :469
        return 1;
    }
    return 2;
}
#endif

// albedo_slater, surface_physics.f90:503-530
DEVI double albedo_slater(double tsurf, const DevPar &P)
{
#if SEMIC_STRICT
    double tm = 0.0;
    if (tsurf >= P.tmin && tsurf <= P.tmax) tm = P.slater_f * (tsurf - P.tmin);
    if (tsurf > c_t0) tm = 1.0;
    return P.alb_smax - (P.smax_m_smin * pow(tm, 3.0));
#else
    const bool in = (tsurf >= P.tmin) & (tsurf <= P.tmax);
    double tm = in ? P.slater_f * (tsurf - P.tmin) : 0.0;
    tm = (tsurf > kc[14]) ? 1.0 : tm;
    return fma(-P.smax_m_smin, tm * tm * tm, P.alb_smax);
#endif
}

// albedo_denby, surface_physics.f90:535-542
DEVI double albedo_denby(double melt, const DevPar &P)
{
#if SEMIC_STRICT
    return P.alb_smin + (P.smax_m_smin * fexp(fdiv(-melt, P.mcrit)));
#else
    return fma(P.smax_m_smin, fexp(dmin1(700.0, dmax1(-700.0, -melt * P.inv_mcrit))), P.alb_smin);
#endif
}

// albedo_isba, surface_physics.f90:476-496 (called with tau = tstic, :345)
DEVI double albedo_isba(double alb, double sf, double melt, const DevPar &P)
{
    const double alb_dry = alb - P.isba_dry;
#if SEMIC_STRICT
    const double alb_wet = ((alb - P.alb_smin) * P.isba_wet) + P.alb_smin;
    const double alb_new = fdiv(sf * P.tstic, P.isba_wcr) * P.smax_m_smin;
    double w_alb = 0.0;
    if (melt > 0.0) w_alb = 1.0 - fdiv(melt, P.mcrit);
    w_alb = dmin1(1.0, dmax1(w_alb, 0.0));
    double a = (((1.0 - w_alb) * alb_dry) + (w_alb * alb_wet)) + alb_new;
#else
    const double alb_wet = fma(alb - P.alb_smin, P.isba_wet, P.alb_smin);
    double w_alb = 0.0;
    if (melt > 0.0) w_alb = fma(-melt, P.inv_mcrit, 1.0);
    w_alb = dmin1(1.0, dmax1(w_alb, 0.0));
    double a = fma(sf, P.isba_new, fma(w_alb, alb_wet, (1.0 - w_alb) * alb_dry));
#endif
    return dmin1(P.alb_smax, dmax1(a, P.alb_smin));
}

// ------------------------------------------------------------------------------------------------
// one point, one day
// ------------------------------------------------------------------------------------------------
struct Pt {
    // carried day to day (surface_physics.f90:56-90)
    double tsurf, hsnow, hice, alb, alb_snow, qmr_res;
    // diagnostics of the current day (also carried when a boundary flag freezes them)
    double t2m, melt, melted_snow, melted_ice, refr, smb, acc, lhf, shf, lwu, subl, evap,
           smb_snow, smb_ice, runoff, qmr, amp;
    int mask;
    // plain fast path: lwu, evap and runoff only matter in the final state (they feed nothing and are not daily output), so
    // point_day leaves their ingredients here and finish_diagnostics() forms them on the day the state is stored
    double aux_rr;          // refrozen_rain (runoff = (melt + rf) - refrozen_rain, :303; s.runoff holds rf, s.lwu holds ts^4,
    bool aux_ice;           //  s.evap holds q of the last sub-step, aux_ice its ts < t0)
};

#define BND(x) (GENERIC && (B.x != 0))
// tsurf and alb overrides are also honoured by the EXT2 instantiations of the plain path (k_run_lean with external rows)
#define BNDX(x) ((GENERIC || EXT2) && (B.x != 0))

// n_ksub x energy_balance (surface_physics.f90:158-214) for one point; state in registers
#if SEMIC_STRICT
template <bool GENERIC, bool ALLG, int UNR, bool EXT2 = false>
DEVI void energy_substeps(double &ts, double &t2m, double &shf, double &lhf, double &subl, double &evap, double &lwu,
                          double &q_last, double &over, bool &ice_last, const double shf_c, const double lhf_c,
                          const double rad, const double qq, const double sp, const double qmr_res, const bool gate,
                          const int nsub, const DevPar &P, const semic_bnd &B, const double *tab)
{
    (void)tab;
#pragma unroll UNR
    for (int k = 0; k < nsub; ++k) {
        if (!BND(shf)) shf = shf_c * (ts - t2m);                               // :173
        if (!BND(lhf)) {                                                       // :179-185
            const bool ice = ts < c_t0;                                        // :415
            const double q = latent_q(ts, lhf_c, qq, sp, ice);
            lhf = q * (ice ? c_cls : c_clv);                                   // :421,:428
            q_last = q;
            ice_last = ice;
            if (GENERIC) { subl = ice ? q : 0.0; evap = ice ? 0.0 : q; }
        }
        if (GENERIC) subl = fdiv(subl, c_rhow);                                // :186 (every sub-step, Q3)
        lwu = longwave_upward(ts, c_sigm);                                     // :191
        const double qsb = (((rad - lwu) - shf) - lhf) - qmr_res;              // :194
        over = 0.0;                                                            // :197
        if (!BNDX(tsurf)) {                                                    // :198
            ts = ts + fdiv(qsb * P.tsticsub, P.ceff);                          // :199
            if (gate && ts > c_t0) {                                           // :201-204
                over = ts - c_t0;
                ts = c_t0;
            }
        }
        if (gate) t2m = t2m + fdiv((shf + lhf) * P.tsticsub, P.ceff);          // :209-211 (Q1: forcing t2m is mutated)
    }
}
#else
// Fast arithmetic.  This translation unit is compiled with -fmad=false: every fused operation below is written as
// fma(), so all instantiations (unroll factors, compile-time n_ksub, bitmap or not) produce the same bits.
// One sub-step of the plain path (no boundary flag but tsurf/alb of EXT2) on the values of one point: SubK holds the day's
// constants, SubV what changes from sub-step to sub-step.  shf, lhf, lwu of the LAST sub-step are formed after the loop
// (substep_diag) from d, q, the latent heat and ts^4.
struct SubK { double shf_c, lat_a, lat_b, sp, radq, cg; bool gate; };
struct SubV { double ts, t2m, ts_raw, lat_next, na, b, d, q, lat, t4; };

template <bool ALLG>
DEVI void substep_init(SubV &v, const double ts, const double t2m)
{
    // ts < t0 as a signed 64-bit integer compare of the bit patterns: identical to the floating-point
    // compares for every non-NaN ts (t0 > 0), and they run on the ALU pipe in the shadow of the FP64 pipe, which is the
    // busy one (DSETP costs it two cycles per warp, profiles/r01_summary.md section 4).  One compare per sub-step serves
    // both decisions: the next sub-step's ts < t0 (:415) is !(ts_raw >= t0) whether or not the clamp fires.
    // The selects that depend on it (latent heat, the ei/ew constants) are made right at the compare, for the NEXT sub-step,
    // and carried as values: the compiler otherwise re-issues the compare at every use.
    const bool ice0 = __double_as_longlong(ts) < __double_as_longlong(kc[14]);
    v.ts = ts; v.t2m = t2m; v.ts_raw = ts;
    v.lat_next = ice0 ? kc[22] : kc[23];
    v.na = (ALLG || ice0) ? kc[24] : kc[25];
    v.b = (ALLG || ice0) ? kc[19] : kc[21];
    v.d = 0.0; v.q = 0.0; v.lat = 0.0; v.t4 = 0.0;
}

template <bool ALLG, bool EXT2>
DEVI void substep_plain(SubV &v, const SubK &K, const double c, const semic_bnd &B, const double *tab)
{
    const double t0 = kc[14], sigm = kc[17];
    const long long t0_bits = __double_as_longlong(t0);
    v.lat = v.lat_next;                                                        // :415 decided on the tsurf this sub-step starts from
    v.q = latent_q_ab(v.ts, K.lat_a, K.lat_b, K.sp, v.na, v.b, tab);           // :179-185
    v.d = v.ts - v.t2m;                                                        // :173
    const double t2 = v.ts * v.ts;
    v.t4 = t2 * t2;                                                            // :191
    const double sl = fma(K.shf_c, v.d, v.lat * v.q);                          // shf + lhf, shared by :194 and :210
    const double qsb = fma(-sigm, v.t4, K.radq) - sl;                          // :194
    v.t2m = fma(sl, K.cg, v.t2m);                                              // :209-211 (Q1: forcing t2m is mutated)
    if (!(EXT2 && B.tsurf != 0)) {                                             // :198 (external tsurf stays as given)
        v.ts_raw = fma(qsb, c, v.ts);                                          // :199
        // warm = !(tsurf < t0) of the next sub-step; tsurf = t0 if gated and warm (:201-204; ts_raw == t0: the same value)
        if (ALLG) clamp_and_latent(v.ts_raw, t0_bits, t0, K.gate, kc[22], kc[23], v.ts, v.lat_next);
        else      clamp_and_latent_ab(v.ts_raw, t0_bits, t0, K.gate, kc[22], kc[23], kc[24], kc[25], kc[19], kc[21], v.ts,
                                      v.lat_next, v.na, v.b);
    }
}

DEVI void substep_diag(const SubV &v, const SubK &K, double &shf, double &lhf, double &lwu, double &q_last, bool &ice_last)
{
    shf = K.shf_c * v.d;
    lhf = v.lat * v.q;
    lwu = v.t4;                                                                // ts^4: finish_diagnostics() scales it (:191)
    q_last = v.q;
    ice_last = __double2hiint(v.lat) == __double2hiint(kc[22]);                // the latent heat in use was cls (:421)
}

// ALLG: every lane of the warp is gated (mask == 2 or snow covered) with ts <= t0; a compile-time true removes selects.
template <bool GENERIC, bool ALLG, int UNR, bool EXT2 = false>
DEVI void energy_substeps(double &ts, double &t2m, double &shf, double &lhf, double &subl, double &evap, double &lwu,
                          double &q_last, double &over, bool &ice_last, const double shf_c, const double lhf_c,
                          const double rad, const double qq, const double sp, const double qmr_res, const bool gate_in,
                          const int nsub, const DevPar &P, const semic_bnd &B, const double *tab)
{
    const bool gate = ALLG ? true : gate_in;
    const double t0 = kc[14], sigm = kc[17], c = P.sub_over_ceff;
    double ts_raw = ts;                                                        // tsurf of the last sub-step before the clamp
    const double lat_a = lhf_c * kc[15], lat_b = lhf_c * qq;                   // day constants of latent_q
    const double radq = rad - qmr_res;                                         // :194 with the day-constant terms first
    // per-lane constant that replaces a select inside the loop: an ungated lane adds (shf+lhf)*0 to t2m
    const double cg = gate ? c : 0.0;                                          // :209-211 as t2m += (shf+lhf)*cg
    const long long t0_bits = __double_as_longlong(t0);
    if (!GENERIC) {
        // the diagnostics shf, lhf, lwu of the last sub-step are formed after the loop from d, q, latent heat and ts^4
        const SubK K = {shf_c, lat_a, lat_b, sp, radq, cg, gate};
        SubV v;
        substep_init<ALLG>(v, ts, t2m);
#pragma unroll UNR
        for (int k = 0; k < nsub; ++k) substep_plain<ALLG, EXT2>(v, K, c, B, tab);
        ts = v.ts; t2m = v.t2m; ts_raw = v.ts_raw;
        if (nsub > 0) substep_diag(v, K, shf, lhf, lwu, q_last, ice_last);
    } else {
        // all boundary flags; the same fused forms as above wherever the flag set allows, so that a run whose flags are
        // a subset of {tsurf, alb} gives the same bits here and in the EXT2 instantiation of the plain path
#pragma unroll UNR
        for (int k = 0; k < nsub; ++k) {
            double lh = lhf;
            if (!BND(lhf)) {                                                   // :179-185
                const bool ice = ts < t0;
                const double q = latent_q<false>(ts, lat_a, lat_b, sp, ice, tab);
                lh = (ice ? kc[22] : kc[23]) * q;
                lhf = lh;
                q_last = q;
                ice_last = ice;
                subl = ice ? q : 0.0;
                evap = ice ? 0.0 : q;
            }
            subl = fdiv(subl, c_rhow);                                         // :186 (every sub-step, Q3)
            const double t2 = ts * ts;
            const double t4 = t2 * t2;
            lwu = sigm * t4;                                                   // :191
            double sl;
            if (!BND(shf)) {
                const double d = ts - t2m;                                     // :173
                shf = shf_c * d;
                sl = fma(shf_c, d, lh);
            } else {
                sl = shf + lh;
            }
            const double qsb = fma(-sigm, t4, radq) - sl;                      // :194
            t2m = fma(sl, cg, t2m);                                            // :209-211
            if (!BNDX(tsurf)) {                                                // :198-204
                ts_raw = fma(qsb, c, ts);
                ts = (gate && ts_raw > t0) ? t0 : ts_raw;
            }
        }
    }
    // only the last sub-step's excess survives (:197, Q2)
    if (nsub > 0) over = (!BNDX(tsurf) && gate && ts_raw > t0) ? ts_raw - t0 : 0.0;
}
#endif


// A day is three pieces: day_head reads the forcing and forms the day's constants, day_substeps runs n_ksub x
// energy_balance, day_tail takes the last sub-step's excess into mass_balance.  DayCtx carries the values from piece to
// piece; point_day is the three in a row.
struct DayCtx {
    double ts, t2m, shf, lhf, subl, evap, lwu;          // running values of energy_balance
    double shf_c, lhf_c, rad, qq, sp;                   // day constants
    double q_last, over;                                // q and clamp excess of the last sub-step
    bool gate, ice_last;
    int nsub;
};

template <bool GENERIC, int NK>
DEVI void day_head(const Pt &s, const double *row, const long long T, const DevPar &P, const int eb_steps, double &swd_out,
                   DayCtx &c)
{
    c.gate = (s.mask == 2) || (s.hsnow > 0.0);                        // :201,:209 (hsnow is constant over the sub-steps)
    c.ts = s.tsurf;
    c.t2m = row[SEMIC_V_T2M * T];
    c.shf = s.shf; c.lhf = s.lhf; c.subl = s.subl; c.evap = s.evap; c.lwu = s.lwu;
    const double wind = row[SEMIC_V_WIND * T], rhoa = row[SEMIC_V_RHOA * T];
    const double swd = row[SEMIC_V_SWD * T];
    swd_out = swd;
    // exact hoists: literal left-to-right prefixes of :387, :420/:427 and :194
    c.shf_c  = (P.csh_cap * rhoa) * wind;
    c.lhf_c  = (P.clh * wind) * rhoa;
#if SEMIC_STRICT
    c.rad    = ((1.0 - s.alb) * swd) + row[SEMIC_V_LWD * T];
#else
    c.rad    = fma(1.0 - s.alb, swd, row[SEMIC_V_LWD * T]);
#endif
    c.qq = row[SEMIC_V_QQ * T];
    c.sp = row[SEMIC_V_SP * T];
    c.q_last = 0.0; c.over = 0.0;
    c.ice_last = false;
    c.nsub = GENERIC ? eb_steps : (NK > 0 ? NK : P.n_ksub);           // NK: n_ksub known at compile time
}

// ---------------- energy_balance x nsub (:149-151, :158-214) ----------------
template <bool GENERIC, int UNR, bool EXT2>
DEVI void day_substeps(DayCtx &c, const double qmr_res, const DevPar &P, const semic_bnd &B, const double *tab)
{
    // gated points never exceed t0 (:201-204), which removes the ei/ew constant selects (see latent_q); taken when
    // every active lane of the warp is gated, e.g. the interior of an ice sheet
    const bool allg = !GENERIC && !SEMIC_STRICT && __all_sync(__activemask(), c.gate && c.ts <= c_t0);
    if (allg)
        energy_substeps<GENERIC, true, UNR, EXT2>(c.ts, c.t2m, c.shf, c.lhf, c.subl, c.evap, c.lwu, c.q_last, c.over, c.ice_last,
                                                  c.shf_c, c.lhf_c, c.rad, c.qq, c.sp, qmr_res, c.gate, c.nsub, P, B, tab);
    else
        energy_substeps<GENERIC, false, UNR, EXT2>(c.ts, c.t2m, c.shf, c.lhf, c.subl, c.evap, c.lwu, c.q_last, c.over, c.ice_last,
                                                   c.shf_c, c.lhf_c, c.rad, c.qq, c.sp, qmr_res, c.gate, c.nsub, P, B, tab);
}

// `row` points at this point's column of the day's forcing (a shared-memory stage in k_run, global
// memory in k_ensemble): variable v is row[v*T].  Values are read where they are first needed (LDS
// is cheap, registers across the sub-step loop are not).
template <int SCHEME, bool GENERIC, bool WANT_BRANCH, bool EXT2>
DEVI unsigned day_tail(Pt &s, const double *row, const long long T, const DevPar &P, const semic_bnd &B, const int do_mb,
                       DayCtx &c, const double *tab)
{
    unsigned branch = 0;
    bool any_over = false;              // fast plain path: a lane of the warp was clamped in the last sub-step (qmr != 0)
    (void)any_over;
    const int mask = s.mask;
    if (WANT_BRANCH) { if (c.gate) branch |= SEMIC_B_GATE; }
    double ts = c.ts, t2m = c.t2m, shf = c.shf, lhf = c.lhf, subl = c.subl, evap = c.evap, lwu = c.lwu, qmr = s.qmr;
    {
        const double q_last = c.q_last, over = c.over;
        const bool ice_last = c.ice_last;
        const bool any = c.nsub > 0;
        if (any) {
            // only the last sub-step's qmr survives (:197, Q2)
#if SEMIC_STRICT
            qmr = fdiv(over * P.ceff, P.tsticsub);                                 // :202 (0 when the clamp did not fire)
#else
            // :202 in the reference's own rounding sequence ((ts-t0)*ceff, then the IEEE quotient): the diurnal cycle below
            // amplifies the last bit of tmean by up to 1e5 (see diurnal_cycle_fast), so qmr and tmean are reproduced exactly
            // whenever the clamp excess itself is; days without a clamped lane (3 of 4) skip it
            qmr = 0.0;
            any_over = __any_sync(__activemask(), over > 0.0);
            if (any_over) qmr = fdiv_rn(__dmul_rn(over, P.ceff), P.tsticsub, P.inv_tsticsub);
#endif
            if (WANT_BRANCH) { if (over > 0.0) branch |= SEMIC_B_CLAMP; }
            if (!BND(lhf)) {
                if (WANT_BRANCH) { if (ice_last) branch |= SEMIC_B_TS_LT_T0; }
                if (!GENERIC) {                                                    // :410-411,:420,:427 then :186
#if SEMIC_STRICT
                    subl = ice_last ? fdiv(q_last, c_rhow) : 0.0;
#else
                    subl = ice_last ? q_last * 1.0e-3 : 0.0;
#endif
#if SEMIC_STRICT
                    evap = ice_last ? 0.0 : q_last;
#else
                    evap = q_last;                                                 // finish_diagnostics(): ice_last ? 0 : q_last
                    s.aux_ice = ice_last;
#endif
                }
            }
        }
    }

    double hsnow = s.hsnow, hice = s.hice, alb = s.alb, alb_snow = s.alb_snow, qmr_res = s.qmr_res;
    double melt = s.melt, melted_snow = s.melted_snow, melted_ice = s.melted_ice, refr = s.refr;
    double smb = s.smb, acc = s.acc, smb_snow = s.smb_snow, smb_ice = s.smb_ice, runoff = s.runoff, amp = s.amp;

    // ---------------- mass_balance (:232-374) ----------------
    if (!GENERIC || do_mb) {
        const double sf = row[SEMIC_V_SF * T], rf = row[SEMIC_V_RF * T];
        if (!BND(amp)) amp = P.amp;                                                // :246-248
#if SEMIC_STRICT
        const double tmean = (ts - c_t0) + (fdiv(qmr, P.ceff) * P.tstic);          // :249
#else
        double tmean = ts - kc[14];                                                // :249, unfused like the reference
        if (GENERIC ? __any_sync(__activemask(), qmr != 0.0) : any_over)           // plain path: qmr != 0 only where the clamp fired
            tmean = __dadd_rn(tmean, __dmul_rn(fdiv_rn(qmr, P.ceff, P.inv_ceff), P.tstic));
#endif
        double above, below;
#if SEMIC_STRICT
        const int dbr = diurnal_cycle(amp, tmean, above, below);
#else
        const double inv_amp = BND(amp) ? frcp(amp) : P.inv_amp;
        const double inv_amp2 = BND(amp) ? frcp(amp * amp) : P.inv_amp2;
        const int dbr = diurnal_cycle_fast(amp, inv_amp, inv_amp2, tmean, above, below);
#endif
        if (WANT_BRANCH) { branch |= (dbr == 1 ? SEMIC_B_DIURNAL_LO : (dbr == 2 ? SEMIC_B_DIURNAL_HI : 0)); }
        qmr = 0.0;                                                                 // :250

        double qmelt = 0.0, qcold = 0.0;
        if (SEMIC_STRICT && mask >= 1) {                                           // :252-261
#if SEMIC_STRICT
            qmelt = dmax1(0.0, fdiv(above * P.ceff, P.tstic));
            qcold = dmax1(0.0, fdiv(fabs(below) * P.ceff, P.tstic));
#endif
        }
#if !SEMIC_STRICT
        // above >= 0 on return from diurnal_cycle_fast and dmax1(0,|x|) == |x|; finite values times 0 are the :252-261 zeros.
        // qmelt and qcold only feed melt (:266) and refr (:286): one multiplication by ceff/tstic/(rhow*clm) each
        const double c_mask = (mask >= 1) ? P.melt_per_kelvin : 0.0;
        (void)qmelt; (void)qcold;
#endif
#if SEMIC_STRICT
        if (!BND(melt)) melt = fdiv(qmelt, P.rhow_clm);                            // :266
        const double hs_rate = fdiv(hsnow, P.tstic);
#else
        // IEEE quotient: when all snow melts, hsnow - (hsnow/tstic)*tstic decides between "bare" and a
        // one-ulp ghost snow layer (which keeps the :201 gate on over land); keep the reference's rounding
        // (one Newton refinement of hsnow*(1/tstic) is the correctly rounded quotient up to ~2^-50 odds: 3 instructions
        //  instead of the ~15 of the general IEEE division routine)
        const double hs_q = hsnow * P.inv_tstic;
        const double hs_rate = fma(fma(-hs_q, P.tstic, hsnow), P.inv_tstic, hs_q);
        // A whole warp below the melting range (tmean + amp < 0 on every lane: 3 days of 4): above = +0, hence melt,
        // melted_snow, melted_ice = +0 and the snow-refreezing term is rcrit*(+0).  The short form below produces the bits of
        // the general statements (x - (+0) = x, min(+0, hs_rate) = +0 for hs_rate >= +0), it only skips them.
        const bool cold_warp = !GENERIC && __all_sync(__activemask(), dbr == 0 && __double2hiint(hsnow) >= 0);
#endif
        double refrozen_rain = 0.0, refrozen_snow = 0.0;                           // Q5: defined as 0 under bnd%refr
#if !SEMIC_STRICT
        if (cold_warp) {
            melt = 0.0; melted_snow = 0.0; melted_ice = 0.0;                       // :266-280
            refr = fabs(below) * c_mask;                                           // :286
            refrozen_rain = P.rcrit * dmin1(hs_rate, dmin1(refr, rf));             // :287,:293
            refrozen_snow = P.rcrit_zero;                                          // :289-294: rcrit*min(hs_rate, min(., +0))
            refr = refrozen_rain + refrozen_snow;                                  // :295
            qmr = -(refr * P.frz_rhow_clm);                                        // :298
            if (GENERIC) runoff = (melt + rf) - refrozen_rain;
            else { runoff = rf; s.aux_rr = refrozen_rain; }
            acc = (sf - subl) + refr;                                              // :307
            smb_snow = (sf - subl) + refrozen_snow;                                // :312 with melted_snow = +0
        } else {
        if (!BND(melt)) melt = above * c_mask;
#endif
        melted_snow = dmin1(melt, hs_rate);                                         // :269
        melted_ice  = melt - melted_snow;                                          // :270
        if (!BND(melt)) {                                                          // :272-280
            if (mask != 2) melted_ice = 0.0;
            melt = melted_snow + melted_ice;                                       // x + 0 == x: the land/ocean branch of :277
        }

        if (!BND(refr)) {                                                          // :283-299
#if SEMIC_STRICT
            refr = fdiv(qcold, P.rhow_clm);                                        // :286
#else
            refr = fabs(below) * c_mask;
#endif
            refrozen_rain = dmin1(refr, rf);                                        // :287
#if SEMIC_STRICT
            refrozen_snow = dmax1(refr - refrozen_rain, 0.0);                       // :289
#else
            refrozen_snow = refr - refrozen_rain;           // refrozen_rain = min(refr, rf) <= refr: the max(.,0) of :289 is the identity
#endif
            refrozen_snow = dmin1(refrozen_snow, melted_snow);                      // :291
            refrozen_rain = P.rcrit * dmin1(hs_rate, refrozen_rain);                // :293
            refrozen_snow = P.rcrit * dmin1(hs_rate, refrozen_snow);                // :294
            refr = refrozen_rain + refrozen_snow;                                  // :295
#if SEMIC_STRICT
            qmr = qmr - (((P.one_m_frz * refr) * c_rhow) * 3.30e5);                // :298
#else
            qmr = -(refr * P.frz_rhow_clm);
#endif
        }

#if SEMIC_STRICT
        runoff = (melt + rf) - refrozen_rain;                                      // :302-303
#else
        if (GENERIC) runoff = (melt + rf) - refrozen_rain;
        else { runoff = rf; s.aux_rr = refrozen_rain; }                            // finish_diagnostics()
#endif
        if (!BND(acc)) acc = (sf - subl) + refr;                                   // :307
        if (!BND(smb)) smb_snow = ((sf - subl) - melted_snow) + refrozen_snow;     // :312
#if !SEMIC_STRICT
        }   // !cold_warp
#endif

#if SEMIC_STRICT
        if (mask == 0) hsnow = 0.0;                                                // :315-320
        else hsnow = dmax1(0.0, hsnow + (smb_snow * P.tstic));
#else
        {
            const double hs_new = __dadd_rn(hsnow, __dmul_rn(smb_snow, P.tstic));  // unfused, see hs_rate
            hsnow = ((hs_new > 0.0) & (mask != 0)) ? hs_new : 0.0;                 // :315-320 with one select
        }
#endif
#if SEMIC_STRICT
        const double snow_to_ice = dmax1(0.0, hsnow - c_hsmax);                     // :323
        hsnow = hsnow - snow_to_ice;                                               // :324
        const double s2i_rate = fdiv(snow_to_ice, P.tstic);
        smb_ice = (s2i_rate - melted_ice) + refrozen_rain;                         // :325
        hice = hice + (smb_ice * P.tstic);                                         // :326
        if (!BND(smb)) {                                                           // :329-335
            if (mask == 2) smb = (smb_snow + smb_ice) - s2i_rate;
            else           smb = smb_snow + dmax1(0.0, smb_ice - s2i_rate);
        }
#else
        // no lane of the warp above hsmax (the usual case): snow_to_ice = +0 and :324-:334 lose their s2i terms exactly
        double snow_to_ice = 0.0;
        const double over5 = hsnow - c_hsmax;                                      // :323
        if (__any_sync(__activemask(), over5 > 0.0)) {
            snow_to_ice = dmax1(0.0, over5);
            hsnow = hsnow - snow_to_ice;                                           // :324
            const double s2i_rate = snow_to_ice * P.inv_tstic;
            smb_ice = (s2i_rate - melted_ice) + refrozen_rain;                     // :325
            if (!BND(smb)) {                                                       // :329-335
                if (mask == 2) smb = (smb_snow + smb_ice) - s2i_rate;
                else           smb = smb_snow + dmax1(0.0, smb_ice - s2i_rate);
            }
        } else {
            smb_ice = (0.0 - melted_ice) + refrozen_rain;
            if (!BND(smb)) {
                if (mask == 2) smb = smb_snow + smb_ice;
                else           smb = smb_snow + dmax1(0.0, smb_ice);
            }
        }
        hice = fma(smb_ice, P.tstic, hice);                                        // :326
#endif
        if (WANT_BRANCH) { if (melt > 0.0) branch |= SEMIC_B_MELT_POS; }
        if (WANT_BRANCH) { if (hsnow == 0.0) branch |= SEMIC_B_HSNOW_ZERO; }
        if (WANT_BRANCH) { if (snow_to_ice > 0.0) branch |= SEMIC_B_SNOW2ICE; }

        if (!BNDX(alb)) {                                                          // :339 (Q6: the whole block is skipped)
#if SEMIC_STRICT
            const double f_alb = 1.0 - fexp(fdiv(-hsnow, P.hcrit_eps));            // :338
#else
            // exact shortcut: exp(-x) < 2^-54 for x > 37.43, so 1 - exp(-x) rounds to 1 (taken when the whole warp is
            // under deep snow).  A tiny hcrit can push the argument out of fexp_t's range; exp(-700) is already 0 next to 1.
            double f_alb = 1.0;
            double xs = hsnow * P.inv_hcrit_eps;
            if (!__all_sync(__activemask(), xs > 38.0)) {
                xs = dmin1(xs, 700.0);
                f_alb = 1.0 - fexp_t(-xs, tab);
            }
#endif
            const int scheme = (SCHEME == SCHEME_RUNTIME) ? P.scheme : SCHEME;
            if (scheme == SCHEME_SLATER)      alb_snow = albedo_slater(ts, P);     // :340-341
            else if (scheme == SCHEME_DENBY)  alb_snow = albedo_denby(melt, P);    // :342-343
            else if (scheme == SCHEME_ISBA)   alb_snow = albedo_isba(alb_snow, sf, melt, P);   // :344-346
            else if (scheme == SCHEME_NONE)   alb_snow = P.alb_smax;               // :347-348
#if SEMIC_STRICT
            if (mask == 2)      alb = P.albi + (f_alb * (alb_snow - P.albi));      // :350-352
            else if (mask == 1) alb = P.albl + (f_alb * (alb_snow - P.albl));      // :353-355
            else if (mask == 0) alb = 0.06;                                        // :356-358
            if (scheme == SCHEME_ALEX) {                                           // :360-363 (uses the mutated t2m)
                alb_snow = P.alb_smin + (P.smax_m_smin * ((0.5 * tanh(P.afac * (t2m - P.tmid))) + 0.5));
                alb = alb_snow;
            }
#else
            {   // :350-358 with the background albedo selected first (a per-lane constant): one blend instead of two
                const double bg = (mask == 2) ? P.albi : P.albl;
                const double blend = fma(f_alb, alb_snow - bg, bg);
                alb = (mask == 1 || mask == 2) ? blend : ((mask == 0) ? 0.06 : alb);     // Q7: other masks leave alb unchanged
            }
            if (scheme == SCHEME_ALEX) {
                alb_snow = fma(P.smax_m_smin, fma(0.5, tanh(P.afac * (t2m - P.tmid)), 0.5), P.alb_smin);
                alb = alb_snow;
            }
#endif
        }
        qmr_res = (mask == 0) ? 0.0 : qmr;                                         // :368-372
    }

    s.tsurf = ts; s.hsnow = hsnow; s.hice = hice; s.alb = alb; s.alb_snow = alb_snow; s.qmr_res = qmr_res;
    s.t2m = t2m; s.melt = melt; s.melted_snow = melted_snow; s.melted_ice = melted_ice; s.refr = refr;
    s.smb = smb; s.acc = acc; s.lhf = lhf; s.shf = shf; s.lwu = lwu; s.subl = subl; s.evap = evap;
    s.smb_snow = smb_snow; s.smb_ice = smb_ice; s.runoff = runoff; s.qmr = qmr; s.amp = amp;
    return branch;
}

// one point, one day; `swd_out` returns the day's shortwave for the swnet diagnostic
template <int SCHEME, bool GENERIC, bool WANT_BRANCH, int UNR = 2, int NK = 0, bool EXT2 = false>
DEVI unsigned point_day(Pt &s, const double *row, const long long T, const DevPar &P, const semic_bnd &B,
                        const int eb_steps, const int do_mb, double &swd_out, const double *tab)
{
    DayCtx c;
    day_head<GENERIC, NK>(s, row, T, P, eb_steps, swd_out, c);
    day_substeps<GENERIC, UNR, EXT2>(c, s.qmr_res, P, B, tab);
    return day_tail<SCHEME, GENERIC, WANT_BRANCH, EXT2>(s, row, T, P, B, do_mb, c, tab);
}

// lwu, evap, runoff of the plain fast path from the ingredients point_day left (see struct Pt); a no-op otherwise
template <bool GENERIC>
DEVI void finish_diagnostics(Pt &s, const int nsub)
{
#if !SEMIC_STRICT
    if (!GENERIC) {
        if (nsub > 0) {
            s.lwu = kc[17] * s.lwu;                                                // :191
            s.evap = s.aux_ice ? 0.0 : s.evap;                                     // :410-411,:427
        }
        s.runoff = (s.melt + s.runoff) - s.aux_rr;                                 // :302-303
    }
#else
    (void)s; (void)nsub;
#endif
}

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA (1-D bulk copy) primitives, inline PTX
// ------------------------------------------------------------------------------------------------
DEVI uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

DEVI void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
DEVI void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DEVI void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
DEVI void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
DEVI void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u)   // suspend-time hint: sleep in hardware, do not spin
        : "memory");
}
// global -> shared bulk copy completing on an mbarrier (TMA unit, SASS UBLKCP); 16-byte aligned, size % 16 == 0
DEVI void tma_load_1d(uint32_t smem_dst, const void *gmem_src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gmem_src), "r"(bytes), "r"(bar)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// the persistent multi-day kernel
// ------------------------------------------------------------------------------------------------
DEVI void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <bool GENERIC, int WPC>
struct RunSmem {
    static constexpr int ROWS = kForcingRows + (GENERIC ? kExtRows : 0);
    static constexpr int STAGES = kStages;
    static constexpr size_t stage_bytes = (size_t)ROWS * kTile * sizeof(double);          // 2304 B (3584 B generic)
    static constexpr size_t ring_bytes = stage_bytes * STAGES * WPC;
    static constexpr size_t bar_bytes = (size_t)WPC * STAGES * sizeof(uint64_t);
    static constexpr size_t bytes = ring_bytes + bar_bytes + (SEMIC_STRICT ? 0 : kExpTab) * sizeof(double);   // rings, barriers, exp table
};

// validation-layout rows that feed the overrides tsurf, alb, smb, melt, acc (example.f90:103-107)
__device__ __constant__ int c_ext_var[kExtRows] = {SEMIC_O_STT, SEMIC_O_ALB, SEMIC_O_SMB, SEMIC_O_MELT, SEMIC_O_ACC};

// lane 0 of a warp: one day's forcing (and override rows) of this warp's tile -> one stage of the warp's ring
template <bool GENERIC>
DEVI void issue_day(const RunArgs &a, const long long tile, const int fd, const long long ntl, const uint32_t dst,
                    const uint32_t bar, const bool (&ext_on)[kExtRows], const int n_ext)
{
    constexpr uint32_t rowb = kTile * (uint32_t)sizeof(double);
    mbar_expect_tx(bar, rowb * (uint32_t)(kForcingRows + (GENERIC ? n_ext : 0)));
    if (a.f_tiled) {
        tma_load_1d(dst, a.forcing + ((long long)fd * ntl + tile) * (kForcingRows * kTile), kForcingRows * rowb, bar);
    } else {
        const double *src = a.forcing + tile * kTile;
#pragma unroll
        for (int v = 0; v < kForcingRows; ++v) tma_load_1d(dst + v * rowb, src + a.f_var_off[v], rowb, bar);
    }
    if (GENERIC) {
        const double *vs = a.vali + ((long long)fd * ntl + tile) * (SEMIC_NOUT * kTile);
#pragma unroll
        for (int e = 0; e < kExtRows; ++e)
            if (ext_on[e]) tma_load_1d(dst + (kForcingRows + e) * rowb, vs + c_ext_var[e] * kTile, rowb, bar);
    }
}

template <int SCHEME, bool GENERIC, int MINB, bool BR, int WPC, int UNR = 2>
__global__ void __launch_bounds__(WPC * 32, MINB) k_run(const __grid_constant__ RunArgs a)
{
    using SM = RunSmem<GENERIC, WPC>;
    constexpr int NST = SM::STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler: ring/barrier/tile arithmetic moves to the uniform datapath
    const int lane = threadIdx.x & 31;
    const uint32_t smem_u = smem_u32(smem_raw);
    const uint32_t ring_u = smem_u + (uint32_t)(warp * NST) * (uint32_t)SM::stage_bytes;     // this warp's ring
    const uint32_t bar_u = smem_u + (uint32_t)SM::ring_bytes + (uint32_t)(warp * NST) * 8u;  // this warp's barriers
    const double *ring = reinterpret_cast<const double *>(smem_raw) + (size_t)(warp * NST) * SM::ROWS * kTile;
    double *tab = reinterpret_cast<double *>(smem_raw + SM::ring_bytes + SM::bar_bytes);

    if (lane == 0) {
        for (int s = 0; s < NST; ++s) mbar_init(bar_u + 8u * s, 1);
        mbar_fence_init();
    }
#if !SEMIC_STRICT
    for (int j = threadIdx.x; j < kExpTab; j += WPC * 32) tab[j] = g_exp2_tab[j];
#endif
    __syncthreads();

    bool ext_on[kExtRows] = {false, false, false, false, false};
    int n_ext = 0;
    if (GENERIC && a.vali != nullptr) {
        ext_on[0] = a.B.tsurf != 0; ext_on[1] = a.B.alb != 0; ext_on[2] = a.B.smb != 0;
        ext_on[3] = a.B.melt != 0;  ext_on[4] = a.B.acc != 0;
        for (int e = 0; e < kExtRows; ++e) n_ext += ext_on[e] ? 1 : 0;
    }

    const long long ld = a.ld;
    const long long ntl = a.sld / kTile;                         // tiles per series row
    const long long ntiles = ((long long)a.nx + kTile - 1) / kTile;
    const long long nwarps = (long long)gridDim.x * WPC;
    int stage = 0;
    uint32_t phase = 0;

    // tile -> warp map interleaved over the CTAs: the ntiles % nwarps tiles of the last round land on ALL SMs (a few warps each)
    // instead of filling the first SMs completely -- 3 % on a 2 Mi-point shard (18.45 tiles per warp)
    for (long long tile = (long long)warp * gridDim.x + blockIdx.x; tile < ntiles; tile += nwarps) {
        const long long i = tile * kTile + lane;
        const bool valid = i < a.nx;
        const long long ii = valid ? i : 0;

        // prefetch the first days of this tile into the (empty) ring, in stage order
        int fd_issue = a.f_day0 % a.f_nwin;
        int d_issue = 0;
        if (lane == 0) {
            int s = stage;
            for (; d_issue < NST && d_issue < a.ndays; ++d_issue) {
                issue_day<GENERIC>(a, tile, fd_issue, ntl, ring_u + (uint32_t)s * (uint32_t)SM::stage_bytes, bar_u + 8u * s,
                                   ext_on, n_ext);
                if (++fd_issue == a.f_nwin) fd_issue = 0;
                if (++s == NST) s = 0;
            }
        }

        Pt s;
        {
            const double *S = a.slab + ii;
            s.mask = valid ? a.mask[ii] : 0;
            s.tsurf    = S[SEMIC_F_TSURF * ld];
            s.hsnow    = S[SEMIC_F_HSNOW * ld];
            s.hice     = S[SEMIC_F_HICE * ld];
            s.alb      = S[SEMIC_F_ALB * ld];
            s.alb_snow = S[SEMIC_F_ALB_SNOW * ld];
            s.qmr_res  = S[SEMIC_F_QMR_RES * ld];
            if (GENERIC) {
                s.melt = S[SEMIC_F_MELT * ld];
                s.melted_snow = S[SEMIC_F_MELTED_SNOW * ld];
                s.melted_ice = S[SEMIC_F_MELTED_ICE * ld];
                s.refr = S[SEMIC_F_REFR * ld];
                s.smb = S[SEMIC_F_SMB * ld];
                s.acc = S[SEMIC_F_ACC * ld];
                s.lhf = S[SEMIC_F_LHF * ld];
                s.shf = S[SEMIC_F_SHF * ld];
                s.lwu = S[SEMIC_F_LWU * ld];
                s.subl = S[SEMIC_F_SUBL * ld];
                s.evap = S[SEMIC_F_EVAP * ld];
                s.smb_snow = S[SEMIC_F_SMB_SNOW * ld];
                s.smb_ice = S[SEMIC_F_SMB_ICE * ld];
                s.runoff = S[SEMIC_F_RUNOFF * ld];
                s.qmr = S[SEMIC_F_QMR * ld];
                s.amp = S[SEMIC_F_AMP * ld];
            } else {
                s.melt = s.melted_snow = s.melted_ice = s.refr = s.smb = s.acc = 0.0;
                s.lhf = s.shf = s.lwu = s.subl = s.evap = s.smb_snow = s.smb_ice = 0.0;
                s.runoff = s.qmr = s.amp = 0.0;
            }
            s.t2m = 0.0;
        }

        // daily output: tiled ring, this lane's column of the warp's block; the slot advances by one per written day
        const long long o_block = (long long)a.o_nvar * kTile;                    // doubles per tile-day
        double *oday = (a.out != nullptr) ? a.out + tile * o_block + lane : nullptr;
        const long long o_day = ntl * o_block;
        int oslot = 0;

        for (int d = 0; d < a.ndays; ++d) {
            mbar_wait(bar_u + 8u * stage, phase);
            const double *row = ring + (size_t)stage * SM::ROWS * kTile + lane;
            const bool write_out = (oday != nullptr) && (d >= a.out_start);
            if (valid) {
                if (GENERIC) {                                         // example.f90:103-107
                    if (ext_on[0]) s.tsurf = row[(kForcingRows + 0) * kTile];
                    if (ext_on[1]) s.alb   = row[(kForcingRows + 1) * kTile];
                    if (ext_on[2]) s.smb   = row[(kForcingRows + 2) * kTile];
                    if (ext_on[3]) s.melt  = row[(kForcingRows + 3) * kTile];
                    if (ext_on[4]) s.acc   = row[(kForcingRows + 4) * kTile];
                }
                double swd;
                const unsigned br = point_day<SCHEME, GENERIC, BR, UNR>(s, row, kTile, a.P, a.B, a.eb_steps, a.do_mb, swd,
                                                                        SEMIC_STRICT ? nullptr : tab);
                if (write_out) {                                       // example.f90:117-122
                    oday[SEMIC_O_STT * kTile]   = s.tsurf;
                    oday[SEMIC_O_ALB * kTile]   = s.alb;
                    oday[SEMIC_O_SWNET * kTile] = swd * (1.0 - s.alb);
                    oday[SEMIC_O_SMB * kTile]   = s.smb;
                    oday[SEMIC_O_MELT * kTile]  = s.melt;
                    oday[SEMIC_O_ACC * kTile]   = s.acc;
                    oday[SEMIC_O_SHF * kTile]   = s.shf;
                    oday[SEMIC_O_LHF * kTile]   = s.lhf;
                    if (BR) oday[SEMIC_O_BRANCH * kTile] = (double)br;  // series with 9 rows
                }
            }
            // all lanes are done reading this stage: refill it with day d + NST (generic-proxy reads are ordered
            // before the async-proxy write by the warp barrier + proxy fence)
            __syncwarp();
            if (lane == 0 && d_issue < a.ndays) {
                fence_proxy_async();
                issue_day<GENERIC>(a, tile, fd_issue, ntl, ring_u + (uint32_t)stage * (uint32_t)SM::stage_bytes,
                                   bar_u + 8u * stage, ext_on, n_ext);
                ++d_issue;
                if (++fd_issue == a.f_nwin) fd_issue = 0;
            }
            if (++stage == NST) { stage = 0; phase ^= 1u; }
            if (write_out) {
                oday += o_day;
                if (++oslot == a.o_nring) { oslot = 0; oday -= (long long)a.o_nring * o_day; }
            }
        }

        // final state of the tile: all 23 non-forcing fields, as after the reference's last call
        if (a.ndays > 0 && valid) {
            if (GENERIC ? a.do_mb : 1) finish_diagnostics<GENERIC>(s, GENERIC ? a.eb_steps : a.P.n_ksub);
            double *S = a.slab + i;
            S[SEMIC_F_T2M * ld] = s.t2m;
            S[SEMIC_F_TSURF * ld] = s.tsurf;
            S[SEMIC_F_HSNOW * ld] = s.hsnow;
            S[SEMIC_F_HICE * ld] = s.hice;
            S[SEMIC_F_ALB * ld] = s.alb;
            S[SEMIC_F_ALB_SNOW * ld] = s.alb_snow;
            S[SEMIC_F_MELT * ld] = s.melt;
            S[SEMIC_F_MELTED_SNOW * ld] = s.melted_snow;
            S[SEMIC_F_MELTED_ICE * ld] = s.melted_ice;
            S[SEMIC_F_REFR * ld] = s.refr;
            S[SEMIC_F_SMB * ld] = s.smb;
            S[SEMIC_F_ACC * ld] = s.acc;
            S[SEMIC_F_LHF * ld] = s.lhf;
            S[SEMIC_F_SHF * ld] = s.shf;
            S[SEMIC_F_LWU * ld] = s.lwu;
            S[SEMIC_F_SUBL * ld] = s.subl;
            S[SEMIC_F_EVAP * ld] = s.evap;
            S[SEMIC_F_SMB_SNOW * ld] = s.smb_snow;
            S[SEMIC_F_SMB_ICE * ld] = s.smb_ice;
            S[SEMIC_F_RUNOFF * ld] = s.runoff;
            S[SEMIC_F_QMR * ld] = s.qmr;
            S[SEMIC_F_QMR_RES * ld] = s.qmr_res;
            S[SEMIC_F_AMP * ld] = s.amp;
        }
    }
}

#if !SEMIC_STRICT
// ------------------------------------------------------------------------------------------------
// the benchmarked instantiation of the persistent kernel: no boundary flags, tiled forcing, 8 daily rows
// ------------------------------------------------------------------------------------------------
// Same physics (point_day) and the same warp-private TMA rings as k_run; what differs is the shell around it:
//  * the warp index is broadcast with a shuffle, so tile, ring, barrier and forcing addresses are warp-uniform to the
//    compiler: they live in uniform registers and UBLKCP takes them directly (k_run needs an ELECT/R2UR waterfall);
//  * every lane counts the issued days (only the copy itself is elected), nothing in the loop re-derives an address
//    from threadIdx;
//  * the final state is stored from inside the day loop (last day only), so the 15 diagnostics that are not daily
//    output are not live across iterations (and one copy of the day body keeps chunked runs bit-identical);
//  * lanes beyond nx compute on the padding columns (ld is a multiple of 1024) instead of branching around the day.
DEVI bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}

// AVG: instead of every day's 8 fields, their means over consecutive periods of a.avg_days days are written (the
// surface_physics_average protocol, surface_physics.f90:565-629: zero at 'init', add at every 'step', divide by the
// number of steps at 'end'), accumulated in registers: the daily 64 B/point of output traffic become 64/avg_days.
// EXT = 1: bnd%tsurf and/or bnd%alb with an external series (BASELINE configs[3]; example.f90:103-104): the day's rows
// stt / alb of the validation-layout series ride in the same stage as the forcing (one more bulk copy of 256 or 512 B).
// EXT = 2: any set of boundary flags (the GENERIC physics: all 23 fields carried, flagged fields keep their state value or
// come from the external rows stt alb smb melt acc, example.f90:103-107) on the same shell; 16 warps x 128 registers.
template <int WPC, int EXT>
struct LeanSmem {
    static constexpr int ROWS = kForcingRows + (EXT == 2 ? 6 : (EXT == 1 ? 2 : 0));
    static constexpr int STAGES = (WPC > 24) ? 2 : kStages;                                  // 227 KB of shared memory per CTA
    static constexpr size_t stage_bytes = (size_t)ROWS * kTile * sizeof(double);           // 2304 B (2816 / 3840 B with external rows)
    static constexpr size_t ring_bytes = stage_bytes * STAGES * WPC;
    static constexpr size_t bar_bytes = (size_t)WPC * STAGES * sizeof(uint64_t);
    static constexpr size_t bytes = ring_bytes + bar_bytes + kExpTab * sizeof(double);      // rings, barriers, exp table
};

// the 23 non-forcing fields of one point into its slab column (the state after the reference's last call)
DEVI void store_state(double *S, const long long ld, const Pt &s)
{
    S[SEMIC_F_T2M * ld] = s.t2m;
    S[SEMIC_F_TSURF * ld] = s.tsurf;
    S[SEMIC_F_HSNOW * ld] = s.hsnow;
    S[SEMIC_F_HICE * ld] = s.hice;
    S[SEMIC_F_ALB * ld] = s.alb;
    S[SEMIC_F_ALB_SNOW * ld] = s.alb_snow;
    S[SEMIC_F_MELT * ld] = s.melt;
    S[SEMIC_F_MELTED_SNOW * ld] = s.melted_snow;
    S[SEMIC_F_MELTED_ICE * ld] = s.melted_ice;
    S[SEMIC_F_REFR * ld] = s.refr;
    S[SEMIC_F_SMB * ld] = s.smb;
    S[SEMIC_F_ACC * ld] = s.acc;
    S[SEMIC_F_LHF * ld] = s.lhf;
    S[SEMIC_F_SHF * ld] = s.shf;
    S[SEMIC_F_LWU * ld] = s.lwu;
    S[SEMIC_F_SUBL * ld] = s.subl;
    S[SEMIC_F_EVAP * ld] = s.evap;
    S[SEMIC_F_SMB_SNOW * ld] = s.smb_snow;
    S[SEMIC_F_SMB_ICE * ld] = s.smb_ice;
    S[SEMIC_F_RUNOFF * ld] = s.runoff;
    S[SEMIC_F_QMR * ld] = s.qmr;
    S[SEMIC_F_QMR_RES * ld] = s.qmr_res;
    S[SEMIC_F_AMP * ld] = s.amp;
}
// the 8 daily rows of one point (example.f90:117-122)
DEVI void store_daily(double *oday, const Pt &s, const double swd)
{
    oday[SEMIC_O_STT * kTile]   = s.tsurf;
    oday[SEMIC_O_ALB * kTile]   = s.alb;
    oday[SEMIC_O_SWNET * kTile] = swd * (1.0 - s.alb);
    oday[SEMIC_O_SMB * kTile]   = s.smb;
    oday[SEMIC_O_MELT * kTile]  = s.melt;
    oday[SEMIC_O_ACC * kTile]   = s.acc;
    oday[SEMIC_O_SHF * kTile]   = s.shf;
    oday[SEMIC_O_LHF * kTile]   = s.lhf;
}

template <int SCHEME, int WPC, int UNR, int NK, bool AVG = false, int EXT = 0>
__global__ void __launch_bounds__(WPC * 32, 1) k_run_lean(const __grid_constant__ RunArgs a)
{
    using SM = LeanSmem<WPC, EXT>;
    constexpr int NST = SM::STAGES;
    constexpr uint32_t SB = (uint32_t)SM::stage_bytes;
    constexpr uint32_t FB = (uint32_t)(kForcingRows * kTile * sizeof(double));   // forcing bytes per tile-day
    constexpr int FBLK = kForcingRows * kTile;                   // doubles per tile-day of forcing
    constexpr int OBLK = SEMIC_NOUT * kTile;                     // doubles per tile-day of output
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t smem_u = smem_u32(smem_raw);
    const uint32_t ring_u = smem_u + (uint32_t)(warp * NST) * SB;
    const uint32_t bar_u = smem_u + (uint32_t)SM::ring_bytes + (uint32_t)(warp * NST) * 8u;
    const double *ring = reinterpret_cast<const double *>(smem_raw + (size_t)(warp * NST) * SB) + lane;
    double *tab = reinterpret_cast<double *>(smem_raw + SM::ring_bytes + SM::bar_bytes);

    if (lane == 0) {
        for (int s = 0; s < NST; ++s) mbar_init(bar_u + 8u * s, 1);
        mbar_fence_init();
    }
    for (int j = threadIdx.x; j < kExpTab; j += WPC * 32) tab[j] = g_exp2_tab[j];
    __syncthreads();

    const long long ld = a.ld;
    const long long ntl = a.sld / kTile;                         // tiles per series row
    const long long ntiles = ((long long)a.nx + kTile - 1) / kTile;
    const long long nwarps = (long long)gridDim.x * WPC;
    const long long f_day = ntl * FBLK;                          // doubles per forcing day
    const long long o_day = ntl * OBLK;
    const int ndays = a.ndays;
    __builtin_assume(a.P.n_ksub >= 1);                           // guaranteed by the launcher (n_ksub < 1 runs k_run)
    int ostart = (a.out != nullptr) ? a.out_start : ndays;       // first day that is written
    asm volatile("" : "+r"(ostart));                             // keep it in a register instead of re-deriving it daily
    uint32_t stage = 0, phase = 0;
    // external rows: stt (row 0) and alb (row 1) of the validation layout are adjacent: one copy covers what is flagged
    // (EXT = 2: also smb, melt, acc; the copy spans first .. last flagged row of stt alb swnet smb melt acc)
    const bool x_on = EXT != 0 && a.vali != nullptr;
    const bool x_ts = x_on && a.B.tsurf != 0, x_alb = x_on && a.B.alb != 0;
    const bool x_smb = EXT == 2 && x_on && a.B.smb != 0, x_melt = EXT == 2 && x_on && a.B.melt != 0;
    const bool x_acc = EXT == 2 && x_on && a.B.acc != 0;
    const int x_row0 = x_ts ? SEMIC_O_STT : x_alb ? SEMIC_O_ALB : x_smb ? SEMIC_O_SMB : x_melt ? SEMIC_O_MELT : SEMIC_O_ACC;
    const int x_row1 = x_acc ? SEMIC_O_ACC : x_melt ? SEMIC_O_MELT : x_smb ? SEMIC_O_SMB : x_alb ? SEMIC_O_ALB : SEMIC_O_STT;
    const bool x_any = x_ts || x_alb || x_smb || x_melt || x_acc;
    const uint32_t x_bytes = x_any ? (uint32_t)(x_row1 - x_row0 + 1) * kTile * (uint32_t)sizeof(double) : 0u;
    const uint32_t x_dst = FB + (uint32_t)x_row0 * kTile * (uint32_t)sizeof(double);       // stage rows 9 (stt) .. 14 (acc)
    const long long v_day = ntl * OBLK;                          // doubles per day of the external series

    // tile -> warp map interleaved over the CTAs: the ntiles % nwarps tiles of the last round land on ALL SMs (a few warps each)
    // instead of filling the first SMs completely -- 3 % on a 2 Mi-point shard (18.45 tiles per warp)
    for (long long tile = (long long)warp * gridDim.x + blockIdx.x; tile < ntiles; tile += nwarps) {
        const double *ftile = a.forcing + tile * FBLK;
        int fd = a.f_day0 % a.f_nwin;                            // next forcing day to fetch and its address
        const double *fsrc = ftile + (long long)fd * f_day;
        const double *vtile = x_any ? a.vali + tile * OBLK + (long long)x_row0 * kTile : nullptr;
        const double *vsrc = x_any ? vtile + (long long)fd * v_day : nullptr;
        int d_issue = 0;
        {
            uint32_t s = stage;
            const int npre = ndays < NST ? ndays : NST;
            for (; d_issue < npre; ++d_issue) {
                if (elect_one()) {
                    mbar_expect_tx(bar_u + 8u * s, FB + x_bytes);
                    tma_load_1d(ring_u + s * SB, fsrc, FB, bar_u + 8u * s);
                    if (EXT && x_any) tma_load_1d(ring_u + s * SB + x_dst, vsrc, x_bytes, bar_u + 8u * s);
                }
                fsrc += f_day;
                if (EXT && x_any) vsrc += v_day;
                if (++fd == a.f_nwin) { fd = 0; fsrc = ftile; if (EXT && x_any) vsrc = vtile; }
                if (++s == NST) s = 0;
            }
        }

        const long long i = tile * kTile + lane;
        int dlast = (i < a.nx) ? ndays - 1 : -1;                 // the day whose full state this lane stores
        asm volatile("" : "+r"(dlast));
        Pt s;
        {
            const double *S = a.slab + i;                        // the slab has ld >= ntiles*32 columns
            s.mask = a.mask[i];
            s.tsurf    = S[SEMIC_F_TSURF * ld];
            s.hsnow    = S[SEMIC_F_HSNOW * ld];
            s.hice     = S[SEMIC_F_HICE * ld];
            s.alb      = S[SEMIC_F_ALB * ld];
            s.alb_snow = S[SEMIC_F_ALB_SNOW * ld];
            s.qmr_res  = S[SEMIC_F_QMR_RES * ld];
            if (EXT == 2) {                                      // flagged fields keep their state value
                s.melt = S[SEMIC_F_MELT * ld];
                s.melted_snow = S[SEMIC_F_MELTED_SNOW * ld];
                s.melted_ice = S[SEMIC_F_MELTED_ICE * ld];
                s.refr = S[SEMIC_F_REFR * ld];
                s.smb = S[SEMIC_F_SMB * ld];
                s.acc = S[SEMIC_F_ACC * ld];
                s.lhf = S[SEMIC_F_LHF * ld];
                s.shf = S[SEMIC_F_SHF * ld];
                s.lwu = S[SEMIC_F_LWU * ld];
                s.subl = S[SEMIC_F_SUBL * ld];
                s.evap = S[SEMIC_F_EVAP * ld];
                s.smb_snow = S[SEMIC_F_SMB_SNOW * ld];
                s.smb_ice = S[SEMIC_F_SMB_ICE * ld];
                s.runoff = S[SEMIC_F_RUNOFF * ld];
                s.qmr = S[SEMIC_F_QMR * ld];
                s.amp = S[SEMIC_F_AMP * ld];
            } else {
                s.melt = s.melted_snow = s.melted_ice = s.refr = s.smb = s.acc = 0.0;
                s.lhf = s.shf = s.lwu = s.subl = s.evap = s.smb_snow = s.smb_ice = 0.0;
                s.runoff = s.qmr = s.amp = 0.0;
            }
            s.t2m = 0.0;
        }
        double *oday = a.out + tile * OBLK + lane;
        int oslot = 0;
        double avg[SEMIC_NOUT] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        int navg = 0;
        (void)avg; (void)navg;

        for (int d = 0; d < ndays; ++d) {
            mbar_wait(bar_u + 8u * stage, phase);
            const double *row = ring + (size_t)stage * (SB / sizeof(double));
            if (EXT) {                                           // example.f90:103-107
                if (x_ts)  s.tsurf = row[(kForcingRows + SEMIC_O_STT) * kTile];
                if (x_alb) s.alb   = row[(kForcingRows + SEMIC_O_ALB) * kTile];
                if (EXT == 2) {
                    if (x_smb)  s.smb  = row[(kForcingRows + SEMIC_O_SMB) * kTile];
                    if (x_melt) s.melt = row[(kForcingRows + SEMIC_O_MELT) * kTile];
                    if (x_acc)  s.acc  = row[(kForcingRows + SEMIC_O_ACC) * kTile];
                }
            }
            double swd;
            if (EXT == 2) point_day<SCHEME, true, false, UNR, NK, false>(s, row, kTile, a.P, a.B, a.eb_steps, a.do_mb, swd, tab);
            else          point_day<SCHEME, false, false, UNR, NK, EXT == 1>(s, row, kTile, a.P, a.B, 0, 1, swd, tab);
            if (d >= ostart) {                                   // example.f90:117-122
                if (!AVG) {
                    store_daily(oday, s, swd);
                } else {
                    avg[SEMIC_O_STT]   += s.tsurf;               // field_average 'step' (:614)
                    avg[SEMIC_O_ALB]   += s.alb;
                    avg[SEMIC_O_SWNET] += swd * (1.0 - s.alb);
                    avg[SEMIC_O_SMB]   += s.smb;
                    avg[SEMIC_O_MELT]  += s.melt;
                    avg[SEMIC_O_ACC]   += s.acc;
                    avg[SEMIC_O_SHF]   += s.shf;
                    avg[SEMIC_O_LHF]   += s.lhf;
                    ++navg;
                }
                if (!AVG || navg == a.avg_days || d == ndays - 1) {
                    if (AVG) {
                        const double nt = (double)navg;          // 'end' (:621): IEEE division by the number of steps
#pragma unroll
                        for (int q = 0; q < SEMIC_NOUT; ++q) { oday[q * kTile] = avg[q] / nt; avg[q] = 0.0; }   // 'init' (:611)
                        navg = 0;
                    }
                    oday += o_day;
                    if (++oslot == a.o_nring) { oslot = 0; oday -= (long long)a.o_nring * o_day; }
                }
            }
            if (d == dlast) {                   // final state: all 23 non-forcing fields, as after the reference's last call
                if (EXT != 2) finish_diagnostics<false>(s, NK > 0 ? NK : a.P.n_ksub);
                store_state(a.slab + i, ld, s);
            }
            // every lane has consumed its column of the stage (the loaded values fed the arithmetic above): refill it
            __syncwarp();
            if (d_issue < ndays) {
                if (elect_one()) {
                    fence_proxy_async();       // generic-proxy reads of the stage (above) before the async-proxy overwrite
                    mbar_expect_tx(bar_u + 8u * stage, FB + x_bytes);
                    tma_load_1d(ring_u + stage * SB, fsrc, FB, bar_u + 8u * stage);
                    if (EXT && x_any) tma_load_1d(ring_u + stage * SB + x_dst, vsrc, x_bytes, bar_u + 8u * stage);
                }
                ++d_issue;
                fsrc += f_day;
                if (EXT && x_any) vsrc += v_day;
                if (++fd == a.f_nwin) { fd = 0; fsrc = ftile; if (EXT && x_any) vsrc = vtile; }
            }
            if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ensemble misfit (optimize/run_particles.f90:104-178 + utils.f90:23-42), config 5
// ------------------------------------------------------------------------------------------------
// A CTA is 32 points x 8 parameter sets ("particles"): lane = point, warp = particle, so parameters
// are warp-uniform (read from shared memory) and the day's forcing/validation rows are shared by the
// 8 warps through L1.  Every trajectory starts from the same initial slab (Q14), runs nloop x ntime
// days and keeps one-pass shifted sums of d = model - validation for stt, melt, smb, swnet in the
// last loop; crmsd^2 = (sum(e^2) - sum(e)^2/n)/n / var(validation).  Per-particle sums over the CTA's
// 32 points go to partial[particle][tile] (fixed-order second stage: bitwise reproducible).
constexpr int kEnsParticles = 8;
constexpr int kEnsStages = 2;
constexpr int kEnsRows = kForcingRows + 4;                          // forcing + validation stt | swnet smb melt
struct EnsSmem {
    static constexpr uint32_t stage_bytes = kEnsRows * kTile * (uint32_t)sizeof(double);          // 3328 B
    static constexpr size_t ring_bytes = (size_t)stage_bytes * kEnsStages * kEnsParticles;
    static constexpr size_t bar_bytes = (size_t)kEnsParticles * kEnsStages * sizeof(uint64_t);
    static constexpr size_t par_bytes = (size_t)kEnsParticles * sizeof(DevPar);
    static constexpr size_t bytes = ring_bytes + bar_bytes + par_bytes + kExpTab * sizeof(double);   // 3 CTAs per SM
};

// SCHEME / NK: albedo scheme and n_ksub as compile-time constants when every particle shares them (optimize/tools.py:31,35
// writes n_ksub = 3 and alb_scheme = "none" into every particle's namelist); SCHEME_RUNTIME / 0 otherwise.
// Like k_run_lean every warp feeds itself: a private 2-stage ring, one bulk copy of the day's 9 forcing rows plus two of
// the validation rows it is scored against (stt; swnet, smb, melt are adjacent), so no load latency sits in the day loop
// (the eight warps of a CTA fetch the same tile-days: L2 hits).
template <int SCHEME, int NK>
__global__ void __launch_bounds__(32 * kEnsParticles, 3) k_ensemble(const __grid_constant__ EnsArgs a, const int ngroups)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr uint32_t SB = EnsSmem::stage_bytes;
    constexpr uint32_t FB = kForcingRows * kTile * (uint32_t)sizeof(double);
    constexpr uint32_t RB = kTile * (uint32_t)sizeof(double);                                     // one row
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t smem_u = smem_u32(smem_raw);
    const uint32_t ring_u = smem_u + (uint32_t)(warp * kEnsStages) * SB;
    const uint32_t bar_u = smem_u + (uint32_t)EnsSmem::ring_bytes + (uint32_t)(warp * kEnsStages) * 8u;
    const double *ring = reinterpret_cast<const double *>(smem_raw + (size_t)(warp * kEnsStages) * SB) + lane;
    DevPar *s_par = reinterpret_cast<DevPar *>(smem_raw + EnsSmem::ring_bytes + EnsSmem::bar_bytes);
    double *tab = reinterpret_cast<double *>(smem_raw + EnsSmem::ring_bytes + EnsSmem::bar_bytes + EnsSmem::par_bytes);

    for (int j = threadIdx.x; j < kExpTab; j += 32 * kEnsParticles) tab[j] = g_exp2_tab[j];
    const int tile = blockIdx.x / ngroups;
    const int group = blockIdx.x % ngroups;
    const int p = group * kEnsParticles + warp;
    const bool pvalid = p < a.npar;
    {   // stage the CTA's parameter sets
        const int nwords = (int)(sizeof(DevPar) / sizeof(double));
        const double *src = reinterpret_cast<const double *>(a.pars + (pvalid ? p : 0));
        double *dst = reinterpret_cast<double *>(&s_par[warp]);
        for (int w = lane; w < nwords; w += 32) dst[w] = src[w];
    }
    if (lane == 0) {
        for (int st = 0; st < kEnsStages; ++st) mbar_init(bar_u + 8u * st, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const DevPar &P = s_par[warp];
    const long long i = (long long)tile * 32 + lane;
    const bool valid = pvalid && i < a.nx;
    semic_bnd B;
    B.t2m = B.tsurf = B.hsnow = B.alb = B.melt = B.refr = B.smb = B.acc = B.lhf = B.shf = B.subl = B.amp = 0;

    const long long ld = a.ld;
    const long long ntl = ld / kTile;
    const double *ftile = a.forcing + (long long)tile * (kForcingRows * kTile);
    const double *vtile = a.vali + (long long)tile * (SEMIC_NOUT * kTile);
    const long long f_day = ntl * (kForcingRows * kTile), v_day = ntl * (SEMIC_NOUT * kTile);
    const int total = a.nloop * a.ntime;                          // days of this trajectory
    int n_issue = 0, d_issue = 0;                                 // forcing row and count of the next day to fetch
    auto issue = [&](const uint32_t st) {
        if (elect_one()) {
            fence_proxy_async();               // generic-proxy reads of the stage before the async-proxy overwrite
            const uint32_t dst = ring_u + st * SB, bar = bar_u + 8u * st;
            const double *v = vtile + (long long)n_issue * v_day;
            mbar_expect_tx(bar, FB + 4u * RB);
            tma_load_1d(dst, ftile + (long long)n_issue * f_day, FB, bar);
            tma_load_1d(dst + FB, v + SEMIC_O_STT * kTile, RB, bar);                              // row 9: stt
            tma_load_1d(dst + FB + RB, v + SEMIC_O_SWNET * kTile, 3u * RB, bar);                  // rows 10-12: swnet smb melt
        }
        ++d_issue;
        if (++n_issue == a.ntime) n_issue = 0;
    };
    for (uint32_t st = 0; st < (uint32_t)kEnsStages && d_issue < total; ++st) issue(st);

    Pt s;
    {
        const double *S = a.slab0 + i;                            // the slab has ld >= ntiles*32 columns
        s.mask = a.mask[i];
        s.tsurf = S[SEMIC_F_TSURF * ld];
        s.hsnow = S[SEMIC_F_HSNOW * ld];
        s.hice = S[SEMIC_F_HICE * ld];
        s.alb = S[SEMIC_F_ALB * ld];
        s.alb_snow = S[SEMIC_F_ALB_SNOW * ld];
        s.qmr_res = S[SEMIC_F_QMR_RES * ld];
        s.melt = s.melted_snow = s.melted_ice = s.refr = s.smb = s.acc = 0.0;
        s.lhf = s.shf = s.lwu = s.subl = s.evap = s.smb_snow = s.smb_ice = s.runoff = s.qmr = s.amp = s.t2m = 0.0;
    }

    double d0[4] = {0, 0, 0, 0}, se[4] = {0, 0, 0, 0}, see[4] = {0, 0, 0, 0};
    uint32_t stage = 0, phase = 0;
    const int first_scored = total - a.ntime;                     // the last loop is scored (run_particles.f90:148)
    for (int day = 0; day < total; ++day) {
        mbar_wait(bar_u + 8u * stage, phase);
        const double *row = ring + (size_t)stage * (SB / sizeof(double));
        double swd;
        point_day<SCHEME, false, false, (NK > 0 ? NK : 2), NK>(s, row, kTile, P, B, P.n_ksub, 1, swd, tab);
        if (day >= first_scored) {                                               // run_particles.f90:148-157
            double d[4];
            d[0] = s.tsurf - row[(kForcingRows + 0) * kTile];                    // stt
            d[1] = s.melt - row[(kForcingRows + 3) * kTile];                     // melt
            d[2] = s.smb - row[(kForcingRows + 2) * kTile];                      // smb
            d[3] = swd * (1.0 - s.alb) - row[(kForcingRows + 1) * kTile];        // swnet
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (day == first_scored) d0[q] = d[q];
                const double e = d[q] - d0[q];
                se[q] += e;
                see[q] = fma(e, e, see[q]);
            }
        }
        __syncwarp();
        if (d_issue < total) issue(stage);
        if (++stage == (uint32_t)kEnsStages) { stage = 0; phase ^= 1u; }
    }
    // sum over the 4 variables of crmsd^2 (run_particles.f90:172-174)
    double c = 0.0;
    const double n = (double)a.ntime;
    const long long ii = (i < a.nx) ? i : 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double var_d = dmax1(0.0, (see[q] - se[q] * se[q] / n) / n);
        c += var_d * a.inv_var[(long long)q * ld + ii];
    }
    if (!valid) c = 0.0;
    // fixed-order warp tree
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
    if (lane == 0 && pvalid) a.partial[(long long)p * a.ntiles + tile] = c;
}

// 1/var of the validation series per point for stt, melt, smb, swnet (utils.f90:33-34: two-pass)
__global__ void k_vali_invvar(const double *vali, int ntime, int nx, int ld, double *inv_var)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    const int vars[4] = {SEMIC_O_STT, SEMIC_O_MELT, SEMIC_O_SMB, SEMIC_O_SWNET};
    const long long day_stride = (long long)(ld / kTile) * SEMIC_NOUT * kTile;
    for (int q = 0; q < 4; ++q) {
        const double *x = vali + ((long long)(i / kTile) * SEMIC_NOUT + vars[q]) * kTile + (i % kTile);
        double sum = 0.0;
        for (int n = 0; n < ntime; ++n) sum += x[(long long)n * day_stride];
        const double avg = sum / ntime;
        double ss = 0.0;
        for (int n = 0; n < ntime; ++n) { const double d = x[(long long)n * day_stride] - avg; ss += d * d; }
        inv_var[(long long)q * ld + i] = 1.0 / (ss / ntime);
    }
}

// second stage: sumsq[p] = sum over tiles, fixed order
__global__ void k_ens_reduce(const double *partial, int ntiles, int npar, double *sumsq)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npar) return;
    double acc = 0.0;
    for (int t = 0; t < ntiles; ++t) acc += partial[(long long)p * ntiles + t];
    sumsq[p] = acc;
}
// ------------------------------------------------------------------------------------------------
// second stage + the one collective of the path in ONE kernel, over NVLink peer memory
// ------------------------------------------------------------------------------------------------
// Replaces k_ens_reduce followed by ncclAllReduce (optimize/run_particles.f90:171-177 + tools.py:58-62: the per-particle
// sums of all shards).  Every CTA reduces its particles over the local tiles into this rank's peer-visible slot; the last
// CTA to finish publishes the slot (release store of the epoch, system scope); every CTA then waits for the epoch flag of
// every rank (acquire loads through the NVLink mapping, nanosleep back-off, bounded by a time-out) and adds the ranks' values
// in rank order -- every rank computes the same bits.  Slots alternate with the epoch: a rank can only be two epochs ahead
// of a peer after that peer has finished reading, so no second handshake is needed.
DEVI void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
DEVI unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
DEVI double ld_relaxed_sys_f64(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void k_ens_reduce_p2p(const double *partial, int ntiles, int npar, double *sumsq, const P2PArgs p)
{
    __shared__ int s_bad;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int slot = (int)(p.epoch & 1ull);
    double *mine = p.data[p.rank] + (size_t)slot * kP2PMaxPar;
    if (threadIdx.x == 0) s_bad = 0;
    if (q < npar) {
        double acc = 0.0;
        for (int t = 0; t < ntiles; ++t) acc += partial[(long long)q * ntiles + t];      // fixed order, as k_ens_reduce
        mine[q] = acc;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(p.done, 1u);
        if (prev == gridDim.x - 1) {                    // every CTA of this rank has written its part
            *p.done = 0u;
            __threadfence_system();
            st_release_sys_u64(p.flags[p.rank] + slot, p.epoch);
        }
    }
    if (threadIdx.x < p.nranks) {
        const unsigned long long *f = p.flags[threadIdx.x] + slot;
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(f) < p.epoch) {
            if (clock64() - t0 > p.timeout_cycles) { s_bad = 1; atomicExch(p.err, 1); break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (s_bad || q >= npar) return;
    double sum = 0.0;
    for (int r = 0; r < p.nranks; ++r) sum += ld_relaxed_sys_f64(p.data[r] + (size_t)slot * kP2PMaxPar + q);
    sumsq[q] = sum;
}
#endif  // !SEMIC_STRICT

#undef BND
#undef BNDX

}  // namespace SEMIC_MODE_NS
}  // namespace semic
