// column_physics.cuh -- per-cell / per-column scalar physics of the land time-step (device code).
//
// Every function cites the reference code it computes (paths relative to the reference root).
// The arithmetic is written in the SAME operation order as the reference kernels so that a
// translation unit compiled with -fmad=false ("faithful" math) reproduces the reference CPU
// path rounding for rounding wherever no transcendental is involved.  `FAST` selects
// algebraically identical shortcuts (van Genuchten n = 2 via cbrt/sqrt, exp10 instead of pow,
// reciprocal multiplies); that translation unit is additionally compiled with FMA contraction.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/terrarium_b200.h"

namespace trm {

// Parameters converted once to the number format NF (like Julia's `Struct{NF}` constructors).
template <class NF>
struct DevParams {
    NF por, org, L;            // homogeneous_strat.jl:34-61 ; rho_w * Lsl (soil_energy_closures.jl:76,113)
    NF sqk[5], hc[5];          // sqrt(kappa_i) and c_i in the order water, ice, air, mineral, organic
    NF Ksat, vg_alpha, vg_n, bc_psis, bc_lambda, theta_res, Omega, vwcf;
    NF rho_a, c_a, Llg, Tref, sigma, eps_mw, albedo, emis, kappa_skin, C_h, Vmin, tau_r, beta;
    // derived constants used by FAST math only
    NF rpor, neg_inv_alpha, vg_k_exp1, vg_k_exp2, vg_inv_m_neg, vg_inv_n, r_thspan, se_off;
    NF hc_wi, hc_ia, hc_base, sqk_wi, sqk_ia, sqk_base;   // regrouped constituent sums (see energy_to_temperature)
    int32_t swrc, unsat_k, sat_halo, skin, ground_res;
    int32_t albedo_kind, rad_kind, turb_kind;   // trm_albedo_kind / trm_radiative_kind / trm_turbulent_kind
    NF th_fc;                  // field capacity (ground evaporation resistance, plant available water)
    int32_t vg_n_is_2;
    int32_t bc_k;              // Brooks-Corey: 1 / lambda when that is a small integer (the default lambda = 0.2 -> 5), else 0
};

template <class NF> struct Lim;
template <> struct Lim<float> {
    __device__ static float eps() { return 1.1920928955078125e-07f; }
    __device__ static float inf() { return CUDART_INF_F; }
    __device__ static float nan() { return CUDART_NAN_F; }
};
template <> struct Lim<double> {
    __device__ static double eps() { return 2.220446049250313e-16; }
    __device__ static double inf() { return CUDART_INF; }
    __device__ static double nan() { return CUDART_NAN; }
};

// Julia `min` / `max` on floats propagate NaN (base/math.jl); fmin/fmax would drop it.
template <class NF> __device__ __forceinline__ NF jmin(NF a, NF b) { return (a != a || b != b) ? Lim<NF>::nan() : (b < a ? b : a); }
template <class NF> __device__ __forceinline__ NF jmax(NF a, NF b) { return (a != a || b != b) ? Lim<NF>::nan() : (a < b ? b : a); }

__device__ __forceinline__ float  tpow(float a, float b)   { return powf(a, b); }
__device__ __forceinline__ double tpow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float  tsqrt(float a)  { return sqrtf(a); }
__device__ __forceinline__ double tsqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float  tcbrt(float a)  { return cbrtf(a); }
__device__ __forceinline__ double tcbrt(double a) { return cbrt(a); }
__device__ __forceinline__ float  texp(float a)   { return expf(a); }
__device__ __forceinline__ double texp(double a)  { return exp(a); }
__device__ __forceinline__ float  texp10(float a)  { return exp10f(a); }
__device__ __forceinline__ double texp10(double a) { return exp10(a); }
__device__ __forceinline__ float  tabs(float a)   { return fabsf(a); }
__device__ __forceinline__ double tabs(double a)  { return fabs(a); }
__device__ __forceinline__ float  tcos(float a)   { return cosf(a); }
__device__ __forceinline__ double tcos(double a)  { return cos(a); }
__device__ __forceinline__ float  fma_(float a, float b, float c)    { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }


// ---- math policy -------------------------------------------------------------------------------
// Faithful: IEEE division / sqrt and Julia's NaN-propagating min / max.
// Fast (FP64): reciprocal / rsqrt seeds (MUFU.RCP64H / MUFU.RSQ64H, ~20 bits) refined by ONE third-order
// (Halley type) step on the FP64 FMA pipe (error <= ~2 ulp, no IEEE fix-up branches), compare+select min / max.
template <class NF, bool FAST>
struct M {
    __device__ static __forceinline__ NF div(NF a, NF b) { return a / b; }
    __device__ static __forceinline__ NF rcp(NF a) { return 1 / a; }
    __device__ static __forceinline__ NF sqrt_(NF a) { return tsqrt(a); }
    __device__ static __forceinline__ NF rsqrt_(NF a) { return 1 / tsqrt(a); }
    __device__ static __forceinline__ NF mn(NF a, NF b) { return jmin(a, b); }
    __device__ static __forceinline__ NF mx(NF a, NF b) { return jmax(a, b); }
    __device__ static __forceinline__ NF pow23(NF x) { NF c = tcbrt(x); return c * c; }
    __device__ static __forceinline__ NF pos(NF e) { return jmax(e, NF(0)); }
    __device__ static __forceinline__ void roots(NF x, NF& x23, NF& x12) { x23 = pow23(x); x12 = sqrt_(x); }
};
static __device__ __noinline__ double pow23_cold(double x) { double c = cbrt(x); return c * c; }

template <>
struct M<double, true> {
    __device__ static __forceinline__ double rcp(double x) {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        // one third-order step: r (1 + e + e^2), e = 1 - x r ; seed error 2^-20 -> 2^-60
        const double e = fma(-x, r, 1.0);
        return fma(r, fma(e, e, e), r);
    }
    __device__ static __forceinline__ double div(double a, double b) { return a * rcp(b); }
    // seed of 1/sqrt(x) (MUFU.RSQ64H), clamped to 2^512 so that x = 0 gives finite products (0 * seed = 0)
    __device__ static __forceinline__ double rsqrt_seed(double x) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        return __hiloint2double(min(__double2hiint(y), 0x5FF00000), 0);
    }
    // one third-order step: y (1 + h + 3/2 h^2), h = 1/2 - 1/2 x y^2 ; rsqrt_(0) = 1.875 * 2^512 (finite)
    __device__ static __forceinline__ double rsqrt_(double x) {
        const double y = rsqrt_seed(x);
        const double h = fma((-0.5 * x) * y, y, 0.5);
        const double w = fma(1.5, h * h, h);
        return fma(y, w, y);
    }
    // same step in Goldschmidt form: g = x y ~ sqrt(x), sqrt(x) = g (1 + h + 3/2 h^2) ; sqrt_(0) = 0 without a branch
    __device__ static __forceinline__ double sqrt_(double x) {
        const double y = rsqrt_seed(x);
        const double g = x * y;
        const double h = fma(-g, 0.5 * y, 0.5);
        const double w = fma(1.5, h * h, h);
        return fma(g, w, g);
    }
    // max(e, 0) on the integer pipe: clear the value when its sign bit is set (-0 -> +0, NaN passes)
    __device__ static __forceinline__ double pos(double e) {
        const int hi = __double2hiint(e), keep = ~(hi >> 31);
        return __hiloint2double(hi & keep, __double2loint(e) & keep);
    }
    // compare + select (NaN in `a` falls through to `b`; fast math makes no promise about NaN states)
    __device__ static __forceinline__ double mn(double a, double b) { return a < b ? a : b; }
    __device__ static __forceinline__ double mx(double a, double b) { return a > b ? a : b; }
    // x^(2/3) for 0 < x <= 1: r ~ x^(-1/3) seeded in FP32 (otherwise idle pipe), two Newton steps
    // r <- r + r (1 - x r^3) / 3 in FP64; x^(2/3) = x r.
    __device__ static __forceinline__ double pow23(double x) {
        if (x < 1.0e-30) return pow23_cold(x);
        double r = (double)rcbrtf((float)x);
        // one third-order step for r ~ x^(-1/3): r (1 + e/3 + 2/9 e^2), e = 1 - x r^3
        const double e = fma(-(x * r) * r, r, 1.0);
        r = fma(r, e * fma(2.0 / 9.0, e, 1.0 / 3.0), r);
        return x * r;
    }
    // x^(2/3) and x^(1/2) for 0 < x <= 1 from ONE root: r ~ x^(-1/6) seeded in FP32 (lg2 / ex2 on the MUFU pipe,
    // relative error < 2^-19 for x >= 1e-30), one third-order step r (1 + e/6 + 7/72 e^2), e = 1 - x r^6, in FP64;
    // x^(2/3) = x r^2, x^(1/2) = x r^3.
    __device__ static __forceinline__ void roots(double x, double& x23, double& x12) {
        if (x < 1.0e-30) { x23 = pow23_cold(x); x12 = sqrt_(x); return; }
        float l, s;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"((float)x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(l * (-1.0f / 6.0f)));
        double r = (double)s;
        const double r3 = (r * r) * r;
        const double e = fma(-x * r3, r3, 1.0);
        r = fma(r, e * fma(7.0 / 72.0, e, 1.0 / 6.0), r);
        x23 = x * (r * r);
        x12 = x23 * r;
    }
};
template <>
struct M<float, true> {
    // MUFU seeds are within ~1 ulp of FP32 already: no refinement, no IEEE fix-up paths (no calls in the layer loop)
    __device__ static __forceinline__ float rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    __device__ static __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
    // seed of 1/sqrt(x), clamped to 2^64 so that x = 0 gives finite products (0 * seed = 0)
    __device__ static __forceinline__ float rsqrt_(float x) {
        float y;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return fminf(y, 1.8446744e19f);
    }
    __device__ static __forceinline__ float sqrt_(float x) { return x * rsqrt_(x); }
    __device__ static __forceinline__ float mn(float a, float b) { return fminf(a, b); }
    __device__ static __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
    __device__ static __forceinline__ float pow23(float x) { float c = cbrtf(x); return c * c; }
    __device__ static __forceinline__ float pos(float e) { return fmaxf(e, 0.0f); }
    // x^(2/3) and x^(1/2) from r ~ x^(-1/6) (lg2 / ex2 seed, one Newton step), like the FP64 variant
    __device__ static __forceinline__ void roots(float x, float& x23, float& x12) {
        if (x < 1.0e-30f) { x23 = pow23(x); x12 = sqrtf(x); return; }
        float l, r;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * (-1.0f / 6.0f)));
        const float r3 = (r * r) * r;
        const float e = fmaf(-x * r3, r3, 1.0f);
        r = fmaf(r, e * (1.0f / 6.0f), r);
        x23 = x * (r * r);
        x12 = x23 * r;
    }
};

// exp for the fast build (per-column surface / vegetation block: thirteen of them per column). Argument reduction
// x = n ln2 + r with the round-to-nearest magic constant, degree-13 Taylor polynomial on |r| <= ln2 / 2 (truncation
// 4e-18), scaling through the exponent field. The result is clamped to [2^-1000, 2^1000] on the INTEGER exponent n (two
// integer instructions instead of a NaN-aware Float64 min / max pair: |x| up to 2^31 ln2 still reduces correctly); no overflow
// to Inf, no denormal handling, NaN / Inf arguments give an unspecified finite or NaN value.
__device__ __forceinline__ double exp_fast(double x) {
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const int n = max(min(__double2loint(t), 1000), -1000);
    const double nf = t - 6755399441055744.0;
    double r = fma(nf, -6.93147180369123816490e-01, x);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;            // 1/13!
    p = fma(p, r, 2.08767569878681e-09);          // 1/12!
    p = fma(p, r, 2.505210838544172e-08);         // 1/11!
    p = fma(p, r, 2.755731922398589e-07);         // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);        // 1/9!
    p = fma(p, r, 2.48015873015873e-05);          // 1/8!
    p = fma(p, r, 1.984126984126984e-04);         // 1/7!
    p = fma(p, r, 1.388888888888889e-03);         // 1/6!
    p = fma(p, r, 8.333333333333333e-03);         // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);        // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);        // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}
// Julia's NaN-propagating max / min in the faithful build, compare + select in the fast build
template <class NF, bool FAST> __device__ __forceinline__ NF xmax(NF a, NF b) { return FAST ? M<NF, FAST>::mx(a, b) : jmax(a, b); }
template <class NF, bool FAST> __device__ __forceinline__ NF xmin(NF a, NF b) { return FAST ? M<NF, FAST>::mn(a, b) : jmin(a, b); }
template <class NF, bool FAST> __device__ __forceinline__ NF xexp(NF a) { return texp(a); }
template <> __device__ __forceinline__ double xexp<double, true>(double a) { return exp_fast(a); }

// a / b: IEEE division in the faithful build, reciprocal seed + one third-order step (<= 2 ulp) in the fast build
template <class NF, bool FAST>
__device__ __forceinline__ NF dv(NF a, NF b) { return FAST ? M<NF, FAST>::div(a, b) : a / b; }

// volumetric fractions, src/processes/soil/stratigraphy/soil_volume.jl:52-67,103-107
template <class NF>
struct Fractions { NF water, ice, air, mineral, organic; };

template <class NF>
__device__ __forceinline__ Fractions<NF> fractions(const DevParams<NF>& p, NF sat, NF liq) {
    Fractions<NF> f;
    NF wi = sat * p.por;
    f.water = wi * liq;
    f.ice = wi * (1 - liq);
    f.air = (1 - sat) * p.por;
    NF solid = 1 - p.por;
    f.organic = solid * p.org;
    f.mineral = solid * (1 - p.org);
    return f;
}

// InverseQuadratic bulk thermal conductivity, soil_thermal_properties.jl:90-108
template <class NF>
__device__ __forceinline__ NF thermal_conductivity(const DevParams<NF>& p, NF sat, NF liq) {
    Fractions<NF> f = fractions(p, sat, liq);
    NF s = p.sqk[0] * f.water + p.sqk[1] * f.ice + p.sqk[2] * f.air + p.sqk[3] * f.mineral + p.sqk[4] * f.organic;
    return s * s;
}
// same sum regrouped as wi (liq (k_w - k_i) + (k_i - k_a)) + (k_a por + solid part), k = sqrt(kappa) (fast math only)
template <class NF>
__device__ __forceinline__ NF thermal_conductivity_fast(const DevParams<NF>& p, NF sat, NF liq) {
    const NF s = fma_(sat * p.por, fma_(liq, p.sqk_wi, p.sqk_ia), p.sqk_base);
    return s * s;
}

// volumetric heat capacity, soil_thermal_properties.jl:110-123
template <class NF>
__device__ __forceinline__ NF heat_capacity(const DevParams<NF>& p, NF sat, NF liq) {
    Fractions<NF> f = fractions(p, sat, liq);
    return p.hc[0] * f.water + p.hc[1] * f.ice + p.hc[2] * f.air + p.hc[3] * f.mineral + p.hc[4] * f.organic;
}

// liquid fraction inside the phase change zone -L theta <= U < 0 (soil_energy_closures.jl:139-159); out of line:
// it is rare, and inlined it would be if-converted into predicated instructions that every cell pays for
template <class NF>
__device__ __noinline__ NF partial_liquid_fraction_cold(NF U, NF Lt) { return 1 - U / (Lim<NF>::eps() - Lt); }

// energy_to_temperature + liquid_water_fraction (free water freeze curve),
// soil_energy_closures.jl:99-159 ; safediv utils/utils.jl:25 ; Bool * x is a strong zero.
template <class NF, bool FAST>
__device__ __forceinline__ void energy_to_temperature(const DevParams<NF>& p, NF U, NF sat, NF& T, NF& liq) {
    if (FAST) {
        // same three regimes with selects; only the (rare) phase change zone takes a branch (out of line).
        // C = sum_i c_i theta_i regrouped as wi (liq (c_w - c_i) + (c_i - c_a)) + (c_a por + solid part).
        const NF wi = sat * p.por;
        const NF Lt = p.L * wi;
        const bool thawed = U >= 0, frozen = U < -Lt;
        const NF num = thawed ? U : (frozen ? U + Lt : NF(0));
        liq = thawed ? NF(1) : NF(0);
        if (!thawed && !frozen) liq = partial_liquid_fraction_cold(U, Lt);
        const NF C = fma_(wi, fma_(liq, p.hc_wi, p.hc_ia), p.hc_base);
        T = num * M<NF, FAST>::rcp(C);
        return;
    }
    NF Lt = p.L * sat * p.por;
    if (U >= 0) {
        liq = 1;
    } else if (U >= -Lt) {
        NF y = -Lt;
        NF sd = (y == 0) ? Lim<NF>::inf() : U / (y + Lim<NF>::eps());
        liq = 1 - sd;
    } else {
        liq = 0;
    }
    NF C = heat_capacity(p, sat, liq);
    if (U < -Lt) T = (U + Lt) / C;
    else if (U >= 0) T = U / C;
    else T = 0;
}

// temperature_to_energy (initialisation), soil_energy_closures.jl:64-97
template <class NF>
__device__ __forceinline__ void temperature_to_energy(const DevParams<NF>& p, NF T, NF sat, NF& U, NF& liq) {
    liq = T >= 0 ? NF(1) : NF(0);
    NF C = heat_capacity(p, sat, liq);
    U = T * C - p.L * sat * p.por * (1 - liq);
}

// hydraulic conductivity at a cell centre, soil_hydraulic_properties.jl:170-221
// (real branch of the complex-valued formula; exponent n/(n+1) as coded)
template <class NF>
__device__ __forceinline__ NF cell_conductivity_reference(const DevParams<NF>& p, NF sat, NF liq) {
    Fractions<NF> f = fractions(p, sat, liq);
    if (p.unsat_k == TRM_UNSATK_LINEAR) {
        NF thsat = f.water + f.ice + f.air;
        return p.Ksat * f.water / thsat;
    }
    NF n = p.vg_n;
    NF x = f.water / p.por;
    NF I_ice = tpow(NF(10), -p.Omega * (1 - liq));
    NF inner = 1 - tpow(x, n / (n + 1));
    NF a = 1 - tpow(inner, (n - 1) / n);
    return tabs(p.Ksat * I_ice * tsqrt(x) * (a * a));
}
// out-of-line copy for the cold branches of the fast path (keeps the hot loop small in the instruction cache)
template <class NF>
__device__ __noinline__ NF cell_conductivity_cold(const DevParams<NF>& p, NF sat, NF liq) { return cell_conductivity_reference(p, sat, liq); }
template <class NF>
__device__ __noinline__ NF ice_impedance_cold(NF Omega, NF liq) { return texp10(-Omega * (1 - liq)); }

// VG2 (compile time): the caller has checked on the host that the soil is van Genuchten with n = 2 for both the
// retention curve and the conductivity, so the run-time tests for the general (cold) formulas drop out of the loop
template <class NF, bool FAST, bool VG2 = false>
__device__ __forceinline__ NF cell_conductivity(const DevParams<NF>& p, NF sat, NF liq) {
    if (FAST) {
        if (!VG2 && (!p.vg_n_is_2 || p.unsat_k == TRM_UNSATK_LINEAR)) return cell_conductivity_cold(p, sat, liq);
        // van Genuchten n = 2. End members are exact in the reference formula too:
        // x = 0 -> K = 0, x = 1 (saturated, thawed) -> K = K_sat
        const NF x = sat * liq;
        if (x == NF(0)) return NF(0);
        NF I_ice = NF(1);
        if (liq != NF(1)) I_ice = ice_impedance_cold(p.Omega, liq);
        if (x == NF(1)) return p.Ksat * I_ice;
        // |.| guards the square root against a -1 ulp residue of the approximate power when x -> 1
        NF x23, x12;
        M<NF, FAST>::roots(x, x23, x12);
        const NF a = 1 - M<NF, FAST>::sqrt_(tabs(1 - x23));
        return tabs(p.Ksat * I_ice * x12 * (a * a));
    }
    return cell_conductivity_reference(p, sat, liq);
}

// inverse soil water retention curve psi_m(theta; theta_sat) [FreezeCurves.jl 0.9 VanGenuchten /
// BrooksCorey], called at soil_hydraulic_closures.jl:115-118
template <class NF>
__device__ __forceinline__ NF swrc_inverse_reference(const DevParams<NF>& p, NF theta, NF thsat) {
    if (p.swrc == TRM_SWRC_VANGENUCHTEN) {
        if (!(theta < thsat)) return NF(0);
        NF se = (theta - p.theta_res) / (thsat - p.theta_res);
        NF n = p.vg_n, m = 1 - 1 / n;
        return -1 / p.vg_alpha * tpow(tpow(se, -1 / m) - NF(1), 1 / n);
    }
    if (!(theta < thsat)) return -p.bc_psis;
    NF se = (theta - p.theta_res) / (thsat - p.theta_res);
    return -p.bc_psis * tpow(se, -1 / p.bc_lambda);
}
template <class NF>
__device__ __noinline__ NF swrc_inverse_cold(const DevParams<NF>& p, NF theta, NF thsat) { return swrc_inverse_reference(p, theta, thsat); }

template <class NF, bool FAST, bool VG2 = false>
__device__ __forceinline__ NF swrc_inverse(const DevParams<NF>& p, NF theta, NF thsat) {
    if (FAST) {
        if (!VG2 && (p.swrc != TRM_SWRC_VANGENUCHTEN || !p.vg_n_is_2)) return swrc_inverse_cold(p, theta, thsat);
        // m = 1/2: psi_m = -(1/alpha) sqrt(se^-2 - 1) ; se = 0 (dry layer) is -Inf as in the reference
        const NF se = fma_(theta, p.r_thspan, p.se_off);   // (theta - theta_res) / (theta_sat - theta_res)
        const NF t = se * se;
        // sqrt((1 - t) / t) = a / sqrt(a t), a = 1 - se^2 from one FMA (no cancellation as se -> 1), one
        // reciprocal square root and no branch: a = 0 gives 0 * (finite seed) = 0 ; |.|: rounding guard as se -> 1
        const NF a = tabs(fma_(-se, se, NF(1)));
        NF r = (p.neg_inv_alpha * a) * M<NF, FAST>::rsqrt_(a * t);
        r = t == NF(0) ? -Lim<NF>::inf() : r;
        return theta < thsat ? r : NF(0);
    }
    return swrc_inverse_reference(p, theta, thsat);
}

// total pressure head, saturation_to_pressure! soil_hydraulic_closures.jl:102-129 ; psiz = zc - zref
template <class NF, bool FAST, bool VG2 = false>
__device__ __forceinline__ NF pressure_head(const DevParams<NF>& p, NF sat, NF wt, NF zc, NF psiz) {
    NF psim = swrc_inverse<NF, FAST, VG2>(p, sat * p.por, p.por);
    NF psih = M<NF, FAST>::pos(wt - zc);
    return psih + psim + psiz;
}

// Julia Base.Math.pow_body(x::Float64, 4) (compensated power by squaring); Float32 `x^4` goes
// through Float64 and rounds once.  Used for (Ts + Tref)^4, physical_constants.jl:67.
__device__ __forceinline__ double pow4(double x) {
    double y = 1.0, xnlo = 0.0, ynlo = 0.0;
#pragma unroll
    for (int it = 0; it < 2; ++it) {   // n = 4 -> 2 -> 1: two squarings, the odd branch is never taken
        double err = x * 2 * xnlo;
        double hi = x * x;
        double lo = fma(x, x, -hi);
        x = hi; xnlo = lo + err;
    }
    double err = fma(y, xnlo, x * ynlo);
    return (isfinite(x) && isfinite(err)) ? fma(x, y, err) : x * y;
}
__device__ __forceinline__ float pow4(float x) { double d = (double)x; double d2 = d * d; return (float)(d2 * d2); }

// saturation vapour pressure (August-Roche-Magnus), physics_utils.jl:54-73
template <class NF, bool FAST = false>
__device__ __forceinline__ NF saturation_vapor_pressure(NF T) {
    if (FAST) {   // the same two formulas with the constants selected first: one division and one exponential per call
        const NF a = T <= 0 ? NF(22.46) : NF(17.62), b = T <= 0 ? NF(272.62) : NF(243.12);
        return NF(611.0) * xexp<NF, FAST>(dv<NF, FAST>(a * T, T + b));
    }
    return T <= 0 ? NF(611.0) * xexp<NF, FAST>(dv<NF, FAST>(NF(22.46) * T, T + NF(272.62))) : NF(611.0) * xexp<NF, FAST>(dv<NF, FAST>(NF(17.62) * T, T + NF(243.12)));
}

}  // namespace trm
