// euler2_kernel.cuh -- Float32 ForwardEuler / Heun stages with TWO adjacent columns per thread and packed f32x2 math.
//
// Float32 is the reference's number format for every global configuration and for its own benchmark
// (test/benchmarks/gpu/soil_heat_hydrology_global.jl:40). The one-column-per-thread kernel (euler_kernel.cuh) halves the
// bytes of the Float64 kernel but not its instruction count, and is issue bound. Here a thread owns the columns c and c+1
// (c even), i.e. one 8-byte element of every [layer][column] row:
//   * every shared-memory access (pipeline strip, cp.async ring), every cp.async and every global store moves 8 bytes and
//     serves two columns -- the per-iteration bookkeeping (addresses, ring slots, loop control, metrics) is paid once;
//   * additions, multiplications and fused multiply-adds run as sm_100 packed instructions (add / mul / fma.rn.f32x2 ->
//     FADD2 / FMUL2 / FFMA2): one issue slot for both columns; comparisons, selects, min / max and the MUFU seeds stay per
//     column;
//   * two independent dependency chains per thread hide the fixed-latency stalls that dominate the scalar kernel.
// Same algorithm, same pipeline and the same shared-memory layout idea as euler_kernel (see its header); the per-cell
// formulas are the FAST (reciprocal / rsqrt based, van Genuchten n = 2) forms of column_physics.cuh restated on pairs.
// Columns whose saturation goes negative take the per-column slow path of the scalar kernel. Fast math and recomputed closure
// fields only, for two soils (template parameter SOIL): van Genuchten n = 2 retention curve + conductivity (the configuration
// of BASELINE.json) and the reference's DEFAULT hydraulics, Brooks-Corey with an integer 1 / lambda + linear conductivity
// (ConstantSoilHydraulics(), the soil of test/benchmarks/gpu/soil_heat_hydrology_global.jl). Every other case runs euler_kernel.
#pragma once

#include "euler_kernel.cuh"

namespace trm {

// A pair of adjacent columns. Float32: one 8-byte value, packed f32x2 arithmetic. Float64: one 16-byte value; the
// arithmetic stays per lane (sm_100 has no packed FP64), the win is everything else -- 16-byte shared-memory accesses,
// cp.async and global stores, one set of addresses / ring slots / metrics / loop control for two columns, and two
// independent dependency chains per thread.
template <class T> struct P2;
template <> struct alignas(8) P2<float> { float x, y; };
template <> struct alignas(16) P2<double> { double x, y; };
using F2 = P2<float>;
using D2 = P2<double>;
struct B2 { bool x, y; };
template <class T> struct SameT { using type = T; };   // (non-deduced scalar operands: `pair * 0.5f` works for both formats)
template <class T> using Scalar = typename SameT<T>::type;

#define TRM_U64(v) reinterpret_cast<uint64_t&>(v)
#define TRM_CU64(v) reinterpret_cast<const uint64_t&>(v)
template <class T> __device__ __forceinline__ P2<T> bc2(T s) { return P2<T>{s, s}; }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { F2 c; asm("add.f32x2 %0, %1, %2;" : "=l"(TRM_U64(c)) : "l"(TRM_CU64(a)), "l"(TRM_CU64(b))); return c; }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { F2 c; asm("mul.f32x2 %0, %1, %2;" : "=l"(TRM_U64(c)) : "l"(TRM_CU64(a)), "l"(TRM_CU64(b))); return c; }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(TRM_U64(d)) : "l"(TRM_CU64(a)), "l"(TRM_CU64(b)), "l"(TRM_CU64(c))); return d; }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return fma2(b, bc2(-1.0f), a); }   // a - b as one FFMA2 (exact product)
__device__ __forceinline__ D2 operator+(D2 a, D2 b) { return D2{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ D2 operator*(D2 a, D2 b) { return D2{a.x * b.x, a.y * b.y}; }
__device__ __forceinline__ D2 fma2(D2 a, D2 b, D2 c) { return D2{fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)}; }
__device__ __forceinline__ D2 operator-(D2 a, D2 b) { return D2{a.x - b.x, a.y - b.y}; }
template <class T> __device__ __forceinline__ P2<T> operator+(P2<T> a, Scalar<T> s) { return a + bc2<T>(s); }
template <class T> __device__ __forceinline__ P2<T> operator*(P2<T> a, Scalar<T> s) { return a * bc2<T>(s); }
template <class T> __device__ __forceinline__ P2<T> fma2(P2<T> a, Scalar<T> b, P2<T> c) { return fma2(a, bc2<T>(b), c); }
template <class T> __device__ __forceinline__ P2<T> fma2(P2<T> a, P2<T> b, Scalar<T> c) { return fma2(a, b, bc2<T>(c)); }
template <class T> __device__ __forceinline__ P2<T> fma2(P2<T> a, Scalar<T> b, Scalar<T> c) { return fma2(a, bc2<T>(b), bc2<T>(c)); }
template <class T> __device__ __forceinline__ P2<T> sel(B2 m, P2<T> a, P2<T> b) { return P2<T>{m.x ? a.x : b.x, m.y ? a.y : b.y}; }
template <class T> __device__ __forceinline__ P2<T> sel(B2 m, P2<T> a, Scalar<T> b) { return P2<T>{m.x ? a.x : b, m.y ? a.y : b}; }
template <class T> __device__ __forceinline__ P2<T> sel(B2 m, Scalar<T> a, P2<T> b) { return P2<T>{m.x ? a : b.x, m.y ? a : b.y}; }
template <class T> __device__ __forceinline__ P2<T> sel2(B2 m, T a, T b) { return P2<T>{m.x ? a : b, m.y ? a : b}; }
template <class T> __device__ __forceinline__ P2<T> min2(P2<T> a, P2<T> b) { return P2<T>{M<T, true>::mn(a.x, b.x), M<T, true>::mn(a.y, b.y)}; }
template <class T> __device__ __forceinline__ P2<T> pos2(P2<T> a) { return P2<T>{M<T, true>::pos(a.x), M<T, true>::pos(a.y)}; }
template <class T> __device__ __forceinline__ P2<T> abs2(P2<T> a) { return P2<T>{tabs(a.x), tabs(a.y)}; }
template <class T> __device__ __forceinline__ P2<T> rcp2(P2<T> a) { return P2<T>{M<T, true>::rcp(a.x), M<T, true>::rcp(a.y)}; }
template <class T> __device__ __forceinline__ P2<T> rsqrt2(P2<T> a) { return P2<T>{M<T, true>::rsqrt_(a.x), M<T, true>::rsqrt_(a.y)}; }   // finite at 0, see M<T, true>
__device__ __forceinline__ B2 operator&&(B2 a, B2 b) { return B2{a.x && b.x, a.y && b.y}; }
__device__ __forceinline__ B2 operator!(B2 a) { return B2{!a.x, !a.y}; }
__device__ __forceinline__ bool any(B2 a) { return a.x || a.y; }

// pair-wide shared-memory accesses through explicit shared addresses (same ordering argument as sts / ldsv in euler_kernel.cuh)
__device__ __forceinline__ void sts2(uint32_t a, F2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(a), "f"(v.x), "f"(v.y)); }
__device__ __forceinline__ void sts2(uint32_t a, D2 v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a), "d"(v.x), "d"(v.y)); }
__device__ __forceinline__ F2 lds2(uint32_t a, float*) { F2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ D2 lds2(uint32_t a, double*) { D2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void stg2(float* p, F2 v) {
#ifdef TRM_NO_STCS
    *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y);
#else
    __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
#endif
}
__device__ __forceinline__ void stg2(double* p, D2 v) {
#ifdef TRM_NO_STCS
    *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y);
#else
    __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
#endif
}
__device__ __forceinline__ F2 ldg2(const float* p) { const float2 v = *reinterpret_cast<const float2*>(p); return F2{v.x, v.y}; }
__device__ __forceinline__ D2 ldg2(const double* p) { const double2 v = *reinterpret_cast<const double2*>(p); return D2{v.x, v.y}; }

// ---- per-cell physics on pairs (FAST forms of column_physics.cuh) --------------------------------------------------
// energy_to_temperature + liquid_water_fraction, soil_energy_closures.jl:99-159 ; wi = sat * por is handed back for the
// conductivity
template <class T>
__device__ __forceinline__ void energy_to_temperature2(const DevParams<T>& p, P2<T> U, P2<T> sat, P2<T>& T_, P2<T>& liq, P2<T>& wi) {
    wi = sat * p.por;
    const P2<T> Lt = wi * p.L;
    const B2 thawed{U.x >= T(0), U.y >= T(0)}, frozen{U.x < -Lt.x, U.y < -Lt.y};
    const P2<T> UL = U + Lt;
    const P2<T> num{thawed.x ? U.x : (frozen.x ? UL.x : T(0)), thawed.y ? U.y : (frozen.y ? UL.y : T(0))};
    liq = sel2<T>(thawed, T(1), T(0));
    const B2 partial{!thawed.x && !frozen.x, !thawed.y && !frozen.y};
    if (any(partial)) {   // phase change zone: rare, out of line, ONE divergent region for the pair
        if (partial.x) liq.x = partial_liquid_fraction_cold(U.x, Lt.x);
        if (partial.y) liq.y = partial_liquid_fraction_cold(U.y, Lt.y);
    }
    const P2<T> C = fma2<T>(wi, fma2<T>(liq, p.hc_wi, p.hc_ia), p.hc_base);
    T_ = num * rcp2(C);
}
// InverseQuadratic bulk thermal conductivity, soil_thermal_properties.jl:90-108 (regrouped constituent sum)
template <class T>
__device__ __forceinline__ P2<T> thermal_conductivity2(const DevParams<T>& p, P2<T> wi, P2<T> liq) {
    const P2<T> s = fma2<T>(wi, fma2<T>(liq, p.sqk_wi, p.sqk_ia), p.sqk_base);
    return s * s;
}
// r ~ x^(-1/6) of both lanes for 0 < x <= 1: lg2 / ex2 seed on the MUFU pipe in FP32 (relative error < 2^-19; the argument
// is clamped away from 0, a lane with x <= 1e-30 ends in the end member K = 0 to within 1e-60 K_sat) and one Newton step
// (Float32: second order ; Float64: the third-order step of M<double, true>::roots)
__device__ __forceinline__ F2 inv_sixth_root2(F2 x) {
    F2 r;
    float lx, ly;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lx) : "f"(fmaxf(x.x, 1.0e-30f)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(ly) : "f"(fmaxf(x.y, 1.0e-30f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(lx * (-1.0f / 6.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(ly * (-1.0f / 6.0f)));
    const F2 r3 = (r * r) * r;
    const F2 ne = fma2<float>(x * r3, r3, -1.0f);            // -(1 - x r^6)
    return fma2(r, ne * (-1.0f / 6.0f), r);
}
__device__ __forceinline__ D2 inv_sixth_root2(D2 x) {
    float lx, ly, sx, sy;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lx) : "f"(fmaxf((float)x.x, 1.0e-30f)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(ly) : "f"(fmaxf((float)x.y, 1.0e-30f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sx) : "f"(lx * (-1.0f / 6.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sy) : "f"(ly * (-1.0f / 6.0f)));
    D2 r{(double)sx, (double)sy};
    const D2 r3 = (r * r) * r;
    const D2 e = D2{fma(-x.x * r3.x, r3.x, 1.0), fma(-x.y * r3.y, r3.y, 1.0)};
    return D2{fma(r.x, e.x * fma(7.0 / 72.0, e.x, 1.0 / 6.0), r.x), fma(r.y, e.y * fma(7.0 / 72.0, e.y, 1.0 / 6.0), r.y)};
}
// sqrt of both lanes, exactly 0 at 0 (Float32: d rsqrt(d) with the clamped seed ; Float64: Goldschmidt form of M<double, true>)
__device__ __forceinline__ F2 sqrt2(F2 d) { return d * rsqrt2(d); }
__device__ __forceinline__ D2 sqrt2(D2 d) { return D2{M<double, true>::sqrt_(d.x), M<double, true>::sqrt_(d.y)}; }
// hydraulic conductivity at the cell centres, van Genuchten n = 2 (soil_hydraulic_properties.jl:170-221, exponent n/(n+1) as
// coded): K = K_sat I_ice sqrt(x) (1 - sqrt(1 - x^(2/3)))^2 with the exact end members x = 0 -> 0, x = 1 -> K_sat I_ice
template <class T>
__device__ __forceinline__ P2<T> cell_conductivity2(const DevParams<T>& p, P2<T> sat, P2<T> liq) {
    const P2<T> x = sat * liq;
    P2<T> KI = bc2<T>(p.Ksat);
    const B2 icy{liq.x != T(1) && x.x != T(0), liq.y != T(1) && x.y != T(0)};
    if (any(icy)) {   // (partly frozen and wet: rare, out of line)
        if (icy.x) KI.x = p.Ksat * ice_impedance_cold(p.Omega, liq.x);
        if (icy.y) KI.y = p.Ksat * ice_impedance_cold(p.Omega, liq.y);
    }
    const B2 zero{x.x == T(0), x.y == T(0)}, one{x.x == T(1), x.y == T(1)};
    P2<T> K = KI;
    if (!((zero.x || one.x) && (zero.y || one.y))) {   // (whole saturated / frozen / dry zones skip the roots)
        // x^(2/3) and x^(1/2) from one r ~ x^(-1/6), as M<T, true>::roots
        const P2<T> r = inv_sixth_root2(x);
        const P2<T> x23 = x * (r * r), x12 = x23 * r;
        const P2<T> d = abs2(bc2<T>(T(1)) - x23);          // |.|: guards the root against a -1 ulp residue as x -> 1
        const P2<T> a = bc2<T>(T(1)) - sqrt2(d);
        K = abs2((KI * x12) * (a * a));
    }
    K = sel<T>(one, KI, K);
    return sel<T>(zero, T(0), K);
}
// UnsatKLinear (soil_hydraulic_properties.jl:170-198): K = K_sat water / (water + ice + air), the three fractions formed as
// in volumetric_fractions (soil_volume.jl:52-67)
template <class T>
__device__ __forceinline__ P2<T> cell_conductivity2_linear(const DevParams<T>& p, P2<T> sat, P2<T> liq) {
    const P2<T> wi = sat * p.por;
    const P2<T> water = wi * liq;
    const P2<T> ice = wi * (bc2<T>(T(1)) - liq);
    const P2<T> air = (bc2<T>(T(1)) - sat) * p.por;
    return (water * p.Ksat) * rcp2((water + ice) + air);
}
enum Soil2 { SOIL2_VG2 = 0, SOIL2_BC_LINEAR = 1 };
template <int SOIL, class T>
__device__ __forceinline__ P2<T> cell_conductivity2s(const DevParams<T>& p, P2<T> sat, P2<T> liq) {
    return SOIL == SOIL2_VG2 ? cell_conductivity2(p, sat, liq) : cell_conductivity2_linear(p, sat, liq);
}
// Brooks-Corey matric head with an integer exponent k = 1 / lambda: psi_m = -psi_s se^-k below saturation, -psi_s at
// saturation (FreezeCurves BrooksCorey); se^k by repeated multiplication (k is launch-uniform, <= 8)
template <class T>
__device__ __forceinline__ P2<T> brooks_corey_psim2(const DevParams<T>& p, P2<T> theta) {
    const P2<T> se = fma2<T>(theta, p.r_thspan, p.se_off);
    P2<T> pw = se;
#pragma unroll 1
    for (int i = 1; i < p.bc_k; ++i) pw = pw * se;
    P2<T> r = rcp2(pw) * -p.bc_psis;          // se = 0 (dry layer): rcp(0) = +Inf -> -Inf as in the reference
    if constexpr (sizeof(T) == 8) r = P2<T>{pw.x == T(0) ? -Lim<T>::inf() : r.x, pw.y == T(0) ? -Lim<T>::inf() : r.y};   // (the refined reciprocal is NaN at 0)
    return P2<T>{theta.x < p.por ? r.x : -p.bc_psis, theta.y < p.por ? r.y : -p.bc_psis};
}
// total pressure head, saturation_to_pressure! (soil_hydraulic_closures.jl:102-129) with the van Genuchten n = 2 retention
// curve: psi_m = -(1/alpha) sqrt(se^-2 - 1) = -(1/alpha) a / sqrt(a t), a = 1 - se^2, t = se^2 (see swrc_inverse)
template <int SOIL = SOIL2_VG2, class T>
__device__ __forceinline__ P2<T> pressure_head2(const DevParams<T>& p, P2<T> sat, P2<T> wt, T zc, T psiz) {
    const P2<T> theta = sat * p.por;
    if (SOIL == SOIL2_BC_LINEAR) return (pos2(wt + (-zc)) + brooks_corey_psim2(p, theta)) + psiz;
    const P2<T> se = fma2<T>(theta, p.r_thspan, p.se_off);
    const P2<T> t = se * se;
    const P2<T> a = abs2(fma2<T>(se * T(-1), se, T(1)));
    P2<T> r = (a * p.neg_inv_alpha) * rsqrt2(a * t);
    const T ninf = -Lim<T>::inf();
    r = P2<T>{t.x == T(0) ? ninf : r.x, t.y == T(0) ? ninf : r.y};
    const P2<T> psim{theta.x < p.por ? r.x : T(0), theta.y < p.por ? r.y : T(0)};
    const P2<T> psih = pos2(wt + (-zc));
    return (psih + psim) + psiz;
}
// Oceananigans halo fill on pairs (halo_value, stage_kernel.cuh); edge iterations only
template <class T>
__device__ __forceinline__ P2<T> halo_value2(int kind, P2<T> edge, P2<T> v, T D, bool top) {
    return P2<T>{halo_value(kind, edge.x, v.x, D, top), halo_value(kind, edge.y, v.y, D, top)};
}

// threads per block of the pair kernels (the stride of their shared-memory strips): Float64 pairs need 448 bytes of shared
// memory per thread, so smaller blocks pack more warps into an SM
#ifndef TRM_EULER2_F64_BLOCK
#define TRM_EULER2_F64_BLOCK 128
#endif
template <class T> __host__ __device__ constexpr int euler2_block() { return sizeof(T) == 8 ? TRM_EULER2_F64_BLOCK : TRM_EULER_BLOCK; }

// Strip slots of the pair kernels: the soil moisture limiting factor slot exists only in the LandModel variants; the prefetched
// top boundary temperature follows in a slot of its own (BCT_SLOT).
template <class T, int MS, int MODE, bool LAND>
struct Euler2Smem {
    static constexpr int PF_ = euler_pf(MODE);
    static constexpr bool CMET = MS == EULER_MS_SMALL;                                 // compact rows: read from the kernel parameters (MetricsC)
    static constexpr int METRICS = CMET ? 0 : (met_rows(LAND) * MS + 1) / 2 * 2;        // scalars (the strips start pair-aligned)
    static constexpr int BCT_SLOT = LAND ? EF_BETA + 1 : EF_BETA;                      // prefetched TEMPERATURE_TOP boundary value
    static constexpr int NSTRIP = BCT_SLOT + 1;
    static constexpr int STRIP = NSTRIP * euler2_block<T>();                          // pair slots
    static constexpr int RING = 2 * EULER_RD * euler2_block<T>();                     // U, sat
    // Heun stage 2: k1U, k1S, bU, bS of the layer to be updated (stored protocol) / k1U, k1S next to U, sat (recompute protocol)
    static constexpr int XRING = (MODE != MODE_HEUN2 ? 0 : (heun_recompute<T>() ? 2 * EULER_RD : 4 * PF_)) * euler2_block<T>();
    static constexpr size_t BYTES = sizeof(T) * (size_t)METRICS + 2 * sizeof(T) * (size_t)(STRIP + RING + XRING);
};

#ifndef TRM_EULER2_BLOCKS
#define TRM_EULER2_BLOCKS 6
#endif
#ifndef TRM_EULER2_H2_BLOCKS
#define TRM_EULER2_H2_BLOCKS 4   // Float32 Heun stage 2
#endif
#ifndef TRM_EULER2_F64_BLOCKS
#define TRM_EULER2_F64_BLOCKS 4   // 16-byte slots: 52 KB of shared memory per 128-thread block (10 strip slots + the rings)
#endif
template <class T, int PHYS, int MODE, int MS>
constexpr int euler2_min_blocks() { return sizeof(T) == 8 ? ((phys_land(PHYS) || MS != EULER_MS_SMALL) ? 3 : TRM_EULER2_F64_BLOCKS) : (MODE == MODE_HEUN2 ? TRM_EULER2_H2_BLOCKS : (phys_land(PHYS) ? 5 : TRM_EULER2_BLOCKS)); }

// slow path of one column: a layer went negative. Downward sweep (soil_hydrology.jl:201-216) top -> bottom on the raw
// profile the thread has just stored, then water table and closures bottom -> top (same code as the scalar kernel).
template <class T, class Met, int MODE, bool LAND, bool VG2>
__device__ __noinline__ void euler2_slow_column(const StageArgs<T>& A, Met met, int64_t c, T Sx_new) {
    constexpr bool H1 = MODE == MODE_HEUN1;
    using NF = T;
    const DevParams<T>& p = A.p;
    const int nz = A.nz;
    const int64_t ld = A.ld;
    NF carry_dn = NF(0);
#pragma unroll 1
    for (int k = nz; k >= 1; --k) {
        const int64_t o = (int64_t)(k - 1) * ld + c;
        NF s = A.yS[o];
        if (k < nz) s -= carry_dn;
        if (k >= 2) {
            const NF d = jmax(-s, NF(0));
            s += d;
            carry_dn = d * met.dzc(k) / met.dzc(k - 1);
        }
        if (k == nz) {
            const NF e = jmax(s - 1, NF(0));
            s -= e;
            Sx_new += e * met.dzc(nz);
        }
        if (k == 1) s = jmax(s, NF(0));
        A.yS[o] = s;
    }
    int idx = 0;
#pragma unroll 1
    for (int k = 1; k <= nz; ++k) if (idx == 0 && A.yS[(int64_t)(k - 1) * ld + c] < 1) idx = k;
    if (idx == 0) idx = nz + 1;
    const NF wt_new = met.zF(idx);
    A.yWt[c] = wt_new;
    if (H1 && !(LAND && has_veg(A))) return;
    if (!H1) A.ySx[c] = Sx_new;
    NF beta = NF(0);
#pragma unroll 1
    for (int k = 1; k <= nz; ++k) {
        const int64_t o = (int64_t)(k - 1) * ld + c;
        NF s = A.yS[o], U = A.yU[o], Tc, lc;
        energy_to_temperature<NF, true>(p, U, s, Tc, lc);
        if (LAND && has_veg(A)) beta += plant_available_water_fast(A.vp, p, s, lc) * met.root(k);
        if (H1) continue;   // the stage state keeps no closure fields
        A.yT[o] = Tc; A.yL[o] = lc;
        if (k == nz && A.hio_out) A.hio_out[c] = Tc;
        A.yP[o] = pressure_head<NF, true, VG2>(p, s, wt_new, met.zC(k), met.psiz(k));
    }
    if (LAND && has_veg(A)) A.ybeta[c] = beta;
}

template <class T, int PHYS, int MS, int MODE, int SOIL = SOIL2_VG2>
__global__ void __launch_bounds__(euler2_block<T>(), (euler2_min_blocks<T, PHYS, MODE, MS>())) euler2_kernel(const __grid_constant__ StageArgs<T> A) {
    using F2 = P2<T>;   // (the pair type of this instantiation: Float32 or Float64)
    constexpr bool RICH = phys_richards(PHYS);
    constexpr bool LAND = phys_land(PHYS);
    constexpr int B = euler2_block<T>();
    constexpr int ES = 2 * (int)sizeof(T);   // bytes per strip / ring slot: one pair
    using SM = Euler2Smem<T, MS, MODE, LAND>;
    constexpr bool H1 = MODE == MODE_HEUN1, H2 = MODE == MODE_HEUN2;
    constexpr int DIST_ = euler_dist(MODE), PF_ = euler_pf(MODE);
    constexpr bool CLOSE = !H1;
    constexpr bool RC = heun_recompute<T>() && (H1 || H2);   // Heun "recompute" protocol (stage_kernel.cuh: heun_recompute)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nz = A.nz;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    using Met = std::conditional_t<SM::CMET, MetricsC<T>, Metrics<T, MS>>;
    Met met;
    if constexpr (SM::CMET) met.A = &A;
    else {
        T* sm = reinterpret_cast<T*>(smem_raw);
        for (int q = 0; q < met_rows(LAND); ++q)
            for (int i = threadIdx.x; i < nz + 3; i += B) sm[q * MS + i] = A.metrics[q * MET_STRIDE + i];
        __syncthreads();
        met.base = smem_base;
    }

    const int64_t c = 2 * ((int64_t)blockIdx.x * B + threadIdx.x);   // columns c (lane x) and c + 1 (lane y)
    if (c >= A.ncol) return;
    const bool vy = c + 1 < A.ncol;        // ragged last pair: lane y computes on the padding of the rows and stores nothing
    const int64_t c1 = vy ? c + 1 : c;     // column the per-column (scalar) reads of lane y use
    const int64_t ld = A.ld;
    const DevParams<T>& p = A.p;
    const T dt = A.dt;

    const uint32_t strip0 = smem_base + (uint32_t)(SM::METRICS * (int)sizeof(T) + threadIdx.x * ES);
    uint32_t kf_cur = strip0 + B * ES, kf_prv = strip0;
    const uint32_t ring0 = smem_base + (uint32_t)(SM::METRICS * (int)sizeof(T) + (SM::STRIP + threadIdx.x) * ES);
    auto ld2 = [](uint32_t a) { return lds2(a, (T*)nullptr); };
    auto rd = [&](int f) { return ld2(strip0 + (uint32_t)(f * B * ES)); };
    auto wr = [&](int f, F2 v) { sts2(strip0 + (uint32_t)(f * B * ES), v); };
    auto ringU = [&](int k) { return ring0 + (uint32_t)((k & (EULER_RD - 1)) * B * ES); };
    auto ringS = [&](int k) { return ring0 + (uint32_t)(((k & (EULER_RD - 1)) + EULER_RD) * B * ES); };
    const F2 zero2 = bc2<T>(T(0));
    sts2(kf_cur, zero2); sts2(kf_prv, zero2); wr(EF_QH, zero2); wr(EF_G, zero2); wr(EF_KC, zero2);

    auto bc_input = [&](int slot) -> F2 {
        const int kind = A.bc[slot].kind;
        if (kind == TRM_BC_DEFAULT) return zero2;
        const T t = kind == TRM_BC_FLUX ? A.t_b : A.t_x;
        const int which = kind == TRM_BC_FLUX ? 1 : 0;
        const InputDesc<T>& s = A.in[A.bc[slot].input];
        return F2{eval_input(s, c, t, which), eval_input(s, c1, t, which)};
    };
    const F2 wtx = RICH ? ldg2(A.xWt + c) : zero2;
    const bool bct_pre = A.bct_pre != 0;
    // per-column surface temperature vector (device memory, or mapped host memory bound with trm_bind_host_io): fetched now,
    // with the first layer, into a strip slot of its own and consumed ~nz iterations later when the halo above the surface
    // is formed -- a read of host memory needs that lead (fetched DIST layers ahead it cost 13.4 instead of 3.4 ms per
    // 10 M-column step). A mapped host ring is not padded: one scalar copy per column.
    const uint32_t bct_slot = strip0 + (uint32_t)(SM::BCT_SLOT * B * ES);
    if (bct_pre) {
        const T* a = A.in[A.bc[TRM_BC_TEMPERATURE_TOP].input].a;
        cp_async<(int)sizeof(T)>(bct_slot, a + c);
        cp_async<(int)sizeof(T)>(bct_slot + (uint32_t)sizeof(T), a + c1);
    }

    uint32_t oin = (uint32_t)c;
    const uint32_t xring0 = ring0 + (uint32_t)(SM::RING * ES);
    uint32_t oext = (uint32_t)c;
    auto k1U_slot = [&](int k) { return xring0 + (uint32_t)((k & (EULER_RD - 1)) * B * ES); };
    auto k1S_slot = [&](int k) { return xring0 + (uint32_t)(((k & (EULER_RD - 1)) + EULER_RD) * B * ES); };
    auto prefetch = [&](int k, bool always = false) {
        if (always || k <= nz) {
            cp_async<ES>(ringU(k), A.xU + oin);
            cp_async<ES>(ringS(k), A.xS + oin);
            // (recompute protocol, stage 2) k1 of layer k lives next to U / sat of layer k, from its prefetch until its update
            if (H2 && RC) { cp_async<ES>(k1U_slot(k), A.k1U + oin); if (RICH) cp_async<ES>(k1S_slot(k), A.k1S + oin); }
            oin += (uint32_t)ld;
        }
        if (H2 && !RC) {
            const int kk = k - 2;
            if (kk >= 1 && kk <= nz) {
                const uint32_t dst = xring0 + (uint32_t)((kk & (PF_ - 1)) * B * ES);
                cp_async<ES>(dst, A.k1U + oext);
                cp_async<ES>(dst + 2 * PF_ * B * ES, A.bU + oext);
                if (RICH) { cp_async<ES>(dst + PF_ * B * ES, A.k1S + oext); cp_async<ES>(dst + 3 * PF_ * B * ES, A.bS + oext); }
                oext += (uint32_t)ld;
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int k = 1; k <= DIST_; ++k) prefetch(k);

    F2 carry = zero2;
    int negx = 0, negy = 0;     // ORed sign words of the updated saturations (a set sign bit -> slow path of that column)
    int idxx = 0, idxy = 0;     // lowest unsaturated layer (compute_water_table!), 0 = not found yet
    F2 wt_new = zero2;
    F2 Sx_new = zero2;
    if (RICH && !H1) Sx_new = ldg2(A.bSx + c);   // (zero tendency, soil_hydrology.jl:260-267)
    uint32_t oout = (uint32_t)c;
    if (LAND && has_veg(A)) wr(EF_BETA, zero2);

    // Flux boundary conditions of the time-n state (compute_z_bcs!, abstract_timestepper.jl:69) on the tendencies of the top /
    // bottom layer. LandModel: ground heat flux and infiltration left by surface_kernel (land_model.jl:56-62)
    auto apply_top_flux = [&](F2& tU, F2& tS) {
        const T dz = met.dzc(nz);
        if (LAND) {
            const F2 G_top = ldg2(A.G + c);
            tU = F2{tU.x - G_top.x / dz, tU.y - G_top.y / dz};
            if (RICH) { const F2 infil_top = ldg2(A.infil + c); tS = F2{tS.x - (-infil_top.x) / dz, tS.y - (-infil_top.y) / dz}; }
        } else {
            if (A.bc[TRM_BC_ENERGY_TOP].kind == TRM_BC_FLUX) { const F2 f = bc_input(TRM_BC_ENERGY_TOP); tU = F2{tU.x - f.x / dz, tU.y - f.y / dz}; }
            if (RICH && A.bc[TRM_BC_SATURATION_TOP].kind == TRM_BC_FLUX) { const F2 f = bc_input(TRM_BC_SATURATION_TOP); tS = F2{tS.x - f.x / dz, tS.y - f.y / dz}; }
        }
    };
    auto apply_bottom_flux = [&](F2& tU, F2& tS) {
        const T dz = met.dzc(1);
        if (A.bc[TRM_BC_ENERGY_BOTTOM].kind == TRM_BC_FLUX) { const F2 f = bc_input(TRM_BC_ENERGY_BOTTOM); tU = F2{tU.x + f.x / dz, tU.y + f.y / dz}; }
        if (RICH && A.bc[TRM_BC_SATURATION_BOTTOM].kind == TRM_BC_FLUX) { const F2 f = bc_input(TRM_BC_SATURATION_BOTTOM); tS = F2{tS.x + f.x / dz, tS.y + f.y / dz}; }
    };
    // recompute protocol, stage 2: upward-sweep carry of the stage state being rebuilt ; columns whose stage state was stored
    F2 carry1 = zero2;
    const bool flagx = (H2 && RC && RICH) ? (A.hflag_in[c] != T(0)) : false;
    const bool flagy = (H2 && RC && RICH) ? (A.hflag_in[c1] != T(0)) : false;
    uint32_t oent = (uint32_t)c;       // element offset of the entering layer (stored stage state of a flagged column)

    F2 Ur_nx = zero2, sr_nx = zero2;   // raw U / sat of the layer that enters next
    // One pipeline iteration (see euler_kernel). The flux slots hold the NEGATED fluxes: EF_QH = kappa_f dT/dz,
    // EF_QD = K* dpsi/dz, so that the tendencies come out of one packed FFMA each without sign flips.
    auto iterate = [&](const int m, auto pos_tag) {
        constexpr int POS = decltype(pos_tag)::value;   // position in the column, see IterPos (euler_kernel.cuh)
        constexpr bool GEN = POS == IP_GEN;
        const bool m_is_1 = GEN ? m == 1 : POS == IP_M1;
        const bool m_is_nz = GEN ? m == nz : POS == IP_NZ;
        const bool enters = GEN ? m <= nz : POS <= IP_NZ;
        const bool is_halo = GEN ? m == nz + 1 : POS == IP_HALO;
        const bool has_face = GEN ? m <= nz + 1 : POS != IP_LAST;
        const bool has_darcy = GEN ? m >= 2 : POS != IP_M1;
        const bool updates = GEN ? m >= 3 : POS >= IP_M3;
        const bool j_is_1 = GEN ? m == 3 : POS == IP_M3;
        const bool j_is_nz = GEN ? m == nz + 2 : POS == IP_LAST;
        if (POS == IP_INNER) prefetch(m + DIST_, true);
        else if (POS >= IP_INNER_NP && !(H2 && !RC)) {   // no layer left to prefetch (stored protocol: Heun stage 2 still fetches k1 / the base state of layer m+DIST-2)
            cp_async_commit();
        }
        else prefetch(m + DIST_, false);
        F2 Tn, Pn = zero2, kapn, Kfn = zero2;
        const F2 Kf1 = RICH ? ld2(kf_prv) : zero2;
        if (enters) {
            // (the raw values of the entering layer were read from the ring at the end of the previous iteration: the
            //  shared-memory latency overlaps the update / closure arithmetic of that iteration instead of heading this one)
            F2 Ur = Ur_nx;
            F2 sr = sr_nx;
            if (H2 && RC) {
                // stage state of layer m: what stage 1 computed from the same base state and k1, and did not store
                // (explicit_step! + upward sweep of adjust_saturation_profile! of the stage copy, heun.jl:45-49) -- the same
                // operations in the same order as the update section of stage 1
                F2 t1U = ld2(k1U_slot(m)), t1S = RICH ? ld2(k1S_slot(m)) : zero2;
                if (m_is_nz) apply_top_flux(t1U, t1S);
                if (m_is_1) apply_bottom_flux(t1U, t1S);
                F2 Us = fma2<T>(t1U, dt, Ur), ss = sr;
                if (RICH) {
                    ss = fma2<T>(t1S, dt, ss) + carry1;
                    if (!m_is_nz) {
                        const F2 e = pos2(ss + T(-1));
                        ss = ss - e;
                        carry1 = e * (met.dzc(m) * met.rdzc(m + 1));
                    } else {
                        ss = ss - pos2(ss + T(-1));   // top excess of the stage copy (its surface excess water is not used)
                    }
                    if (flagx) { Us.x = A.sU[oent]; ss.x = A.sS[oent]; }
                    if (flagy) { Us.y = A.sU[oent + 1]; ss.y = A.sS[oent + 1]; }
                }
                oent += (uint32_t)ld;
                Ur = Us; sr = ss;
            }
            F2 ln, wi;
            energy_to_temperature2(p, Ur, sr, Tn, ln, wi);
            if (RICH) Pn = pressure_head2<SOIL, T>(p, sr, wtx, met.zC(m), met.psiz(m));
            kapn = thermal_conductivity2(p, wi, ln);
            if (RICH) {
                const F2 Kcn = cell_conductivity2s<SOIL, T>(p, sr, ln);
                Kfn = (m_is_1 || m_is_nz) ? Kcn : min2(Kcn, rd(EF_KC));   // Kf[1] = Kc[1], Kf[Nz] = Kc[Nz]
                wr(EF_KC, Kcn);
            }
        } else if (is_halo) {   // halo above the surface
            if (bct_pre) cp_async_wait<DIST_>();   // (its group is the first one: long complete unless the column has fewer layers than the prefetch distance)
            Tn = halo_value2(A.bc[TRM_BC_TEMPERATURE_TOP].kind, rd(EF_T), bct_pre ? ld2(bct_slot) : bc_input(TRM_BC_TEMPERATURE_TOP), met.dzf(nz + 1), true);
            const bool copy = RICH || p.sat_halo == TRM_HALO_COPY;
            kapn = copy ? rd(EF_KAP) : bc2<T>(thermal_conductivity_fast(p, T(0), T(1)));
            if (RICH) Pn = halo_value2(A.bc[TRM_BC_PRESSURE_TOP].kind, rd(EF_P), bc_input(TRM_BC_PRESSURE_TOP), met.dzf(nz + 1), true);
            Kfn = Kf1;
        } else {
            Tn = zero2; kapn = zero2;
        }
        F2 Tp, kapp, Pp = zero2;
        if (m_is_1) {
            Tp = halo_value2(A.bc[TRM_BC_TEMPERATURE_BOTTOM].kind, Tn, bc_input(TRM_BC_TEMPERATURE_BOTTOM), met.dzf(1), false);
            const bool copy = RICH || p.sat_halo == TRM_HALO_COPY;
            kapp = copy ? kapn : bc2<T>(thermal_conductivity_fast(p, T(0), T(1)));
            if (RICH) Pp = halo_value2(A.bc[TRM_BC_PRESSURE_BOTTOM].kind, Pn, bc_input(TRM_BC_PRESSURE_BOTTOM), met.dzf(1), false);
        } else {
            Tp = rd(EF_T); kapp = rd(EF_KAP);
            if (RICH) Pp = rd(EF_P);
        }
        // ---- (negated) heat flux and head gradient at face m ----
        F2 nqh = zero2, gn = zero2;
        if (has_face) {
            const T rz = met.rdzf(m);
            nqh = ((kapn + kapp) * (T(0.5) * rz)) * (Tn - Tp);
            if (RICH) gn = (Pn - Pp) * rz;
        }
        const F2 dnqh = nqh - rd(EF_QH);
        // ---- (negated) Darcy flux at face m-1 ----
        F2 nqd = zero2;
        if (RICH && has_darcy) {
            const F2 g = rd(EF_G);
            const F2 Kf2 = ld2(kf_cur);
            const F2 Kk = min2(Kf1, sel<T>(B2{g.x < T(0), g.y < T(0)}, Kf2, Kfn));
            nqd = Kk * g;
        }

        if (updates) {
            const int j = m - 2;
            const uint32_t o = oout;
            oout += (uint32_t)ld;
            const T rzc = met.rdzc(j);
            F2 tU = rd(EF_DQH) * rzc;
            F2 tS = zero2;
            if (RICH) tS = fma2<T>(nqd - rd(EF_QD), rzc, p.vwcf) * p.rpor;
            F2 Ub, sb;
            if (H2 && RC) {   // average_tendencies! (heun.jl:27-35) ; k1 and the base state sit in the 8-deep rings
                tU = (ld2(k1U_slot(j)) + tU) * T(0.5);
                Ub = ld2(ringU(j)); sb = ld2(ringS(j));
                if (RICH) tS = (ld2(k1S_slot(j)) + tS) * T(0.5);
            } else if (H2) {   // average_tendencies! (heun.jl:27-35) ; the base is the state at time n
                const uint32_t x = xring0 + (uint32_t)((j & (PF_ - 1)) * B * ES);
                tU = (ld2(x) + tU) * T(0.5);
                Ub = ld2(x + 2 * PF_ * B * ES);
                if (RICH) { tS = (ld2(x + PF_ * B * ES) + tS) * T(0.5); sb = ld2(x + 3 * PF_ * B * ES); }
                else sb = ld2(ringS(j));
            } else {
                if (H1) { stg2(A.oTU + o, tU); if (RICH) stg2(A.oTS + o, tS); }   // k1, before the Flux BCs
                Ub = ld2(ringU(j)); sb = ld2(ringS(j));
            }
            if (j_is_nz) apply_top_flux(tU, tS);
            if (j_is_1) apply_bottom_flux(tU, tS);
            // ---- explicit step ----
            const F2 Un = fma2<T>(tU, dt, Ub);
            F2 sn = sb;
            if (RICH) {
                sn = fma2<T>(tS, dt, sn) + carry;
                if (!j_is_nz) {   // upward sweep of adjust_saturation_profile! (soil_hydrology.jl:192-199)
                    const F2 e = pos2(sn + T(-1));
                    sn = sn - e;
                    carry = e * (met.dzc(j) * met.rdzc(j + 1));
                }
                negx |= sign_word(sn.x); negy |= sign_word(sn.y);
            }
            // A column that has gone negative keeps storing its raw values (no surface excess); the closure stores of
            // such a column are overwritten by its slow path after the sweep.
            if (RICH) {
                if (j_is_nz) {                         // top excess -> surface_excess_water (:210-214)
                    F2 e = pos2(sn + T(-1));
                    e = F2{negx < 0 ? T(0) : e.x, negy < 0 ? T(0) : e.y};
                    sn = sn - e;
                    Sx_new = fma2<T>(e, met.dzc(nz), Sx_new);
                }
                // (recompute protocol, Heun stage 1: the stage state is not stored -- except its top layer, which the vegetation
                //  block of stage 2 reads through surface_kernel -- and a column that goes negative rebuilds it in its slow path)
                if (!(H1 && RC) || j_is_nz) stg2(A.yS + o, sn);
                if (idxx == 0 && below_one(sn.x)) { idxx = j; wt_new.x = met.zF(j); }   // compute_water_table!, kernel_utils.jl:7-16
                if (idxy == 0 && below_one(sn.y)) { idxy = j; wt_new.y = met.zF(j); }
            }
            if (!(H1 && RC) || j_is_nz) stg2(A.yU + o, Un);
            if (CLOSE || (LAND && has_veg(A))) {
                F2 Tc, lc, wi;
                energy_to_temperature2(p, Un, sn, Tc, lc, wi);
                if (LAND && has_veg(A)) {   // soil moisture limiting factor of the NEW state (plant_available_water.jl:31-35)
                    F2 x = ((wi * lc) + (-A.vp.th_wp)) * A.vp.r_paw_span;
                    x = F2{M<T, true>::mn(M<T, true>::mx(x.x, T(0)), T(1)), M<T, true>::mn(M<T, true>::mx(x.y, T(0)), T(1))};
                    wr(EF_BETA, fma2<T>(x, met.root(j), rd(EF_BETA)));
                }
                if (CLOSE) {
                    stg2(A.yT + o, Tc); stg2(A.yL + o, lc);
                    if (j_is_nz && A.hio_out) { A.hio_out[c] = Tc.x; if (vy) A.hio_out[c + 1] = Tc.y; }
                    // layers below the water table wait for it (written after the sweep); in a pair with only one column
                    // still below its water table that column's value is overwritten there
                    if (RICH && (idxx | idxy) != 0) stg2(A.yP + o, pressure_head2<SOIL, T>(p, sn, wt_new, met.zC(j), met.psiz(j)));
                }
            }
        }
        wr(EF_T, Tn); wr(EF_KAP, kapn); wr(EF_QH, nqh); wr(EF_DQH, dnqh);
        if (RICH) { wr(EF_P, Pn); sts2(kf_cur, Kfn); wr(EF_G, gn); wr(EF_QD, nqd); }
        const uint32_t t = kf_cur; kf_cur = kf_prv; kf_prv = t;
        if (GEN ? m + 1 <= nz : POS < IP_NZ) {   // layer m+1 enters next: all but the DIST-1 most recent groups have landed
            cp_async_wait<DIST_ - 1>();
            Ur_nx = ld2(ringU(m + 1));
            sr_nx = ld2(ringS(m + 1));
        }
    };
    cp_async_wait<DIST_ - 1>();   // layer 1
    Ur_nx = ld2(ringU(1));
    sr_nx = ld2(ringS(1));
    if (TRM_EULER_SPEC && nz >= 4) {
        iterate(1, IterTag<IP_M1>{}); iterate(2, IterTag<IP_M2>{}); iterate(3, IterTag<IP_M3>{});
        int m = 4;
#pragma unroll 1
        for (; m <= nz - DIST_; ++m) iterate(m, IterTag<IP_INNER>{});
#pragma unroll 1
        for (; m < nz; ++m) iterate(m, IterTag<IP_INNER_NP>{});
        iterate(nz, IterTag<IP_NZ>{}); iterate(nz + 1, IterTag<IP_HALO>{}); iterate(nz + 2, IterTag<IP_LAST>{});
    } else {
        int m = 1;
#pragma unroll 1
        while (m <= nz + 2) {
            if (m >= 4 && m <= nz - DIST_) {
#pragma unroll 1
                do { iterate(m, IterTag<IP_INNER>{}); ++m; } while (m <= nz - DIST_);
            } else {
                iterate(m, IterTag<IP_GEN>{});
                ++m;
            }
        }
    }
    const bool nx = RICH && negx < 0, ny = RICH && negy < 0;
    if (LAND && has_veg(A)) {
        const F2 b = rd(EF_BETA);
        if (!nx) A.ybeta[c] = b.x;
        if (vy && !ny) A.ybeta[c + 1] = b.y;
    }
    if (!RICH) return;

    // ---- after the sweep: water table, surface excess water, pressure head of the saturated zone (per column) ----
    if (idxx == 0) { idxx = nz + 1; wt_new.x = met.zF(nz + 1); }
    if (idxy == 0) { idxy = nz + 1; wt_new.y = met.zF(nz + 1); }
    if (!nx) A.yWt[c] = wt_new.x;
    if (vy && !ny) A.yWt[c + 1] = wt_new.y;
    if (!H1) {
        if (!nx) A.ySx[c] = Sx_new.x;
        if (vy && !ny) A.ySx[c + 1] = Sx_new.y;
        // psi_m(sat >= 1) is a constant: (wt - zC) + psat + (zC - zref) is one value for the whole saturated zone
        const T psat = swrc_inverse<T, true, SOIL == SOIL2_VG2>(p, p.por, p.por);
        const T zref = met.zF(nz + 1);
        const T px = (wt_new.x - zref) + psat, py = (wt_new.y - zref) + psat;
        const int kx = nx ? 0 : (idxx <= nz ? idxx : nz + 1), ky = (ny || !vy) ? 0 : (idxy <= nz ? idxy : nz + 1);   // layers 1 .. k-1 are rewritten
        const int kmin = kx < ky ? kx : ky, kmax = kx < ky ? ky : kx;
        int64_t o = c;
        int k = 1;
#pragma unroll 1
        for (; k < kmin; ++k, o += ld) stg2(A.yP + o, F2{px, py});
        T* const rest = A.yP + (kx < ky ? 1 : 0);
        const T pr = kx < ky ? py : px;
#pragma unroll 1
        for (; k < kmax; ++k, o += ld) rest[o] = pr;
    }
    if (H1 && RC) {
        // stage 2 rebuilds the stage state of a column from the base state and k1 -- unless a layer went negative: the stage
        // copy then needs the downward sweep, is formed here in full, stored, and the column is flagged (euler_kernel.cuh)
        if (nx) heun1_slow_column<T, true, Met, LAND>(A, met, c); else A.hflag_out[c] = T(0);
        if (vy) { if (ny) heun1_slow_column<T, true, Met, LAND>(A, met, c + 1); else A.hflag_out[c + 1] = T(0); }
        return;
    }
    if (nx) euler2_slow_column<T, Met, MODE, LAND, SOIL == SOIL2_VG2>(A, met, c, Sx_new.x);
    if (ny && vy) euler2_slow_column<T, Met, MODE, LAND, SOIL == SOIL2_VG2>(A, met, c + 1, Sx_new.y);
}

}  // namespace trm
