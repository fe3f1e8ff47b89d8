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

struct alignas(8) F2 { float x, y; };
struct B2 { bool x, y; };

#define TRM_U64(v) reinterpret_cast<uint64_t&>(v)
#define TRM_CU64(v) reinterpret_cast<const uint64_t&>(v)
__device__ __forceinline__ F2 bc2(float s) { return F2{s, s}; }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { F2 c; asm("add.f32x2 %0, %1, %2;" : "=l"(TRM_U64(c)) : "l"(TRM_CU64(a)), "l"(TRM_CU64(b))); return c; }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { F2 c; asm("mul.f32x2 %0, %1, %2;" : "=l"(TRM_U64(c)) : "l"(TRM_CU64(a)), "l"(TRM_CU64(b))); return c; }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(TRM_U64(d)) : "l"(TRM_CU64(a)), "l"(TRM_CU64(b)), "l"(TRM_CU64(c))); return d; }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return fma2(b, bc2(-1.0f), a); }   // a - b as one FFMA2 (exact product)
__device__ __forceinline__ F2 operator+(F2 a, float s) { return a + bc2(s); }
__device__ __forceinline__ F2 operator*(F2 a, float s) { return a * bc2(s); }
__device__ __forceinline__ F2 fma2(F2 a, float b, F2 c) { return fma2(a, bc2(b), c); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, float c) { return fma2(a, b, bc2(c)); }
__device__ __forceinline__ F2 fma2(F2 a, float b, float c) { return fma2(a, bc2(b), bc2(c)); }
__device__ __forceinline__ F2 sel(B2 m, F2 a, F2 b) { return F2{m.x ? a.x : b.x, m.y ? a.y : b.y}; }
__device__ __forceinline__ F2 sel(B2 m, F2 a, float b) { return F2{m.x ? a.x : b, m.y ? a.y : b}; }
__device__ __forceinline__ F2 sel(B2 m, float a, F2 b) { return F2{m.x ? a : b.x, m.y ? a : b.y}; }
__device__ __forceinline__ F2 sel(B2 m, float a, float b) { return F2{m.x ? a : b, m.y ? a : b}; }
__device__ __forceinline__ F2 min2(F2 a, F2 b) { return F2{fminf(a.x, b.x), fminf(a.y, b.y)}; }
__device__ __forceinline__ F2 pos2(F2 a) { return F2{fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f)}; }
__device__ __forceinline__ F2 abs2(F2 a) { return F2{fabsf(a.x), fabsf(a.y)}; }
__device__ __forceinline__ F2 rcp2(F2 a) { return F2{M<float, true>::rcp(a.x), M<float, true>::rcp(a.y)}; }
__device__ __forceinline__ F2 rsqrt2(F2 a) { return F2{M<float, true>::rsqrt_(a.x), M<float, true>::rsqrt_(a.y)}; }   // clamped seed, see M<float, true>
__device__ __forceinline__ B2 operator&&(B2 a, B2 b) { return B2{a.x && b.x, a.y && b.y}; }
__device__ __forceinline__ B2 operator!(B2 a) { return B2{!a.x, !a.y}; }
__device__ __forceinline__ bool any(B2 a) { return a.x || a.y; }

// 8-byte shared-memory accesses through explicit shared addresses (same ordering argument as sts / ldsv in euler_kernel.cuh)
__device__ __forceinline__ void sts2(uint32_t a, F2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(a), "f"(v.x), "f"(v.y)); }
__device__ __forceinline__ F2 lds2(uint32_t a) { F2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void stg2(float* p, F2 v) {
#ifdef TRM_NO_STCS
    *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y);
#else
    __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
#endif
}
__device__ __forceinline__ F2 ldg2(const float* p) { const float2 v = *reinterpret_cast<const float2*>(p); return F2{v.x, v.y}; }

// ---- per-cell physics on pairs (FAST forms of column_physics.cuh) --------------------------------------------------
// energy_to_temperature + liquid_water_fraction, soil_energy_closures.jl:99-159 ; wi = sat * por is handed back for the
// conductivity
__device__ __forceinline__ void energy_to_temperature2(const DevParams<float>& p, F2 U, F2 sat, F2& T, F2& liq, F2& wi) {
    wi = sat * p.por;
    const F2 Lt = wi * p.L;
    const B2 thawed{U.x >= 0.0f, U.y >= 0.0f}, frozen{U.x < -Lt.x, U.y < -Lt.y};
    const F2 UL = U + Lt;
    const F2 num{thawed.x ? U.x : (frozen.x ? UL.x : 0.0f), thawed.y ? U.y : (frozen.y ? UL.y : 0.0f)};
    liq = sel(thawed, 1.0f, 0.0f);
    if (!thawed.x && !frozen.x) liq.x = partial_liquid_fraction_cold(U.x, Lt.x);   // phase change zone: rare, out of line
    if (!thawed.y && !frozen.y) liq.y = partial_liquid_fraction_cold(U.y, Lt.y);
    const F2 C = fma2(wi, fma2(liq, p.hc_wi, p.hc_ia), p.hc_base);
    T = num * rcp2(C);
}
// InverseQuadratic bulk thermal conductivity, soil_thermal_properties.jl:90-108 (regrouped constituent sum)
__device__ __forceinline__ F2 thermal_conductivity2(const DevParams<float>& p, F2 wi, F2 liq) {
    const F2 s = fma2(wi, fma2(liq, p.sqk_wi, p.sqk_ia), p.sqk_base);
    return s * s;
}
// hydraulic conductivity at the cell centres, van Genuchten n = 2 (soil_hydraulic_properties.jl:170-221, exponent n/(n+1) as
// coded): K = K_sat I_ice sqrt(x) (1 - sqrt(1 - x^(2/3)))^2 with the exact end members x = 0 -> 0, x = 1 -> K_sat I_ice
__device__ __forceinline__ F2 cell_conductivity2(const DevParams<float>& p, F2 sat, F2 liq) {
    const F2 x = sat * liq;
    F2 KI = bc2(p.Ksat);
    if (liq.x != 1.0f && x.x != 0.0f) KI.x = p.Ksat * ice_impedance_cold(p.Omega, liq.x);
    if (liq.y != 1.0f && x.y != 0.0f) KI.y = p.Ksat * ice_impedance_cold(p.Omega, liq.y);
    const B2 zero{x.x == 0.0f, x.y == 0.0f}, one{x.x == 1.0f, x.y == 1.0f};
    F2 K = KI;
    if (!((zero.x || one.x) && (zero.y || one.y))) {   // (whole saturated / frozen / dry zones skip the roots)
        // x^(2/3) and x^(1/2) from one r ~ x^(-1/6) (lg2 / ex2 seed, one Newton step), as M<float, true>::roots ; the
        // argument is clamped away from 0 (the result of a lane with x <= 1e-30 underflows to the exact end member 0)
        F2 r;
        {
            float lx, ly;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lx) : "f"(fmaxf(x.x, 1.0e-30f)));
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(ly) : "f"(fmaxf(x.y, 1.0e-30f)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(lx * (-1.0f / 6.0f)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(ly * (-1.0f / 6.0f)));
        }
        const F2 r3 = (r * r) * r;
        const F2 ne = fma2(x * r3, r3, -1.0f);            // -(1 - x r^6)
        r = fma2(r, ne * (-1.0f / 6.0f), r);
        const F2 x23 = x * (r * r), x12 = x23 * r;
        const F2 d = abs2(bc2(1.0f) - x23);            // |.|: guards the root against a -1 ulp residue as x -> 1
        const F2 a = fma2(d, rsqrt2(d) * -1.0f, 1.0f);  // 1 - sqrt(d), sqrt(d) = d rsqrt(d) with the clamped seed (d = 0 -> 0)
        K = abs2((KI * x12) * (a * a));
    }
    K = sel(one, KI, K);
    return sel(zero, 0.0f, K);
}
// UnsatKLinear (soil_hydraulic_properties.jl:170-198): K = K_sat water / (water + ice + air), the three fractions formed as
// in volumetric_fractions (soil_volume.jl:52-67)
__device__ __forceinline__ F2 cell_conductivity2_linear(const DevParams<float>& p, F2 sat, F2 liq) {
    const F2 wi = sat * p.por;
    const F2 water = wi * liq;
    const F2 ice = wi * (bc2(1.0f) - liq);
    const F2 air = (bc2(1.0f) - sat) * p.por;
    return (water * p.Ksat) * rcp2((water + ice) + air);
}
enum Soil2 { SOIL2_VG2 = 0, SOIL2_BC_LINEAR = 1 };
template <int SOIL>
__device__ __forceinline__ F2 cell_conductivity2s(const DevParams<float>& p, F2 sat, F2 liq) {
    return SOIL == SOIL2_VG2 ? cell_conductivity2(p, sat, liq) : cell_conductivity2_linear(p, sat, liq);
}
// Brooks-Corey matric head with an integer exponent k = 1 / lambda: psi_m = -psi_s se^-k below saturation, -psi_s at
// saturation (FreezeCurves BrooksCorey); se^k by repeated multiplication (k is launch-uniform, <= 8)
__device__ __forceinline__ F2 brooks_corey_psim2(const DevParams<float>& p, F2 theta) {
    const F2 se = fma2(theta, p.r_thspan, p.se_off);
    F2 pw = se;
#pragma unroll 1
    for (int i = 1; i < p.bc_k; ++i) pw = pw * se;
    const F2 r = rcp2(pw) * -p.bc_psis;          // se = 0 (dry layer): rcp(0) = +Inf -> -Inf as in the reference
    return F2{theta.x < p.por ? r.x : -p.bc_psis, theta.y < p.por ? r.y : -p.bc_psis};
}
// total pressure head, saturation_to_pressure! (soil_hydraulic_closures.jl:102-129) with the van Genuchten n = 2 retention
// curve: psi_m = -(1/alpha) sqrt(se^-2 - 1) = -(1/alpha) a / sqrt(a t), a = 1 - se^2, t = se^2 (see swrc_inverse)
template <int SOIL = SOIL2_VG2>
__device__ __forceinline__ F2 pressure_head2(const DevParams<float>& p, F2 sat, F2 wt, float zc, float psiz) {
    const F2 theta = sat * p.por;
    if (SOIL == SOIL2_BC_LINEAR) return (pos2(wt + (-zc)) + brooks_corey_psim2(p, theta)) + psiz;
    const F2 se = fma2(theta, p.r_thspan, p.se_off);
    const F2 t = se * se;
    const F2 a = abs2(fma2(se * -1.0f, se, 1.0f));
    F2 r = (a * p.neg_inv_alpha) * rsqrt2(a * t);
    const float ninf = -Lim<float>::inf();
    r = F2{t.x == 0.0f ? ninf : r.x, t.y == 0.0f ? ninf : r.y};
    const F2 psim{theta.x < p.por ? r.x : 0.0f, theta.y < p.por ? r.y : 0.0f};
    const F2 psih = pos2(wt + (-zc));
    return (psih + psim) + psiz;
}
// Oceananigans halo fill on pairs (halo_value, stage_kernel.cuh); edge iterations only
__device__ __forceinline__ F2 halo_value2(int kind, F2 edge, F2 v, float D, bool top) {
    return F2{halo_value(kind, edge.x, v.x, D, top), halo_value(kind, edge.y, v.y, D, top)};
}

template <int MS, int MODE>
struct Euler2Smem {
    static constexpr int PF_ = euler_pf(MODE);
    static constexpr int METRICS = MET_COUNT * MS;                                  // floats
    static constexpr int STRIP = EF_COUNT * TRM_EULER_BLOCK;                        // 8-byte slots
    static constexpr int RING = 2 * EULER_RD * TRM_EULER_BLOCK;                     // U, sat
    static constexpr int XRING = (MODE == MODE_HEUN2 ? 4 : 0) * PF_ * TRM_EULER_BLOCK;   // k1U, k1S, bU, bS
    static constexpr size_t BYTES = 4 * (size_t)METRICS + 8 * (size_t)(STRIP + RING + XRING);
};
static_assert((MET_COUNT * EULER_MS_SMALL * 4) % 8 == 0 && (MET_COUNT * MET_STRIDE * 4) % 8 == 0, "the strips must be 8-byte aligned");

#ifndef TRM_EULER2_BLOCKS
#define TRM_EULER2_BLOCKS 6
#endif
template <int PHYS, int MODE>
constexpr int euler2_min_blocks() { return MODE == MODE_HEUN2 ? 4 : (phys_land(PHYS) ? 5 : TRM_EULER2_BLOCKS); }

// slow path of one column: a layer went negative. Downward sweep (soil_hydrology.jl:201-216) top -> bottom on the raw
// profile the thread has just stored, then water table and closures bottom -> top (same code as the scalar kernel).
template <int MS, int MODE, bool LAND, bool VG2>
__device__ __noinline__ void euler2_slow_column(const StageArgs<float>& A, Metrics<float, MS> met, int64_t c, float Sx_new) {
    constexpr bool H1 = MODE == MODE_HEUN1;
    using NF = float;
    const DevParams<float>& p = A.p;
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

template <int PHYS, int MS, int MODE, int SOIL = SOIL2_VG2>
__global__ void __launch_bounds__(TRM_EULER_BLOCK, (euler2_min_blocks<PHYS, MODE>())) euler2_kernel(const __grid_constant__ StageArgs<float> A) {
    constexpr bool RICH = phys_richards(PHYS);
    constexpr bool LAND = phys_land(PHYS);
    constexpr int B = TRM_EULER_BLOCK;
    constexpr int ES = 8;   // bytes per strip / ring slot: one pair
    using SM = Euler2Smem<MS, MODE>;
    constexpr bool H1 = MODE == MODE_HEUN1, H2 = MODE == MODE_HEUN2;
    constexpr int DIST_ = euler_dist(MODE), PF_ = euler_pf(MODE);
    constexpr bool CLOSE = !H1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nz = A.nz;
    {
        float* sm = reinterpret_cast<float*>(smem_raw);
        for (int q = 0; q < met_rows(LAND); ++q)
            for (int i = threadIdx.x; i < nz + 3; i += B) sm[q * MS + i] = A.metrics[q * MET_STRIDE + i];
    }
    __syncthreads();
    Metrics<float, MS> met;
    met.base = (uint32_t)__cvta_generic_to_shared(smem_raw);

    const int64_t c = 2 * ((int64_t)blockIdx.x * B + threadIdx.x);   // columns c (lane x) and c + 1 (lane y)
    if (c >= A.ncol) return;
    const bool vy = c + 1 < A.ncol;        // ragged last pair: lane y computes on the padding of the rows and stores nothing
    const int64_t c1 = vy ? c + 1 : c;     // column the per-column (scalar) reads of lane y use
    const int64_t ld = A.ld;
    const DevParams<float>& p = A.p;
    const float dt = A.dt;

    const uint32_t strip0 = met.base + (uint32_t)(SM::METRICS * 4 + threadIdx.x * ES);
    uint32_t kf_cur = strip0 + B * ES, kf_prv = strip0;
    const uint32_t ring0 = met.base + (uint32_t)(SM::METRICS * 4 + (SM::STRIP + threadIdx.x) * ES);
    auto rd = [&](int f) { return lds2(strip0 + (uint32_t)(f * B * ES)); };
    auto wr = [&](int f, F2 v) { sts2(strip0 + (uint32_t)(f * B * ES), v); };
    auto ringU = [&](int k) { return ring0 + (uint32_t)((k & (EULER_RD - 1)) * B * ES); };
    auto ringS = [&](int k) { return ring0 + (uint32_t)(((k & (EULER_RD - 1)) + EULER_RD) * B * ES); };
    const F2 zero2 = bc2(0.0f);
    sts2(kf_cur, zero2); sts2(kf_prv, zero2); wr(EF_QH, zero2); wr(EF_G, zero2); wr(EF_KC, zero2);

    auto bc_input = [&](int slot) -> F2 {
        const int kind = A.bc[slot].kind;
        if (kind == TRM_BC_DEFAULT) return zero2;
        const float t = kind == TRM_BC_FLUX ? A.t_b : A.t_x;
        const int which = kind == TRM_BC_FLUX ? 1 : 0;
        const InputDesc<float>& s = A.in[A.bc[slot].input];
        return F2{eval_input(s, c, t, which), eval_input(s, c1, t, which)};
    };
    const F2 wtx = RICH ? ldg2(A.xWt + c) : zero2;
    const bool bct_pre = A.bct_pre != 0;
    if (bct_pre) {   // (a mapped host ring is not padded: one 4-byte copy per column)
        const float* a = A.in[A.bc[TRM_BC_TEMPERATURE_TOP].input].a;
        cp_async<4>(strip0 + (uint32_t)(EF_BCT * B * ES), a + c);
        cp_async<4>(strip0 + (uint32_t)(EF_BCT * B * ES) + 4, a + c1);
    }

    uint32_t oin = (uint32_t)c;
    const uint32_t xring0 = ring0 + (uint32_t)(SM::RING * ES);
    uint32_t oext = (uint32_t)c;
    auto prefetch = [&](int k, bool always = false) {
        if (always || k <= nz) {
            cp_async<ES>(ringU(k), A.xU + oin);
            cp_async<ES>(ringS(k), A.xS + oin);
            oin += (uint32_t)ld;
        }
        if (H2) {
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

    // One pipeline iteration (see euler_kernel). The flux slots hold the NEGATED fluxes: EF_QH = kappa_f dT/dz,
    // EF_QD = K* dpsi/dz, so that the tendencies come out of one packed FFMA each without sign flips.
    auto iterate = [&](const int m, auto inner_tag) {
        constexpr bool inner = decltype(inner_tag)::value;
        prefetch(m + DIST_, inner);
        F2 Tn, Pn = zero2, kapn, Kfn = zero2;
        const F2 Kf1 = RICH ? lds2(kf_prv) : zero2;
        if (inner || m <= nz) {
            cp_async_wait<DIST_>();
            const F2 Ur = lds2(ringU(m));
            const F2 sr = lds2(ringS(m));
            F2 ln, wi;
            energy_to_temperature2(p, Ur, sr, Tn, ln, wi);
            if (RICH) Pn = pressure_head2<SOIL>(p, sr, wtx, met.zC(m), met.psiz(m));
            kapn = thermal_conductivity2(p, wi, ln);
            if (RICH) {
                const F2 Kcn = cell_conductivity2s<SOIL>(p, sr, ln);
                Kfn = (!inner && (m == 1 || m == nz)) ? Kcn : min2(Kcn, rd(EF_KC));   // Kf[1] = Kc[1], Kf[Nz] = Kc[Nz]
                wr(EF_KC, Kcn);
            }
        } else if (!inner && m == nz + 1) {   // halo above the surface
            Tn = halo_value2(A.bc[TRM_BC_TEMPERATURE_TOP].kind, rd(EF_T), bct_pre ? rd(EF_BCT) : bc_input(TRM_BC_TEMPERATURE_TOP), met.dzf(nz + 1), true);
            const bool copy = RICH || p.sat_halo == TRM_HALO_COPY;
            kapn = copy ? rd(EF_KAP) : bc2(thermal_conductivity_fast(p, 0.0f, 1.0f));
            if (RICH) Pn = halo_value2(A.bc[TRM_BC_PRESSURE_TOP].kind, rd(EF_P), bc_input(TRM_BC_PRESSURE_TOP), met.dzf(nz + 1), true);
            Kfn = Kf1;
        } else {
            Tn = zero2; kapn = zero2;
        }
        F2 Tp, kapp, Pp = zero2;
        if (!inner && m == 1) {
            Tp = halo_value2(A.bc[TRM_BC_TEMPERATURE_BOTTOM].kind, Tn, bc_input(TRM_BC_TEMPERATURE_BOTTOM), met.dzf(1), false);
            const bool copy = RICH || p.sat_halo == TRM_HALO_COPY;
            kapp = copy ? kapn : bc2(thermal_conductivity_fast(p, 0.0f, 1.0f));
            if (RICH) Pp = halo_value2(A.bc[TRM_BC_PRESSURE_BOTTOM].kind, Pn, bc_input(TRM_BC_PRESSURE_BOTTOM), met.dzf(1), false);
        } else {
            Tp = rd(EF_T); kapp = rd(EF_KAP);
            if (RICH) Pp = rd(EF_P);
        }
        // ---- (negated) heat flux and head gradient at face m ----
        F2 nqh = zero2, gn = zero2;
        if (inner || m <= nz + 1) {
            const float rz = met.rdzf(m);
            nqh = ((kapn + kapp) * (0.5f * rz)) * (Tn - Tp);
            if (RICH) gn = (Pn - Pp) * rz;
        }
        const F2 dnqh = nqh - rd(EF_QH);
        // ---- (negated) Darcy flux at face m-1 ----
        F2 nqd = zero2;
        if (RICH && (inner || m >= 2)) {
            const F2 g = rd(EF_G);
            const F2 Kf2 = lds2(kf_cur);
            const F2 Kk = min2(Kf1, sel(B2{g.x < 0.0f, g.y < 0.0f}, Kf2, Kfn));
            nqd = Kk * g;
        }
        F2 G_top = zero2, infil_top = zero2;
        if (LAND && !inner && m == nz + 2) { G_top = ldg2(A.G + c); infil_top = ldg2(A.infil + c); }

        if (inner || m >= 3) {
            const int j = m - 2;
            const uint32_t o = oout;
            oout += (uint32_t)ld;
            const float rzc = met.rdzc(j);
            F2 tU = rd(EF_DQH) * rzc;
            F2 tS = zero2;
            if (RICH) tS = fma2(nqd - rd(EF_QD), rzc, p.vwcf) * p.rpor;
            F2 Ub, sb;
            if (H2) {   // average_tendencies! (heun.jl:27-35) ; the base is the state at time n
                const uint32_t x = xring0 + (uint32_t)((j & (PF_ - 1)) * B * ES);
                tU = (lds2(x) + tU) * 0.5f;
                Ub = lds2(x + 2 * PF_ * B * ES);
                if (RICH) { tS = (lds2(x + PF_ * B * ES) + tS) * 0.5f; sb = lds2(x + 3 * PF_ * B * ES); }
                else sb = lds2(ringS(j));
            } else {
                if (H1) { stg2(A.oTU + o, tU); if (RICH) stg2(A.oTS + o, tS); }   // k1, before the Flux BCs
                Ub = lds2(ringU(j)); sb = lds2(ringS(j));
            }
            if (!inner && j == nz) {   // Flux boundary conditions (compute_z_bcs!, abstract_timestepper.jl:69)
                const float dz = met.dzc(nz);
                if (LAND) { tU = F2{tU.x - G_top.x / dz, tU.y - G_top.y / dz}; if (RICH) tS = F2{tS.x - (-infil_top.x) / dz, tS.y - (-infil_top.y) / dz}; }
                else {
                    if (A.bc[TRM_BC_ENERGY_TOP].kind == TRM_BC_FLUX) { const F2 f = bc_input(TRM_BC_ENERGY_TOP); tU = F2{tU.x - f.x / dz, tU.y - f.y / dz}; }
                    if (RICH && A.bc[TRM_BC_SATURATION_TOP].kind == TRM_BC_FLUX) { const F2 f = bc_input(TRM_BC_SATURATION_TOP); tS = F2{tS.x - f.x / dz, tS.y - f.y / dz}; }
                }
            }
            if (!inner && j == 1) {
                const float dz = met.dzc(1);
                if (A.bc[TRM_BC_ENERGY_BOTTOM].kind == TRM_BC_FLUX) { const F2 f = bc_input(TRM_BC_ENERGY_BOTTOM); tU = F2{tU.x + f.x / dz, tU.y + f.y / dz}; }
                if (RICH && A.bc[TRM_BC_SATURATION_BOTTOM].kind == TRM_BC_FLUX) { const F2 f = bc_input(TRM_BC_SATURATION_BOTTOM); tS = F2{tS.x + f.x / dz, tS.y + f.y / dz}; }
            }
            // ---- explicit step ----
            const F2 Un = fma2(tU, dt, Ub);
            F2 sn = sb;
            if (RICH) {
                sn = fma2(tS, dt, sn) + carry;
                if (inner || j < nz) {   // upward sweep of adjust_saturation_profile! (soil_hydrology.jl:192-199)
                    const F2 e = pos2(sn + -1.0f);
                    sn = sn - e;
                    carry = e * (met.dzc(j) * met.rdzc(j + 1));
                }
                negx |= __float_as_int(sn.x); negy |= __float_as_int(sn.y);
            }
            // A column that has gone negative keeps storing its raw values (no surface excess); the closure stores of
            // such a column are overwritten by its slow path after the sweep.
            if (RICH) {
                if (!inner && j == nz) {                         // top excess -> surface_excess_water (:210-214)
                    F2 e = pos2(sn + -1.0f);
                    e = F2{negx < 0 ? 0.0f : e.x, negy < 0 ? 0.0f : e.y};
                    sn = sn - e;
                    Sx_new = fma2(e, met.dzc(nz), Sx_new);
                }
                stg2(A.yS + o, sn);
                if (idxx == 0 && below_one(sn.x)) { idxx = j; wt_new.x = met.zF(j); }   // compute_water_table!, kernel_utils.jl:7-16
                if (idxy == 0 && below_one(sn.y)) { idxy = j; wt_new.y = met.zF(j); }
            }
            stg2(A.yU + o, Un);
            if (CLOSE || (LAND && has_veg(A))) {
                F2 Tc, lc, wi;
                energy_to_temperature2(p, Un, sn, Tc, lc, wi);
                if (LAND && has_veg(A)) {   // soil moisture limiting factor of the NEW state (plant_available_water.jl:31-35)
                    F2 x = ((wi * lc) + (-A.vp.th_wp)) * A.vp.r_paw_span;
                    x = F2{fminf(fmaxf(x.x, 0.0f), 1.0f), fminf(fmaxf(x.y, 0.0f), 1.0f)};
                    wr(EF_BETA, fma2(x, met.root(j), rd(EF_BETA)));
                }
                if (CLOSE) {
                    stg2(A.yT + o, Tc); stg2(A.yL + o, lc);
                    if (!inner && j == nz && A.hio_out) { A.hio_out[c] = Tc.x; if (vy) A.hio_out[c + 1] = Tc.y; }
                    // layers below the water table wait for it (written after the sweep); in a pair with only one column
                    // still below its water table that column's value is overwritten there
                    if (RICH && (idxx | idxy) != 0) stg2(A.yP + o, pressure_head2<SOIL>(p, sn, wt_new, met.zC(j), met.psiz(j)));
                }
            }
        }
        wr(EF_T, Tn); wr(EF_KAP, kapn); wr(EF_QH, nqh); wr(EF_DQH, dnqh);
        if (RICH) { wr(EF_P, Pn); sts2(kf_cur, Kfn); wr(EF_G, gn); wr(EF_QD, nqd); }
        const uint32_t t = kf_cur; kf_cur = kf_prv; kf_prv = t;
    };
    {
        int m = 1;
#pragma unroll 1
        while (m <= nz + 2) {
            if (m >= 4 && m <= nz - DIST_) {
#pragma unroll 1
                do { iterate(m, std::true_type{}); ++m; } while (m <= nz - DIST_);
            } else {
                iterate(m, std::false_type{});
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
        const float psat = swrc_inverse<float, true, SOIL == SOIL2_VG2>(p, p.por, p.por);
        const float zref = met.zF(nz + 1);
        const float px = (wt_new.x - zref) + psat, py = (wt_new.y - zref) + psat;
        const int kx = nx ? 0 : (idxx <= nz ? idxx : nz + 1), ky = (ny || !vy) ? 0 : (idxy <= nz ? idxy : nz + 1);   // layers 1 .. k-1 are rewritten
        const int kmin = kx < ky ? kx : ky, kmax = kx < ky ? ky : kx;
        int64_t o = c;
        int k = 1;
#pragma unroll 1
        for (; k < kmin; ++k, o += ld) stg2(A.yP + o, F2{px, py});
        float* const rest = A.yP + (kx < ky ? 1 : 0);
        const float pr = kx < ky ? py : px;
#pragma unroll 1
        for (; k < kmax; ++k, o += ld) rest[o] = pr;
    }
    if (nx) euler2_slow_column<MS, MODE, LAND, SOIL == SOIL2_VG2>(A, met, c, Sx_new.x);
    if (ny && vy) euler2_slow_column<MS, MODE, LAND, SOIL == SOIL2_VG2>(A, met, c + 1, Sx_new.y);
}

}  // namespace trm
