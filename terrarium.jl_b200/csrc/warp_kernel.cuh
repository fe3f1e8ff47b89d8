// warp_kernel.cuh -- small domains: one WARP per column, one lane per layer, `nsteps` time steps inside one launch.
//
// The streaming kernels (euler_kernel.cuh) give a column to one thread: on the reference's real grids (N72: 14 017
// columns, N145: 56 951) the GPU then holds less than one wave of warps and a step costs the latency of one thread's
// sweep over the nz layers (23 us Float32, 33 us Float64), whatever the column count. The explicit scheme is parallel
// over the layers of a column, so here the 32 lanes of a warp hold the layers of ONE column (lane l = layer l + 1, lane nz
// = the halo cell above the surface, nz <= 31): the vertical stencil is five warp shuffles per evaluation, the water
// table is a ballot, the upward saturation sweep is a five-round scan (fast math; the reference's sequential chain in faithful
// math and for the rare downward sweep), and because columns never interact the warp
// advances its column `nsteps` steps with U, sat, the water table and the surface excess water in registers -- global
// memory is touched at the start and at the end of the launch only. Heun runs both stages back to back in the same
// registers (no stage state, no k1 in memory). Per-cell arithmetic: the functions of column_physics.cuh in the order of
// the streaming kernels (same reference citations), so both kernels agree to rounding and, where no transcendental and no
// contraction is involved (faithful math, heat only), bit for bit.
//
// LandModel (LAND): the per-column surface block stays a launch of its own (surface_kernel, stage_kernel.cuh: per-column scalar
// work that a lane-per-layer layout cannot spread), so a LandModel step is surface_kernel + ONE launch of this kernel with
// nsteps = 1: the ground heat flux and the infiltration it left behind are the top Flux BCs, both Heun stages run here (stage 2
// applies the same time-n fluxes, heun.jl:63-66), and for the vegetated model the kernel leaves what the surface launches need:
// the soil moisture limiting factor of the new state and -- Heun -- of the stage state, plus the stage state's top layer, on
// which the stage-2 surface launch (k2 of the vegetation prognostics only) runs AFTER this kernel.
#pragma once

#include "stage_kernel.cuh"

namespace trm {

#ifndef TRM_WARP_BLOCK
#define TRM_WARP_BLOCK 128   // four adjacent columns per block (measured equal to eight: profiles/r02_summary.md)
#endif
#ifndef TRM_WARP_F64_THREADS
// resident threads per SM the Float64 fast-math instantiations must allow. Measured on 14 017 / 49 152 columns (us per step,
// ForwardEuler): 512 threads (<= 128 registers) 8.51 / 27.9 ; 640 (96) 7.65 / 24.6 ; 768 (80, 48-96 bytes of spills) 7.28 / 22.9
#define TRM_WARP_F64_THREADS 768
#endif
#ifndef TRM_WARP_F32_THREADS
#define TRM_WARP_F32_THREADS 1024   // Float32: <= 64 registers (1280 threads / 48 registers measured 2-5 % slower; Float64 with 1024 / 64: +5 % on full domains, -12 % on a single column)
#endif
constexpr int WARP_MAX_NZ = 31;   // lane nz is the halo cell above the surface

enum WarpSoil { WSOIL_GENERIC = 0 /* run-time tests, general formulas out of line */, WSOIL_VG2 = 1 /* van Genuchten n = 2 */,
                WSOIL_BC_LINEAR = 2 /* Brooks-Corey with integer 1 / lambda + linear conductivity: the reference's defaults */ };

__device__ __forceinline__ bool sign_set(double v) { return __double2hiint(v) < 0; }
__device__ __forceinline__ bool sign_set(float v)  { return __float_as_int(v) < 0; }
__device__ __forceinline__ bool lt_one(double v) { return __double2hiint(v) < 0x3FF00000; }
__device__ __forceinline__ bool lt_one(float v)  { return __float_as_int(v) < 0x3F800000; }

// Brooks-Corey matric head with an integer exponent k = 1 / lambda (fast math): psi_m = -psi_s se^-k below saturation,
// -psi_s at saturation (FreezeCurves BrooksCorey, SURVEY.md A.9 ; soil_hydraulic_closures.jl:115-118)
template <class NF>
__device__ __forceinline__ NF brooks_corey_psim_fast(const DevParams<NF>& p, NF theta) {
    const NF se = fma_(theta, p.r_thspan, p.se_off);
    NF pw = se;
#pragma unroll 1
    for (int i = 1; i < p.bc_k; ++i) pw = pw * se;
    NF r = M<NF, true>::rcp(pw) * -p.bc_psis;
    r = pw == NF(0) ? -Lim<NF>::inf() : r;   // dry layer: -Inf as in the reference
    return theta < p.por ? r : -p.bc_psis;
}
// UnsatKLinear (soil_hydraulic_properties.jl:170-198) with the reciprocal of fast math
template <class NF>
__device__ __forceinline__ NF cell_conductivity_linear_fast(const DevParams<NF>& p, NF sat, NF liq) {
    const NF wi = sat * p.por;
    const NF water = wi * liq, ice = wi * (1 - liq), air = (1 - sat) * p.por;
    return (water * p.Ksat) * M<NF, true>::rcp((water + ice) + air);
}

template <class NF, bool RICH, bool FAST, int SOIL, bool LAND = false>
__global__ void __launch_bounds__(TRM_WARP_BLOCK, (sizeof(NF) == 4 ? TRM_WARP_F32_THREADS : (FAST ? TRM_WARP_F64_THREADS : 512)) / TRM_WARP_BLOCK)   // <= 64 / 80 (faithful: 128) registers
column_warp_kernel(const __grid_constant__ StageArgs<NF> A, const int nsteps, const int heun) {
    using Mx = M<NF, FAST>;
    constexpr bool VG2 = SOIL == WSOIL_VG2;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // A block owns CW adjacent columns. Fields are [layer][column]: a lane of a column's warp would touch one 4-byte element
    // of a sector per access, so the block moves its nz x CW tile of every field between global and shared memory with all
    // threads (CW consecutive threads = one contiguous row segment: 16 / 32 bytes of a 32-byte sector in Float32 / Float64
    // instead of 4 / 8) and the warps pick their column out of the tile. That matters when a launch advances one step only
    // (LandModel, per-step callers): loads and stores are then the larger part of the launch.
    constexpr int CW = TRM_WARP_BLOCK / 32, TS = CW + 1;   // (odd row stride: the column reads are bank-conflict free)
    __shared__ NF tile[5][32 * TS];
    const int wi = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * CW;
    // Warps past the last column of a ragged last block repeat the work of the last column and store nothing: every warp runs
    // the same code, so the compiler keeps the shuffles free of re-convergence code (a loop that depends on `active`, which it
    // cannot prove warp-uniform, cost a WARPSYNC sequence per shuffle and 15-20 % of the step time).
    const bool active = c0 + wi < A.ncol;
    const int wc = active ? wi : (int)(A.ncol - 1 - c0);   // column of the tile this warp computes
    const int64_t c = c0 + wc;
    const int tr = threadIdx.x / CW, tc = threadIdx.x % CW;   // this thread's element of the tile in the cooperative copies
    const int64_t tg = (int64_t)tr * A.ld + c0 + tc;          // (rows are padded to a multiple of 64 columns: reads stay in bounds)
    const DevParams<NF>& p = A.p;
    const int nz = A.nz;
    const int k = lane + 1;                               // 1-based layer (lanes < nz) / face index (lanes <= nz)
    const bool isL = lane < nz, isH = lane == nz, isTop = lane == nz - 1, isBot = lane == 0;
    const NF dt = A.dt;
    const int64_t o = (int64_t)lane * A.ld + c;           // this lane's cell in a [layer][column] field
    const int to = lane * TS + wc;                        // ... and in the shared-memory tile

    // grid metrics of this lane's layer / lower face, read once (rows of MET_STRIDE values, 1-based + halos)
    const NF* met = A.metrics;
    const int kc = min(k, nz), kf = min(k, nz + 1);
    const NF zC_k = met[MET_ZC * MET_STRIDE + kc], psiz_k = met[MET_PSIZ * MET_STRIDE + kc];
    const NF dzc_k = met[MET_DZC * MET_STRIDE + kc], rdzc_k = met[MET_RDZC * MET_STRIDE + kc];
    const NF zF_k = met[MET_ZF * MET_STRIDE + kf];
    const NF dzf_k = met[MET_DZF * MET_STRIDE + kf], rdzf_k = met[MET_RDZF * MET_STRIDE + kf];
    const NF dzc_up = __shfl_down_sync(FULL, dzc_k, 1), rdzc_up = __shfl_down_sync(FULL, rdzc_k, 1);   // layer k + 1
    const NF dzc_dn = __shfl_up_sync(FULL, dzc_k, 1);                                                    // layer k - 1
    const NF zsurf = met[MET_ZF * MET_STRIDE + nz + 1];

    // Boundary values. Evaluating an input is per-column scalar work (a Float64 `sin` for the sinusoid form) that one lane would
    // do while 31 wait, every step: instead, every 31 steps lane i evaluates every boundary input at the clock time of step
    // i of the block (t advanced i times by dt in the clock's number format, exactly as the step loop does) and a step reads
    // its values with a shuffle: slot value at the step's start time from lane `si`, at start + dt (what Heun stage 2 needs
    // for Value / Gradient BCs; Flux BCs always belong to the time-n state, abstract_timestepper.jl:69, heun.jl:63-66) from
    // lane `si + 1`. Inputs whose descriptor is per step on the host side (tables / rasters: time bracket ; host evaluated
    // functions: value pair) run one step per launch, with descriptor index 0 = step start (lane 0), 1 = start + dt (lane 1).
    constexpr int BC_BLOCK = 31;
    NF t = A.t_x;
    NF bcv[TRM_BC_NSLOTS];
#pragma unroll
    for (int q = 0; q < TRM_BC_NSLOTS; ++q) bcv[q] = NF(0);
    int si = 0;   // step index inside the current block of BC_BLOCK steps
    auto refresh_bcs = [&](int remaining) {   // `remaining` steps of the launch: lanes beyond it hold values nobody reads
        NF tl = t;
        const int adds = min(lane, remaining);
#pragma unroll 1
        for (int i = 0; i < adds; ++i) tl = tl + dt;
#pragma unroll
        for (int q = 0; q < TRM_BC_NSLOTS; ++q)
            if (A.bc[q].kind != TRM_BC_DEFAULT) bcv[q] = eval_input(A.in[A.bc[q].input], c, tl, lane == 0 ? 0 : 1);
    };
    auto bc_val = [&](int slot, bool stage2) -> NF {   // (call from converged code only)
        const int kind = A.bc[slot].kind;
        if (kind == TRM_BC_DEFAULT) return NF(0);
        const bool late = stage2 && kind != TRM_BC_FLUX;
        return __shfl_sync(FULL, bcv[slot], si + (late ? 1 : 0));
    };
    auto pressure = [&](NF s, NF wt) -> NF {
        if (FAST && SOIL == WSOIL_BC_LINEAR) return (Mx::pos(wt + (-zC_k)) + brooks_corey_psim_fast(p, s * p.por)) + psiz_k;
        return pressure_head<NF, FAST, VG2>(p, s, wt, zC_k, psiz_k);
    };
    const bool halo_copy = RICH || p.sat_halo == TRM_HALO_COPY;
    // boundary condition kinds are launch uniform: read once, not once per step
    const int kT_top = A.bc[TRM_BC_TEMPERATURE_TOP].kind, kT_bot = A.bc[TRM_BC_TEMPERATURE_BOTTOM].kind;
    const int kP_top = A.bc[TRM_BC_PRESSURE_TOP].kind, kP_bot = A.bc[TRM_BC_PRESSURE_BOTTOM].kind;
    const bool fE_top = !LAND && A.bc[TRM_BC_ENERGY_TOP].kind == TRM_BC_FLUX, fE_bot = A.bc[TRM_BC_ENERGY_BOTTOM].kind == TRM_BC_FLUX;
    const bool fS_top = !LAND && RICH && A.bc[TRM_BC_SATURATION_TOP].kind == TRM_BC_FLUX, fS_bot = RICH && A.bc[TRM_BC_SATURATION_BOTTOM].kind == TRM_BC_FLUX;
    const NF G_top = LAND ? A.G[c] : NF(0), infil_top = (LAND && RICH) ? A.infil[c] : NF(0);   // (LandModel launches advance one step)
    const bool veg = LAND && has_veg(A);
    const NF root_k = veg ? met[MET_ROOT * MET_STRIDE + kc] : NF(0);
    // soil moisture limiting factor of a state: Integral(PAW * root_fraction / dz, dims = 3) (plant_available_water.jl:31-35),
    // a warp sum here (the streaming kernels and the oracle add bottom -> top: equal to rounding)
    auto beta_of = [&](NF sat, NF liq) -> NF {
        NF b = NF(0);
        if (isL) b = FAST ? plant_available_water_fast(A.vp, p, sat, liq) * root_k : plant_available_water(A.vp, p, sat, liq) * root_k / dzc_k * dzc_k;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) b += __shfl_xor_sync(FULL, b, d);
        return b;
    };
    // Oceananigans halo fill (halo_value, stage_kernel.cuh). Fast math: every form is linear in the edge cell and the boundary
    // value -- Value: edge + (v - edge) / (D / 2) * D = 2 v - edge ; Gradient: edge +- v D ; none: edge -- so the lane that forms a
    // halo (lane nz: above the surface ; lane 0: below the bottom layer) keeps its two coefficients per field and a halo is one
    // multiply + one FMA, selected into place (no division, no branch)
    auto halo_coef = [&](int kind, bool top, NF& ce, NF& cv) {
        ce = kind == TRM_BC_VALUE ? NF(-1) : NF(1);
        cv = kind == TRM_BC_VALUE ? NF(2) : (kind == TRM_BC_GRADIENT ? (top ? dzf_k : -dzf_k) : NF(0));
    };
    NF ceT, cvT, ceP, cvP;
    halo_coef(isH ? kT_top : kT_bot, isH, ceT, cvT);
    halo_coef(isH ? kP_top : kP_bot, isH, ceP, cvP);
    auto kappa_dry = [&]() -> NF { return FAST ? thermal_conductivity_fast(p, NF(0), NF(1)) : thermal_conductivity(p, NF(0), NF(1)); };

    // compute_auxiliary! + compute_tendencies! on the state (Ux, sx, wtx): tendencies of this lane's layer, before Flux BCs
    auto evaluate = [&](NF Ux, NF sx, NF wtx, bool stage2, bool loaded, NF& tU, NF& tS, NF& liq_out) {
        const NF bT_top = kT_top != TRM_BC_DEFAULT ? bc_val(TRM_BC_TEMPERATURE_TOP, stage2) : NF(0);
        const NF bT_bot = kT_bot != TRM_BC_DEFAULT ? bc_val(TRM_BC_TEMPERATURE_BOTTOM, stage2) : NF(0);
        const NF bP_top = (RICH && kP_top != TRM_BC_DEFAULT) ? bc_val(TRM_BC_PRESSURE_TOP, stage2) : NF(0);
        const NF bP_bot = (RICH && kP_bot != TRM_BC_DEFAULT) ? bc_val(TRM_BC_PRESSURE_BOTTOM, stage2) : NF(0);
        NF T = NF(0), liq = NF(1), P = NF(0), kap = NF(0), Kc = NF(0);
        if (isL) {
            if (loaded) { T = tile[2][to]; liq = tile[3][to]; if (RICH) P = tile[4][to]; }   // (the tile holds the inputs until the end)
            else {
                energy_to_temperature<NF, FAST>(p, Ux, sx, T, liq);
                if (RICH) P = pressure(sx, wtx);
            }
            kap = FAST ? thermal_conductivity_fast(p, sx, liq) : thermal_conductivity(p, sx, liq);
            if (RICH) Kc = (FAST && SOIL == WSOIL_BC_LINEAR) ? cell_conductivity_linear_fast(p, sx, liq) : cell_conductivity<NF, FAST, VG2>(p, sx, liq);
        }
        liq_out = liq;
        const NF T_dn = __shfl_up_sync(FULL, T, 1), kap_dn = __shfl_up_sync(FULL, kap, 1);
        NF P_dn = NF(0), Kf = NF(0);
        if (RICH) {
            P_dn = __shfl_up_sync(FULL, P, 1);
            const NF Kc_dn = __shfl_up_sync(FULL, Kc, 1);
            // face conductivity Kf[k] (soil_hydrology.jl:249-276): Kf[1] = Kc[1], Kf[Nz] = Kc[Nz], Kf[Nz+1] = Kf[Nz]
            if (isL) Kf = (isBot || isTop) ? Kc : Mx::mn(Kc, Kc_dn);
            const NF Kf_dn = __shfl_up_sync(FULL, Kf, 1);
            if (isH) Kf = Kf_dn;
        }
        NF Tp = T_dn, kapp = kap_dn, Pp = P_dn;
        if (FAST) {
            // halo cell above the surface (lane nz, from layer nz) and below the bottom layer (lower neighbour of lane 0)
            // (fill_halo_regions!), branch free
            const NF hT = fma_(cvT, isH ? bT_top : bT_bot, ceT * (isH ? T_dn : T));
            const NF kd = halo_copy ? NF(0) : kappa_dry();
            T = isH ? hT : T;
            kap = isH ? (halo_copy ? kap_dn : kd) : kap;
            Tp = isBot ? hT : Tp;
            kapp = isBot ? (halo_copy ? kap : kd) : kapp;
            if (RICH) {
                const NF hP = fma_(cvP, isH ? bP_top : bP_bot, ceP * (isH ? P_dn : P));
                P = isH ? hP : P;
                Pp = isBot ? hP : Pp;
            }
        } else {
            if (isH) {
                T = halo_value(kT_top, T_dn, bT_top, dzf_k, true);
                kap = halo_copy ? kap_dn : kappa_dry();
                if (RICH) P = halo_value(kP_top, P_dn, bP_top, dzf_k, true);
            }
            if (isBot) {
                Tp = halo_value(kT_bot, T, bT_bot, dzf_k, false);
                kapp = halo_copy ? kap : kappa_dry();
                if (RICH) Pp = halo_value(kP_bot, P, bP_bot, dzf_k, false);
            }
        }
        // heat flux and head gradient at face k (diffusive_heat_flux, soil_energy.jl:134-149), faces 1 .. nz + 1
        NF qh = NF(0), g = NF(0);
        if (lane <= nz) {
            qh = -((kap + kapp) / 2) * ((T - Tp) * rdzf_k);
            if (RICH) g = (P - Pp) * rdzf_k;
        }
        const NF qh_up = __shfl_down_sync(FULL, qh, 1);
        tU = -((qh_up - qh) * rdzc_k);                                   // soil_energy.jl:112-131
        tS = NF(0);
        if (RICH) {
            // Darcy flux at face k (darcy_flux, soil_hydrology_rre.jl:119-131): upwinded face conductivity, Kf[0] = Kf[Nz+2] = 0
            NF Kf_lo = __shfl_up_sync(FULL, Kf, 1), Kf_hi = __shfl_down_sync(FULL, Kf, 1);
            if (isBot) Kf_lo = NF(0);
            if (isH) Kf_hi = NF(0);
            NF Kk;
            if (FAST) Kk = Mx::mn(Kf, g < 0 ? Kf_lo : Kf_hi);
            else Kk = (g < 0 ? jmin(Kf_lo, Kf) : NF(0)) + (g >= 0 ? jmin(Kf, Kf_hi) : NF(0));
            const NF qd = lane <= nz ? -Kk * g : NF(0);
            const NF qd_up = __shfl_down_sync(FULL, qd, 1);
            const NF dth = -((qd_up - qd) * rdzc_k) + NF(0) + p.vwcf;     // soil_hydrology_rre.jl:95-117
            tS = FAST ? dth * p.rpor : dth / p.por;                       // soil_hydrology.jl:222-237
        }
    };
    // Flux boundary conditions of the time-n state on the tendencies of the top / bottom layer (compute_z_bcs!)
    auto apply_flux_bcs = [&](NF& tU, NF& tS) {
        if (LAND && isTop) {   // ground heat flux and infiltration left by surface_kernel (land_model.jl:56-62)
            tU -= G_top / dzc_k;
            if (RICH) tS -= (-infil_top) / dzc_k;
        }
        if (!(fE_top || fE_bot || fS_top || fS_bot)) return;
        const NF vE_top = fE_top ? bc_val(TRM_BC_ENERGY_TOP, false) : NF(0), vS_top = fS_top ? bc_val(TRM_BC_SATURATION_TOP, false) : NF(0);
        const NF vE_bot = fE_bot ? bc_val(TRM_BC_ENERGY_BOTTOM, false) : NF(0), vS_bot = fS_bot ? bc_val(TRM_BC_SATURATION_BOTTOM, false) : NF(0);
        if (isTop) {
            if (fE_top) tU -= vE_top / dzc_k;
            if (fS_top) tS -= vS_top / dzc_k;
        }
        if (isBot) {
            if (fE_bot) tU += vE_bot / dzc_k;
            if (fS_bot) tS += vS_bot / dzc_k;
        }
    };
    // adjust_saturation_profile! (soil_hydrology.jl:185-219) + compute_water_table! (:170-175) on the updated saturations of
    // the column. Returns this lane's saturation; `Sx` receives the top excess (surface_excess_water), `wt` the water table,
    // `idx` the lowest unsaturated layer (0: none), `slow` whether the downward sweep ran.
    auto adjust = [&](NF sn, NF& Sx, NF& wt, int& idx, bool& slow) -> NF {
        // upward sweep (:192-199): excess of a layer is handed to the layer above, scaled by the thickness ratio
        const bool over = isL && !isTop && sn > NF(1);
        if (FAST && __any_sync(FULL, over)) {
            // The carry out of a layer as a function of the carry in is h(c) = max(r (s + c - 1), 0) = max(alpha c + beta, gamma)
            // with alpha = r (thickness ratio), beta = r (s - 1), gamma = 0 -- a family closed under composition,
            // (a2, b2, g2) o (a1, b1, g1) = (a2 a1, a2 b1 + b2, max(a2 g1 + b2, g2)) -- so the carries of the whole column are
            // an inclusive scan over the lanes (five shuffle rounds) evaluated at c = 0 instead of a chain through nz layers.
            // Same recurrence, different association: the carries agree with the sequential form to rounding.
            const bool mid = isL && !isTop;
            const NF r = dzc_k * rdzc_up;
            NF al = mid ? r : NF(0), be = mid ? r * (sn - 1) : NF(0), ga = NF(0);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const NF a1 = __shfl_up_sync(FULL, al, d), b1 = __shfl_up_sync(FULL, be, d), g1 = __shfl_up_sync(FULL, ga, d);
                if (lane >= d) { ga = Mx::mx(fma_(al, g1, be), ga); be = fma_(al, b1, be); al = al * a1; }
            }
            NF cin = __shfl_up_sync(FULL, Mx::mx(be, ga), 1);
            cin = isBot ? NF(0) : cin;
            sn = sn + cin;
            if (mid) { const NF e = Mx::pos(sn - 1); sn -= e; }
        } else if (__any_sync(FULL, over)) {
            NF carry = NF(0);
#pragma unroll 1
            for (int l = 0; l < nz; ++l) {
                const NF cin = __shfl_sync(FULL, carry, (l + 31) & 31);   // (lane 31 is never a layer: 0 for the bottom layer)
                if (lane == l) {
                    sn = sn + cin;
                    if (!isTop) {
                        const NF e = Mx::pos(sn - 1);
                        sn -= e;
                        carry = FAST ? e * dzc_k * rdzc_up : e * dzc_k / dzc_up;
                    }
                }
            }
        } else {
            sn = sn + NF(0);   // (the carry every layer adds in the sequential form)
        }
        const bool neg = isL && (FAST ? sign_set(sn) : sn < 0);
        slow = __any_sync(FULL, neg) != 0;
        unsigned unsat;
        if (!slow) {
            // downward sweep with no deficit anywhere: sat += max(-sat, 0) is the identity (:201-208)
            if (!FAST && isL && k >= 2) sn = sn + jmax(-sn, NF(0));
            if (isTop) {                                   // top excess -> surface_excess_water (:210-214)
                const NF e = Mx::pos(sn - 1);
                sn -= e;
                Sx += e * dzc_k;
            }
            if (!FAST && isBot) sn = jmax(sn, NF(0));      // :216
            unsat = __ballot_sync(FULL, isL && (FAST ? lt_one(sn) : sn < 1));
        } else {
            // a layer went negative: downward sweep top -> bottom (:201-216), deficits are taken from the layer below
            NF carry_dn = NF(0);
#pragma unroll 1
            for (int l = nz - 1; l >= 0; --l) {
                const NF cin = __shfl_sync(FULL, carry_dn, (l + 1) & 31);
                if (lane == l) {
                    NF s = sn;
                    if (k < nz) s -= cin;
                    if (k >= 2) {
                        const NF d = jmax(-s, NF(0));
                        s += d;
                        carry_dn = d * dzc_k / dzc_dn;
                    }
                    if (k == nz) {
                        const NF e = jmax(s - 1, NF(0));
                        s -= e;
                        Sx += e * dzc_k;
                    }
                    if (k == 1) s = jmax(s, NF(0));
                    sn = s;
                }
            }
            unsat = __ballot_sync(FULL, isL && sn < 1);
        }
        Sx = __shfl_sync(FULL, Sx, nz - 1);
        idx = __ffs((int)unsat);                          // compute_water_table!, kernel_utils.jl:7-16 ; 0 = all saturated
        wt = __shfl_sync(FULL, zF_k, idx ? idx - 1 : nz);  // zF(idx), or z of the surface
        return sn;
    };

    // ---- state of the column: registers for the whole launch ----
    if (tr < A.nz) {
        tile[0][tr * TS + tc] = A.xU[tg]; tile[1][tr * TS + tc] = A.xS[tg];
        if (A.load_aux) { tile[2][tr * TS + tc] = A.xT[tg]; tile[3][tr * TS + tc] = A.xL[tg]; if (RICH) tile[4][tr * TS + tc] = A.xP[tg]; }
    }
    __syncthreads();
    NF U = NF(0), s = NF(0);
    if (isL) { U = tile[0][to]; s = tile[1][to]; }
    NF wt = RICH ? A.xWt[c] : NF(0);
    NF Sx = RICH ? A.bSx[c] : NF(0);
    int idx = 0;
    bool slow = false;
    bool loaded = A.load_aux != 0;   // first evaluation after initialize / a user write: stored closure fields (see stage_kernel.cuh)

#pragma unroll 1
    for (int step = 0; step < nsteps; ++step) {
        if (si == BC_BLOCK) si = 0;
        if (si == 0) refresh_bcs(nsteps - step);
        NF tU, tS, liq_x;
        evaluate(U, s, wt, false, loaded, tU, tS, liq_x);
        loaded = false;
        if (heun) {
            // heun.jl:37-71: stage state = explicit step with k1 (+ Flux BCs) and its closure, k2 on the stage state at t + dt,
            // averaged tendencies (+ the Flux BCs of the time-n state) applied to the base state
            NF t1U = tU, t1S = tS;
            apply_flux_bcs(t1U, t1S);
            const NF Us = U + t1U * dt;
            NF ss = s, wts = wt;
            if (RICH) {
                NF Sx_stage = NF(0); int idx_s; bool slow_s;
                ss = adjust(s + t1S * dt, Sx_stage, wts, idx_s, slow_s);   // (the stage copy's surface excess water is not used)
            }
            NF k2U, k2S, liq_s;
            evaluate(Us, ss, wts, true, false, k2U, k2S, liq_s);
            if (veg) {
                // for the stage-2 surface launch (vegetation block on the stage state, heun.jl:45-58): top layer of the stage
                // state and its soil moisture limiting factor
                const NF bs = beta_of(ss, liq_s);
                if (active && isTop) { A.stU[o] = Us; if (RICH) A.stS[o] = ss; }
                if (active && lane == 0) A.sbeta[c] = bs;
            }
            tU = (tU + k2U) / 2;                                          // average_tendencies!, heun.jl:27-35
            if (RICH) tS = (tS + k2S) / 2;
        }
        apply_flux_bcs(tU, tS);
        U = U + tU * dt;                                                  // explicit_step!, abstract_timestepper.jl:113-141
        if (RICH) {
            Sx = Sx + NF(0) * dt;                                         // surface_excess_water tendency is zero (soil_hydrology.jl:260-267)
            s = adjust(s + tS * dt, Sx, wt, idx, slow);
        }
        t = t + dt;                                                       // tick!(clock, dt) in the clock's number format
        ++si;
    }
    if (nsteps <= 0) return;

    // ---- closure! of the final state, results into the tile, cooperative stores ----
    __syncthreads();   // (every warp is through with the inputs in the tile: it now takes the results)
    NF lc_new = NF(1);
    if (isL) {   // (a repeating warp writes the same values to the same tile column)
        NF Tc, lc;
        energy_to_temperature<NF, FAST>(p, U, s, Tc, lc);
        tile[0][to] = U; tile[2][to] = Tc; tile[3][to] = lc;
        if (active && isTop && A.hio_out) A.hio_out[c] = Tc;              // ground temperature -> mapped host memory
        lc_new = lc;
        if (RICH) {
            tile[1][to] = s;
            NF Pc;
            if (slow || (idx != 0 && k >= idx)) Pc = pressure(s, wt);
            else {
                // saturated zone below the water table: psi_m(sat >= 1) is a constant (see euler_kernel.cuh)
                const NF psat = (FAST && SOIL == WSOIL_BC_LINEAR) ? brooks_corey_psim_fast(p, p.por) : swrc_inverse<NF, FAST, VG2>(p, p.por, p.por);
                Pc = FAST ? (wt - zsurf) + psat : Mx::mx(NF(0), wt - zC_k) + psat + psiz_k;
            }
            tile[4][to] = Pc;
        }
    }
    __syncthreads();
    if (tr < A.nz && c0 + tc < A.ncol) {
        A.yU[tg] = tile[0][tr * TS + tc]; A.yT[tg] = tile[2][tr * TS + tc]; A.yL[tg] = tile[3][tr * TS + tc];
        if (RICH) { A.yS[tg] = tile[1][tr * TS + tc]; A.yP[tg] = tile[4][tr * TS + tc]; }
    }
    if (RICH && active && lane == 0) { A.yWt[c] = wt; A.ySx[c] = Sx; }
    if (veg) {   // soil moisture limiting factor of the NEW state, for the surface launch of the next step
        const NF b = beta_of(s, lc_new);
        if (active && lane == 0) A.ybeta[c] = b;
    }
}

}  // namespace trm
