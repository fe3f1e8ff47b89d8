// stage_kernel.cuh -- the fused per-column stage kernel (one launch per timestepper stage).
//
// One thread owns one column and streams it bottom -> top exactly once.  In that single sweep it
// performs what the reference does in ~25 separate KernelAbstractions launches per Euler step
// (SURVEY.md 3.2):
//
//   update_state!            src/state_variables.jl:72-80
//     reset_tendencies!        (tendencies never exist in memory here)
//     update_inputs!           inputs are evaluated in-kernel at clock time t (device resident forcing)
//     fill_halo_regions!       halo values are formed in registers from the BC descriptors
//     compute_auxiliary!       hydraulic conductivity at centres/faces, soil_hydrology.jl:145-163,249-276;
//                              LandModel: evaporation, runoff/infiltration, 2 x surface energy balance
//                              (land_model.jl:79-88)
//     compute_tendencies!      Richards/Darcy soil_hydrology_rre.jl:95-131 ; heat conduction soil_energy.jl:112-149
//   explicit_step!           src/timesteppers/abstract_timestepper.jl:65-141 (+ Flux BCs)
//   closure!                 adjust_saturation_profile! soil_hydrology.jl:185-219, compute_water_table! :170-175,
//                            saturation_to_pressure! soil_hydraulic_closures.jl:102-129,
//                            energy_to_temperature! soil_energy_closures.jl:99-159
//
// Memory traffic per column-layer-step (Richards + energy, Euler, LOAD_AUX = false): read U, sat;
// write U, sat, T, liq, psi = 7 * sizeof(NF).  Temperature, liquid fraction and pressure head of the
// state at time n are NOT re-read: they are recomputed in registers from (U, sat, water_table) with
// the same closure code that wrote them at the end of step n-1.  LOAD_AUX = true reads them from
// HBM instead (first step after initialize / after the user overwrote a field, where the stored
// closure fields are not a function of the stored prognostic fields).
//
// Data layout: [layer][column], column fastest: a warp touches 32 adjacent columns of one layer
// per load instruction (fully coalesced 256 B for FP64, 128 B for FP32); raw loads are software
// pipelined PF layers ahead in registers.
#pragma once

#include "column_physics.cuh"
#include "vegetation.cuh"

namespace trm {

enum StageMode { MODE_EULER = 0, MODE_HEUN1 = 1, MODE_HEUN2 = 2, MODE_TEND = 3, MODE_AUX = 4 };
// soil hydrology / model combination a kernel is compiled for: SoilModel with immobile water or Richards flow,
// LandModel (bare ground) on top of either (default_soil(grid, nothing) is the immobile variant, land_model.jl:111)
enum Phys { PHYS_NOFLOW = 0, PHYS_RICHARDS = 1, PHYS_LAND = 2 /* LandModel + Richards */, PHYS_LAND_NOFLOW = 3, PHYS_COUNT = 4 };
__host__ __device__ constexpr bool phys_richards(int phys) { return phys == PHYS_RICHARDS || phys == PHYS_LAND; }
__host__ __device__ constexpr bool phys_land(int phys) { return phys == PHYS_LAND || phys == PHYS_LAND_NOFLOW; }

// Grid metrics (reference 1-based layer / face indices + halos), one row of MET_STRIDE values per quantity.
// The fixed stride lets the kernels address `quantity[k]` as (k-dependent register) + immediate.
constexpr int MET_STRIDE = TRM_MAX_NZ + 3;
enum Metric { MET_ZF = 0, MET_ZC = 1, MET_DZC = 2, MET_RDZC = 3, MET_DZF = 4, MET_RDZF = 5, MET_PSIZ = 6 /* zC - zF[nz+1] */,
              MET_ROOT = 7 /* static root fraction, root_distribution.jl:47-56 ; only staged by the LandModel kernels */, MET_COUNT = 8 };
// rows a kernel stages in shared memory
__host__ __device__ constexpr int met_rows(bool land) { return land ? MET_COUNT : MET_COUNT - 1; }

// Shared-memory reads through an explicit 32-bit shared address: the generic-pointer form makes the compiler
// rebuild the shared window base (S2R SR_CgaCtaId + LEA) in front of every access when registers are tight.
__device__ __forceinline__ float  lds(uint32_t a, float*)  { float v;  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds(uint32_t a, double*) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
#define TRM_METRIC_ACCESSORS \
    __device__ __forceinline__ NF zF(int k) const { return get(MET_ZF, k); } \
    __device__ __forceinline__ NF zC(int k) const { return get(MET_ZC, k); } \
    __device__ __forceinline__ NF dzc(int k) const { return get(MET_DZC, k); } \
    __device__ __forceinline__ NF rdzc(int k) const { return get(MET_RDZC, k); } \
    __device__ __forceinline__ NF dzf(int k) const { return get(MET_DZF, k); } \
    __device__ __forceinline__ NF rdzf(int k) const { return get(MET_RDZF, k); } \
    __device__ __forceinline__ NF psiz(int k) const { return get(MET_PSIZ, k); } \
    __device__ __forceinline__ NF root(int k) const { return get(MET_ROOT, k); }
template <class NF, int STRIDE = MET_STRIDE>
struct Metrics {
    uint32_t base;   // shared address of the metric table (rows of STRIDE values)
    __device__ __forceinline__ NF get(int q, int k) const { return lds(base + (uint32_t)((q * STRIDE + k) * (int)sizeof(NF)), (NF*)nullptr); }
    TRM_METRIC_ACCESSORS
};
// Compact metric rows (nz + 3 <= CMET_STRIDE) inside the kernel parameters: the staged kernels read `quantity[k]` with a
// launch-uniform k straight from the constant bank (uniform loads: no shared-memory staging, no traffic on the LSU data
// pipe, which is what bounds those kernels -- profiles/r02_summary.md).
constexpr int CMET_STRIDE = 40;
template <class NF> struct StageArgs;
template <class NF>
struct MetricsC {
    const StageArgs<NF>* A;
    __device__ __forceinline__ NF get(int q, int k) const { return A->cmet[q * CMET_STRIDE + k]; }
    TRM_METRIC_ACCESSORS
};

#ifndef TRM_MAX_BLOCK
#define TRM_MAX_BLOCK 256   // largest block the stage kernel may be launched with
#endif
#ifndef TRM_MIN_BLOCKS
#define TRM_MIN_BLOCKS 1    // resident blocks per SM the register allocator must allow
#endif

// where a clock time falls on the time axis of a table input (searched once per launch on the host)
struct TimeBracket {
    int32_t i1, i2;      // rows of the two bracketing snapshots (i1 only when flat)
    int32_t flat, pad_;  // 1: before / after the axis or exactly on a node of a raster -> the value of row i1
    double e, dtt;       // t - times[i1], times[i2] - times[i1]
};

template <class NF>
struct InputDesc {
    int32_t kind, nt;
    NF cval;
    double period, lo, hi;
    const NF *a, *b, *c;     // FIELD: a ; SINUSOID: a = mean, b = amp, c = phase ; TABLE / RASTER: a = values[nt][ld]
    int64_t ld;
    TimeBracket br[2];       // TABLE / RASTER: bracket of t_x (0) and of t_b (1)
};

template <class NF>
struct StageArgs {
    int64_t ncol, ld;
    int32_t nz, mode, load_aux, richards;
    NF dt;
    NF t_x;   // clock time of the state the tendencies are evaluated on (halo BC inputs, forcing)
    NF t_b;   // clock time of the base state (Flux BC inputs); differs from t_x only in Heun stage 2
    // X: state the tendencies are evaluated on
    const NF *xU, *xS, *xT, *xL, *xP, *xWt;
    // B: base state the update is applied to
    const NF *bU, *bS, *bSx;
    // Y: updated state
    NF *yU, *yS, *yT, *yL, *yP, *yWt, *ySx;
    // tendencies: read in Heun stage 2, written in Heun stage 1 (without Flux BCs) and TEND (with)
    const NF *k1U, *k1S;
    NF *oTU, *oTS;
    // Heun, "recompute" protocol of the Float64 staged kernels (euler_kernel.cuh): stage 1 stores k1 and the stage water table
    // but not the stage state, which stage 2 rebuilds from the base state and k1. A column whose stage state needs the downward
    // sweep (negative saturation) is flagged by stage 1, which then stores its stage state in full (sU / sS, read by stage 2).
    const NF *sU, *sS, *hflag_in;
    NF* hflag_out;
    NF* Kf;   // z-face hydraulic conductivity [nz+1][ld], written in AUX / TEND
    // LandModel 2-D fields
    NF *Ts, *G, *SWup, *LWup, *Rnet, *Hs, *Hl, *Egnd, *infil, *runoff;
    // vegetated LandModel: 2-D fields (VegField order; auxiliaries are written by every evaluation on the model
    // state), the prognostic triple (C_veg, nu, w_can) of the state the tendencies are evaluated on (vx), of the base
    // state (vb) and of the updated state (vy), Heun k1 in / out, plant available water [nz][ld] (MODE_AUX only)
    int32_t veg, pad2_;
    NF* veg2d[VF_COUNT];
    const NF *vx[3], *vb[3], *vk1[3];
    NF *vy[3], *vok1[3];
    NF* paw;
    // soil moisture limiting factor of the state the surface block is evaluated on (read by surface_kernel) and of the
    // state a stage kernel writes (accumulated there while the closure fields of the new state are formed)
    const NF* xbeta;
    NF* ybeta;
    // (warp-per-column kernel, vegetated LandModel under Heun) what the stage-2 surface launch reads: top layer of the stage
    // state (stored into the stage fields) and the factor of the stage state
    NF *stU, *stS, *sbeta;
    // per-step exchange with a host-side coupler (trm_bind_host_io): when non-null, the temperature of the top layer of
    // the state this launch writes (ground_temperature) is also stored here -- a device pointer to page-locked, mapped
    // host memory, written straight from the stage kernel (no copy engine, no staging buffer)
    NF* hio_out;
    // the TEMPERATURE_TOP boundary value is a per-column vector (TRM_SRC_FIELD, possibly mapped host memory): the staged
    // kernel fetches it with cp.async at kernel start instead of a synchronous load at the top of the column
    int32_t bct_pre, pad3_;
    VegParams<NF> vp;
    const NF* metrics;   // [MET_COUNT][MET_STRIDE], see enum Metric
    NF cmet[MET_COUNT * CMET_STRIDE];   // the same rows with stride CMET_STRIDE when nz + 3 <= CMET_STRIDE (see MetricsC)
    DevParams<NF> p;
    trm_bc bc[TRM_BC_NSLOTS];
    InputDesc<NF> in[TRM_IN_COUNT];
};

// Heun traffic: with the stage state stored (stage 1: read U, sat; write U*, sat*, k1U, k1S -- stage 2: read U*, sat*, k1U, k1S,
// U, sat; write U, sat, T, liq, psi) a step moves 17 values per cell. The "recompute" protocol drops the stage state from
// memory: stage 2 rebuilds U* = U + dt (k1U + Flux BCs), sat* = adjust(sat + dt (k1S + Flux BCs)) of the entering layer from
// the base state and k1, which it needs anyway for the update two iterations later (one 8-deep ring serves both uses):
// 13 values per cell. Staged kernels of both number formats (the generic streaming kernel keeps the stored form);
// TRM_HEUN_STORE_STAGE switches it off.
template <class NF> __host__ __device__ constexpr bool heun_recompute() {
#ifdef TRM_HEUN_STORE_STAGE
    return false;
#else
    return true;
#endif
}

// vegetated LandModel? (TRM_NO_VEG: tuning builds that compile the vegetation code out)
template <class NF>
__device__ __forceinline__ bool has_veg(const StageArgs<NF>& A) {
#ifdef TRM_NO_VEG
    return false;
#else
    return A.veg != 0;
#endif
}

// update_inputs! (input_sources.jl:165-171) and function valued BCs, evaluated at clock time t.
// `which`: 0 when t is the time the tendencies are evaluated at (t_x), 1 for the time of the base state (t_b, Flux BCs)
template <class NF>
__device__ __forceinline__ NF eval_input_inline(const InputDesc<NF>& s, int64_t c, NF t, int which = 0) {
    // (an if-chain on the launch-uniform kind, cheapest kinds first: the jump table of a `switch` costs an indirect branch
    //  per input, and the surface block evaluates a dozen inputs per column)
    const int kind = s.kind;
    if (kind == TRM_SRC_CONST) return s.cval;
    if (kind == TRM_SRC_FIELD) return s.a[c];
    // host-evaluated function of time: the launcher points `a` at the values of the time the tendencies are evaluated
    // at (t_x) and `b` at those of the base state's time (t_b)
    if (kind == TRM_SRC_FIELD_PAIR) return which == 0 ? s.a[c] : s.b[c];
    switch (kind) {
        case TRM_SRC_SINUSOID: {
            // `2pi * t / period - lon` is Float64 arithmetic in Julia whatever NF is
            // (examples/simulations/soil_heat_global.jl:79-88); rounded when stored in the NF field.
            double ph = 6.283185307179586 * (double)t / s.period - (double)s.c[c];
            double v = (double)s.a[c] + (double)s.b[c] * sin(ph);
            v = fmin(fmax(v, s.lo), s.hi);
            return (NF)v;
        }
        case TRM_SRC_TABLE: {
            // FieldTimeSeries[Time(t)]: linear between bracketing snapshots, flat outside. The bracket of the clock
            // time is the same for every column: the host has searched the time axis (Handle::bracket) and passes
            // the two rows, e = t - t1 and dt = t2 - t1 ; frac = (NF)(e / dt) exactly as input_sources.jl:165-171 would
            const TimeBracket& k = s.br[which];
            if (k.flat) return s.a[(int64_t)k.i1 * s.ld + c];
            const NF frac = (NF)(k.e / k.dtt);
            return s.a[(int64_t)k.i2 * s.ld + c] * frac + s.a[(int64_t)k.i1 * s.ld + c] * (1 - frac);
        }
        case TRM_SRC_RASTER: {
            // update_from_raster! (ext/TerrariumRastersExt/TerrariumRastersExt.jl:96-121): x1 + eps (x2 - x1) / dt in
            // Float64 between nodes, the node value on a node, flat beyond either end of the time axis
            const TimeBracket& k = s.br[which];
            if (k.flat) return s.a[(int64_t)k.i1 * s.ld + c];
            const NF x1 = s.a[(int64_t)k.i1 * s.ld + c], x2 = s.a[(int64_t)k.i2 * s.ld + c];
            return (NF)((double)x1 + k.e * (double)(x2 - x1) / k.dtt);
        }
    }
    return NF(0);
}
// out-of-line copy for the layer loop (boundary condition inputs: rare, and inlined they would bloat the loop)
template <class NF>
__device__ __noinline__ NF eval_input(const InputDesc<NF>& s, int64_t c, NF t, int which = 0) { return eval_input_inline(s, c, t, which); }

// input evaluation of the surface block: inlined (independent loads overlap; measured 1.5 % faster than one shared
// out-of-line copy, TRM_SURFACE_SHARED_INPUTS)
template <class NF>
__device__ __forceinline__ NF surface_input(const InputDesc<NF>& s, int64_t c, NF t) {
#ifdef TRM_SURFACE_SHARED_INPUTS
    return eval_input(s, c, t);
#else
    return eval_input_inline(s, c, t);
#endif
}

// Oceananigans halo fill for one side.  `D` is the face spacing at the
// boundary, `top` selects the sign convention.
template <class NF>
__device__ __forceinline__ NF halo_value(int kind, NF edge, NF v, NF D, bool top) {
    if (kind == TRM_BC_VALUE) {
        if (top) { NF g = (v - edge) / (D / 2); return edge + g * D; }
        NF g = (edge - v) / (D / 2); return edge + g * (-D);
    }
    if (kind == TRM_BC_GRADIENT) return top ? edge + v * D : edge + v * (-D);
    return edge;
}

// ---- LandModel per-column surface processes ---------------------------------------------------
template <class NF>
struct Surface {
    NF SWd, LWd, Ta, pres, q, V, rain, Tskin_in;
    NF alb, emi;                     // ConstantAlbedo parameters or the PrescribedAlbedo inputs (albedo.jl:7-44)
    NF swu_in, lwu_in, hs_in, hl_in; // inputs of PrescribedRadiativeFluxes / PrescribedTurbulentFluxes
    double ra;
};

// compute_surface_energy_fluxes! (one evaluation), surface_energy_balance.jl:119-144:
// radiative_fluxes.jl:85-100,196-209 ; turbulent_fluxes.jl:85-105,137-150 ; skin_temperature.jl:76-80
template <class NF, bool FAST>
__device__ __forceinline__ void seb_fluxes(const DevParams<NF>& p, const Surface<NF>& a, NF Tsurf, NF Egnd,
                                           NF& swu, NF& lwu, NF& rnet, NF& hs, NF& hl, NF& G) {
    if (p.rad_kind == TRM_RADIATIVE_PRESCRIBED) {   // radiative_fluxes.jl:46-50
        swu = a.swu_in; lwu = a.lwu_in;
    } else {
        swu = a.alb * a.SWd;
        NF TK = Tsurf + p.Tref;
        lwu = a.emi * p.sigma * pow4(TK) + (1 - a.emi) * a.LWd;
    }
    rnet = swu - a.SWd + lwu - a.LWd;
    if (p.turb_kind == TRM_TURBULENT_PRESCRIBED) {  // turbulent_fluxes.jl:9-16
        hs = a.hs_in; hl = a.hl_in;
    } else {
        hs = (NF)((double)(p.c_a * p.rho_a) * dv<double, FAST>((double)(Tsurf - a.Ta), a.ra));
        hl = p.Llg * p.rho_a * Egnd;
    }
    G = rnet - hs - hl;
}

// Vegetation and canopy block of the vegetated LandModel for one column (see land_surface). Its own out-of-line
// function: the bare-ground LandModel must not pay for the registers this block needs around the call. Scalars come
// in by value; the results land_surface needs (ground / canopy evaporation, transpiration, rain reaching the ground)
// are read back from the auxiliary fields this function writes.
// results of the vegetation block that the rest of the surface block needs (kept in registers by the inlined variant)
template <class NF> struct VegOut { NF Egnd, rain_ground, E_can, transp; };

template <class NF, bool FAST>
__device__ __forceinline__ void vegetation_surface_impl(const StageArgs<NF>& A, int64_t c, NF Ta, NF SWd, NF pres, NF ea, NF rain, NF Vc, double ra,
                                                        NF dq, NF T_top, NF beta_sm, NF beta_g, bool stage2, VegOut<NF>& out) {
    const DevParams<NF>& p = A.p;
    const VegParams<NF>& v = A.vp;
    const NF co2 = surface_input(A.in[TRM_IN_CO2], c, A.t_x);
    const NF SAI = surface_input(A.in[TRM_IN_SAI], c, A.t_x);
    const NF Rdl = surface_input(A.in[TRM_IN_DAILY_LEAF_RESPIRATION], c, A.t_x);
    const NF Cv = A.vx[0][c], nu = A.vx[1][c], wcan = A.vx[2][c];
    const NF An_prev = A.veg2d[VF_AN][c];
    // PALADYNCarbonDynamics / PALADYNPhenology auxiliaries (carbon_dynamics.jl:82-85, phenology.jl:33-70)
    const NF LAIb = dv<NF, FAST>(Cv, (NF(2.0) / v.SLA) + v.awl);
    const NF fdec = NF(0), phen = NF(1.0);
    const NF LAI = (fdec * phen + (NF(1.0) - fdec)) * LAIb;
    // MedlynStomatalConductance (stomatal_conductance.jl:45-82): vapour pressure deficit at the air temperature,
    // net assimilation of the PREVIOUS evaluation (vegetation_carbon.jl:89-91)
    const NF vpd_air = xmax<NF, FAST>(saturation_vapor_pressure<NF, FAST>(Ta) - ea, NF(0.1));
    const NF fapar = 1 - xexp<NF, FAST>(-v.k_ext * LAI);   // also the absorbed fraction of PAR (photosynthesis.jl:124-128)
    const NF g0 = (v.g_min / 1000) * fapar * beta_sm;
    const NF gw = g0 + dv<NF, FAST>(NF(1.6) * (1 + dv<NF, FAST>(v.g1, M<NF, FAST>::sqrt_(vpd_air))) * An_prev, co2) * NF(1.0e6);
    const NF lamc = NF(1.0) - dv<NF, FAST>(NF(1.0), NF(1.0) + dv<NF, FAST>(v.g1, M<NF, FAST>::sqrt_(vpd_air * NF(1.0e-3))));
    // LUEPhotosynthesis (photosynthesis.jl:284-344)
    NF Rd, An;
    photosynthesis<NF, FAST>(v, Ta, SWd, pres, co2, LAI, fapar, lamc, beta_sm, Rd, An);
    const NF GPP = An * NF(1.0e-3);
    // PALADYNAutotrophicRespiration (autotrophic_respiration.jl:46-154) ; T_soil = ground temperature
    const NF f_soil = (T_top > 7) ? xexp<NF, FAST>(NF(308.56) * (NF(1.0) / NF(56.02) - dv<NF, FAST>(NF(1.0), NF(46.02) + T_top))) : NF(0);
    const NF f_air = xexp<NF, FAST>(NF(308.56) * (NF(1.0) / NF(56.02) - dv<NF, FAST>(NF(1.0), NF(46.02) + Ta)));
    const NF resp10 = NF(0.066);
    const NF R_leaf = Rdl / NF(1000.0);
    const NF R_stem = dv<NF, FAST>(resp10 * f_air * (v.awl * ((NF(2.0) / v.SLA) + v.awl)), Cv * v.aws * v.cn_sapwood);
    const NF R_root = dv<NF, FAST>(resp10 * f_soil * phen * (NF(2.0) / v.SLA), v.SLA * Cv * v.cn_root);
    const NF Rm = R_leaf + R_stem + R_root;
    const NF Rg = NF(0.25) * (GPP - Rm);
    const NF Ra = Rm + Rg;
    const NF NPP = GPP - Ra;
    // PALADYNCanopyInterception (canopy_interception.jl:79-187)
    const NF wmax = v.w_can_max * (LAI + SAI);
    const NF f_can = wmax > 0 ? dv<NF, FAST>(wcan, wmax) : NF(0);
    const NF I_can = v.alpha_int * rain * (NF(1) - xexp<NF, FAST>(-v.k_ext_can * (LAI + SAI)));
    const NF R_can = xmax<NF, FAST>(wcan, NF(0)) / v.tau_w;
    const NF rain_ground = rain - I_can + R_can;
    // PALADYNCanopyEvapotranspiration (canopy_evapotranspiration.jl:51-177): humidity gradients at the skin and at
    // the ground temperature, resistance between ground and canopy, stomatal resistance
    const NF esg = saturation_vapor_pressure<NF, FAST>(T_top);
    const NF dqg = dv<NF, FAST>(p.eps_mw * xmax<NF, FAST>(esg - ea, NF(0.1)), pres);
    const NF re = dv<NF, FAST>(1 - xexp<NF, FAST>(-LAI - SAI), v.C_can * Vc);
    const NF rs = dv<NF, FAST>(NF(1), xmax<NF, FAST>(gw, tsqrt(Lim<NF>::eps())));
    const NF transp = (NF)dv<double, FAST>((double)dq, ra + (double)rs);
    const NF Egnd = (NF)dv<double, FAST>((double)(beta_g * dqg), ra + (double)re);
    const NF E_can = (NF)dv<double, FAST>((double)(f_can * dq), ra);
    // tendencies: canopy water (canopy_interception.jl:121-127), vegetation carbon (carbon_dynamics.jl:107-112),
    // vegetation area fraction (vegetation_dynamics.jl:60-75)
    const NF lam = lambda_NPP(v, LAIb);
    NF k[3];
    k[2] = I_can - E_can - R_can;
    k[0] = (NF(1.0) - lam) * NPP - (v.gamma_L / v.SLA + v.gamma_R / v.SLA + v.gamma_S * v.awl) * LAIb;
    const NF nus = xmax<NF, FAST>(nu, v.nu_seed);
    k[1] = dv<NF, FAST>(lam * NPP, Cv) * nus * (NF(1.0) - nu) - v.gamma_v * nus;
    if (A.mode == MODE_EULER || A.mode == MODE_HEUN1 || A.mode == MODE_HEUN2) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            NF ki = k[i];
            if (A.mode == MODE_HEUN1) A.vok1[i][c] = ki;
            if (A.mode == MODE_HEUN2) ki = (A.vk1[i][c] + ki) / 2;   // average_tendencies!, heun.jl:27-35
            A.vy[i][c] = A.vb[i][c] + ki * A.dt;
        }
    }
    out.Egnd = Egnd; out.rain_ground = rain_ground; out.E_can = E_can; out.transp = transp;
    if (stage2) return;   // the auxiliaries of a stage-2 evaluation belong to the stage copy (heun.jl:45-58)
    NF* const* o = A.veg2d;
    o[VF_LAIB][c] = LAIb; o[VF_LAI][c] = LAI; o[VF_PHEN][c] = phen; o[VF_GWCAN][c] = gw; o[VF_LAMC][c] = lamc;
    o[VF_AN][c] = An; o[VF_RD][c] = Rd; o[VF_GPP][c] = GPP; o[VF_RA][c] = Ra; o[VF_NPP][c] = NPP; o[VF_BETASM][c] = beta_sm;
    o[VF_ICAN][c] = I_can; o[VF_RCAN][c] = R_can; o[VF_FCAN][c] = f_can; o[VF_RAING][c] = rain_ground;
    o[VF_ECAN][c] = E_can; o[VF_TRANSP][c] = transp;
}
// out-of-line copy for the generic stage kernel, which must not pay for the registers of this block around the call
template <class NF, bool FAST>
__device__ __noinline__ void vegetation_surface(const StageArgs<NF>& A, int64_t c, NF Ta, NF SWd, NF pres, NF ea, NF rain, NF Vc, double ra,
                                                NF dq, NF T_top, NF beta_sm, NF beta_g, bool stage2, VegOut<NF>& out) {
    vegetation_surface_impl<NF, FAST>(A, c, Ta, SWd, pres, ea, rain, Vc, ra, dq, T_top, beta_sm, beta_g, stage2, out);
}

// LandModel per-column surface processes for one update_state! (land_model.jl:79-88): bare-ground evaporation,
// runoff / infiltration, then the surface energy balance kernel twice. Reads the atmospheric inputs at clock time
// t, the skin temperature and the surface excess water; writes every 2-D surface field; returns the two fluxes
// that couple to the top soil layer. Out of line: it runs once per column and step, and inlined it would set the
// register budget of the whole layer loop.
// Vegetated LandModel (A.veg): beta_sm is the soil moisture limiting factor accumulated over the column by the caller;
// the block additionally evaluates the vegetation auxiliaries (vegetation_carbon.jl:71-104), canopy interception and
// evapotranspiration (surface_hydrology.jl:36-50) and advances canopy water, vegetation carbon and area fraction.
// `stage2` (Heun stage 2): only what feeds the k2 tendencies of those three variables is evaluated -- on the stage
// state, at t + dt -- and no auxiliary field is written (they belong to the stage copy in the reference, heun.jl:45-58).
template <class NF, bool FAST, bool INLINE_VEG>
__device__ __forceinline__ void land_surface_impl(const StageArgs<NF>& A, int64_t c, bool richards, NF T_top, NF sat_top, NF liq_top, NF K_top, NF dz_top,
                                                  NF beta_sm, bool stage2, NF& G_out, NF& inf_out) {
    const DevParams<NF>& p = A.p;
    const bool veg = has_veg(A);
    // (inputs: see surface_input)
    Surface<NF> a;
    // surface_excess_water(...) is the prognostic field under RichardsEq (soil_hydrology_rre.jl:28) and identically
    // zero for immobile soil water (soil_hydrology.jl:138)
    const NF Ts0 = A.Ts[c], S = richards ? A.bSx[c] : NF(0);
    a.SWd = surface_input(A.in[TRM_IN_SHORTWAVE_DOWN], c, A.t_x);
    a.LWd = surface_input(A.in[TRM_IN_LONGWAVE_DOWN], c, A.t_x);
    a.Ta = surface_input(A.in[TRM_IN_AIR_TEMPERATURE], c, A.t_x);
    a.pres = surface_input(A.in[TRM_IN_AIR_PRESSURE], c, A.t_x);
    a.q = surface_input(A.in[TRM_IN_SPECIFIC_HUMIDITY], c, A.t_x);
    a.V = surface_input(A.in[TRM_IN_WINDSPEED], c, A.t_x);
    a.rain = surface_input(A.in[TRM_IN_RAINFALL], c, A.t_x);
    const bool prescribed = p.skin == TRM_SKIN_PRESCRIBED;
    a.Tskin_in = prescribed ? surface_input(A.in[TRM_IN_SKIN_TEMPERATURE], c, A.t_x) : NF(0);
    a.alb = p.albedo; a.emi = p.emis;
    if (p.albedo_kind == TRM_ALBEDO_PRESCRIBED) { a.alb = surface_input(A.in[TRM_IN_ALBEDO], c, A.t_x); a.emi = surface_input(A.in[TRM_IN_EMISSIVITY], c, A.t_x); }
    a.swu_in = a.lwu_in = a.hs_in = a.hl_in = NF(0);
    if (p.rad_kind == TRM_RADIATIVE_PRESCRIBED) { a.swu_in = surface_input(A.in[TRM_IN_SHORTWAVE_UP], c, A.t_x); a.lwu_in = surface_input(A.in[TRM_IN_LONGWAVE_UP], c, A.t_x); }
    if (p.turb_kind == TRM_TURBULENT_PRESCRIBED) { a.hs_in = surface_input(A.in[TRM_IN_SENSIBLE_HEAT_FLUX], c, A.t_x); a.hl_in = surface_input(A.in[TRM_IN_LATENT_HEAT_FLUX], c, A.t_x); }
    // aerodynamic_resistance, prescribed_atmosphere.jl:110-116,137 (Float64 literal 1.0e-6 promotes)
    NF Vc = xmax<NF, FAST>(a.V, p.Vmin);
    double Va = fmax((double)Vc, 1.0e-6);
    a.ra = dv<double, FAST>(1.0, (double)p.C_h * Va);
    NF Ts = Ts0;
    // BareGroundEvaporation, bare_ground_evaporation.jl:49-62 ; compute_humidity_vpd
    // prescribed_atmosphere.jl:160-182, physical_constants.jl:83-97, physics_utils.jl:38
    NF Tsurf = prescribed ? a.Tskin_in : Ts;
    NF es = saturation_vapor_pressure<NF, FAST>(Tsurf);
    NF ea = dv<NF, FAST>(a.q * a.pres, p.eps_mw + (1 - p.eps_mw) * a.q);
    NF vpd = xmax<NF, FAST>(es - ea, NF(0.1));
    NF dq = dv<NF, FAST>(p.eps_mw * vpd, a.pres);
    // ground evaporation resistance factor, ground_resistance_factor.jl:6-11 (constant) / :32-57 (soil moisture limited)
    NF beta_g = p.beta;
    if (p.ground_res == TRM_GROUND_RES_SOIL_MOISTURE) {
        const NF thw = sat_top * p.por * liq_top;
        beta_g = NF(1);
        if (thw < p.th_fc) { const NF d = 1 - tcos(NF(3.141592653589793) * thw / p.th_fc); beta_g = d * d / 4; }
    }
    NF Egnd = (NF)dv<double, FAST>((double)(beta_g * dq), a.ra);
    NF rain_ground = a.rain;   // NoCanopyInterception: rainfall_ground aliases rainfall (canopy_interception.jl:11-15)
    NF Qh = Egnd;              // surface_humidity_flux of the evapotranspiration scheme
    if (veg) {
        VegOut<NF> vo;
        if (INLINE_VEG) vegetation_surface_impl<NF, FAST>(A, c, a.Ta, a.SWd, a.pres, ea, a.rain, Vc, a.ra, dq, T_top, beta_sm, beta_g, stage2, vo);
        else vegetation_surface<NF, FAST>(A, c, a.Ta, a.SWd, a.pres, ea, a.rain, Vc, a.ra, dq, T_top, beta_sm, beta_g, stage2, vo);
        if (stage2) { G_out = A.G[c]; inf_out = A.infil[c]; return; }   // Flux BCs use the time-n fluxes of stage 1 (heun.jl:63-66)
        Egnd = vo.Egnd;
        rain_ground = vo.rain_ground;
        Qh = Egnd + vo.E_can + vo.transp;   // surface_humidity_flux, canopy_evapotranspiration.jl:75-80
    }
    // DirectSurfaceRunoff, direct_surface_runoff.jl:87-117 ; K_top = Kf[Nz]
    NF drain, inf;
    if (S > 0) { drain = dv<NF, FAST>(xmax<NF, FAST>(S, NF(0)), p.tau_r); inf = (sat_top < 1) ? xmin<NF, FAST>(drain, K_top) : NF(0); }
    else { drain = 0; inf = (sat_top < 1) ? xmin<NF, FAST>(rain_ground, K_top) : NF(0); }
    NF runoff = rain_ground + drain - inf;
    // surface energy balance kernel, executed twice (land_model.jl:85-86) ; the latent heat flux follows the humidity
    // flux of the evapotranspiration scheme (turbulent_fluxes.jl:137-150)
    NF swu, lwu, rnet, hs, hl, G;
    // Each call of the fused SEB kernel is: fluxes(Ts) -> Ts = Tg - G dz / (2 kappa_s) -> fluxes(Ts). The first evaluation of
    // the second call sees exactly the inputs of the last evaluation of the first call and reproduces its results bit for
    // bit, so it is not repeated: three flux evaluations instead of four.
    seb_fluxes<NF, FAST>(p, a, prescribed ? a.Tskin_in : Ts, Qh, swu, lwu, rnet, hs, hl, G);
    if (!prescribed) {
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
            Ts = T_top - dv<NF, FAST>(G * dz_top, 2 * p.kappa_skin);   // ImplicitSkinTemperature, skin_temperature.jl:62-68,138-150
            seb_fluxes<NF, FAST>(p, a, Ts, Qh, swu, lwu, rnet, hs, hl, G);
        }
    }
    A.Egnd[c] = Egnd; A.infil[c] = inf; A.runoff[c] = runoff;
    A.SWup[c] = swu; A.LWup[c] = lwu; A.Rnet[c] = rnet; A.Hs[c] = hs; A.Hl[c] = hl; A.G[c] = G;
    if (!prescribed) A.Ts[c] = Ts;
    G_out = G; inf_out = inf;
}
// out of line for the generic stage kernel: it runs once per column and step, and inlined it would set the register
// budget of the whole layer loop
template <class NF, bool FAST>
__device__ __noinline__ void land_surface(const StageArgs<NF>& A, int64_t c, bool richards, NF T_top, NF sat_top, NF liq_top, NF K_top, NF dz_top,
                                          NF beta_sm, bool stage2, NF& G_out, NF& inf_out) {
    land_surface_impl<NF, FAST, false>(A, c, richards, T_top, sat_top, liq_top, K_top, dz_top, beta_sm, stage2, G_out, inf_out);
}

// ---- surface block as its own launch (staged ForwardEuler / Heun kernels, euler_kernel.cuh) ----------------------
// The surface processes of one update_state! only need the top soil layer of the state they are evaluated on, the
// per-column 2-D fields and (vegetated model) the soil moisture limiting factor, which the stage kernel that produced
// that state has left in `xbeta`. Running them as a separate one-thread-per-column launch keeps the layer loop of
// the stage kernel free of the call (and of its register demand); the stage kernel then reads G and the infiltration
// like any other Flux boundary condition.
// resident blocks of 128 threads the register allocator must allow. The block is one inlined flow (input loads up front, the
// vegetation results in registers): Float64 wants ~116 registers to be spill free; measured per 10 M vegetated columns with
// 8 / 6 / 5 / 4 blocks: 4.92 / 4.85 / 4.84 / 4.97 ms per step (bare ground 4.01 / - / 4.08 / 4.18), Float32 2.44 (8) / 2.51 (4);
// the out-of-line form of round 1 (results handed over through global memory, 328-byte stack frame): 5.22 / 4.07 / 2.65 ms
#ifndef TRM_SURFACE_BLOCKS
#define TRM_SURFACE_BLOCKS 0
#endif
template <class NF> constexpr int surface_blocks() { return TRM_SURFACE_BLOCKS ? TRM_SURFACE_BLOCKS : (sizeof(NF) == 8 ? 6 : 8); }
template <class NF, bool FAST>
__global__ void __launch_bounds__(128, (surface_blocks<NF>())) surface_kernel(const __grid_constant__ StageArgs<NF> A) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.ncol) return;
    const DevParams<NF>& p = A.p;
    const int nz = A.nz;
    const int64_t o = (int64_t)(nz - 1) * A.ld + c;
    const NF sat_top = A.xS[o];
    NF T_top, liq_top;
    if (A.load_aux) { T_top = A.xT[o]; liq_top = A.xL[o]; }
    else energy_to_temperature<NF, FAST>(p, A.xU[o], sat_top, T_top, liq_top);
    const NF K_top = cell_conductivity<NF, FAST>(p, sat_top, liq_top);   // Kf[Nz] = Kc[Nz], soil_hydrology.jl:249-276
    const NF dz_top = A.metrics[MET_DZC * MET_STRIDE + nz];
    NF G, inf;
#ifdef TRM_SURFACE_OUTLINE
    land_surface<NF, FAST>(A, c, A.richards != 0, T_top, sat_top, liq_top, K_top, dz_top, has_veg(A) ? A.xbeta[c] : NF(0), A.mode == MODE_HEUN2, G, inf);
#else
    // one inlined flow: every input load can be issued before the first dependent instruction, the results of the vegetation
    // block stay in registers, no stack frame
    land_surface_impl<NF, FAST, true>(A, c, A.richards != 0, T_top, sat_top, liq_top, K_top, dz_top, has_veg(A) ? A.xbeta[c] : NF(0), A.mode == MODE_HEUN2, G, inf);
#endif
}

// soil moisture limiting factor from the stored saturation / liquid fraction (first step after initialize or after the
// user overwrote a field; afterwards the stage kernels keep it current)
template <class NF, bool FAST>
__global__ void __launch_bounds__(128) beta_kernel(const __grid_constant__ StageArgs<NF> A) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.ncol) return;
    const NF* root = A.metrics + MET_ROOT * MET_STRIDE;
    const NF* dzc = A.metrics + MET_DZC * MET_STRIDE;
    NF b = 0;
    int64_t o = c;
#pragma unroll 1
    for (int k = 1; k <= A.nz; ++k, o += A.ld) {
        const NF s = A.xS[o], l = A.xL[o];
        b += FAST ? plant_available_water_fast(A.vp, A.p, s, l) * root[k] : plant_available_water(A.vp, A.p, s, l) * root[k] / dzc[k] * dzc[k];
    }
    A.ybeta[c] = b;
}

// Values produced while layer m enters the pipeline and consumed one or two iterations later.  Two
// instances alternate roles (the main loop is unrolled by two), so nothing is ever shifted between
// registers: "cur" is overwritten by iteration m, "prv" was written by iteration m-1, and the old
// content of "cur" is what iteration m-2 left behind.
template <class NF>
struct Stage {
    NF U, s;           // prognostic values of layer m (consumed when the layer is updated, at m+2)
    NF T, l, P;        // closure fields of layer m
    NF kap, Kc;        // thermal / hydraulic conductivity at the centre of layer m
    NF Kf;             // hydraulic conductivity at face m
    NF qh, g;          // heat flux and pressure-head gradient at face m (between layers m-1 and m)
    NF dqh;            // qh[m] - qh[m-1]   : heat flux divergence numerator of layer m-1
    NF qd;             // Darcy flux at face m-1
    NF rU, rS, rT, rL, rP;   // raw loads of layer m+2 (software prefetch, two iterations ahead)
};

template <class NF, int PHYS, int MODE_CT, int LOAD_CT, bool FAST>
__global__ void __launch_bounds__(TRM_MAX_BLOCK, TRM_MIN_BLOCKS) stage_kernel(const __grid_constant__ StageArgs<NF> A) {
    constexpr bool RICH = phys_richards(PHYS);
    constexpr bool LAND = phys_land(PHYS);
    using Mx = M<NF, FAST>;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nz = A.nz;
    {
        NF* sm = reinterpret_cast<NF*>(smem_raw);
        for (int q = 0; q < met_rows(LAND); ++q)
            for (int i = threadIdx.x; i < nz + 3; i += blockDim.x) sm[q * MET_STRIDE + i] = A.metrics[q * MET_STRIDE + i];
    }
    __syncthreads();
    Metrics<NF> met;
    met.base = (uint32_t)__cvta_generic_to_shared(smem_raw);

    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.ncol) return;
    const int64_t ld = A.ld;
    const DevParams<NF>& p = A.p;
    const int mode = MODE_CT >= 0 ? MODE_CT : A.mode;
    const bool load_aux = LOAD_CT >= 0 ? (LOAD_CT != 0) : (A.load_aux != 0);
    const bool need_K = RICH || LAND || mode == MODE_AUX || mode == MODE_TEND;   // compute_hydraulics! always runs in the
                                                                        // reference; only materialised when asked
    const bool write_K = mode == MODE_AUX || mode == MODE_TEND;
    const bool do_update = mode == MODE_EULER || mode == MODE_HEUN1 || mode == MODE_HEUN2;
    const bool full_closure = mode == MODE_EULER || mode == MODE_HEUN2;
    const NF dt = A.dt;

    // ---- boundary condition inputs at the times the reference evaluates them ----
    auto bc_input = [&](int slot) -> NF {
        const int kind = A.bc[slot].kind;
        if (kind == TRM_BC_DEFAULT) return NF(0);
        return eval_input(A.in[A.bc[slot].input], c, kind == TRM_BC_FLUX ? A.t_b : A.t_x, kind == TRM_BC_FLUX ? 1 : 0);
    };
    const NF bc_T_top = bc_input(TRM_BC_TEMPERATURE_TOP);
    const NF wtx = RICH && !load_aux ? A.xWt[c] : NF(0);

    // running element offsets: `oin` addresses the layer being prefetched, `oout` the layer being updated
    int64_t oin = c;
    auto prefetch = [&](Stage<NF>& st, int k) {
        if (k <= nz) {
            st.rU = A.xU[oin]; st.rS = A.xS[oin];
            if (load_aux) { st.rT = A.xT[oin]; st.rL = A.xL[oin]; if (RICH) st.rP = A.xP[oin]; }
            oin += ld;
        }
    };

    Stage<NF> s0, s1;
    {
        Stage<NF> z;
        z.U = z.s = z.T = z.l = z.P = z.kap = z.Kc = z.Kf = z.qh = z.g = z.dqh = z.qd = NF(0);
        z.rU = z.rS = z.rT = z.rL = z.rP = NF(0);
        s0 = z; s1 = z;
    }
    prefetch(s1, 1);   // iteration m reads the raw values parked in slot (m & 1)
    prefetch(s0, 2);

    NF carry = NF(0);          // over-saturation handed to the layer above (upward sweep of adjust_saturation_profile!)
    bool any_neg = false;      // a negative saturation needs the downward sweep -> slow path
    int idx = 0;               // lowest unsaturated layer (compute_water_table!), 0 = not found yet
    NF wt_new = NF(0);
    NF Sx_new = NF(0);
    if (RICH && do_update) Sx_new = A.bSx[c] + NF(0) * dt;   // surface_excess_water tendency is zero (soil_hydrology.jl:260-267)
    NF G_top = NF(0), infil_top = NF(0);   // LandModel: fluxes coupling the surface to the top soil layer
    NF beta_sm = NF(0);                    // vegetated LandModel: soil moisture limiting factor (plant_available_water.jl:31-35)
    int64_t oout = c;                      // element offset of layer m-2

    // One pipeline iteration: layer m (or the halo above the surface for m = nz+1) enters, the fluxes
    // through face m (heat) and face m-1 (Darcy, needs Kf[m]) are formed, layer m-2 is updated and closed.
    auto iterate = [&](const int m, Stage<NF>& cur, const Stage<NF>& prv) {
        // ---- old content of `cur`: iteration m-2 ----
        const NF U2 = cur.U, s2 = cur.s, T2 = cur.T, l2 = cur.l, Kf2 = cur.Kf;
        // ---- layer m ----
        // (for m = nz + 2 nothing enters the window: the slots keep their old values, which nobody reads)
        NF Tn = cur.T, ln = cur.l, Pn = cur.P, kapn = cur.kap, Kcn = cur.Kc;
        const NF Ur = cur.rU, sr = cur.rS;
        if (m <= nz) {
            if (load_aux) { Tn = cur.rT; ln = cur.rL; Pn = cur.rP; }
            else {
                energy_to_temperature<NF, FAST>(p, Ur, sr, Tn, ln);
                if (RICH) Pn = pressure_head<NF, FAST>(p, sr, wtx, met.zC(m), met.psiz(m));
            }
            kapn = FAST ? thermal_conductivity_fast(p, sr, ln) : thermal_conductivity(p, sr, ln);
            if (need_K) Kcn = cell_conductivity<NF, FAST>(p, sr, ln);
            if (LAND && has_veg(A)) {   // Integral(PAW * root_fraction / dz, dims = 3), accumulated bottom -> top
                const NF paw = FAST ? plant_available_water_fast(A.vp, p, sr, ln) : plant_available_water(A.vp, p, sr, ln);
                beta_sm += FAST ? paw * met.root(m) : paw * met.root(m) / met.dzc(m) * met.dzc(m);
                if (mode == MODE_AUX) A.paw[(int64_t)(m - 1) * ld + c] = paw;
            }
        } else if (m == nz + 1) {   // halo above the surface, built from layer nz (prv)
            Tn = halo_value(A.bc[TRM_BC_TEMPERATURE_TOP].kind, prv.T, bc_T_top, met.dzf(nz + 1), true);
            const NF sh = (RICH || p.sat_halo == TRM_HALO_COPY) ? prv.s : NF(0);
            kapn = FAST ? thermal_conductivity_fast(p, sh, prv.l) : thermal_conductivity(p, sh, prv.l);
            if (RICH) Pn = halo_value(A.bc[TRM_BC_PRESSURE_TOP].kind, prv.P, bc_input(TRM_BC_PRESSURE_TOP), met.dzf(nz + 1), true);
        }
        prefetch(cur, m + 2);
        // ---- lower neighbour of layer m: layer m-1, or the halo below the bottom layer for m = 1 ----
        NF Tp = prv.T, kapp = prv.kap, Pp = prv.P;
        if (m == 1) {
            Tp = halo_value(A.bc[TRM_BC_TEMPERATURE_BOTTOM].kind, Tn, bc_input(TRM_BC_TEMPERATURE_BOTTOM), met.dzf(1), false);
            const NF sb0 = (RICH || p.sat_halo == TRM_HALO_COPY) ? sr : NF(0);
            kapp = FAST ? thermal_conductivity_fast(p, sb0, ln) : thermal_conductivity(p, sb0, ln);
            if (RICH) Pp = halo_value(A.bc[TRM_BC_PRESSURE_BOTTOM].kind, Pn, bc_input(TRM_BC_PRESSURE_BOTTOM), met.dzf(1), false);
        }
        // ---- face conductivity Kf[m], soil_hydrology.jl:249-276 ----
        NF Kfn = NF(0);
        if (need_K) {
            if (m == 1 || m == nz) Kfn = Kcn;                   // Kf[1] = Kc[1], Kf[Nz] = Kc[Nz]
            else if (m < nz) Kfn = Mx::mn(Kcn, prv.Kc);
            else if (m == nz + 1) Kfn = prv.Kf;                 // Kf[Nz+1] = Kf[Nz] ; Kf[Nz+2] is a halo face (0)
            if (write_K && m <= nz + 1) A.Kf[(int64_t)(m - 1) * ld + c] = Kfn;
        }
        // ---- heat flux and head gradient at face m (diffusive_heat_flux, soil_energy.jl:134-149) ----
        NF qhn = NF(0), gn = NF(0);
        if (m <= nz + 1) {
            qhn = -((kapn + kapp) / 2) * ((Tn - Tp) * met.rdzf(m));
            if (RICH) gn = (Pn - Pp) * met.rdzf(m);
        }
        const NF dqhn = qhn - prv.qh;
        // ---- Darcy flux at face m-1 (darcy_flux, soil_hydrology_rre.jl:119-131) ----
        NF qdn = NF(0);
        if (RICH && m >= 2 && m <= nz + 2) {
            const NF g = prv.g;
            NF Kk;
            if (FAST) Kk = Mx::mn(prv.Kf, g < 0 ? Kf2 : Kfn);
            else Kk = (g < 0 ? jmin(Kf2, prv.Kf) : NF(0)) + (g >= 0 ? jmin(prv.Kf, Kfn) : NF(0));
            qdn = -Kk * g;
        }

        // ---- LandModel surface processes, once the top layer is the one about to be updated ----
        if (LAND && m == nz + 2) {
            if (mode == MODE_HEUN2 && !has_veg(A)) { G_top = A.G[c]; infil_top = A.infil[c]; }   // time-n fluxes (heun.jl:63-66)
            else land_surface<NF, FAST>(A, c, RICH, T2, s2, l2, Kf2, met.dzc(nz), beta_sm, mode == MODE_HEUN2, G_top, infil_top);
        }

        if (m >= 3 && m <= nz + 2 && mode != MODE_AUX) {
            // ---- tendencies of layer j = m-2 ----
            const int j = m - 2;
            const int64_t o = oout;
            oout += ld;
            NF tU = -(prv.dqh * met.rdzc(j));                                      // soil_energy.jl:112-131
            NF tS = NF(0);
            if (RICH) {
                const NF dth = -((qdn - prv.qd) * met.rdzc(j)) + NF(0) + p.vwcf;     // soil_hydrology_rre.jl:95-117
                tS = FAST ? dth * p.rpor : dth / p.por;                          // soil_hydrology.jl:222-237
            }
            if (mode == MODE_HEUN1) { A.oTU[o] = tU; if (RICH) A.oTS[o] = tS; }
            if (mode == MODE_HEUN2) {                                           // average_tendencies! heun.jl:27-35
                tU = (A.k1U[o] + tU) / 2;
                if (RICH) tS = (A.k1S[o] + tS) / 2;
            }
            // Flux boundary conditions (compute_z_bcs!, abstract_timestepper.jl:69 )
            if (j == nz) {
                if (LAND) { tU -= G_top / met.dzc(nz); tS -= (-infil_top) / met.dzc(nz); }           // land_model.jl:56-62
                else {
                    if (A.bc[TRM_BC_ENERGY_TOP].kind == TRM_BC_FLUX) tU -= bc_input(TRM_BC_ENERGY_TOP) / met.dzc(nz);
                    if (RICH && A.bc[TRM_BC_SATURATION_TOP].kind == TRM_BC_FLUX) tS -= bc_input(TRM_BC_SATURATION_TOP) / met.dzc(nz);
                }
            }
            if (j == 1) {
                if (A.bc[TRM_BC_ENERGY_BOTTOM].kind == TRM_BC_FLUX) tU += bc_input(TRM_BC_ENERGY_BOTTOM) / met.dzc(1);
                if (RICH && A.bc[TRM_BC_SATURATION_BOTTOM].kind == TRM_BC_FLUX) tS += bc_input(TRM_BC_SATURATION_BOTTOM) / met.dzc(1);
            }
            if (mode == MODE_TEND) { A.oTU[o] = tU; if (RICH) A.oTS[o] = tS; }

            if (do_update) {
                // ---- explicit step, abstract_timestepper.jl:113-141 ----
                const NF Ub = (mode == MODE_HEUN2) ? A.bU[o] : U2;
                const NF Un = Ub + tU * dt;
                NF sn = s2;
                if (RICH) {
                    const NF sb = (mode == MODE_HEUN2) ? A.bS[o] : s2;
                    sn = sb + tS * dt;
                    // ---- adjust_saturation_profile!, upward sweep (soil_hydrology.jl:192-199) ----
                    sn = sn + carry;
                    if (j < nz) {
                        const NF e = Mx::mx(sn - 1, NF(0));
                        sn -= e;
                        carry = FAST ? e * met.dzc(j) * met.rdzc(j + 1) : e * met.dzc(j) / met.dzc(j + 1);
                    }
                    if (sn < 0) any_neg = true;
                }
                if (RICH && any_neg) {
                    // raw values for the slow path below (the downward sweep needs the whole profile)
                    A.yU[o] = Un; A.yS[o] = sn;
                } else {
                    if (RICH) {
                        // downward sweep with no deficit anywhere: sat += max(-sat, 0) (:201-208) is the identity
                        if (!FAST && j >= 2) sn = sn + jmax(-sn, NF(0));
                        if (j == nz) {                                   // top excess -> surface_excess_water (:210-214)
                            const NF e = Mx::mx(sn - 1, NF(0));
                            sn -= e;
                            Sx_new += e * met.dzc(nz);
                        }
                        if (!FAST && j == 1) sn = jmax(sn, NF(0));       // :216
                        A.yS[o] = sn;
                        if (idx == 0 && sn < 1) { idx = j; wt_new = met.zF(j); }   // compute_water_table!, kernel_utils.jl:7-16
                    }
                    A.yU[o] = Un;
                    if (full_closure) {
                        NF Tc, lc;
                        energy_to_temperature<NF, FAST>(p, Un, sn, Tc, lc);
                        A.yT[o] = Tc; A.yL[o] = lc;
                        if (j == nz && A.hio_out) A.hio_out[c] = Tc;
                        // layers below the water table wait for it (written after the sweep)
                        if (RICH && idx != 0) A.yP[o] = pressure_head<NF, FAST>(p, sn, wt_new, met.zC(j), met.psiz(j));
                    }
                }
            }
        }
        // ---- what later iterations need from this one ----
        cur.U = Ur; cur.s = sr; cur.T = Tn; cur.l = ln; cur.P = Pn; cur.kap = kapn; cur.Kc = Kcn;
        cur.Kf = Kfn; cur.qh = qhn; cur.g = gn; cur.dqh = dqhn; cur.qd = qdn;
    };

#pragma unroll 1
    for (int m = 1; m <= nz + 2; m += 2) {
        iterate(m, s1, s0);
        if (m + 1 <= nz + 2) iterate(m + 1, s0, s1);
    }
    if (mode == MODE_AUX || mode == MODE_TEND || !do_update) return;
    if (!RICH) return;

    if (!any_neg) {
        if (idx == 0) { idx = nz + 1; wt_new = met.zF(nz + 1); }   // all saturated: z of the surface (halo cell / fallback give the same)
        A.yWt[c] = wt_new;
        if (A.ySx) A.ySx[c] = Sx_new;
        if (full_closure) {
            // pressure head of the saturated zone below the water table: psi_m(sat >= 1) is a constant
            const NF psat = swrc_inverse<NF, FAST>(p, p.por, p.por);
            int64_t o = c;
#pragma unroll 1
            for (int k = 1; k < idx && k <= nz; ++k, o += ld) {
                const NF z = met.zC(k);
                A.yP[o] = Mx::mx(NF(0), wt_new - z) + psat + met.psiz(k);
            }
        }
        return;
    }
    // ---- slow path: a layer went negative. Downward sweep (soil_hydrology.jl:201-216) top -> bottom on the raw
    //      profile this thread just stored, then water table and closures bottom -> top. ----
    {
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
        idx = 0;
#pragma unroll 1
        for (int k = 1; k <= nz; ++k) if (idx == 0 && A.yS[(int64_t)(k - 1) * ld + c] < 1) idx = k;
        if (idx == 0) idx = nz + 1;
        wt_new = met.zF(idx);
        A.yWt[c] = wt_new;
        if (A.ySx) A.ySx[c] = Sx_new;
        if (full_closure) {
#pragma unroll 1
            for (int k = 1; k <= nz; ++k) {
                const int64_t o = (int64_t)(k - 1) * ld + c;
                NF s = A.yS[o], U = A.yU[o], Tc, lc;
                energy_to_temperature<NF, FAST>(p, U, s, Tc, lc);
                A.yT[o] = Tc; A.yL[o] = lc;
                if (k == nz && A.hio_out) A.hio_out[c] = Tc;
                A.yP[o] = pressure_head<NF, FAST>(p, s, wt_new, met.zC(k), met.psiz(k));
            }
        }
    }
}

// ---- initialisation kernels (not on the hot path) ---------------------------------------------
// initialize!(state, model): Richards closure! (adjust, water table, psi) then the inverse energy
// closure T -> U (soil_coupled.jl:45-54, soil_hydrology_rre.jl:33-47, soil_energy.jl:64-77).
template <class NF, bool FAST>
__global__ void init_kernel(int64_t ncol, int64_t ld, int nz, int richards, const NF* __restrict__ metrics, DevParams<NF> p,
                            NF* U, NF* S, NF* T, NF* Lq, NF* P, NF* Wt, NF* Sx) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const NF* zF = metrics + MET_ZF * MET_STRIDE;
    const NF* zC = metrics + MET_ZC * MET_STRIDE;
    const NF* dzc = metrics + MET_DZC * MET_STRIDE;
    const NF* psiz = metrics + MET_PSIZ * MET_STRIDE;
    if (richards) {
        for (int k = 1; k <= nz - 1; ++k) {
            NF s = S[(int64_t)(k - 1) * ld + c];
            NF e = jmax(s - 1, NF(0));
            S[(int64_t)(k - 1) * ld + c] = s - e;
            S[(int64_t)k * ld + c] += e * dzc[k] / dzc[k + 1];
        }
        for (int k = nz; k >= 2; --k) {
            NF s = S[(int64_t)(k - 1) * ld + c];
            NF d = jmax(-s, NF(0));
            S[(int64_t)(k - 1) * ld + c] = s + d;
            S[(int64_t)(k - 2) * ld + c] -= d * dzc[k] / dzc[k - 1];
        }
        NF st = S[(int64_t)(nz - 1) * ld + c];
        NF e = jmax(st - 1, NF(0));
        S[(int64_t)(nz - 1) * ld + c] = st - e;
        Sx[c] += e * dzc[nz];
        S[c] = jmax(S[c], NF(0));
    }
    int idx = 0;
    for (int k = 1; k <= nz; ++k) if (idx == 0 && S[(int64_t)(k - 1) * ld + c] < 1) idx = k;
    // k = nz + 1 is the halo cell of the saturation field; found or not the result is zF[nz + 1]
    if (idx == 0) idx = nz + 1;
    NF wt = zF[idx];
    Wt[c] = wt;
    for (int k = 1; k <= nz; ++k) {
        const int64_t o = (int64_t)(k - 1) * ld + c;
        NF s = S[o];
        if (richards) P[o] = pressure_head<NF, FAST>(p, s, wt, zC[k], psiz[k]);
        NF Uo, lo;
        temperature_to_energy(p, T[o], s, Uo, lo);
        U[o] = Uo; Lq[o] = lo;
    }
}

// ---- masked columns <-> full ring grid (RingGrids.Field(field, grid) / Field(ring_field, grid),
//      src/grids/column_ring_grid.jl:102-149): scatter / gather by the ring position of every column ----
template <class NF>
__global__ void ring_fill_kernel(int64_t n, NF* __restrict__ out, NF fill) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = fill;
}
template <class NF>
__global__ void ring_scatter_kernel(int64_t ncol, int64_t ld, int nrows, int64_t nring, const int64_t* __restrict__ ring_index,
                                    const NF* __restrict__ field, NF* __restrict__ ring) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const int64_t r = ring_index[c];
    for (int k = 0; k < nrows; ++k) ring[(int64_t)k * nring + r] = field[(int64_t)k * ld + c];
}
template <class NF>
__global__ void ring_gather_kernel(int64_t ncol, int64_t ld, int nrows, int64_t nring, const int64_t* __restrict__ ring_index,
                                   const NF* __restrict__ ring, NF* __restrict__ field) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const int64_t r = ring_index[c];
    for (int k = 0; k < nrows; ++k) field[(int64_t)k * ld + c] = ring[(int64_t)k * nring + r];
}

// ---- diagnostics: column budgets and extrema (one partial per block, reduced by finish_diag) ----
template <class NF>
__global__ void __launch_bounds__(256) diag_kernel(int64_t ncol, int64_t ld, int nz, const NF* __restrict__ metrics, NF por,
                                                   const NF* __restrict__ U, const NF* __restrict__ T, const NF* __restrict__ S,
                                                   const NF* __restrict__ Sx, double* __restrict__ partial) {
    const NF* dzc = metrics + MET_DZC * MET_STRIDE;
    double e = 0, w = 0, tmin = CUDART_INF, tmax = -CUDART_INF, smin = CUDART_INF, smax = -CUDART_INF, nan = 0;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += (int64_t)gridDim.x * blockDim.x) {
        for (int k = 1; k <= nz; ++k) {
            const int64_t o = (int64_t)(k - 1) * ld + c;
            double u = U[o], t = T[o], s = S[o], dz = dzc[k];
            e += u * dz; w += s * (double)por * dz;
            tmin = fmin(tmin, t); tmax = fmax(tmax, t); smin = fmin(smin, s); smax = fmax(smax, s);
            nan += (!isfinite(u)) + (!isfinite(t)) + (!isfinite(s));
        }
        if (Sx) w += (double)Sx[c];
    }
    __shared__ double red[7][256];
    double v[7] = {e, w, tmin, tmax, smin, smax, nan};
#pragma unroll
    for (int i = 0; i < 7; ++i) red[i][threadIdx.x] = v[i];
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
            red[2][threadIdx.x] = fmin(red[2][threadIdx.x], red[2][threadIdx.x + s]);
            red[3][threadIdx.x] = fmax(red[3][threadIdx.x], red[3][threadIdx.x + s]);
            red[4][threadIdx.x] = fmin(red[4][threadIdx.x], red[4][threadIdx.x + s]);
            red[5][threadIdx.x] = fmax(red[5][threadIdx.x], red[5][threadIdx.x + s]);
            red[6][threadIdx.x] += red[6][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int i = 0; i < 7; ++i) partial[(int64_t)blockIdx.x * 8 + i] = red[i][0];
}

// second level of the reduction: one block, every thread folds the partials b = t, t + 256, ... in that fixed order, then a
// fixed tree over the 256 threads -- parallel and still deterministic (the summation order does not depend on timing)
static __global__ void __launch_bounds__(256) finish_diag(int nblocks, const double* __restrict__ partial, double ncol, double* __restrict__ out) {
    double e = 0, w = 0, tmin = CUDART_INF, tmax = -CUDART_INF, smin = CUDART_INF, smax = -CUDART_INF, nan = 0;
    for (int b = threadIdx.x; b < nblocks; b += 256) {
        const double* q = partial + (int64_t)b * 8;
        e += q[0]; w += q[1]; tmin = fmin(tmin, q[2]); tmax = fmax(tmax, q[3]); smin = fmin(smin, q[4]); smax = fmax(smax, q[5]); nan += q[6];
    }
    __shared__ double red[7][256];
    const double v[7] = {e, w, tmin, tmax, smin, smax, nan};
#pragma unroll
    for (int i = 0; i < 7; ++i) red[i][threadIdx.x] = v[i];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
            red[2][threadIdx.x] = fmin(red[2][threadIdx.x], red[2][threadIdx.x + s]);
            red[3][threadIdx.x] = fmax(red[3][threadIdx.x], red[3][threadIdx.x + s]);
            red[4][threadIdx.x] = fmin(red[4][threadIdx.x], red[4][threadIdx.x + s]);
            red[5][threadIdx.x] = fmax(red[5][threadIdx.x], red[5][threadIdx.x + s]);
            red[6][threadIdx.x] += red[6][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 7; ++i) out[i] = red[i][0];
        out[7] = ncol;
    }
}

}  // namespace trm
