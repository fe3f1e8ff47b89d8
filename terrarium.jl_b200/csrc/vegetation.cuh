// vegetation.cuh -- per-column PALADYN vegetation and canopy hydrology of the vegetated LandModel (device code).
//
// Everything here is a per-column scalar computation executed once per column and stage, inside the surface block
// of the stage kernels (land_surface, stage_kernel.cuh). Operation order follows the reference kernels:
//   src/processes/vegetation/{photosynthesis,stomatal_conductance,autotrophic_respiration,carbon_dynamics,
//   vegetation_dynamics,phenology,plant_available_water,root_distribution}.jl and
//   src/processes/surface_hydrology/{canopy_interception/canopy_interception,evapotranspiration/canopy_evapotranspiration}.jl
#pragma once

#include "column_physics.cuh"

namespace trm {

__device__ __forceinline__ float  tlog(float a)  { return logf(a); }
__device__ __forceinline__ double tlog(double a) { return log(a); }

// parameters of VegetationCarbon + PALADYN canopy hydrology converted once to NF (Struct{NF} constructors)
template <class NF>
struct VegParams {
    NF th_fc, th_wp, C_mass;
    NF tau25, Kc25, Ko25, q10_tau, q10_Kc, q10_Ko, alpha_leaf, alpha_a, alpha_C3, cq, k_ext;
    NF T_CO2_high, T_CO2_low, T_photos_high, T_photos_low, theta_r;
    NF g1, g_min, cn_sapwood, cn_root, aws;
    NF SLA, awl, LAI_min, LAI_max, gamma_L, gamma_R, gamma_S, nu_seed, gamma_v;
    NF alpha_int, k_ext_can, w_can_max, tau_w, C_can;
    // derived on the host, used by FAST math only: 1 / (th_fc - th_wp), ln(q10_*), the three temperature stress constants
    NF r_paw_span, ln_q10_tau, ln_q10_Kc, ln_q10_Ko, ts_k1, ts_k2, ts_k3;
};

// 2-D vegetation fields addressed through StageArgs::veg[]: three prognostic variables, then the auxiliaries
enum VegField {
    VF_CVEG = 0, VF_NU, VF_WCAN,
    VF_LAIB, VF_LAI, VF_PHEN, VF_GWCAN, VF_LAMC, VF_AN, VF_RD, VF_GPP, VF_RA, VF_NPP, VF_BETASM,
    VF_ICAN, VF_RCAN, VF_FCAN, VF_RAING, VF_ECAN, VF_TRANSP, VF_COUNT
};

// plant available water of one cell, plant_available_water.jl:64-79 (theta_w = liquid water fraction of the volume)
template <class NF>
__device__ __forceinline__ NF plant_available_water(const VegParams<NF>& v, const DevParams<NF>& p, NF sat, NF liq) {
    const NF thw = sat * p.por * liq;
    return jmax(jmin(NF(1), (thw - v.th_wp) / (v.th_fc - v.th_wp)), NF(0));
}
// same with the reciprocal span (fast math: no division in the layer loop)
template <class NF>
__device__ __forceinline__ NF plant_available_water_fast(const VegParams<NF>& v, const DevParams<NF>& p, NF sat, NF liq) {
    const NF x = (sat * p.por * liq - v.th_wp) * v.r_paw_span;
    return x < NF(0) ? NF(0) : (x > NF(1) ? NF(1) : x);
}

// compute_respiration_assimilation, photosynthesis.jl:212-275 (with compute_kinetic_parameters :86-91, compute_PAR
// :110-113, compute_APAR :124-128, compute_temperature_stress :143-169, compute_assimilation_factors :185-194,
// compute_Vc_max :208-211 -- called with APAR --, compute_Rd :225-228, compute_Ag :241-248)
template <class NF, bool FAST>
__device__ __forceinline__ void photosynthesis(const VegParams<NF>& v, NF T_air, NF swdown, NF pres, NF co2, NF LAI, NF fapar /* 1 - exp(-k_ext LAI) */, NF lamc, NF beta_sm,
                                               NF& Rd, NF& An) {
    const NF pres_O2 = NF(0.209) * pres;          // physics_utils.jl:16-20
    const NF pres_a = co2 * NF(1.0e-6) * pres;    // physics_utils.jl:27-30
    Rd = 0; An = 0;
    if (!(swdown > 0 && T_air > NF(-3.0))) return;
    const NF ex = (T_air - NF(25.0)) * NF(0.1);
    // fast math: q10^ex = exp(ex ln q10) with the logarithms taken once on the host
    const NF tau = v.tau25 * (FAST ? xexp<NF, FAST>(ex * v.ln_q10_tau) : tpow(v.q10_tau, ex));
    const NF Kc = v.Kc25 * (FAST ? xexp<NF, FAST>(ex * v.ln_q10_Kc) : tpow(v.q10_Kc, ex));
    const NF Ko = v.Ko25 * (FAST ? xexp<NF, FAST>(ex * v.ln_q10_Ko) : tpow(v.q10_Ko, ex));
    const NF Gs = dv<NF, FAST>(pres_O2, NF(2.0) * tau);
    if (!(LAI > 0)) return;
    const NF PAR = NF(0.5) * swdown * (NF(1.0) - v.alpha_leaf) * v.cq;
    const NF APAR = v.alpha_a * PAR * fapar;
    const NF pres_i = lamc * pres_a;
    const NF k1 = FAST ? v.ts_k1 : NF(2.0) * tlog(NF(1.0) / NF(0.99) - NF(1.0)) / (v.T_CO2_low - v.T_photos_low);
    const NF k2 = FAST ? v.ts_k2 : NF(0.5) * (v.T_CO2_low + v.T_photos_low);
    const NF k3 = FAST ? v.ts_k3 : tlog(NF(0.99) / NF(0.01)) / (v.T_CO2_high - v.T_photos_high);
    NF T_stress = 0;
    if (v.T_CO2_low < T_air && T_air < v.T_CO2_high) {
        const NF low = dv<NF, FAST>(NF(1.0), NF(1.0) + xexp<NF, FAST>(k1 * (k2 - T_air)));
        const NF high = NF(1.0) - NF(0.01) * xexp<NF, FAST>(k3 * (T_air - v.T_photos_high));
        T_stress = low * high;
    }
    const NF c_1 = dv<NF, FAST>(v.alpha_C3 * T_stress * v.C_mass * (pres_i - Gs), pres_i + NF(2.0) * Gs);
    const NF c_2 = dv<NF, FAST>(pres_i - Gs, pres_i + Kc * (NF(1.0) + dv<NF, FAST>(pres_O2, Ko)));
    const NF Vc_max = dv<NF, FAST>(c_1 * APAR * (pres_i + Kc * (NF(1.0) + dv<NF, FAST>(pres_O2, Ko))), pres_i - Gs);
    Rd = v.alpha_C3 * Vc_max * beta_sm;
    const NF JE = c_1 * APAR, JC = c_2 * Vc_max;
    const NF sJ = JE + JC;
    const NF Ag = dv<NF, FAST>(sJ - M<NF, FAST>::sqrt_(sJ * sJ - NF(4) * v.theta_r * JE * JC), NF(2) * v.theta_r) * beta_sm;
    An = Ag - Rd;
}

// compute_lambda_NPP, carbon_dynamics.jl:64-74
template <class NF>
__device__ __forceinline__ NF lambda_NPP(const VegParams<NF>& v, NF LAI_b) {
    return LAI_b < v.LAI_min ? NF(0) : (LAI_b <= v.LAI_max ? (LAI_b - v.LAI_min) / (v.LAI_max - v.LAI_min) : NF(1.0));
}

}  // namespace trm
