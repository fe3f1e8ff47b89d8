// terrarium_oracle.cpp -- CPU restatement of Terrarium.jl's per-column land time-step.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under oracle/ is part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or
// call it, and only as the checker (or as the timed CPU baseline), never as a fallback.
//
// PARITY PINNING STATUS: "partially pinned".  The reference is pure Julia and neither `julia` nor
// its un-vendored dependencies (Oceananigans 0.100-0.106, FreezeCurves 0.9; Project.toml:31-48,
// there is no Manifest.toml) exist in this environment, so the reference cannot be executed.
// This file restates the algorithm from the reference sources (file:line cited per function,
// paths relative to /root/reference) and is pinned against every known-answer test the reference
// holds for the path (tests/test_golden.py; vegetation / canopy: test/vegetation/*.jl,
// test/surface_hydrology/canopy_*_tests.jl, tests/test_vegetation.py, which also holds an independent numpy
// restatement of the per-column vegetation formulas; the soil energy + Richards step, ForwardEuler and Heun, and the
// bare-ground LandModel step are cross-checked against a second, separately written numpy restatement in tests/test_numpy_cross_check.py).  Semantics that live
// in the absent dependencies are restated from their published behaviour and are marked [OCN]
// (Oceananigans) or [FC] (FreezeCurves); each is listed in DESIGN.md as "unpinned at rounding
// level".
//
// Structure: "reference-structured" -- one loop nest per reference kernel launch, in the reference
// order, tendencies / face conductivities / halo planes materialised in memory exactly like the
// Julia code does (~29 array passes per Euler step).  This is deliberate: the same code is timed as
// the CPU baseline ("restated reference CPU path (C++/OpenMP, N cores)").
//
// Layout: [k][column] (column fastest, like Oceananigans' parent arrays), k = 0 and k = Nz+1 are
// the z-halo planes of centre fields, faces are k = 1..Nz+1 with halo planes 0 and Nz+2.
// k = 1 is the BOTTOM cell (docs/src/introduction/numerical_core.md:21-22).

#include "../include/terrarium_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& m) { g_err = m; return code; }

constexpr int64_t CHUNK = 2048;  // columns per OpenMP work item

// Julia `min` / `max` for floats (NaN propagating).
template <class T> inline T jmin(T a, T b) { return (a != a || b != b) ? std::numeric_limits<T>::quiet_NaN() : (b < a ? b : a); }
template <class T> inline T jmax(T a, T b) { return (a != a || b != b) ? std::numeric_limits<T>::quiet_NaN() : (a < b ? b : a); }

// Julia Base.Math.pow_body(x::Float64, n::Integer) for n = 4 (base/math.jl): compensated
// power-by-squaring.  Float32 goes through Float64 power_by_squaring and rounds once.
inline double two_mul_hi(double a, double b, double& lo) { double hi = a * b; lo = std::fma(a, b, -hi); return hi; }
inline double jl_pow4(double x) {
    double y = 1.0, xnlo = 0.0, ynlo = 0.0;
    int n = 4;
    while (n > 1) {
        if (n & 1) { double err = std::fma(y, xnlo, x * ynlo); double lo; y = two_mul_hi(x, y, lo); ynlo = lo + err; }
        double err = x * 2 * xnlo; double lo; x = two_mul_hi(x, x, lo); xnlo = lo + err;
        n >>= 1;
    }
    double err = std::fma(y, xnlo, x * ynlo);
    return (std::isfinite(x) && std::isfinite(err)) ? std::fma(x, y, err) : x * y;
}
inline float jl_pow4(float x) { double d = (double)x; double d2 = d * d; return (float)(d2 * d2); }

struct InputSource {
    int kind = TRM_SRC_CONST;
    double cval = 0.0;
    double period = 1.0, lo = -INFINITY, hi = INFINITY;
    int nt = 0;
    std::vector<double> times;
    double pair_t0 = 0.0;   // TRM_SRC_FIELD_PAIR: start time of the step the pair was handed over for
};

template <class NF>
struct State {  // one StateVariables instance (src/state_variables.jl:16-54)
    int nz = 0; int64_t nc = 0;
    // centre fields with z-halos: (nz+2) x nc
    std::vector<NF> U, T, liq, sat, psi, tendU, tendsat;
    // z-face field with halos: (nz+3) x nc   (faces 1..nz+1)
    std::vector<NF> Kf;
    // 2-D
    std::vector<NF> Sx, tendSx, wt, Ts, tendTs, G, SWup, LWup, Rnet, Hs, Hl, Egnd, infil, runoff;
    // vegetated LandModel: prognostic 2-D (+ tendencies), auxiliary 2-D, plant available water (nz+2) x nc
    std::vector<NF> Cveg, nu, wcan, tendCveg, tendnu, tendwcan;
    std::vector<NF> LAIb, LAI, phen, gwcan, lamc, An, Rd, GPP, Ra, NPP, betasm, Ican, Rcan, fcan, raing, Ecan, transp;
    std::vector<NF> PAW;
    std::vector<std::vector<NF>> in;  // materialised input fields [TRM_IN_COUNT][nc]
    NF time = 0; int64_t iteration = 0;

    void alloc(int nz_, int64_t nc_, bool land, bool veg = false) {
        nz = nz_; nc = nc_;
        size_t c3 = (size_t)(nz + 2) * nc, f3 = (size_t)(nz + 3) * nc;
        for (auto* v : {&U, &T, &liq, &sat, &psi, &tendU, &tendsat}) v->assign(c3, NF(0));
        Kf.assign(f3, NF(0));
        for (auto* v : {&Sx, &tendSx, &wt}) v->assign(nc, NF(0));
        if (land) for (auto* v : {&Ts, &tendTs, &G, &SWup, &LWup, &Rnet, &Hs, &Hl, &Egnd, &infil, &runoff}) v->assign(nc, NF(0));
        if (veg) {
            for (auto* v : {&Cveg, &nu, &wcan, &tendCveg, &tendnu, &tendwcan, &LAIb, &LAI, &phen, &gwcan, &lamc, &An, &Rd, &GPP, &Ra, &NPP,
                            &betasm, &Ican, &Rcan, &fcan, &raing, &Ecan, &transp}) v->assign(nc, NF(0));
            PAW.assign(c3, NF(0));
        }
        in.assign(TRM_IN_COUNT, std::vector<NF>());
    }
    inline size_t ix(int k, int64_t c) const { return (size_t)k * nc + c; }
};

template <class NF>
struct Oracle {
    trm_config cfg{};
    int nz = 0; int64_t nc = 0;
    bool land = false, richards = false, heun = false, veg = false;
    trm_params vp{};          // vegetation parameters are converted to NF where they are used (Struct{NF} fields)
    std::vector<NF> rootf;    // static root fraction per layer (root_distribution.jl:47-56), index 1..nz
    // grid metrics in NF  [OCN] generate_coordinate: halo faces extend with the edge spacing,
    // centres are face midpoints, dzf are centre differences.
    std::vector<NF> zF, zC, dzc, dzf, rdzc, rdzf;  // indices as in the reference (1-based + halos)
    // parameters in NF
    NF por, org, L, sqk[5], hc[5], Ksat, vg_alpha, vg_n, bc_psis, bc_lambda, theta_res, Omega, vwcf;
    NF rho_a, c_a, Llg, Tref, sigma, eps_mw, albedo, emis, kappa_skin, C_h, Vmin, tau_r, beta;
    State<NF> st, stage;
    std::vector<InputSource> src;
    std::vector<std::vector<NF>> src_field, src_mean, src_amp, src_phase, src_table;
    bool initialized = false;
    int64_t passes = 0;  // 3-D array passes executed (reported for the baseline description)

    // ---------------------------------------------------------------- construction
    int setup(const trm_config& c) {
        cfg = c; nz = c.nz; nc = c.ncol;
        land = c.model == TRM_MODEL_LAND; richards = c.hydrology == TRM_RICHARDS; heun = c.timestepper == TRM_HEUN;
        veg = land && c.vegetation == TRM_VEG_CARBON;
        const trm_params& p = c.params;
        vp = p;
        // ---- grid (column_grid.jl:30-31: z_coords converted to NF, then [OCN] metrics in NF)
        zF.assign(nz + 3, NF(0)); zC.assign(nz + 2, NF(0)); dzc.assign(nz + 2, NF(0)); dzf.assign(nz + 3, NF(0));
        rdzc.assign(nz + 2, NF(0)); rdzf.assign(nz + 3, NF(0));
        for (int k = 1; k <= nz + 1; ++k) zF[k] = (NF)c.z_faces[k - 1];
        zF[0] = zF[1] - (zF[2] - zF[1]);
        zF[nz + 2] = zF[nz + 1] + (zF[nz + 1] - zF[nz]);
        for (int k = 0; k <= nz + 1; ++k) { zC[k] = (zF[k + 1] + zF[k]) / 2; dzc[k] = zF[k + 1] - zF[k]; rdzc[k] = 1 / dzc[k]; }
        for (int k = 1; k <= nz + 1; ++k) { dzf[k] = zC[k] - zC[k - 1]; rdzf[k] = 1 / dzf[k]; }
        // ---- parameters
        NF por_m = (NF)p.mineral_porosity, por_o = (NF)p.organic_porosity;
        // homogeneous_strat.jl:34-45 : organic = rho_soc / ((1 - por_o) * rho_org)
        org = (NF)p.rho_soc / ((1 - por_o) * (NF)p.rho_org);
        // homogeneous_strat.jl:52-61 : (1 - organic) * por_m + organic * por_o
        por = (1 - org) * por_m + org * por_o;
        L = (NF)p.rho_w * (NF)p.Lsl;                       // soil_energy_closures.jl:76,113
        for (int i = 0; i < 5; ++i) { sqk[i] = std::sqrt((NF)p.kappa[i]); hc[i] = (NF)p.heatcap[i]; }
        Ksat = (NF)p.K_sat; vg_alpha = (NF)p.vg_alpha; vg_n = (NF)p.vg_n; bc_psis = (NF)p.bc_psis; bc_lambda = (NF)p.bc_lambda;
        theta_res = (NF)p.theta_res; Omega = (NF)p.impedance; vwcf = (NF)p.vwc_forcing;
        rho_a = (NF)p.rho_a; c_a = (NF)p.c_a; Llg = (NF)p.Llg; Tref = (NF)p.Tref; sigma = (NF)p.sigma; eps_mw = (NF)p.eps_mw;
        albedo = (NF)p.albedo; emis = (NF)p.emissivity; kappa_skin = (NF)p.kappa_skin; C_h = (NF)p.C_h; Vmin = (NF)p.min_windspeed;
        tau_r = (NF)p.tau_r; beta = (NF)p.evap_beta;
        st.alloc(nz, nc, land, veg);
        if (veg) {
            // root_fraction(rootdist, grid, ...), root_distribution.jl:47-56: density at the cell centres times the
            // layer thickness, normalised by its sum over the column
            rootf.assign(nz + 2, NF(0));
            NF a = (NF)p.root_a, b = (NF)p.root_b, sum = 0;
            for (int k = 1; k <= nz; ++k) { rootf[k] = NF(0.5) * (a * std::exp(a * zC[k]) + b * std::exp(b * zC[k])) * dzc[k]; sum += rootf[k]; }
            for (int k = 1; k <= nz; ++k) rootf[k] = rootf[k] / sum;
        }
        src.assign(TRM_IN_COUNT, InputSource());
        src_field.assign(TRM_IN_COUNT, {}); src_mean.assign(TRM_IN_COUNT, {}); src_amp.assign(TRM_IN_COUNT, {});
        src_phase.assign(TRM_IN_COUNT, {}); src_table.assign(TRM_IN_COUNT, {});
        // input defaults, prescribed_atmosphere.jl:89-99,147-149,192-195,220-224,10-14
        src[TRM_IN_AIR_TEMPERATURE].cval = 10; src[TRM_IN_AIR_PRESSURE].cval = 101325; src[TRM_IN_WINDSPEED].cval = 0.1;
        src[TRM_IN_SPECIFIC_HUMIDITY].cval = 1.0e-3; src[TRM_IN_SHORTWAVE_DOWN].cval = 300; src[TRM_IN_LONGWAVE_DOWN].cval = 50;
        src[TRM_IN_DAYTIME_LENGTH].cval = 12; src[TRM_IN_CO2].cval = 380;
        return TRM_OK;
    }

    // ---------------------------------------------------------------- inputs
    // update_inputs! (state_variables.jl:154-162, input_sources.jl:165-171) + evaluation of
    // function-valued boundary conditions at clock time t [OCN getbc(f, x, t)].
    NF eval_input(int id, int64_t c, NF t) const {
        const InputSource& s = src[id];
        switch (s.kind) {
            case TRM_SRC_CONST: return (NF)s.cval;
            case TRM_SRC_FIELD: return src_field[id][c];
            // host-evaluated function of time handed over for one step: values at the step's start time `pair_t0` and at
            // t + dt (the stage clock of Heun, heun.jl:53)
            case TRM_SRC_FIELD_PAIR: return (double)t > s.pair_t0 ? src_amp[id][c] : src_field[id][c];
            case TRM_SRC_SINUSOID: {
                // examples/simulations/soil_heat_global.jl:79-88: `2pi * t / period - lon` promotes to
                // Float64 in Julia whatever NF is; the result is rounded when stored in the NF field.
                double ph = 6.283185307179586 * (double)t / s.period - (double)src_phase[id][c];
                double v = (double)src_mean[id][c] + (double)src_amp[id][c] * std::sin(ph);
                v = std::min(std::max(v, s.lo), s.hi);
                return (NF)v;
            }
            case TRM_SRC_TABLE: {
                // [OCN] FieldTimeSeries[Time(t)]: v2*n + v1*(1-n), n = (t-t1)/(t2-t1); flat outside
                // the table (rule of ext/TerrariumRastersExt/TerrariumRastersExt.jl:104-120).
                const std::vector<double>& tt = s.times; const std::vector<NF>& v = src_table[id];
                double td = (double)t;
                if (td <= tt.front()) return v[c];
                if (td >= tt.back()) return v[(size_t)(s.nt - 1) * nc + c];
                int n2 = (int)(std::upper_bound(tt.begin(), tt.end(), td) - tt.begin()); int n1 = n2 - 1;
                NF frac = (NF)((td - tt[n1]) / (tt[n2] - tt[n1]));
                return v[(size_t)n2 * nc + c] * frac + v[(size_t)n1 * nc + c] * (1 - frac);
            }
            case TRM_SRC_RASTER: {
                // update_from_raster!, ext/TerrariumRastersExt/TerrariumRastersExt.jl:96-121: searchsorted on the time
                // axis; between nodes x1 + eps * (x2 - x1) / dt (eps, dt Float64 seconds), the node value on a node,
                // flat beyond either end
                const std::vector<double>& tt = s.times; const std::vector<NF>& v = src_table[id];
                double td = (double)t;
                int right = (int)(std::lower_bound(tt.begin(), tt.end(), td) - tt.begin()) + 1;   // first(searchsorted), 1-based
                int left = (int)(std::upper_bound(tt.begin(), tt.end(), td) - tt.begin());        // last(searchsorted), 1-based
                if (left >= 1 && right <= s.nt) {
                    NF x1 = v[(size_t)(left - 1) * nc + c], x2 = v[(size_t)(right - 1) * nc + c];
                    double dtt = tt[right - 1] - tt[left - 1], e = td - tt[left - 1];
                    return dtt > 0 ? (NF)((double)x1 + e * (double)(x2 - x1) / dtt) : x2;
                }
                return v[(size_t)(std::min(right, s.nt) - 1) * nc + c];
            }
        }
        return NF(0);
    }
    bool input_used(int id) const {
        if (land && id >= TRM_IN_AIR_TEMPERATURE) return true;
        for (int s = 0; s < TRM_BC_NSLOTS; ++s) if (cfg.bc[s].kind != TRM_BC_DEFAULT && cfg.bc[s].input == id) return true;
        return false;
    }
    NF t_inputs = 0;   // clock time of the last update_inputs! on the model state
    int get_input(int id, void* host, int64_t count) {
        if (count != nc) return fail(TRM_ERR_INVALID, "get_input: count != ncol");
        NF* h = (NF*)host;
        for (int64_t c = 0; c < nc; ++c) h[c] = eval_input(id, c, t_inputs);
        return TRM_OK;
    }
    void update_inputs(State<NF>& s) {
        if (&s == &st) t_inputs = s.time;
        for (int id = 0; id < TRM_IN_COUNT; ++id) {
            if (!input_used(id)) continue;
            auto& f = s.in[id]; if ((int64_t)f.size() != nc) f.assign(nc, NF(0));
            NF t = s.time;
#pragma omp parallel for schedule(static)
            for (int64_t c = 0; c < nc; ++c) f[c] = eval_input(id, c, t);
        }
    }

    // ---------------------------------------------------------------- soil constituents
    struct Fr { NF water, ice, air, mineral, organic; };
    // soil_volume.jl:52-67,103-107
    inline Fr fractions(NF sat, NF liq) const {
        NF wi = sat * por; Fr f;
        f.water = wi * liq; f.ice = wi * (1 - liq); f.air = (1 - sat) * por;
        NF solid = 1 - por; f.organic = solid * org; f.mineral = solid * (1 - org);
        return f;
    }
    // soil_thermal_properties.jl:90-123 (InverseQuadratic; sum order water, ice, air, mineral, organic)
    inline NF conductivity(NF sat, NF liq) const {
        Fr f = fractions(sat, liq);
        NF s = sqk[0] * f.water + sqk[1] * f.ice + sqk[2] * f.air + sqk[3] * f.mineral + sqk[4] * f.organic;
        return s * s;
    }
    inline NF heat_capacity(NF sat, NF liq) const {
        Fr f = fractions(sat, liq);
        return hc[0] * f.water + hc[1] * f.ice + hc[2] * f.air + hc[3] * f.mineral + hc[4] * f.organic;
    }
    // soil_hydraulic_properties.jl:170-221 (real branch of the complex formula: theta_w/theta_sat in [0,1])
    inline NF cell_K(NF sat, NF liq) const {
        Fr f = fractions(sat, liq);
        if (cfg.unsat_k == TRM_UNSATK_LINEAR) {
            NF thsat = f.water + f.ice + f.air;
            return Ksat * f.water / thsat;
        }
        NF n = vg_n;
        NF I_ice = std::pow(NF(10), -Omega * (1 - liq));
        NF x = f.water / por;
        NF inner = 1 - std::pow(x, n / (n + 1));
        NF a = 1 - std::pow(inner, (n - 1) / n);
        return std::fabs(Ksat * I_ice * std::sqrt(x) * (a * a));
    }
    // [FC] inverse soil water retention curve psi_m(theta; theta_sat)
    inline NF swrc_inv(NF theta, NF thsat) const {
        if (cfg.swrc == TRM_SWRC_VANGENUCHTEN) {
            NF n = vg_n, m = 1 - 1 / n;
            if (!(theta < thsat)) return NF(0);
            NF se = (theta - theta_res) / (thsat - theta_res);
            return -1 / vg_alpha * std::pow(std::pow(se, -1 / m) - NF(1), 1 / n);
        }
        if (!(theta < thsat)) return -bc_psis;
        NF se = (theta - theta_res) / (thsat - theta_res);
        return -bc_psis * std::pow(se, -1 / bc_lambda);
    }

    // ---------------------------------------------------------------- halos [OCN] (Appendix B.4)
    NF bc_value(const State<NF>& s, int slot, int64_t c) const { return s.in[cfg.bc[slot].input][c]; }
    void fill_halo_field(State<NF>& s, std::vector<NF>& f, int slot_top, int slot_bottom) {
        int kt = slot_top >= 0 ? cfg.bc[slot_top].kind : TRM_BC_DEFAULT;
        int kb = slot_bottom >= 0 ? cfg.bc[slot_bottom].kind : TRM_BC_DEFAULT;
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF ct = f[s.ix(nz, c)], cb = f[s.ix(1, c)];
            // top: c[N+1] = c[N] + grad * dzf[N+1]; Value: grad = (v - c[N]) / (dzf/2)
            if (kt == TRM_BC_VALUE) { NF D = dzf[nz + 1]; NF g = (bc_value(s, slot_top, c) - ct) / (D / 2); f[s.ix(nz + 1, c)] = ct + g * D; }
            else if (kt == TRM_BC_GRADIENT) { NF D = dzf[nz + 1]; f[s.ix(nz + 1, c)] = ct + bc_value(s, slot_top, c) * D; }
            else f[s.ix(nz + 1, c)] = ct;
            // bottom: c[0] = c[1] - grad * dzf[1]; Value: grad = (c[1] - v) / (dzf/2)
            if (kb == TRM_BC_VALUE) { NF D = dzf[1]; NF g = (cb - bc_value(s, slot_bottom, c)) / (D / 2); f[s.ix(0, c)] = cb + g * (-D); }
            else if (kb == TRM_BC_GRADIENT) { NF D = dzf[1]; f[s.ix(0, c)] = cb + bc_value(s, slot_bottom, c) * (-D); }
            else f[s.ix(0, c)] = cb;
        }
    }
    // state_variables.jl:85-100: every prognostic field, then every closure field.
    void fill_halo_regions(State<NF>& s) {
        fill_halo_field(s, s.U, -1, -1);  // Flux / default BCs on internal_energy: no-gradient halo
        if (richards) {
            fill_halo_field(s, s.sat, -1, -1);
            fill_halo_field(s, s.psi, TRM_BC_PRESSURE_TOP, TRM_BC_PRESSURE_BOTTOM);
        } else if (cfg.sat_halo == TRM_HALO_COPY) {
            fill_halo_field(s, s.sat, -1, -1);  // switch B.6: constant initializer filled the halos
        }
        fill_halo_field(s, s.T, TRM_BC_TEMPERATURE_TOP, TRM_BC_TEMPERATURE_BOTTOM);
        fill_halo_field(s, s.liq, -1, -1);
    }

    // ---------------------------------------------------------------- kernels, reference order
    void reset_tendencies(State<NF>& s) {  // state_variables.jl:127-136
        std::fill(s.tendU.begin(), s.tendU.end(), NF(0)); passes++;
        if (richards) { std::fill(s.tendsat.begin(), s.tendsat.end(), NF(0)); std::fill(s.tendSx.begin(), s.tendSx.end(), NF(0)); passes++; }
        if (land) std::fill(s.tendTs.begin(), s.tendTs.end(), NF(0));
    }
    // compute_hydraulics! soil_hydrology.jl:145-163, kernel :249-276
    void compute_hydraulics(State<NF>& s) {
#pragma omp parallel for schedule(static)
        for (int64_t c0 = 0; c0 < nc; c0 += CHUNK) {
            int64_t c1 = std::min(nc, c0 + CHUNK);
            for (int k = 1; k <= nz; ++k)
                for (int64_t c = c0; c < c1; ++c) {
                    if (k <= 1) s.Kf[s.ix(k, c)] = cell_K(s.sat[s.ix(1, c)], s.liq[s.ix(1, c)]);
                    else if (k >= nz) { NF K = cell_K(s.sat[s.ix(nz, c)], s.liq[s.ix(nz, c)]); s.Kf[s.ix(k, c)] = K; s.Kf[s.ix(k + 1, c)] = K; }
                    else s.Kf[s.ix(k, c)] = jmin(cell_K(s.sat[s.ix(k, c)], s.liq[s.ix(k, c)]), cell_K(s.sat[s.ix(k - 1, c)], s.liq[s.ix(k - 1, c)]));
                }
        }
        passes += 3;
    }
    // darcy_flux soil_hydrology_rre.jl:119-131
    inline NF darcy(const State<NF>& s, int k, int64_t c) const {
        NF g = (s.psi[s.ix(k, c)] - s.psi[s.ix(k - 1, c)]) * rdzf[k];
        NF Kk = (g < 0 ? jmin(s.Kf[s.ix(k - 1, c)], s.Kf[s.ix(k, c)]) : NF(0)) + (g >= 0 ? jmin(s.Kf[s.ix(k, c)], s.Kf[s.ix(k + 1, c)]) : NF(0));
        return -Kk * g;
    }
    // compute_tendencies! (Richards) soil_hydrology_rre.jl:76-93,150-162; soil_hydrology.jl:222-237
    void richards_tendency(State<NF>& s) {
#pragma omp parallel for schedule(static)
        for (int64_t c0 = 0; c0 < nc; c0 += CHUNK) {
            int64_t c1 = std::min(nc, c0 + CHUNK);
            for (int k = 1; k <= nz; ++k)
                for (int64_t c = c0; c < c1; ++c) {
                    NF div = (darcy(s, k + 1, c) - darcy(s, k, c)) * rdzc[k];
                    NF dth = -div + NF(0) + vwcf;          // ET forcing is `nothing` -> zero (soil_coupled.jl:86)
                    s.tendsat[s.ix(k, c)] += dth / por;
                    if (k == 1) s.tendSx[c] += NF(0);       // runoff is `nothing` -> zero(S)
                }
        }
        passes += 4;
    }
    // compute_tendencies! (energy) soil_energy.jl:83-149
    inline NF heat_flux(const State<NF>& s, int k, int64_t c) const {
        NF kap = (conductivity(s.sat[s.ix(k, c)], s.liq[s.ix(k, c)]) + conductivity(s.sat[s.ix(k - 1, c)], s.liq[s.ix(k - 1, c)])) / 2;
        return -kap * ((s.T[s.ix(k, c)] - s.T[s.ix(k - 1, c)]) * rdzf[k]);
    }
    void energy_tendency(State<NF>& s) {
#pragma omp parallel for schedule(static)
        for (int64_t c0 = 0; c0 < nc; c0 += CHUNK) {
            int64_t c1 = std::min(nc, c0 + CHUNK);
            for (int k = 1; k <= nz; ++k)
                for (int64_t c = c0; c < c1; ++c)
                    s.tendU[s.ix(k, c)] += -((heat_flux(s, k + 1, c) - heat_flux(s, k, c)) * rdzc[k]);
        }
        passes += 5;
    }

    // ---- LandModel auxiliaries --------------------------------------------------------------
    // aerodynamic_resistance prescribed_atmosphere.jl:110-116,137 ; the literal 1.0e-6 promotes the
    // expression to Float64 for NF = Float32 -> computed in double here.
    inline double r_a(const State<NF>& s, int64_t c) const {
        NF V = jmax(s.in[TRM_IN_WINDSPEED][c], Vmin);
        double Va = std::max((double)V, 1.0e-6);
        return 1.0 / ((double)C_h * Va);
    }
    // physics_utils.jl:54-73
    inline NF e_sat(NF T) const {
        return T <= 0 ? NF(611.0) * std::exp(NF(22.46) * T / (T + NF(272.62))) : NF(611.0) * std::exp(NF(17.62) * T / (T + NF(243.12)));
    }
    // compute_humidity_vpd prescribed_atmosphere.jl:160-182 + physical_constants.jl:83-97 + physics_utils.jl:38
    inline NF humidity_vpd(const State<NF>& s, int64_t c, NF Tsurf) const {
        NF q = s.in[TRM_IN_SPECIFIC_HUMIDITY][c], p = s.in[TRM_IN_AIR_PRESSURE][c];
        NF es = e_sat(Tsurf);
        NF ea = q * p / (eps_mw + (1 - eps_mw) * q);
        NF vpd = jmax(es - ea, NF(0.1));
        return eps_mw * vpd / p;
    }
    // ground_evaporation_resistance_factor, ground_resistance_factor.jl:6-11 (constant) / :32-57 (Lee & Pielke 1992)
    inline NF ground_resistance(const State<NF>& s, int64_t c) const {
        if (cfg.ground_resistance != TRM_GROUND_RES_SOIL_MOISTURE) return beta;
        Fr f = fractions(s.sat[s.ix(nz, c)], s.liq[s.ix(nz, c)]);
        NF fc = (NF)vp.field_capacity;
        if (f.water < fc) { NF d = 1 - std::cos(NF(3.141592653589793) * f.water / fc); return d * d / 4; }
        return NF(1);
    }
    // bare_ground_evaporation.jl:49-62
    void compute_evaporation(State<NF>& s) {
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF Tsurf = cfg.skin == TRM_SKIN_PRESCRIBED ? s.in[TRM_IN_SKIN_TEMPERATURE][c] : s.Ts[c];
            NF dq = humidity_vpd(s, c, Tsurf);
            s.Egnd[c] = (NF)((double)(ground_resistance(s, c) * dq) / r_a(s, c));
        }
    }
    // direct_surface_runoff.jl:87-117
    void compute_runoff(State<NF>& s) {
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF rain = veg ? s.raing[c] : s.in[TRM_IN_RAINFALL][c];  // rainfall_ground aliases rainfall without canopy (canopy_interception.jl:11-15)
            // surface_excess_water(i, j, grid, fields, hydrology): the prognostic field under RichardsEq
            // (soil_hydrology_rre.jl:28), identically zero for immobile soil water (soil_hydrology.jl:138)
            NF S = richards ? s.Sx[c] : NF(0), Kt = s.Kf[s.ix(nz, c)], sat_top = s.sat[s.ix(nz, c)];
            NF drain, inf;
            if (S > 0) { drain = jmax(S, NF(0)) / tau_r; inf = (sat_top < 1) ? jmin(drain, Kt) : NF(0); }
            else { drain = 0; inf = (sat_top < 1) ? jmin(rain, Kt) : NF(0); }
            s.infil[c] = inf;
            s.runoff[c] = rain + drain - inf;
        }
    }
    // compute_surface_energy_fluxes! surface_energy_balance.jl:119-144 (one evaluation)
    inline void seb_fluxes(State<NF>& s, int64_t c) const {
        NF SWd = s.in[TRM_IN_SHORTWAVE_DOWN][c], LWd = s.in[TRM_IN_LONGWAVE_DOWN][c];
        NF Tsurf = cfg.skin == TRM_SKIN_PRESCRIBED ? s.in[TRM_IN_SKIN_TEMPERATURE][c] : s.Ts[c];
        // albedo / emissivity: ConstantAlbedo parameters or the PrescribedAlbedo inputs (albedo.jl:7-44, abstract_types.jl:120-131)
        const bool alb_in = cfg.albedo_kind == TRM_ALBEDO_PRESCRIBED;
        NF alb = alb_in ? s.in[TRM_IN_ALBEDO][c] : albedo, emi = alb_in ? s.in[TRM_IN_EMISSIVITY][c] : emis;
        NF swu, lwu;
        if (cfg.radiative == TRM_RADIATIVE_PRESCRIBED) {        // radiative_fluxes.jl:46-50
            swu = s.in[TRM_IN_SHORTWAVE_UP][c]; lwu = s.in[TRM_IN_LONGWAVE_UP][c];
        } else {
            swu = alb * SWd;                                     // radiative_fluxes.jl:85-88
            NF TK = Tsurf + Tref;
            lwu = emi * sigma * jl_pow4(TK) + (1 - emi) * LWd;   // radiative_fluxes.jl:95-100, physical_constants.jl:67
        }
        s.SWup[c] = swu; s.LWup[c] = lwu;
        NF rnet = swu - SWd + lwu - LWd;                         // radiative_fluxes.jl:196-209
        s.Rnet[c] = rnet;
        double ra = r_a(s, c);
        NF Ta = s.in[TRM_IN_AIR_TEMPERATURE][c];
        NF hs = (NF)((double)(c_a * rho_a) * ((double)(Tsurf - Ta) / ra));  // turbulent_fluxes.jl:85-105
        // turbulent_fluxes.jl:137-150 (coupled to ET): surface_humidity_flux = E_gnd (+ E_can + T_can, canopy_evapotranspiration.jl:75-80)
        NF Qh = veg ? s.Egnd[c] + s.Ecan[c] + s.transp[c] : s.Egnd[c];
        NF hl = Llg * rho_a * Qh;
        if (cfg.turbulent == TRM_TURBULENT_PRESCRIBED) { hs = s.in[TRM_IN_SENSIBLE_HEAT_FLUX][c]; hl = s.in[TRM_IN_LATENT_HEAT_FLUX][c]; }   // turbulent_fluxes.jl:9-16
        s.Hs[c] = hs; s.Hl[c] = hl;
        s.G[c] = rnet - hs - hl;                                  // skin_temperature.jl:76-80
    }
    // compute_surface_energy_fluxes_kernel! surface_energy_balance.jl:82-110
    void compute_seb(State<NF>& s) {
        bool implicit = cfg.skin == TRM_SKIN_IMPLICIT;
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            seb_fluxes(s, c);
            if (implicit) {
                NF Tg = s.T[s.ix(nz, c)];  // ground_temperature is a view of T[:, :, Nz] (soil_energy.jl:52-57)
                s.Ts[c] = Tg - s.G[c] * dzc[nz] / (2 * kappa_skin);  // skin_temperature.jl:62-68,138-150
                seb_fluxes(s, c);
            }
        }
    }
    // ---- vegetated LandModel (VegetationCarbon + PALADYN canopy hydrology) ---------------------------------
    inline NF skin_T(const State<NF>& s, int64_t c) const { return cfg.skin == TRM_SKIN_PRESCRIBED ? s.in[TRM_IN_SKIN_TEMPERATURE][c] : s.Ts[c]; }
    // compute_vpd physical_constants.jl:83-97 [Pa]
    inline NF vpd_pa(const State<NF>& s, int64_t c, NF Tsurf) const {
        NF q = s.in[TRM_IN_SPECIFIC_HUMIDITY][c], p = s.in[TRM_IN_AIR_PRESSURE][c];
        NF ea = q * p / (eps_mw + (1 - eps_mw) * q);
        return jmax(e_sat(Tsurf) - ea, NF(0.1));
    }
    // compute_auxiliary!(state, grid, veg::VegetationCarbon, ...), vegetation_carbon.jl:71-104 : one loop per launch
    void compute_vegetation(State<NF>& s) {
        const NF th_fc = (NF)vp.field_capacity, th_wp = (NF)vp.wilting_point;
        // FieldCapacityLimitedPAW: XYZ kernel (plant_available_water.jl:64-89) ...
#pragma omp parallel for schedule(static)
        for (int64_t c0 = 0; c0 < nc; c0 += CHUNK) {
            int64_t c1 = std::min(nc, c0 + CHUNK);
            for (int k = 1; k <= nz; ++k)
                for (int64_t c = c0; c < c1; ++c) {
                    Fr f = fractions(s.sat[s.ix(k, c)], s.liq[s.ix(k, c)]);
                    s.PAW[s.ix(k, c)] = jmax(jmin(NF(1), (f.water - th_wp) / (th_fc - th_wp)), NF(0));
                }
        }
        // ... then compute!(soil_moisture_limiting_factor) = Integral(PAW * root_fraction / dz, dims = 3) (:31-35)
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF b = 0;
            for (int k = 1; k <= nz; ++k) b += s.PAW[s.ix(k, c)] * rootf[k] / dzc[k] * dzc[k];
            s.betasm[c] = b;
        }
        passes += 3;
        const NF SLA = (NF)vp.SLA, awl = (NF)vp.awl;
        // PALADYNCarbonDynamics auxiliary: LAI_b (carbon_dynamics.jl:82-85,167-170)
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) s.LAIb[c] = s.Cveg[c] / ((NF(2.0) / SLA) + awl);
        // PALADYNPhenology (phenology.jl:33-70): f_deciduous = 0, phen = 1
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF fdec = 0, ph = NF(1.0);
            s.phen[c] = ph;
            s.LAI[c] = (fdec * ph + (NF(1.0) - fdec)) * s.LAIb[c];
        }
        // MedlynStomatalConductance (stomatal_conductance.jl:45-82,106-118): reads the net assimilation of the
        // PREVIOUS evaluation (the circular dependency noted at vegetation_carbon.jl:89-91)
        const NF g1 = (NF)vp.g1, g_min = (NF)vp.g_min / 1000, k_ext = (NF)vp.k_ext;
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF vpd = vpd_pa(s, c, s.in[TRM_IN_AIR_TEMPERATURE][c]);
            NF co2 = s.in[TRM_IN_CO2][c];
            NF g0 = g_min * (1 - std::exp(-k_ext * s.LAI[c])) * s.betasm[c];
            s.gwcan[c] = g0 + NF(1.6) * (1 + g1 / std::sqrt(vpd)) * s.An[c] / co2 * NF(1.0e6);
            s.lamc[c] = NF(1.0) - NF(1.0) / (NF(1.0) + g1 / std::sqrt(vpd * NF(1.0e-3)));
        }
        // LUEPhotosynthesis (photosynthesis.jl:284-344)
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF Rd, An;
            photosynthesis(s.in[TRM_IN_AIR_TEMPERATURE][c], s.in[TRM_IN_SHORTWAVE_DOWN][c], s.in[TRM_IN_AIR_PRESSURE][c],
                           s.in[TRM_IN_CO2][c], s.LAI[c], s.lamc[c], s.betasm[c], Rd, An);
            s.Rd[c] = Rd; s.An[c] = An; s.GPP[c] = An * NF(1.0e-3);
        }
        // PALADYNAutotrophicRespiration (autotrophic_respiration.jl:46-154)
        const NF cn_sap = (NF)vp.cn_sapwood, cn_root = (NF)vp.cn_root, aws = (NF)vp.aws;
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF Ta = s.in[TRM_IN_AIR_TEMPERATURE][c], Tsoil = s.T[s.ix(nz, c)];
            NF Rdl = s.in[TRM_IN_DAILY_LEAF_RESPIRATION][c], ph = s.phen[c], Cv = s.Cveg[c], GPP = s.GPP[c];
            auto f_temp = [](NF T) { return std::exp(NF(308.56) * (NF(1.0) / NF(56.02) - NF(1.0) / (NF(46.02) + T))); };
            NF f_soil = (Tsoil > 7) ? f_temp(Tsoil) : NF(0);
            NF f_air = f_temp(Ta);
            NF resp10 = NF(0.066);
            NF R_leaf = Rdl / NF(1000.0);
            NF R_stem = resp10 * f_air * (awl * ((NF(2.0) / SLA) + awl)) / (Cv * aws * cn_sap);
            NF R_root = resp10 * f_soil * ph * (NF(2.0) / SLA) / (SLA * Cv * cn_root);
            NF Rm = R_leaf + R_stem + R_root;
            NF Rg = NF(0.25) * (GPP - Rm);
            NF Ra = Rm + Rg;
            s.Ra[c] = Ra; s.NPP[c] = GPP - Ra;
        }
    }
    // compute_respiration_assimilation, photosynthesis.jl:212-275
    void photosynthesis(NF T_air, NF swdown, NF pres, NF co2, NF LAI, NF lamc, NF beta_sm, NF& Rd, NF& An) const {
        NF pres_O2 = NF(0.209) * pres;               // physics_utils.jl:16-20
        NF pres_a = co2 * NF(1.0e-6) * pres;         // physics_utils.jl:27-30
        Rd = 0; An = 0;
        if (!(swdown > 0 && T_air > NF(-3.0))) return;
        NF ex = (T_air - NF(25.0)) * NF(0.1);
        NF tau = (NF)vp.tau25 * std::pow((NF)vp.q10_tau, ex);
        NF Kc = (NF)vp.Kc25 * std::pow((NF)vp.q10_Kc, ex);
        NF Ko = (NF)vp.Ko25 * std::pow((NF)vp.q10_Ko, ex);
        NF Gs = pres_O2 / (NF(2.0) * tau);
        if (!(LAI > 0)) return;
        NF PAR = NF(0.5) * swdown * (NF(1.0) - (NF)vp.alpha_leaf) * (NF)vp.cq;
        NF APAR = (NF)vp.alpha_a * PAR * (NF(1.0) - std::exp(-(NF)vp.k_ext * LAI));
        NF pres_i = lamc * pres_a;
        // compute_temperature_stress :143-169
        NF Tl = (NF)vp.T_CO2_low, Th = (NF)vp.T_CO2_high, Pl = (NF)vp.T_photos_low, Ph = (NF)vp.T_photos_high;
        NF k1 = NF(2.0) * std::log(NF(1.0) / NF(0.99) - NF(1.0)) / (Tl - Pl);
        NF k2 = NF(0.5) * (Tl + Pl);
        NF k3 = std::log(NF(0.99) / NF(0.01)) / (Th - Ph);
        NF T_stress = 0;
        if (Tl < T_air && T_air < Th) {
            NF low = NF(1.0) / (NF(1.0) + std::exp(k1 * (k2 - T_air)));
            NF high = NF(1.0) - NF(0.01) * std::exp(k3 * (T_air - Ph));
            T_stress = low * high;
        }
        // compute_assimilation_factors :185-194, compute_Vc_max :208-211 (called with APAR), compute_Rd, compute_Ag
        NF aC3 = (NF)vp.alpha_C3, th = (NF)vp.theta_r;
        NF c_1 = aC3 * T_stress * (NF)vp.C_mass * (pres_i - Gs) / (pres_i + NF(2.0) * Gs);
        NF c_2 = (pres_i - Gs) / (pres_i + Kc * (NF(1.0) + pres_O2 / Ko));
        NF Vc_max = c_1 * APAR * (pres_i + Kc * (NF(1.0) + pres_O2 / Ko)) / (pres_i - Gs);
        Rd = aC3 * Vc_max * beta_sm;
        NF JE = c_1 * APAR, JC = c_2 * Vc_max;
        NF sJ = JE + JC;
        NF Ag = (sJ - std::sqrt(sJ * sJ - NF(4) * th * JE * JC)) / (NF(2) * th) * beta_sm;
        An = Ag - Rd;
    }
    // PALADYNCanopyInterception auxiliary, canopy_interception.jl:161-187
    void compute_canopy_interception(State<NF>& s) {
        const NF a_int = (NF)vp.alpha_int, k_ext = (NF)vp.k_ext_can, wmax0 = (NF)vp.w_can_max, tau_w = (NF)vp.tau_w;
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF rain = s.in[TRM_IN_RAINFALL][c], LAI = s.LAI[c], SAI = s.in[TRM_IN_SAI][c], w = s.wcan[c];
            NF wmax = wmax0 * (LAI + SAI);
            NF f_can = wmax > 0 ? w / wmax : NF(0);
            NF I_can = a_int * rain * (NF(1) - std::exp(-k_ext * (LAI + SAI)));
            NF R_can = jmax(w, NF(0)) / tau_w;
            s.Ican[c] = I_can; s.Rcan[c] = R_can; s.fcan[c] = f_can;
            s.raing[c] = rain - I_can + R_can;
        }
    }
    // PALADYNCanopyEvapotranspiration, canopy_evapotranspiration.jl:51-177
    void compute_evapotranspiration(State<NF>& s) {
        const NF C_can = (NF)vp.C_can;
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF Tsk = s.Ts[c];   // fields.skin_temperature: the state variable (the prescribed variant reads its input field)
            if (cfg.skin == TRM_SKIN_PRESCRIBED) Tsk = s.in[TRM_IN_SKIN_TEMPERATURE][c];
            NF Tg = s.T[s.ix(nz, c)];
            NF dqs = humidity_vpd(s, c, Tsk), dqg = humidity_vpd(s, c, Tg);
            double ra = r_a(s, c);
            NF V = jmax(s.in[TRM_IN_WINDSPEED][c], Vmin);
            NF re = (1 - std::exp(-s.LAI[c] - s.in[TRM_IN_SAI][c])) / (C_can * V);
            NF rs = 1 / jmax(s.gwcan[c], std::sqrt(std::numeric_limits<NF>::epsilon()));
            s.transp[c] = (NF)((double)dqs / (ra + (double)rs));
            s.Egnd[c] = (NF)((double)(ground_resistance(s, c) * dqg) / (ra + (double)re));
            s.Ecan[c] = (NF)((double)(s.fcan[c] * dqs) / ra);
        }
    }
    // compute_tendencies!: canopy water (canopy_interception.jl:189-200), vegetation carbon (carbon_dynamics.jl:107-112,
    // 152-158), vegetation fraction (vegetation_dynamics.jl:60-75,110-120)
    void vegetation_tendencies(State<NF>& s) {
        const NF SLA = (NF)vp.SLA, awl = (NF)vp.awl, Lmin = (NF)vp.LAI_min, Lmax = (NF)vp.LAI_max;
        const NF gL = (NF)vp.gamma_L, gR = (NF)vp.gamma_R, gS = (NF)vp.gamma_S, nu_seed = (NF)vp.nu_seed, gv = (NF)vp.gamma_v_min;
        auto lambda_NPP = [&](NF LAI_b) { return LAI_b < Lmin ? NF(0) : (LAI_b <= Lmax ? (LAI_b - Lmin) / (Lmax - Lmin) : NF(1.0)); };
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) s.tendwcan[c] = s.Ican[c] - s.Ecan[c] - s.Rcan[c];
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF lam = lambda_NPP(s.LAIb[c]);
            NF Lloc = (gL / SLA + gR / SLA + gS * awl) * s.LAIb[c];
            s.tendCveg[c] = (NF(1.0) - lam) * s.NPP[c] - Lloc;
        }
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            NF lam = lambda_NPP(s.LAIb[c]);
            NF nus = jmax(s.nu[c], nu_seed);
            s.tendnu[c] = (lam * s.NPP[c] / s.Cveg[c]) * nus * (NF(1.0) - s.nu[c]) - gv * nus;
        }
    }
    // compute_auxiliary! soil_coupled.jl:62-72 / land_model.jl:79-88
    void compute_auxiliary(State<NF>& s) {
        compute_hydraulics(s);
        if (veg) { compute_vegetation(s); compute_canopy_interception(s); compute_evapotranspiration(s); compute_runoff(s); compute_seb(s); compute_seb(s); }
        else if (land) { compute_evaporation(s); compute_runoff(s); compute_seb(s); compute_seb(s); }
    }
    // compute_tendencies! soil_coupled.jl:80-90 / land_model.jl:90-96
    void compute_tendencies(State<NF>& s) {
        if (richards) richards_tendency(s);
        energy_tendency(s);
        if (veg) vegetation_tendencies(s);
    }
    // update_state! state_variables.jl:72-80
    void update_state(State<NF>& s) {
        reset_tendencies(s);
        update_inputs(s);
        fill_halo_regions(s);
        compute_auxiliary(s);
        compute_tendencies(s);
    }
    // explicit_step! abstract_timestepper.jl:65-141 with [OCN] compute_z_bcs! (Flux BCs only; Appendix A.8)
    void explicit_step(State<NF>& s, NF dt) {
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            if (land) {  // land_model.jl:56-62 : top Flux BCs G and -infiltration
                s.tendU[s.ix(nz, c)] -= s.G[c] / dzc[nz];
                // the infiltration Flux BC sits on saturation_water_ice, which is only stepped when it is prognostic (RichardsEq)
                if (richards) s.tendsat[s.ix(nz, c)] -= (-s.infil[c]) / dzc[nz];
            } else {
                if (cfg.bc[TRM_BC_ENERGY_TOP].kind == TRM_BC_FLUX) s.tendU[s.ix(nz, c)] -= bc_value(s, TRM_BC_ENERGY_TOP, c) / dzc[nz];
                if (richards && cfg.bc[TRM_BC_SATURATION_TOP].kind == TRM_BC_FLUX) s.tendsat[s.ix(nz, c)] -= bc_value(s, TRM_BC_SATURATION_TOP, c) / dzc[nz];
            }
            if (cfg.bc[TRM_BC_ENERGY_BOTTOM].kind == TRM_BC_FLUX) s.tendU[s.ix(1, c)] += bc_value(s, TRM_BC_ENERGY_BOTTOM, c) / dzc[1];
            if (richards && cfg.bc[TRM_BC_SATURATION_BOTTOM].kind == TRM_BC_FLUX) s.tendsat[s.ix(1, c)] += bc_value(s, TRM_BC_SATURATION_BOTTOM, c) / dzc[1];
        }
#pragma omp parallel for schedule(static)
        for (int64_t c0 = 0; c0 < nc; c0 += CHUNK) {
            int64_t c1 = std::min(nc, c0 + CHUNK);
            for (int k = 1; k <= nz; ++k)
                for (int64_t c = c0; c < c1; ++c) {
                    s.U[s.ix(k, c)] += s.tendU[s.ix(k, c)] * dt;
                    if (richards) s.sat[s.ix(k, c)] += s.tendsat[s.ix(k, c)] * dt;
                }
            if (richards) for (int64_t c = c0; c < c1; ++c) s.Sx[c] += s.tendSx[c] * dt;
            if (land && cfg.skin == TRM_SKIN_IMPLICIT) for (int64_t c = c0; c < c1; ++c) s.Ts[c] += s.tendTs[c] * dt;
            if (veg) for (int64_t c = c0; c < c1; ++c) {
                s.wcan[c] += s.tendwcan[c] * dt; s.Cveg[c] += s.tendCveg[c] * dt; s.nu[c] += s.tendnu[c] * dt;
            }
        }
        passes += richards ? 6 : 3;
    }
    // adjust_saturation_profile! soil_hydrology.jl:185-219
    void adjust_saturation(State<NF>& s) {
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            for (int k = 1; k <= nz - 1; ++k) {
                NF e = jmax(s.sat[s.ix(k, c)] - 1, NF(0));
                s.sat[s.ix(k, c)] -= e;
                s.sat[s.ix(k + 1, c)] += e * dzc[k] / dzc[k + 1];
            }
            for (int k = nz; k >= 2; --k) {
                NF d = jmax(-s.sat[s.ix(k, c)], NF(0));
                s.sat[s.ix(k, c)] += d;
                s.sat[s.ix(k - 1, c)] -= d * dzc[k] / dzc[k - 1];
            }
            NF e = jmax(s.sat[s.ix(nz, c)] - 1, NF(0));
            s.sat[s.ix(nz, c)] -= e;
            s.Sx[c] += e * dzc[nz];
            s.sat[s.ix(1, c)] = jmax(s.sat[s.ix(1, c)], NF(0));
        }
        passes += 2;
    }
    // compute_water_table! soil_hydrology.jl:170-175 + findfirst_z kernel_utils.jl:7-16
    // (scans k = 1..Nz+1 over the face nodes, i.e. reads the halo cell Nz+1)
    void compute_water_table(State<NF>& s) {
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            int idx = -1;
            for (int k = 1; k <= nz + 1; ++k) if (idx < 0 && s.sat[s.ix(k, c)] < 1) idx = k;
            s.wt[c] = idx > 0 ? zF[idx] : zF[nz + 1];
        }
        passes += 1;
    }
    // saturation_to_pressure! soil_hydraulic_closures.jl:102-129
    void saturation_to_pressure(State<NF>& s) {
        NF zref = zF[nz + 1];
#pragma omp parallel for schedule(static)
        for (int64_t c0 = 0; c0 < nc; c0 += CHUNK) {
            int64_t c1 = std::min(nc, c0 + CHUNK);
            for (int k = 1; k <= nz; ++k)
                for (int64_t c = c0; c < c1; ++c) {
                    NF sat = s.sat[s.ix(k, c)], z = zC[k];
                    NF psim = swrc_inv(sat * por, por);
                    NF psiz = z - zref;
                    NF psih = jmax(NF(0), s.wt[c] - z);
                    s.psi[s.ix(k, c)] = psih + psim + psiz;
                }
        }
        passes += 2;
    }
    // energy_to_temperature! soil_energy_closures.jl:99-159 ; safediv utils.jl:25
    void energy_to_temperature(State<NF>& s) {
        const NF epsNF = std::numeric_limits<NF>::epsilon();
#pragma omp parallel for schedule(static)
        for (int64_t c0 = 0; c0 < nc; c0 += CHUNK) {
            int64_t c1 = std::min(nc, c0 + CHUNK);
            for (int k = 1; k <= nz; ++k)
                for (int64_t c = c0; c < c1; ++c) {
                    NF U = s.U[s.ix(k, c)], sat = s.sat[s.ix(k, c)];
                    NF Lt = L * sat * por;
                    NF liq;
                    if (U >= 0) liq = 1;
                    else if (U >= -Lt) { NF y = -Lt; NF sd = (y == 0) ? std::numeric_limits<NF>::infinity() : U / (y + epsNF); liq = 1 - sd; }
                    else liq = 0;  // Bool * x is a strong zero in Julia
                    s.liq[s.ix(k, c)] = liq;
                    NF C = heat_capacity(sat, liq);
                    NF T;
                    if (U < -Lt) T = (U + Lt) / C; else if (U >= 0) T = U / C; else T = 0;
                    s.T[s.ix(k, c)] = T;
                }
        }
        passes += 4;
    }
    // temperature_to_energy! soil_energy_closures.jl:64-97 (initialisation only)
    void temperature_to_energy(State<NF>& s) {
#pragma omp parallel for schedule(static)
        for (int64_t c0 = 0; c0 < nc; c0 += CHUNK) {
            int64_t c1 = std::min(nc, c0 + CHUNK);
            for (int k = 1; k <= nz; ++k)
                for (int64_t c = c0; c < c1; ++c) {
                    NF T = s.T[s.ix(k, c)], sat = s.sat[s.ix(k, c)];
                    NF liq = T >= 0 ? NF(1) : NF(0);
                    s.liq[s.ix(k, c)] = liq;
                    NF C = heat_capacity(sat, liq);
                    s.U[s.ix(k, c)] = T * C - L * sat * por * (1 - liq);
                }
        }
    }
    // closure! soil_coupled.jl:99-107 (hydrology closure only exists for Richards)
    void closure(State<NF>& s) {
        if (richards) { adjust_saturation(s); compute_water_table(s); saturation_to_pressure(s); }
        energy_to_temperature(s);
    }
    void tick(State<NF>& s, NF dt) { s.time = s.time + dt; s.iteration += 1; }

    // ---------------------------------------------------------------- public operations
    // initialize!(integrator) tail, model_integrator.jl:96-109 -> soil_model.jl:31-37 / land_model.jl:68-77
    int initialize() {
        // reset!(clock) and the reset of every non user-initialised field (model_integrator.jl:96-100)
        st.time = 0; st.iteration = 0;
        std::fill(st.liq.begin(), st.liq.end(), NF(0));
        update_inputs(st);
        // hydrology: Richards closure! then hydraulics (soil_hydrology_rre.jl:33-47);
        // NoFlow hydraulics + water table (soil_hydrology.jl:113-117). liquid fraction is still the
        // freshly reset field (zero) at this point in the reference; hydraulics are recomputed at the
        // first update_state! so only the ordering of the writes matters.
        if (richards) { adjust_saturation(st); compute_water_table(st); saturation_to_pressure(st); compute_hydraulics(st); }
        else { compute_hydraulics(st); compute_water_table(st); }
        temperature_to_energy(st);  // soil_energy.jl:64-77
        initialized = true;
        return TRM_OK;
    }
    void copy_state(State<NF>& dst, const State<NF>& src_) { dst = src_; }  // copyto! state_variables.jl:505-523
    // forward_euler.jl:19-31
    void step_euler(NF dt) {
        update_state(st);
        explicit_step(st, dt);
        closure(st);
        tick(st, dt);
    }
    // heun.jl:37-71
    void step_heun(NF dt) {
        update_state(st);
        copy_state(stage, st);
        explicit_step(stage, dt);
        closure(stage);
        tick(stage, dt);
        update_state(stage);
        // average_tendencies! heun.jl:27-35
        size_t n3 = st.tendU.size();
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n3; ++i) {
            st.tendU[i] = (st.tendU[i] + stage.tendU[i]) / 2;
            if (richards) st.tendsat[i] = (st.tendsat[i] + stage.tendsat[i]) / 2;
        }
        if (richards) for (int64_t c = 0; c < nc; ++c) st.tendSx[c] = (st.tendSx[c] + stage.tendSx[c]) / 2;
        if (land && cfg.skin == TRM_SKIN_IMPLICIT) for (int64_t c = 0; c < nc; ++c) st.tendTs[c] = (st.tendTs[c] + stage.tendTs[c]) / 2;
        if (veg) for (int64_t c = 0; c < nc; ++c) {
            st.tendwcan[c] = (st.tendwcan[c] + stage.tendwcan[c]) / 2;
            st.tendCveg[c] = (st.tendCveg[c] + stage.tendCveg[c]) / 2;
            st.tendnu[c] = (st.tendnu[c] + stage.tendnu[c]) / 2;
        }
        explicit_step(st, dt);
        closure(st);
        tick(st, dt);
    }
    int step(double dt, int64_t nsteps) {
        if (!initialized) return fail(TRM_ERR_STATE, "orc_step before orc_initialize");
        for (int64_t i = 0; i < nsteps; ++i) { if (heun) step_heun((NF)dt); else step_euler((NF)dt); }
        return TRM_OK;
    }
    int aux() { compute_auxiliary(st); return TRM_OK; }
    // reset!(integrator.state), model_integrator.jl:98: fields and clock back to zero (inputs are re-read by update_inputs!)
    int reset() {
        auto in_keep = st.in;
        st = State<NF>(); st.alloc(nz, nc, land, veg); st.in = in_keep;
        stage = State<NF>();   // (Heun copies the whole state into its stage at every step, heun.jl:45)
        for (auto& a : acc) std::fill(a.begin(), a.end(), 0.0);
        initialized = false;
        return TRM_OK;
    }
    int tendencies() {
        update_state(st);
        // show what explicit_step! would integrate: add the flux BCs on a scratch copy
        State<NF> tmp = st; NF zero = 0; explicit_step(tmp, zero);
        st.tendU = tmp.tendU; st.tendsat = tmp.tendsat;
        return TRM_OK;
    }

    // ---------------------------------------------------------------- field access
    std::vector<NF>* field3(int id) {
        switch (id) {
            case TRM_F_INTERNAL_ENERGY: return &st.U; case TRM_F_TEMPERATURE: return &st.T;
            case TRM_F_LIQUID_WATER_FRACTION: return &st.liq; case TRM_F_SATURATION_WATER_ICE: return &st.sat;
            case TRM_F_PRESSURE_HEAD: return &st.psi; case TRM_F_TEND_INTERNAL_ENERGY: return &st.tendU;
            case TRM_F_TEND_SATURATION: return &st.tendsat;
            case TRM_F_PLANT_AVAILABLE_WATER: return veg ? &st.PAW : nullptr;
        }
        return nullptr;
    }
    std::vector<NF>* field2(int id) {
        switch (id) {
            case TRM_F_SURFACE_EXCESS_WATER: return &st.Sx; case TRM_F_WATER_TABLE: return &st.wt;
            case TRM_F_SKIN_TEMPERATURE: return &st.Ts; case TRM_F_GROUND_HEAT_FLUX: return &st.G;
            case TRM_F_SHORTWAVE_UP: return &st.SWup; case TRM_F_LONGWAVE_UP: return &st.LWup;
            case TRM_F_NET_RADIATION: return &st.Rnet; case TRM_F_SENSIBLE_HEAT_FLUX: return &st.Hs;
            case TRM_F_LATENT_HEAT_FLUX: return &st.Hl; case TRM_F_EVAPORATION_GROUND: return &st.Egnd;
            case TRM_F_INFILTRATION: return &st.infil; case TRM_F_SURFACE_RUNOFF: return &st.runoff;
            case TRM_F_CARBON_VEGETATION: return &st.Cveg; case TRM_F_VEGETATION_AREA_FRACTION: return &st.nu;
            case TRM_F_CANOPY_WATER: return &st.wcan; case TRM_F_BALANCED_LEAF_AREA_INDEX: return &st.LAIb;
            case TRM_F_LEAF_AREA_INDEX: return &st.LAI; case TRM_F_PHENOLOGY_FACTOR: return &st.phen;
            case TRM_F_CANOPY_WATER_CONDUCTANCE: return &st.gwcan; case TRM_F_LEAF_TO_AIR_CO2_RATIO: return &st.lamc;
            case TRM_F_NET_ASSIMILATION: return &st.An; case TRM_F_LEAF_RESPIRATION: return &st.Rd;
            case TRM_F_GROSS_PRIMARY_PRODUCTION: return &st.GPP; case TRM_F_AUTOTROPHIC_RESPIRATION: return &st.Ra;
            case TRM_F_NET_PRIMARY_PRODUCTION: return &st.NPP; case TRM_F_SOIL_MOISTURE_LIMITING_FACTOR: return &st.betasm;
            case TRM_F_CANOPY_WATER_INTERCEPTION: return &st.Ican; case TRM_F_CANOPY_WATER_REMOVAL: return &st.Rcan;
            case TRM_F_SATURATION_CANOPY_WATER: return &st.fcan; case TRM_F_RAINFALL_GROUND: return &st.raing;
            case TRM_F_EVAPORATION_CANOPY: return &st.Ecan; case TRM_F_TRANSPIRATION: return &st.transp;
        }
        return nullptr;
    }
    int set_field(int id, const void* host, int64_t count) {
        const NF* h = (const NF*)host;
        if (auto* f = field3(id)) {
            if (count != (int64_t)nz * nc) return fail(TRM_ERR_INVALID, "set_field: count != nz*ncol");
            for (int k = 1; k <= nz; ++k) std::memcpy(&(*f)[st.ix(k, 0)], h + (size_t)(k - 1) * nc, sizeof(NF) * nc);  // set! writes the interior only
            return TRM_OK;
        }
        if (auto* f = field2(id)) {
            if (f->empty()) return fail(TRM_ERR_INVALID, "field not defined for this model");
            if (count != nc) return fail(TRM_ERR_INVALID, "set_field: count != ncol");
            std::memcpy(f->data(), h, sizeof(NF) * nc); return TRM_OK;
        }
        return fail(TRM_ERR_INVALID, "set_field: unknown or read-only field");
    }
    int get_field(int id, void* host, int64_t count) {
        NF* h = (NF*)host;
        if (auto* f = field3(id)) {
            if (count != (int64_t)nz * nc) return fail(TRM_ERR_INVALID, "get_field: count != nz*ncol");
            for (int k = 1; k <= nz; ++k) std::memcpy(h + (size_t)(k - 1) * nc, &(*f)[st.ix(k, 0)], sizeof(NF) * nc);
            return TRM_OK;
        }
        if (id == TRM_F_HYDRAULIC_CONDUCTIVITY) {
            if (count != (int64_t)(nz + 1) * nc) return fail(TRM_ERR_INVALID, "get_field: count != (nz+1)*ncol");
            for (int k = 1; k <= nz + 1; ++k) std::memcpy(h + (size_t)(k - 1) * nc, &st.Kf[st.ix(k, 0)], sizeof(NF) * nc);
            return TRM_OK;
        }
        if (id == TRM_F_ROOT_FRACTION && veg) {
            if (count != (int64_t)nz * nc) return fail(TRM_ERR_INVALID, "get_field: count != nz*ncol");
            for (int k = 1; k <= nz; ++k) for (int64_t c = 0; c < nc; ++c) h[(size_t)(k - 1) * nc + c] = rootf[k];
            return TRM_OK;
        }
        if (id == TRM_F_GROUND_TEMPERATURE) {
            if (count != nc) return fail(TRM_ERR_INVALID, "get_field: count != ncol");
            std::memcpy(h, &st.T[st.ix(nz, 0)], sizeof(NF) * nc); return TRM_OK;
        }
        if (auto* f = field2(id)) {
            if (f->empty()) return fail(TRM_ERR_INVALID, "field not defined for this model");
            if (count != nc) return fail(TRM_ERR_INVALID, "get_field: count != ncol");
            std::memcpy(h, f->data(), sizeof(NF) * nc); return TRM_OK;
        }
        return fail(TRM_ERR_INVALID, "get_field: unknown field");
    }
    // time-averaged output: per-field accumulators in the host layout of get_field
    std::vector<std::vector<double>> acc = std::vector<std::vector<double>>(TRM_F_COUNT);
    int accumulate(int id, double w) {
        if (id < 0 || id >= TRM_F_COUNT) return fail(TRM_ERR_INVALID, "accumulate: bad field id");
        int64_t count = field3(id) ? (int64_t)nz * nc : (id == TRM_F_HYDRAULIC_CONDUCTIVITY ? (int64_t)(nz + 1) * nc : nc);
        std::vector<NF> tmp((size_t)count);
        if (int rc = get_field(id, tmp.data(), count)) return rc;
        auto& a = acc[id];
        if ((int64_t)a.size() != count) a.assign((size_t)count, 0.0);
        for (int64_t i = 0; i < count; ++i) a[(size_t)i] = (double)(NF)((NF)a[(size_t)i] + (NF)w * tmp[(size_t)i]);
        return TRM_OK;
    }
    int get_accumulated(int id, void* host, int64_t count, double scale, int reset) {
        if (id < 0 || id >= TRM_F_COUNT || (int64_t)acc[id].size() != count) return fail(TRM_ERR_INVALID, "get_accumulated: nothing accumulated for this field / wrong count");
        NF* h = (NF*)host;
        for (int64_t i = 0; i < count; ++i) h[i] = (NF)scale * (NF)acc[id][(size_t)i];
        if (reset) std::fill(acc[id].begin(), acc[id].end(), 0.0);
        return TRM_OK;
    }
    int diagnostics(trm_diag* d) {
        double e = 0, w = 0, tmin = INFINITY, tmax = -INFINITY, smin = INFINITY, smax = -INFINITY, nan = 0;
        for (int64_t c = 0; c < nc; ++c) {
            for (int k = 1; k <= nz; ++k) {
                double U = st.U[st.ix(k, c)], T = st.T[st.ix(k, c)], s = st.sat[st.ix(k, c)];
                e += U * (double)dzc[k]; w += s * (double)por * (double)dzc[k];
                tmin = std::min(tmin, T); tmax = std::max(tmax, T); smin = std::min(smin, s); smax = std::max(smax, s);
                nan += (!std::isfinite(U)) + (!std::isfinite(T)) + (!std::isfinite(s));
            }
            w += (double)st.Sx[c];
        }
        d->energy = e; d->water = w; d->t_min = tmin; d->t_max = tmax; d->sat_min = smin; d->sat_max = smax; d->nan_count = nan; d->ncol = (double)nc;
        return TRM_OK;
    }
};

struct Handle {
    int dtype;
    Oracle<float>* f32 = nullptr;
    Oracle<double>* f64 = nullptr;
    ~Handle() { delete f32; delete f64; }
};

#define DISPATCH(h, expr) ((h)->dtype == TRM_F32 ? (h)->f32->expr : (h)->f64->expr)

template <class NF>
int set_sinusoid(Oracle<NF>* o, int id, const void* mean, const void* amp, const void* phase, double period, double lo, double hi) {
    o->src[id].kind = TRM_SRC_SINUSOID; o->src[id].period = period; o->src[id].lo = lo; o->src[id].hi = hi;
    o->src_mean[id].assign((const NF*)mean, (const NF*)mean + o->nc);
    o->src_amp[id].assign((const NF*)amp, (const NF*)amp + o->nc);
    o->src_phase[id].assign((const NF*)phase, (const NF*)phase + o->nc);
    return TRM_OK;
}
template <class NF>
int set_table(Oracle<NF>* o, int id, int nt, const double* times, const void* values) {
    o->src[id].kind = TRM_SRC_TABLE; o->src[id].nt = nt; o->src[id].times.assign(times, times + nt);
    o->src_table[id].assign((const NF*)values, (const NF*)values + (size_t)nt * o->nc);
    return TRM_OK;
}
template <class NF>
int set_raster(Oracle<NF>* o, int id, int nt, const double* times, const void* values) {
    int rc = set_table(o, id, nt, times, values);
    o->src[id].kind = TRM_SRC_RASTER;
    return rc;
}
template <class NF>
int set_infield(Oracle<NF>* o, int id, const void* v) {
    o->src[id].kind = TRM_SRC_FIELD; o->src_field[id].assign((const NF*)v, (const NF*)v + o->nc);
    // set!(state.inputs.<name>, values) writes the input Field itself: a following compute_auxiliary! sees it without update_inputs!
    if ((int64_t)o->st.in[id].size() == o->nc) o->st.in[id] = o->src_field[id];
    return TRM_OK;
}

}  // namespace

template <class NF> int set_infield_pair(Oracle<NF>* o, int id, const void* v0, const void* v1) {
    o->src[id].kind = TRM_SRC_FIELD_PAIR; o->src[id].pair_t0 = (double)o->st.time;
    o->src_field[id].assign((const NF*)v0, (const NF*)v0 + o->nc);
    o->src_amp[id].assign((const NF*)v1, (const NF*)v1 + o->nc);
    return TRM_OK;
}

extern "C" {

void orc_default_params(trm_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->mineral_porosity = 0.49; p->organic_porosity = 0.9; p->rho_soc = 0.0; p->rho_org = 1300.0;
    const double k[5] = {0.57, 2.2, 0.025, 3.8, 0.25}; const double c[5] = {4.2e6, 1.9e6, 0.00125e6, 2.0e6, 2.5e6};
    for (int i = 0; i < 5; ++i) { p->kappa[i] = k[i]; p->heatcap[i] = c[i]; }
    p->rho_w = 1000.0; p->Lsl = 3.34e5; p->Llg = 2.257e6; p->rho_a = 1.293; p->c_a = 1005.7; p->Tref = 273.15;
    p->sigma = 5.6704e-8; p->eps_mw = 0.622;
    p->K_sat = 1.0e-5; p->vg_alpha = 1.0; p->vg_n = 2.0; p->bc_psis = 0.01; p->bc_lambda = 0.2; p->theta_res = 0.0;
    p->impedance = 7.0; p->vwc_forcing = 0.0;
    p->albedo = 0.3; p->emissivity = 0.97; p->kappa_skin = 2.0; p->C_h = 1.2e-3; p->min_windspeed = 0.01;
    p->tau_r = 3600.0; p->evap_beta = 1.0;
    p->field_capacity = 0.25; p->wilting_point = 0.05; p->C_mass = 12.0;
    p->tau25 = 2600.0; p->Kc25 = 30.0; p->Ko25 = 3.0e4; p->q10_tau = 0.57; p->q10_Kc = 2.1; p->q10_Ko = 1.2;
    p->alpha_leaf = 0.17; p->alpha_a = 0.5; p->alpha_C3 = 0.08; p->cq = 4.6e-6; p->k_ext = 0.5;
    p->T_CO2_high = 42.0; p->T_CO2_low = -4.0; p->T_photos_high = 30.0; p->T_photos_low = 15.0; p->theta_r = 0.7;
    p->g1 = 2.3; p->g_min = 0.5; p->cn_sapwood = 330.0; p->cn_root = 29.0; p->aws = 10.0;
    p->SLA = 10.0; p->awl = 2.0; p->LAI_min = 1.0; p->LAI_max = 6.0; p->gamma_L = 0.3; p->gamma_R = 0.3; p->gamma_S = 0.05;
    p->nu_seed = 0.001; p->gamma_v_min = 0.002; p->root_a = 7.0; p->root_b = 2.0;
    p->alpha_int = 0.2; p->k_ext_can = 0.5; p->w_can_max = 2.0e-4; p->tau_w = 86400.0; p->C_can = 0.006;
}
void orc_default_config(trm_config* c) {
    std::memset(c, 0, sizeof(*c));
    c->abi_version = TRM_ABI_VERSION; c->dtype = TRM_F64; c->model = TRM_MODEL_SOIL; c->timestepper = TRM_EULER;
    c->hydrology = TRM_NOFLOW; c->swrc = TRM_SWRC_BROOKSCOREY; c->unsat_k = TRM_UNSATK_LINEAR; c->sat_halo = TRM_HALO_ZERO;
    c->skin = TRM_SKIN_IMPLICIT; c->math = TRM_MATH_FAITHFUL;
    orc_default_params(&c->params);
}
const char* orc_last_error(void) { return g_err.c_str(); }
int orc_abi_version(void) { return TRM_ABI_VERSION; }

int orc_create(const trm_config* cfg, trm_handle** out) {
    if (!cfg || !out) return fail(TRM_ERR_INVALID, "null argument");
    if (cfg->abi_version != TRM_ABI_VERSION) return fail(TRM_ERR_INVALID, "ABI version mismatch");
    if (cfg->nz < 2 || cfg->nz > TRM_MAX_NZ || cfg->ncol < 1 || !cfg->z_faces) return fail(TRM_ERR_INVALID, "bad nz / ncol / z_faces");
    for (int k = 0; k < cfg->nz; ++k) if (!(cfg->z_faces[k + 1] > cfg->z_faces[k])) return fail(TRM_ERR_INVALID, "z_faces must increase");
    Handle* h = new Handle(); h->dtype = cfg->dtype; int rc;
    if (cfg->dtype == TRM_F32) { h->f32 = new Oracle<float>(); rc = h->f32->setup(*cfg); }
    else if (cfg->dtype == TRM_F64) { h->f64 = new Oracle<double>(); rc = h->f64->setup(*cfg); }
    else { delete h; return fail(TRM_ERR_INVALID, "bad dtype"); }
    if (rc != TRM_OK) { delete h; return rc; }
    *out = (trm_handle*)h; return TRM_OK;
}
int orc_destroy(trm_handle* h) { delete (Handle*)h; return TRM_OK; }
int orc_sync(trm_handle*) { return TRM_OK; }
int orc_set_field(trm_handle* h, int id, const void* host, int64_t count) { return DISPATCH((Handle*)h, set_field(id, host, count)); }
int orc_get_field(trm_handle* h, int id, void* host, int64_t count) { return DISPATCH((Handle*)h, get_field(id, host, count)); }
int orc_set_input_const(trm_handle* h_, int id, double v) {
    Handle* h = (Handle*)h_; if (id < 0 || id >= TRM_IN_COUNT) return fail(TRM_ERR_INVALID, "bad input id");
    if (h->dtype == TRM_F32) { h->f32->src[id].kind = TRM_SRC_CONST; h->f32->src[id].cval = v; std::fill(h->f32->st.in[id].begin(), h->f32->st.in[id].end(), (float)v); }
    else { h->f64->src[id].kind = TRM_SRC_CONST; h->f64->src[id].cval = v; std::fill(h->f64->st.in[id].begin(), h->f64->st.in[id].end(), v); }
    return TRM_OK;
}
int orc_set_input_field(trm_handle* h_, int id, const void* v) {
    Handle* h = (Handle*)h_; if (id < 0 || id >= TRM_IN_COUNT) return fail(TRM_ERR_INVALID, "bad input id");
    return h->dtype == TRM_F32 ? set_infield(h->f32, id, v) : set_infield(h->f64, id, v);
}
int orc_set_input_field_pair(trm_handle* h_, int id, const void* v0, const void* v1) {
    Handle* h = (Handle*)h_; if (id < 0 || id >= TRM_IN_COUNT) return fail(TRM_ERR_INVALID, "bad input id");
    return h->dtype == TRM_F32 ? set_infield_pair(h->f32, id, v0, v1) : set_infield_pair(h->f64, id, v0, v1);
}
int orc_set_input_sinusoid(trm_handle* h_, int id, const void* mean, const void* amp, const void* phase, double period, double lo, double hi) {
    Handle* h = (Handle*)h_; if (id < 0 || id >= TRM_IN_COUNT) return fail(TRM_ERR_INVALID, "bad input id");
    return h->dtype == TRM_F32 ? set_sinusoid(h->f32, id, mean, amp, phase, period, lo, hi) : set_sinusoid(h->f64, id, mean, amp, phase, period, lo, hi);
}
int orc_set_input_table(trm_handle* h_, int id, int32_t nt, const double* times, const void* values) {
    Handle* h = (Handle*)h_; if (id < 0 || id >= TRM_IN_COUNT || nt < 1) return fail(TRM_ERR_INVALID, "bad input id / nt");
    return h->dtype == TRM_F32 ? set_table(h->f32, id, nt, times, values) : set_table(h->f64, id, nt, times, values);
}
int orc_set_input_raster(trm_handle* h_, int id, int32_t nt, const double* times, const void* values) {
    Handle* h = (Handle*)h_; if (id < 0 || id >= TRM_IN_COUNT || nt < 1) return fail(TRM_ERR_INVALID, "bad input id / nt");
    return h->dtype == TRM_F32 ? set_raster(h->f32, id, nt, times, values) : set_raster(h->f64, id, nt, times, values);
}
int orc_get_input(trm_handle* h, int id, void* host, int64_t count) {
    if (id < 0 || id >= TRM_IN_COUNT || !host) return fail(TRM_ERR_INVALID, "bad input id");
    return DISPATCH((Handle*)h, get_input(id, host, count));
}
int orc_accumulate(trm_handle* h, int id, double w) { return DISPATCH((Handle*)h, accumulate(id, w)); }
int orc_get_accumulated(trm_handle* h, int id, void* host, int64_t count, double scale, int32_t reset) {
    if (!host) return fail(TRM_ERR_INVALID, "null argument");
    return DISPATCH((Handle*)h, get_accumulated(id, host, count, scale, reset));
}
int orc_initialize(trm_handle* h) { return DISPATCH((Handle*)h, initialize()); }
int orc_step(trm_handle* h, double dt, int64_t n) { return DISPATCH((Handle*)h, step(dt, n)); }
int orc_compute_auxiliary(trm_handle* h) { return DISPATCH((Handle*)h, aux()); }
int orc_compute_tendencies(trm_handle* h) { return DISPATCH((Handle*)h, tendencies()); }
int orc_get_clock(trm_handle* h_, double* t, int64_t* it) {
    Handle* h = (Handle*)h_;
    if (h->dtype == TRM_F32) { *t = h->f32->st.time; *it = h->f32->st.iteration; } else { *t = h->f64->st.time; *it = h->f64->st.iteration; }
    return TRM_OK;
}
int orc_set_clock(trm_handle* h_, double t, int64_t it) {
    Handle* h = (Handle*)h_;
    if (h->dtype == TRM_F32) { h->f32->st.time = (float)t; h->f32->st.iteration = it; } else { h->f64->st.time = t; h->f64->st.iteration = it; }
    return TRM_OK;
}
int orc_diagnostics(trm_handle* h, trm_diag* d) { return DISPATCH((Handle*)h, diagnostics(d)); }
int orc_reset(trm_handle* h) { return DISPATCH((Handle*)h, reset()); }
int64_t orc_array_passes(trm_handle* h_) { Handle* h = (Handle*)h_; return h->dtype == TRM_F32 ? h->f32->passes : h->f64->passes; }
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
// bench.py --impl reference under torchrun: the launcher exports OMP_NUM_THREADS=1 to every rank, the reference arm
// (rank 0 alone) is asked to use all host threads
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}  // extern "C"
