// tile_kernel.cuh -- the ForwardEuler stage as a shared-memory tile kernel (the hot path).
//
// A block of 8 warps owns a tile of 32 adjacent columns x all nz layers.  Lanes map to columns
// (every global access of a warp is one fully coalesced row segment of the [layer][column]
// arrays), warps map to layers (warp w handles layers w+1, w+9, w+17, ...), so every layer
// dependent special case (boundary faces, halos, flux boundary conditions) is warp-uniform and the
// cells a thread owns are independent of each other (instruction level parallelism instead of a
// serial walk up the column).  The column state lives in shared memory for the whole stage:
//
//   phase 1  (cell parallel)  load U, sat -> closure fields T, liq, psi of the state at time n
//            (recomputed, or read when LOAD), thermal / hydraulic conductivities -> smem
//   phase 2  (cell parallel)  face conductivities, Fourier and Darcy fluxes, tendencies, Flux BCs,
//            explicit update; LandModel surface processes run in the warp that owns the top layer
//   phase 3  (column serial, one warp, only for tiles in which some layer left [0, 1])
//            adjust_saturation_profile! sweeps (soil_hydrology.jl:185-219) on the smem profile
//   phase 4  (cell parallel)  water table, closures of the new state, stores
//
// The reference functions computed are the same as in stage_kernel.cuh (the generic streaming
// implementation that serves Heun stages, tendencies and auxiliaries); the per-cell arithmetic is
// shared through column_physics.cuh.  HBM traffic per column-layer-step is unchanged: read U, sat;
// write U, sat, T, liq, psi.
//
// NZCAP (32 / 64 / 128) fixes the shared-memory spacing of the per-field tiles at compile time so
// that every shared-memory access is `row offset + immediate`; nz <= NZCAP is a runtime value.
#pragma once

#include "stage_kernel.cuh"

namespace trm {

constexpr int TILE_COLS = 32;     // columns per tile = lanes of a warp
constexpr int TILE_LD = 33;       // padded row stride of the shared-memory tiles: conflict free along columns AND along layers
constexpr int TILE_WARPS = 8;
constexpr int TILE_THREADS = TILE_COLS * TILE_WARPS;

#ifndef TRM_TILE_MIN_BLOCKS
#define TRM_TILE_MIN_BLOCKS 4    // resident blocks per SM the register allocator must allow (<= 64 registers)
#endif

template <int NZCAP>
struct TileLayout {
    static constexpr int ROWS_H = NZCAP + 2;                 // rows of a tile with z-halos
    // element offsets (units of NF) of the tiles; every tile row holds TILE_COLS columns
    static constexpr int T = 0;
    static constexpr int KAP = T + ROWS_H * TILE_LD;
    static constexpr int P = KAP + ROWS_H * TILE_LD;
    static constexpr int KC = P + ROWS_H * TILE_LD;        // row k = layer k, rows 0 and nz+1.. unused
    static constexpr int U = KC + ROWS_H * TILE_LD;        // row k-1 = layer k
    static constexpr int S = U + NZCAP * TILE_LD;
    static constexpr int COL = S + NZCAP * TILE_LD;        // 4 per-column rows: G_top / excess, infiltration, idx, flags
    static constexpr int MET = COL + 4 * TILE_LD;          // 6 metric arrays of NZCAP + 3
    static constexpr int TOTAL = MET + MET_COUNT * (NZCAP + 3);
};

template <class NF, int PHYS, int LOAD_CT, bool FAST, int NZCAP>
__global__ void __launch_bounds__(TILE_THREADS, (NZCAP <= 32 ? TRM_TILE_MIN_BLOCKS : (NZCAP <= 64 ? 2 : 1))) tile_kernel(const __grid_constant__ StageArgs<NF> A) {
    constexpr bool RICH = PHYS != PHYS_NOFLOW;
    constexpr bool LAND = PHYS == PHYS_LAND;
    constexpr bool LOAD = LOAD_CT != 0;
    constexpr int R = NZCAP / TILE_WARPS;      // layers per thread (upper bound)
    constexpr int CH = 4;                      // layers whose loads are issued together
    using Mx = M<NF, FAST>;
    using L = TileLayout<NZCAP>;
    constexpr int NZP = NZCAP + 3;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    NF* sm = reinterpret_cast<NF*>(smem_raw);
    const int nz = A.nz;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    NF* const zF = sm + L::MET;
    NF* const zC = zF + NZP;
    NF* const dzc = zC + NZP;
    NF* const rdzc = dzc + NZP;
    NF* const dzf = rdzc + NZP;
    NF* const rdzf = dzf + NZP;
    NF* const psiz = rdzf + NZP;
    NF* const col = sm + lane;                 // this thread's column inside every tile
    int* const sIdx = reinterpret_cast<int*>(sm + L::COL + 2 * TILE_LD);
    int* const sFlag = reinterpret_cast<int*>(sm + L::COL + 3 * TILE_LD);

    // metrics: host rows of MET_STRIDE -> compact smem rows of NZCAP + 3
    for (int a = 0; a < MET_COUNT; ++a)
        for (int j = threadIdx.x; j < nz + 3; j += TILE_THREADS) zF[a * NZP + j] = A.metrics[a * MET_STRIDE + j];
    const int64_t c0 = (int64_t)blockIdx.x * TILE_COLS + lane;
    const bool valid = c0 < A.ncol;
    const int64_t c = valid ? c0 : A.ncol - 1;   // lanes beyond the last column shadow it (loads only, never stored)
    const int64_t ld = A.ld;
    const DevParams<NF>& p = A.p;
    const NF dt = A.dt;
    if (warp == 0) { sIdx[lane] = nz + 1; sFlag[lane] = 0; }
    __syncthreads();
    const NF wtx = (RICH && !LOAD) ? A.xWt[c] : NF(0);

    auto bc_input = [&](int slot) -> NF {
        const int kind = A.bc[slot].kind;
        if (kind == TRM_BC_DEFAULT) return NF(0);
        return eval_input(A.in[A.bc[slot].input], c, kind == TRM_BC_FLUX ? A.t_b : A.t_x);
    };

    // ---- phase 1: closure fields and conductivities of the state at time n ----
#pragma unroll 1
    for (int i0 = 0; i0 < R; i0 += CH) {
        NF Ur[CH], sr[CH], Tr[CH], lr[CH], Pr[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int k = warp + 1 + TILE_WARPS * (i0 + j);
            Ur[j] = NF(0); sr[j] = NF(0); Tr[j] = NF(0); lr[j] = NF(0); Pr[j] = NF(0);
            if (k <= nz) {
                const int64_t o = (int64_t)(k - 1) * ld + c;
                Ur[j] = A.xU[o]; sr[j] = A.xS[o];
                if (LOAD) { Tr[j] = A.xT[o]; lr[j] = A.xL[o]; if (RICH) Pr[j] = A.xP[o]; }
            }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int k = warp + 1 + TILE_WARPS * (i0 + j);
            if (k > nz) continue;
            NF* const q = col + k * TILE_LD;     // row k of the halo tiles; row k-1 of the U / S tiles is q - TILE_LD
            const NF U = Ur[j], s = sr[j];
            NF T = Tr[j], l = lr[j], P = Pr[j];
            if (!LOAD) {
                energy_to_temperature<NF, FAST>(p, U, s, T, l);
                if (RICH) P = pressure_head<NF, FAST>(p, s, wtx, zC[k], psiz[k]);
            }
            q[L::T] = T;
            if (RICH) q[L::P] = P;
            q[L::U - TILE_LD] = U;
            q[L::S - TILE_LD] = s;
            q[L::KAP] = FAST ? thermal_conductivity_fast(p, s, l) : thermal_conductivity(p, s, l);
            if (RICH) q[L::KC] = cell_conductivity<NF, FAST>(p, s, l);
        }
    }
    // ---- z-halos (fill_halo_regions!) by the warps that own the boundary layers; they
    //      re-read their own shared-memory rows, so no barrier is needed before this point ----
    if (warp == 0) {           // halo below the bottom layer
        NF* const q = col + TILE_LD;
        const NF T = q[L::T], s = q[L::S - TILE_LD];
        col[L::T] = halo_value(A.bc[TRM_BC_TEMPERATURE_BOTTOM].kind, T, bc_input(TRM_BC_TEMPERATURE_BOTTOM), dzf[1], false);
        // conductivity of the halo cell: same (sat, liq) as layer 1 when the saturation halo is a copy, else sat = 0
        // (SURVEY.md Appendix B.6), for which the liquid fraction drops out of the constituent sum
        const bool copy = RICH || p.sat_halo == TRM_HALO_COPY;
        col[L::KAP] = copy ? q[L::KAP] : (FAST ? thermal_conductivity_fast(p, NF(0), NF(1)) : thermal_conductivity(p, NF(0), NF(1)));
        if (RICH) col[L::P] = halo_value(A.bc[TRM_BC_PRESSURE_BOTTOM].kind, q[L::P], bc_input(TRM_BC_PRESSURE_BOTTOM), dzf[1], false);
        (void)s;
    }
    if (warp == ((nz - 1) & (TILE_WARPS - 1))) {   // halo above the surface (+ LandModel surface processes)
        NF* const q = col + nz * TILE_LD;
        NF* const qt = q + TILE_LD;
        const NF T = q[L::T], s = q[L::S - TILE_LD];
        qt[L::T] = halo_value(A.bc[TRM_BC_TEMPERATURE_TOP].kind, T, bc_input(TRM_BC_TEMPERATURE_TOP), dzf[nz + 1], true);
        const bool copy = RICH || p.sat_halo == TRM_HALO_COPY;
        qt[L::KAP] = copy ? q[L::KAP] : (FAST ? thermal_conductivity_fast(p, NF(0), NF(1)) : thermal_conductivity(p, NF(0), NF(1)));
        if (RICH) qt[L::P] = halo_value(A.bc[TRM_BC_PRESSURE_TOP].kind, q[L::P], bc_input(TRM_BC_PRESSURE_TOP), dzf[nz + 1], true);
        if (LAND) {
            // ---- LandModel surface processes (land_model.jl:79-88), per column ----
            const NF Kc = q[L::KC];
            Surface<NF> a;
            a.SWd = eval_input(A.in[TRM_IN_SHORTWAVE_DOWN], c, A.t_x);
            a.LWd = eval_input(A.in[TRM_IN_LONGWAVE_DOWN], c, A.t_x);
            a.Ta = eval_input(A.in[TRM_IN_AIR_TEMPERATURE], c, A.t_x);
            a.pres = eval_input(A.in[TRM_IN_AIR_PRESSURE], c, A.t_x);
            a.q = eval_input(A.in[TRM_IN_SPECIFIC_HUMIDITY], c, A.t_x);
            a.V = eval_input(A.in[TRM_IN_WINDSPEED], c, A.t_x);
            a.rain = eval_input(A.in[TRM_IN_RAINFALL], c, A.t_x);
            const bool prescribed = p.skin == TRM_SKIN_PRESCRIBED;
            a.Tskin_in = prescribed ? eval_input(A.in[TRM_IN_SKIN_TEMPERATURE], c, A.t_x) : NF(0);
            // aerodynamic_resistance, prescribed_atmosphere.jl:110-116,137 (Float64 literal 1.0e-6 promotes)
            NF Vc = jmax(a.V, p.Vmin);
            double Va = fmax((double)Vc, 1.0e-6);
            a.ra = 1.0 / ((double)p.C_h * Va);
            NF Ts = A.Ts[c];
            // BareGroundEvaporation, bare_ground_evaporation.jl:49-62 ; compute_humidity_vpd
            // prescribed_atmosphere.jl:160-182, physical_constants.jl:83-97, physics_utils.jl:38
            NF Tsurf = prescribed ? a.Tskin_in : Ts;
            NF es = saturation_vapor_pressure(Tsurf);
            NF ea = a.q * a.pres / (p.eps_mw + (1 - p.eps_mw) * a.q);
            NF vpd = jmax(es - ea, NF(0.1));
            NF dq = p.eps_mw * vpd / a.pres;
            NF Egnd = (NF)((double)(p.beta * dq) / a.ra);
            // DirectSurfaceRunoff, direct_surface_runoff.jl:87-117 ; Kf[Nz] = Kc[Nz] (soil_hydrology.jl:270-273)
            NF S = A.bSx[c], Kt = Kc, sat_top = s;
            NF drain, inf;
            if (S > 0) { drain = jmax(S, NF(0)) / p.tau_r; inf = (sat_top < 1) ? jmin(drain, Kt) : NF(0); }
            else { drain = 0; inf = (sat_top < 1) ? jmin(a.rain, Kt) : NF(0); }
            NF runoff = a.rain + drain - inf;
            // surface energy balance kernel, executed twice (land_model.jl:85-86)
            NF swu, lwu, rnet, hs, hl, G;
#pragma unroll 1
            for (int rep = 0; rep < 2; ++rep) {
                seb_fluxes(p, a, prescribed ? a.Tskin_in : Ts, Egnd, swu, lwu, rnet, hs, hl, G);
                if (!prescribed) {
                    Ts = T - G * dzc[nz] / (2 * p.kappa_skin);   // ImplicitSkinTemperature, skin_temperature.jl:62-68,138-150
                    seb_fluxes(p, a, Ts, Egnd, swu, lwu, rnet, hs, hl, G);
                }
            }
            if (valid) {
                A.Egnd[c] = Egnd; A.infil[c] = inf; A.runoff[c] = runoff;
                A.SWup[c] = swu; A.LWup[c] = lwu; A.Rnet[c] = rnet; A.Hs[c] = hs; A.Hl[c] = hl; A.G[c] = G;
                if (!prescribed) A.Ts[c] = Ts;
            }
            col[L::COL] = G; col[L::COL + TILE_LD] = inf;
        }
    }
    __syncthreads();

    // ---- phase 2: fluxes, tendencies, explicit step ----
    auto kf_at = [&](int k) -> NF {   // face conductivity Kf[k], soil_hydrology.jl:249-276 (k is warp uniform)
        if (k <= 0 || k >= nz + 2) return NF(0);           // halo faces are never written by the reference
        if (k == 1) return col[L::KC + TILE_LD];
        if (k >= nz) return col[L::KC + nz * TILE_LD];   // Kf[Nz] = Kc[Nz], Kf[Nz+1] = Kf[Nz]
        return Mx::mn(col[L::KC + k * TILE_LD], col[L::KC + (k - 1) * TILE_LD]);
    };
    int flagged = 0;
#pragma unroll 2
    for (int i = 0; i < R; ++i) {
        const int k = warp + 1 + TILE_WARPS * i;
        if (k > nz) break;
        NF* const q = col + k * TILE_LD;
        const NF Tm = q[L::T - TILE_LD], T0 = q[L::T], Tp = q[L::T + TILE_LD];
        const NF km = q[L::KAP - TILE_LD], k0 = q[L::KAP], kp = q[L::KAP + TILE_LD];
        // diffusive_heat_flux at faces k and k+1, soil_energy.jl:134-149
        const NF qh_lo = -((k0 + km) / 2) * ((T0 - Tm) * rdzf[k]);
        const NF qh_hi = -((kp + k0) / 2) * ((Tp - T0) * rdzf[k + 1]);
        NF tU = -((qh_hi - qh_lo) * rdzc[k]);                                   // soil_energy.jl:112-131
        NF tS = NF(0);
        if (RICH) {
            const NF Pm = q[L::P - TILE_LD], P0 = q[L::P], Pp = q[L::P + TILE_LD];
            const NF Kf_m = kf_at(k - 1), Kf_0 = kf_at(k), Kf_p = kf_at(k + 1), Kf_pp = kf_at(k + 2);
            // darcy_flux at faces k and k+1, soil_hydrology_rre.jl:119-131
            const NF g_lo = (P0 - Pm) * rdzf[k], g_hi = (Pp - P0) * rdzf[k + 1];
            NF K_lo, K_hi;
            if (FAST) {
                K_lo = Mx::mn(Kf_0, g_lo < 0 ? Kf_m : Kf_p);
                K_hi = Mx::mn(Kf_p, g_hi < 0 ? Kf_0 : Kf_pp);
            } else {
                K_lo = (g_lo < 0 ? jmin(Kf_m, Kf_0) : NF(0)) + (g_lo >= 0 ? jmin(Kf_0, Kf_p) : NF(0));
                K_hi = (g_hi < 0 ? jmin(Kf_0, Kf_p) : NF(0)) + (g_hi >= 0 ? jmin(Kf_p, Kf_pp) : NF(0));
            }
            const NF qd_lo = -K_lo * g_lo, qd_hi = -K_hi * g_hi;
            const NF dth = -((qd_hi - qd_lo) * rdzc[k]) + NF(0) + p.vwcf;         // soil_hydrology_rre.jl:95-117
            tS = FAST ? dth * p.rpor : dth / p.por;                               // soil_hydrology.jl:222-237
        }
        // Flux boundary conditions (compute_z_bcs!, abstract_timestepper.jl:69 )
        if (k == nz) {
            if (LAND) { tU -= col[L::COL] / dzc[nz]; tS -= (-col[L::COL + TILE_LD]) / dzc[nz]; }   // land_model.jl:56-62
            else {
                if (A.bc[TRM_BC_ENERGY_TOP].kind == TRM_BC_FLUX) tU -= bc_input(TRM_BC_ENERGY_TOP) / dzc[nz];
                if (RICH && A.bc[TRM_BC_SATURATION_TOP].kind == TRM_BC_FLUX) tS -= bc_input(TRM_BC_SATURATION_TOP) / dzc[nz];
            }
        }
        if (k == 1) {
            if (A.bc[TRM_BC_ENERGY_BOTTOM].kind == TRM_BC_FLUX) tU += bc_input(TRM_BC_ENERGY_BOTTOM) / dzc[1];
            if (RICH && A.bc[TRM_BC_SATURATION_BOTTOM].kind == TRM_BC_FLUX) tS += bc_input(TRM_BC_SATURATION_BOTTOM) / dzc[1];
        }
        // explicit step, abstract_timestepper.jl:113-141
        q[L::U - TILE_LD] = q[L::U - TILE_LD] + tU * dt;
        if (RICH) {
            const NF sn = q[L::S - TILE_LD] + tS * dt;
            q[L::S - TILE_LD] = sn;
            if (!(sn >= 0)) flagged |= 2;                                 // negative (or NaN): needs the downward sweep too
            else if (sn > 1) flagged |= 1;                                // over-saturated: needs the upward sweep
            else if (sn < 1) atomicMin(&sIdx[lane], k);                   // compute_water_table!: lowest unsaturated layer
        }
    }
    if (RICH) {
        // ---- phase 3: adjust_saturation_profile! + compute_water_table! for the columns that need it ----
        const int any = __syncthreads_or(flagged);
        if (any) {
            if (flagged) atomicOr(&sFlag[lane], flagged);
            __syncthreads();
            NF* const sS0 = sm + L::S;
            if (FAST && NZCAP <= 32) {
                // Warp-cooperative sweep: the warp walks its 4 columns, lanes = layers (the padded row stride makes
                // the transposed access conflict free).  The upward sweep (soil_hydrology.jl:192-199) becomes a
                // relaxation -- every over-saturated layer hands its excess to the layer above until none is left --
                // which reaches the same profile as the serial sweep up to the rounding of regrouped additions.
                const int k = lane + 1;
                const NF ratio = (k < nz) ? dzc[k] * rdzc[k + 1] : NF(0);
#pragma unroll 1
                for (int cc = warp * (TILE_COLS / TILE_WARPS); cc < (warp + 1) * (TILE_COLS / TILE_WARPS); ++cc) {
                    const int flag = sFlag[cc];
                    if (!flag) continue;
                    NF* const colS = sS0 + cc;
                    if (flag & 2) {
                        // a layer went negative: serial reference sweeps (up, down, top, bottom clamp) by one lane
                        if (lane == 0) {
                            NF carry = NF(0);
                            for (int kk = 1; kk <= nz - 1; ++kk) {
                                NF v = colS[(kk - 1) * TILE_LD] + carry;
                                const NF e = jmax(v - 1, NF(0));
                                v -= e; carry = e * dzc[kk] / dzc[kk + 1];
                                colS[(kk - 1) * TILE_LD] = v;
                            }
                            colS[(nz - 1) * TILE_LD] += carry;
                            for (int kk = nz; kk >= 2; --kk) {
                                NF v = colS[(kk - 1) * TILE_LD];
                                const NF d = jmax(-v, NF(0));
                                colS[(kk - 1) * TILE_LD] = v + d;
                                colS[(kk - 2) * TILE_LD] -= d * dzc[kk] / dzc[kk - 1];
                            }
                            colS[0] = jmax(colS[0], NF(0));
                        }
                        __syncwarp();
                    }
                    NF v = (k <= nz) ? colS[lane * TILE_LD] : NF(0);
                    if (!(flag & 2)) {
#pragma unroll 1
                        for (;;) {
                            const NF e = (k < nz && v > 1) ? v - 1 : NF(0);
                            if (!__any_sync(0xffffffffu, e > 0)) break;
                            v -= e;
                            NF recv = __shfl_up_sync(0xffffffffu, e * ratio, 1);
                            if (lane == 0) recv = NF(0);
                            v += recv;
                        }
                    }
                    // top excess -> surface_excess_water, :210-216
                    NF ex = NF(0);
                    if (k == nz) { ex = Mx::mx(v - 1, NF(0)); v -= ex; sm[L::COL + cc] = ex * dzc[nz]; }   // G_top slot is dead by now
                    if (k <= nz) colS[lane * TILE_LD] = v;
                    const unsigned below = __ballot_sync(0xffffffffu, k <= nz && v < 1);
                    if (lane == 0) sIdx[cc] = below ? __ffs(below) : nz + 1;   // compute_water_table!
                }
            } else {
                const int flag = sFlag[lane];
                if (warp == 0 && flag) {
                    // serial sweeps, one lane per column ; the excess travels in a register instead of through
                    // sat[k+1] (same additions in the same order as soil_hydrology.jl:192-199)
                    NF* const sS = col + L::S;
                    NF carry = NF(0);
                    int idx = nz + 1;
#pragma unroll 1
                    for (int k = 1; k <= nz - 1; ++k) {
                        NF v = sS[(k - 1) * TILE_LD] + carry;
                        const NF e = Mx::mx(v - 1, NF(0));
                        v -= e;
                        carry = FAST ? e * dzc[k] * rdzc[k + 1] : e * dzc[k] / dzc[k + 1];
                        sS[(k - 1) * TILE_LD] = v;
                        if (v < 1) idx = min(idx, k);
                    }
                    NF st = sS[(nz - 1) * TILE_LD] + carry;
                    if (flag & 2) {
                        // downward sweep, :201-208 (only when some layer went negative)
                        sS[(nz - 1) * TILE_LD] = st;
#pragma unroll 1
                        for (int k = nz; k >= 2; --k) {
                            NF v = sS[(k - 1) * TILE_LD];
                            const NF d = jmax(-v, NF(0));
                            sS[(k - 1) * TILE_LD] = v + d;
                            sS[(k - 2) * TILE_LD] -= d * dzc[k] / dzc[k - 1];
                        }
                        st = sS[(nz - 1) * TILE_LD];
                    } else if (!FAST) {
                        st = st + jmax(-st, NF(0));   // the downward sweep is the identity (up to the sign of zero)
                    }
                    // top excess -> surface_excess_water, :210-216
                    const NF e = Mx::mx(st - 1, NF(0));
                    st -= e;
                    sS[(nz - 1) * TILE_LD] = st;
                    col[L::COL] = e * dzc[nz];   // G_top is dead by now: the slot carries the excess to phase 4
                    if (flag & 2) {
                        sS[0] = jmax(sS[0], NF(0));
                        idx = nz + 1;
#pragma unroll 1
                        for (int k = nz; k >= 1; --k) if (sS[(k - 1) * TILE_LD] < 1) idx = k;
                    } else if (st < 1) idx = min(idx, nz);
                    sIdx[lane] = idx;
                }
            }
            __syncthreads();
        }
    }

    // ---- phase 4: water table, closures of the new state, stores ----
    NF wt_new = NF(0);
    if (RICH) {
        wt_new = zF[sIdx[lane]];               // zF[nz+1] when every layer is saturated (halo cell / fallback agree)
        if (warp == 0 && valid) {
            A.yWt[c] = wt_new;
            NF Sx = A.bSx[c] + NF(0) * dt;     // surface_excess_water tendency is zero (soil_hydrology.jl:260-267)
            if (sFlag[lane]) Sx += col[L::COL];
            A.ySx[c] = Sx;
        }
    }
#pragma unroll 2
    for (int i = 0; i < R; ++i) {
        const int k = warp + 1 + TILE_WARPS * i;
        if (k > nz) break;
        NF* const q = col + k * TILE_LD;
        const NF Un = q[L::U - TILE_LD], sn = q[L::S - TILE_LD];
        NF Tn, ln;
        energy_to_temperature<NF, FAST>(p, Un, sn, Tn, ln);
        NF Pn = NF(0);
        if (RICH) Pn = pressure_head<NF, FAST>(p, sn, wt_new, zC[k], psiz[k]);
        if (valid) {
            const int64_t o = (int64_t)(k - 1) * ld + c;
            A.yU[o] = Un; A.yT[o] = Tn; A.yL[o] = ln;
            if (RICH) { A.yS[o] = sn; A.yP[o] = Pn; }
        }
    }
}

}  // namespace trm
