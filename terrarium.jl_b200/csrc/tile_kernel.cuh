// tile_kernel.cuh -- the ForwardEuler stage as a shared-memory tile kernel (the hot path).
//
// A block owns a tile of 32 adjacent columns x all nz layers.  Lanes map to columns (every global
// access of a warp is one fully coalesced row segment of the [layer][column] arrays), warps map to
// layers (warp w handles layers w+1, w+1+W, ...), so every layer-dependent special case (boundary
// faces, halos, flux boundary conditions) is warp-uniform.  The column state lives in shared
// memory for the whole stage:
//
//   phase 1  (cell parallel)  load U, sat -> closure fields T, liq, psi of the state at time n
//            (recomputed, or read when LOAD_AUX), thermal / hydraulic conductivities -> smem
//   phase 2  (cell parallel)  face conductivities, Fourier and Darcy fluxes, tendencies, Flux BCs,
//            explicit update; LandModel surface processes run in the warp that owns the top layer
//   phase 3  (column serial, one warp, only if some column of the tile needs it)
//            adjust_saturation_profile! sweeps (soil_hydrology.jl:185-219) on the smem profile
//   phase 4  (cell parallel)  water table, closures of the new state, stores
//
// The reference functions computed are the same as in stage_kernel.cuh (which remains the generic
// streaming implementation used for Heun stages, tendencies and auxiliaries); the per-cell
// arithmetic is shared through column_physics.cuh.  HBM traffic per column-layer-step is unchanged:
// read U, sat; write U, sat, T, liq, psi.
#pragma once

#include "stage_kernel.cuh"

namespace trm {

constexpr int TILE_COLS = 32;

#ifndef TRM_TILE_THREADS
#define TRM_TILE_THREADS 256     // 8 warps: warp w owns layers w+1, w+9, ...
#endif
#ifndef TRM_TILE_MIN_BLOCKS
#define TRM_TILE_MIN_BLOCKS 4    // resident blocks per SM the register allocator must allow (<= 64 registers)
#endif

// number of NF elements of dynamic shared memory the tile kernel needs
__host__ __device__ inline size_t tile_smem_elems(int nz) {
    // sU, sS, sKc: nz rows ; sT, sKap, sP: nz+2 rows (z-halos) ; 6 metric arrays ; 4 per-column rows
    return (size_t)(3 * nz + 3 * (nz + 2) + 4) * TILE_COLS + 6 * (size_t)(nz + 3);
}

template <class NF, int PHYS, int LOAD_CT, bool FAST>
__global__ void __launch_bounds__(TRM_TILE_THREADS, TRM_TILE_MIN_BLOCKS) tile_kernel(const __grid_constant__ StageArgs<NF> A) {
    constexpr bool RICH = PHYS != PHYS_NOFLOW;
    constexpr bool LAND = PHYS == PHYS_LAND;
    constexpr bool LOAD = LOAD_CT != 0;
    using Mx = M<NF, FAST>;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    NF* sm = reinterpret_cast<NF*>(smem_raw);
    const int nz = A.nz, nzp = nz + 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    // metrics
    NF* zF = sm;
    NF* zC = zF + nzp;
    NF* dzc = zC + nzp;
    NF* rdzc = dzc + nzp;
    NF* dzf = rdzc + nzp;
    NF* rdzf = dzf + nzp;
    // tiles, row stride TILE_COLS
    NF* sU = rdzf + nzp;                         // [nz]    internal energy (time n, then n+1)
    NF* sS = sU + (size_t)nz * TILE_COLS;        // [nz]    saturation (time n, then n+1)
    NF* sKc = sS + (size_t)nz * TILE_COLS;       // [nz]    hydraulic conductivity at centres
    NF* sT = sKc + (size_t)nz * TILE_COLS;       // [nz+2]  temperature with z-halos (row 0 = below bottom)
    NF* sKap = sT + (size_t)(nz + 2) * TILE_COLS;   // [nz+2]  thermal conductivity with z-halos
    NF* sP = sKap + (size_t)(nz + 2) * TILE_COLS;   // [nz+2]  pressure head with z-halos
    NF* sCol = sP + (size_t)(nz + 2) * TILE_COLS;   // [4]     per column: G_top, infil_top, (int) idx, (int) flags
    int* sIdx = reinterpret_cast<int*>(sCol + 2 * TILE_COLS);
    int* sFlag = reinterpret_cast<int*>(sCol + 3 * TILE_COLS);

    for (int i = threadIdx.x; i < 6 * nzp; i += blockDim.x) sm[i] = A.metrics[i];
    const int64_t c = (int64_t)blockIdx.x * TILE_COLS + lane;
    const bool valid = c < A.ncol;
    const int64_t ld = A.ld;
    const DevParams<NF>& p = A.p;
    const NF dt = A.dt;
    if (warp == 0) { sIdx[lane] = nz + 1; sFlag[lane] = 0; }

    // ---- phase 1a: stream the tile into shared memory (independent loads, all in flight together) ----
    for (int k = warp; k < nz; k += nwarp) {
        const int64_t o = (int64_t)k * ld + c;
        sU[k * TILE_COLS + lane] = valid ? A.xU[o] : NF(0);
        sS[k * TILE_COLS + lane] = valid ? A.xS[o] : NF(1);
        if (LOAD) {
            sT[(k + 1) * TILE_COLS + lane] = valid ? A.xT[o] : NF(0);
            sKap[(k + 1) * TILE_COLS + lane] = valid ? A.xL[o] : NF(1);   // liquid fraction parked in the kappa tile
            if (RICH) sP[(k + 1) * TILE_COLS + lane] = valid ? A.xP[o] : NF(0);
        }
    }
    __syncthreads();   // metrics visible (the tile rows are read back by the thread that wrote them)
    const NF zref = zF[nz + 1];
    const NF wtx = (RICH && !LOAD && valid) ? A.xWt[c] : NF(0);

    auto bc_input = [&](int slot) -> NF {
        const int kind = A.bc[slot].kind;
        if (kind == TRM_BC_DEFAULT || !valid) return NF(0);
        return eval_input(A.in[A.bc[slot].input], c, kind == TRM_BC_FLUX ? A.t_b : A.t_x);
    };

    // ---- phase 1b: closure fields and conductivities of the state at time n ----
    for (int k = warp + 1; k <= nz; k += nwarp) {
        const int r = (k - 1) * TILE_COLS + lane, rh = k * TILE_COLS + lane;
        const NF U = sU[r], s = sS[r];
        NF T, l, P = NF(0);
        if (LOAD) { T = sT[rh]; l = sKap[rh]; if (RICH) P = sP[rh]; }
        else {
            energy_to_temperature<NF, FAST>(p, U, s, T, l);
            if (RICH) P = pressure_head<NF, FAST>(p, s, wtx, zC[k], zref);
            sT[rh] = T;
            if (RICH) sP[rh] = P;
        }
        sKap[rh] = FAST ? thermal_conductivity_fast(p, s, l) : thermal_conductivity(p, s, l);
        if (RICH) sKc[r] = cell_conductivity<NF, FAST>(p, s, l);
        if (k == 1) {          // halo below the bottom layer (fill_halo_regions!)
            sT[lane] = halo_value(A.bc[TRM_BC_TEMPERATURE_BOTTOM].kind, T, bc_input(TRM_BC_TEMPERATURE_BOTTOM), dzf[1], false);
            const NF s0 = (RICH || p.sat_halo == TRM_HALO_COPY) ? s : NF(0);
            sKap[lane] = FAST ? thermal_conductivity_fast(p, s0, l) : thermal_conductivity(p, s0, l);
            if (RICH) sP[lane] = halo_value(A.bc[TRM_BC_PRESSURE_BOTTOM].kind, P, bc_input(TRM_BC_PRESSURE_BOTTOM), dzf[1], false);
        }
        if (k == nz) {         // halo above the surface
            const int rt = (nz + 1) * TILE_COLS + lane;
            sT[rt] = halo_value(A.bc[TRM_BC_TEMPERATURE_TOP].kind, T, bc_input(TRM_BC_TEMPERATURE_TOP), dzf[nz + 1], true);
            const NF sh = (RICH || p.sat_halo == TRM_HALO_COPY) ? s : NF(0);
            sKap[rt] = FAST ? thermal_conductivity_fast(p, sh, l) : thermal_conductivity(p, sh, l);
            if (RICH) sP[rt] = halo_value(A.bc[TRM_BC_PRESSURE_TOP].kind, P, bc_input(TRM_BC_PRESSURE_TOP), dzf[nz + 1], true);
            if (LAND) {
                // ---- LandModel surface processes (land_model.jl:79-88), per column, in the top layer's warp ----
                NF G = NF(0), inf = NF(0);
                if (valid) {
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
                    NF S = A.bSx[c], Kt = sKc[r], sat_top = s;
                    NF drain;
                    if (S > 0) { drain = jmax(S, NF(0)) / p.tau_r; inf = (sat_top < 1) ? jmin(drain, Kt) : NF(0); }
                    else { drain = 0; inf = (sat_top < 1) ? jmin(a.rain, Kt) : NF(0); }
                    NF runoff = a.rain + drain - inf;
                    // surface energy balance kernel, executed twice (land_model.jl:85-86)
                    NF swu, lwu, rnet, hs, hl;
#pragma unroll 1
                    for (int rep = 0; rep < 2; ++rep) {
                        seb_fluxes(p, a, prescribed ? a.Tskin_in : Ts, Egnd, swu, lwu, rnet, hs, hl, G);
                        if (!prescribed) {
                            Ts = T - G * dzc[nz] / (2 * p.kappa_skin);   // ImplicitSkinTemperature, skin_temperature.jl:62-68,138-150
                            seb_fluxes(p, a, Ts, Egnd, swu, lwu, rnet, hs, hl, G);
                        }
                    }
                    A.Egnd[c] = Egnd; A.infil[c] = inf; A.runoff[c] = runoff;
                    A.SWup[c] = swu; A.LWup[c] = lwu; A.Rnet[c] = rnet; A.Hs[c] = hs; A.Hl[c] = hl; A.G[c] = G;
                    if (!prescribed) A.Ts[c] = Ts;
                }
                sCol[lane] = G; sCol[TILE_COLS + lane] = inf;
            }
        }
    }
    __syncthreads();

    // ---- phase 2: fluxes, tendencies, explicit step ----
    auto kc_at = [&](int k) -> NF { return sKc[(k - 1) * TILE_COLS + lane]; };
    auto kf_at = [&](int k) -> NF {   // face conductivity Kf[k], soil_hydrology.jl:249-276 (k is warp uniform)
        if (k <= 0 || k >= nz + 2) return NF(0);           // halo faces are never written by the reference
        if (k == 1) return kc_at(1);
        if (k >= nz) return kc_at(nz);                     // Kf[Nz] = Kc[Nz], Kf[Nz+1] = Kf[Nz]
        return Mx::mn(kc_at(k), kc_at(k - 1));
    };
    int flagged = 0;
    for (int k = warp + 1; k <= nz; k += nwarp) {
        const int r = (k - 1) * TILE_COLS + lane, rh = k * TILE_COLS + lane;
        const NF Tm = sT[rh - TILE_COLS], T0 = sT[rh], Tp = sT[rh + TILE_COLS];
        const NF km = sKap[rh - TILE_COLS], k0 = sKap[rh], kp = sKap[rh + TILE_COLS];
        // diffusive_heat_flux at faces k and k+1, soil_energy.jl:134-149
        const NF qh_lo = -((k0 + km) / 2) * ((T0 - Tm) * rdzf[k]);
        const NF qh_hi = -((kp + k0) / 2) * ((Tp - T0) * rdzf[k + 1]);
        NF tU = -((qh_hi - qh_lo) * rdzc[k]);                                   // soil_energy.jl:112-131
        NF tS = NF(0);
        if (RICH) {
            const NF Pm = sP[rh - TILE_COLS], P0 = sP[rh], Pp = sP[rh + TILE_COLS];
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
            if (LAND) { tU -= sCol[lane] / dzc[nz]; tS -= (-sCol[TILE_COLS + lane]) / dzc[nz]; }   // land_model.jl:56-62
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
        sU[r] = sU[r] + tU * dt;
        if (RICH) {
            const NF sn = sS[r] + tS * dt;
            sS[r] = sn;
            if (!(sn <= 1) || sn < 0) flagged = 1;                       // needs adjust_saturation_profile! (also catches NaN)
            else if (sn < 1) atomicMin(&sIdx[lane], k);                  // compute_water_table!: lowest unsaturated layer
        }
    }
    if (RICH) {
        // ---- phase 3: adjust_saturation_profile! + compute_water_table! for tiles that need it ----
        const int any = __syncthreads_or(flagged);
        if (any) {
            if (flagged) sFlag[lane] = 1;   // benign race: every writer stores 1
            __syncthreads();
            if (warp == 0 && sFlag[lane]) {
                // upward sweep, soil_hydrology.jl:192-199
                for (int k = 1; k <= nz - 1; ++k) {
                    NF s = sS[(k - 1) * TILE_COLS + lane];
                    const NF e = jmax(s - 1, NF(0));
                    sS[(k - 1) * TILE_COLS + lane] = s - e;
                    sS[k * TILE_COLS + lane] += e * dzc[k] / dzc[k + 1];
                }
                // downward sweep, :201-208
                for (int k = nz; k >= 2; --k) {
                    NF s = sS[(k - 1) * TILE_COLS + lane];
                    const NF d = jmax(-s, NF(0));
                    sS[(k - 1) * TILE_COLS + lane] = s + d;
                    sS[(k - 2) * TILE_COLS + lane] -= d * dzc[k] / dzc[k - 1];
                }
                // top excess -> surface_excess_water, :210-216
                NF st = sS[(nz - 1) * TILE_COLS + lane];
                const NF e = jmax(st - 1, NF(0));
                sS[(nz - 1) * TILE_COLS + lane] = st - e;
                sCol[lane] = e * dzc[nz];   // G_top is dead by now: reuse the slot for the excess
                sS[lane] = jmax(sS[lane], NF(0));
                int idx = nz + 1;
                for (int k = nz; k >= 1; --k) if (sS[(k - 1) * TILE_COLS + lane] < 1) idx = k;
                sIdx[lane] = idx;
            }
            __syncthreads();
        }
    }

    // ---- phase 4: water table, closures of the new state, stores ----
    NF wt_new = NF(0);
    if (RICH) {
        const int idx = sIdx[lane];
        wt_new = zF[idx];                      // zF[nz+1] when every layer is saturated (halo cell / fallback agree)
        if (warp == 0 && valid) {
            A.yWt[c] = wt_new;
            NF Sx = A.bSx[c] + NF(0) * dt;     // surface_excess_water tendency is zero (soil_hydrology.jl:260-267)
            if (sFlag[lane]) Sx += sCol[lane];
            A.ySx[c] = Sx;
        }
    }
    for (int k = warp + 1; k <= nz; k += nwarp) {
        const int r = (k - 1) * TILE_COLS + lane;
        const NF Un = sU[r], sn = sS[r];
        NF Tn, ln;
        energy_to_temperature<NF, FAST>(p, Un, sn, Tn, ln);
        if (valid) {
            const int64_t o = (int64_t)(k - 1) * ld + c;
            A.yU[o] = Un; A.yT[o] = Tn; A.yL[o] = ln;
            if (RICH) { A.yS[o] = sn; A.yP[o] = pressure_head<NF, FAST>(p, sn, wt_new, zC[k], zref); }
        }
    }
}

}  // namespace trm
