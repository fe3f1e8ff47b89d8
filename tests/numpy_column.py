"""A second, independently written restatement of the coupled soil energy + Richards ForwardEuler step — plain numpy,
vectorised over columns, written from the reference sources and SURVEY Appendix A without looking at
``oracle/terrarium_oracle.cpp`` — used only to cross-check the C++ oracle (tests/test_numpy_cross_check.py). Two
restatements that agree to rounding do not prove either right, but a transcription slip in one of them shows up.

Arrays are ``[layer, column]`` with layer 0 = bottom cell (numerical_core.md:21-22). Float64 only. Scope: SoilModel,
RichardsEq or NoFlow hydrology, van Genuchten retention curve, UnsatKVanGenuchten or UnsatKLinear, Value boundary condition
on the surface temperature, default (zero-flux) boundaries elsewhere, reference default parameters unless passed."""
import numpy as np

EPS = np.finfo(np.float64).eps


class Column:
    def __init__(self, z_faces, T, sat, richards=True, alpha=2.0, n=2.0, K_sat=1e-5, unsat="vg", impedance=7.0, porosity=0.49,
                 sat_halo_copy=None):
        self.zf = np.asarray(z_faces, dtype=np.float64)
        self.dz = np.diff(self.zf)[:, None]                                    # Δz[k]
        self.zc = (0.5 * (self.zf[:-1] + self.zf[1:]))[:, None]
        dzf = np.empty(self.zf.size)                                           # Δzf[k], k = 0..Nz (faces), halo extension at the ends
        dzf[1:-1] = self.zc[1:, 0] - self.zc[:-1, 0]
        dzf[0], dzf[-1] = self.dz[0, 0], self.dz[-1, 0]
        self.dzf = dzf[:, None]
        self.richards, self.alpha, self.n, self.K_sat, self.unsat, self.omega, self.por = richards, alpha, n, K_sat, unsat, impedance, porosity
        # NoFlow: the saturation halo is never filled (zero) in the reference; Richards: copy
        self.sat_halo_copy = richards if sat_halo_copy is None else sat_halo_copy
        self.kappa = dict(water=0.57, ice=2.2, air=0.025, mineral=3.8, organic=0.25)
        self.cap = dict(water=4.2e6, ice=1.9e6, air=1.25e3, mineral=2.0e6, organic=2.5e6)
        self.L = 1000.0 * 3.34e5
        self.t = 0.0
        self.sat = np.array(sat, dtype=np.float64)
        self.S_excess = np.zeros(self.sat.shape[1])
        self.T = np.array(T, dtype=np.float64)
        # initialize!: hydrology closure first (adjusted saturation, water table, pressure head), then T -> U (soil_model_init order)
        if richards:
            self.hydrology_closure()
        self.liq = np.where(self.T >= 0, 1.0, 0.0)                             # soil_energy_closures.jl:64-97
        self.U = self.T * self.heat_capacity(self.sat, self.liq) - self.L * self.sat * self.por * (1.0 - self.liq)

    # -- constituents (soil_volume.jl:52-67; rho_soc = 0: no organic solid) -------------------------------------------
    def fractions(self, sat, liq):
        wi = sat * self.por
        return dict(water=wi * liq, ice=wi * (1.0 - liq), air=(1.0 - sat) * self.por, mineral=(1.0 - self.por) * 1.0, organic=(1.0 - self.por) * 0.0)

    def heat_capacity(self, sat, liq):
        f = self.fractions(sat, liq)
        return sum(self.cap[k] * f[k] for k in ("water", "ice", "air", "mineral", "organic"))

    def conductivity(self, sat, liq):                                          # soil_thermal_properties.jl:90-123
        f = self.fractions(sat, liq)
        return sum(np.sqrt(self.kappa[k]) * f[k] for k in ("water", "ice", "air", "mineral", "organic")) ** 2

    # -- hydrology ------------------------------------------------------------------------------------------------------
    def cell_conductivity(self, sat, liq):                                     # soil_hydraulic_properties.jl:170-221
        f = self.fractions(sat, liq)
        if self.unsat == "linear":
            return self.K_sat * f["water"] / (f["water"] + f["ice"] + f["air"])
        x = f["water"] / self.por
        imp = 10.0 ** (-self.omega * (1.0 - liq))
        n = self.n
        return np.abs(self.K_sat * imp * np.sqrt(x) * (1.0 - (1.0 - x ** (n / (n + 1.0))) ** ((n - 1.0) / n)) ** 2)

    def face_conductivity(self):                                               # soil_hydrology.jl:145-163
        Kc = self.cell_conductivity(self.sat, self.liq)
        nz = Kc.shape[0]
        Kf = np.zeros((nz + 3, Kc.shape[1]))                                   # index j = face k (1-based) ; 0 and nz+2 stay zero
        Kf[1] = Kc[0]
        for k in range(2, nz):
            Kf[k] = np.minimum(Kc[k - 1], Kc[k - 2])
        Kf[nz] = Kc[nz - 1]
        Kf[nz + 1] = Kf[nz]
        return Kf

    def hydrology_closure(self):                                               # soil_hydrology.jl:185-219, :170-175, closures :102-129
        sat, dz, nz = self.sat, self.dz, self.sat.shape[0]
        for k in range(nz - 1):
            e = np.maximum(sat[k] - 1.0, 0.0)
            sat[k] -= e
            sat[k + 1] += e * dz[k, 0] / dz[k + 1, 0]
        for k in range(nz - 1, 0, -1):
            d = np.maximum(-sat[k], 0.0)
            sat[k] += d
            sat[k - 1] -= d * dz[k, 0] / dz[k - 1, 0]
        e = np.maximum(sat[-1] - 1.0, 0.0)
        sat[-1] -= e
        self.S_excess = self.S_excess + e * dz[-1, 0]
        sat[0] = np.maximum(sat[0], 0.0)
        below = sat < 1.0
        first = np.where(below.any(axis=0), below.argmax(axis=0), nz)          # first layer (from the bottom) that is not saturated
        self.water_table = self.zf[first]
        m = 1.0 - 1.0 / self.n
        with np.errstate(divide="ignore", invalid="ignore"):
            psi_m = np.where(sat < 1.0, -(1.0 / self.alpha) * (sat ** (-1.0 / m) - 1.0) ** (1.0 / self.n), 0.0)
        self.psi = np.maximum(0.0, self.water_table[None, :] - self.zc) + psi_m + (self.zc - self.zf[-1])

    # -- tendencies at the current state (compute_tendencies!) ----------------------------------------------
    def tendencies(self, T_top):
        sat, liq, T, dz, dzf = self.sat, self.liq, self.T, self.dz, self.dzf
        nz = sat.shape[0]
        dsat = np.zeros_like(sat)
        if self.richards:
            Kf = self.face_conductivity()
            psi_h = np.vstack([self.psi[:1], self.psi, self.psi[-1:]])       # zero-flux halos
            g = (psi_h[1:] - psi_h[:-1]) / dzf                                 # faces 1..nz+1 -> rows 0..nz
            q = np.empty_like(g)
            for j in range(nz + 1):
                k = j + 1
                Kstar = np.where(g[j] < 0, np.minimum(Kf[k - 1], Kf[k]), np.minimum(Kf[k], Kf[k + 1]))
                q[j] = -Kstar * g[j]
            dsat = (-(q[1:] - q[:-1]) / dz) / self.por
        # energy tendency with halos: T by boundary condition, liq copied, sat copied (Richards) or zero (NoFlow)
        T_h = np.vstack([T[:1], T, 2.0 * np.asarray(T_top, dtype=np.float64)[None, :] - T[-1:]])
        liq_h = np.vstack([liq[:1], liq, liq[-1:]])
        sat_h = np.vstack([sat[:1], sat, sat[-1:]]) if self.sat_halo_copy else np.vstack([0 * sat[:1], sat, 0 * sat[-1:]])
        kc = self.conductivity(sat_h, liq_h)
        kf = 0.5 * (kc[1:] + kc[:-1])
        qh = -kf * (T_h[1:] - T_h[:-1]) / dzf
        dU = -(qh[1:] - qh[:-1]) / dz
        return dsat, dU

    # -- explicit_step! + closures (abstract_timestepper.jl:65-141, soil closures) --------------------------------------------
    def advance(self, dt, dsat, dU):
        if self.richards:
            self.sat = self.sat + dt * dsat
            self.hydrology_closure()
        self.U = self.U + dt * dU
        Lt = self.L * self.sat * self.por
        with np.errstate(divide="ignore", invalid="ignore"):
            frac = np.where(Lt == 0, np.inf, self.U / (-Lt + EPS))            # safediv (utils.jl:25)
            self.liq = np.where(self.U >= 0, 1.0, np.where(self.U >= -Lt, 1.0 - frac, 0.0))
        C = self.heat_capacity(self.sat, self.liq)
        self.T = np.where(self.U < -Lt, (self.U + Lt) / C, np.where(self.U >= 0, self.U / C, 0.0))
        self.t += dt

    def step(self, dt, T_top):
        """ForwardEuler (forward_euler.jl:19-31; order of SURVEY A.10)."""
        self.advance(dt, *self.tendencies(T_top))

    def heun_step(self, dt, T_top_now, T_top_next):
        """Heun (heun.jl:37-71): tendencies at the state, an Euler stage with its closures, tendencies of the stage at
        t + dt (its own inputs), the state advanced with the averaged tendencies."""
        import copy
        k1 = self.tendencies(T_top_now)
        stage = copy.copy(self)
        stage.sat, stage.U, stage.S_excess = self.sat.copy(), self.U.copy(), self.S_excess.copy()
        stage.advance(dt, *k1)
        k2 = stage.tendencies(T_top_next)
        self.advance(dt, (k1[0] + k2[0]) / 2, (k1[1] + k2[1]) / 2)


class LandColumn(Column):
    """Bare-ground ``LandModel`` (land_model.jl:45-108): the soil column above with the surface energy balance
    (surface_energy_balance.jl:60-144, radiative_fluxes.jl, turbulent_fluxes.jl, skin_temperature.jl:62-80), bare-ground
    evaporation (bare_ground_evaporation.jl:49-62) and direct runoff / infiltration (direct_surface_runoff.jl:87-117),
    coupled to the soil through Flux boundary conditions on ``internal_energy`` (G) and ``saturation_water_ice``
    (-infiltration), land_model.jl:56-62. Reference default parameters."""
    rho_a, c_a, Llg, sigma, eps_mw, Tref = 1.293, 1005.7, 2.257e6, 5.6704e-8, 0.622, 273.15
    albedo, emissivity, kappa_skin, C_h, V_min, tau_r, beta = 0.3, 0.97, 2.0, 1.2e-3, 0.01, 3600.0, 1.0

    def __init__(self, z_faces, T, sat, Ts, **kw):
        super().__init__(z_faces, T, sat, richards=True, **kw)
        self.Ts = np.array(Ts, dtype=np.float64)

    @staticmethod
    def e_sat(T):
        return np.where(T <= 0, 611.0 * np.exp(22.46 * T / (T + 272.62)), 611.0 * np.exp(17.62 * T / (T + 243.12)))

    def humidity_deficit(self, Ts, q, p):
        e_air = q * p / (self.eps_mw + (1 - self.eps_mw) * q)
        vpd = np.maximum(self.e_sat(Ts) - e_air, 0.1)
        return self.eps_mw * vpd / p

    def surface(self, f):
        """compute_auxiliary!(state, LandModel) after the soil hydraulics: surface hydrology, then the energy balance twice."""
        r_a = 1.0 / (self.C_h * np.maximum(np.maximum(f["V"], self.V_min), 1.0e-6))
        Kf = self.face_conductivity()
        nz = self.sat.shape[0]
        K_top, sat_top = Kf[nz], self.sat[-1]
        self.E = self.beta * self.humidity_deficit(self.Ts, f["q"], f["p"]) / r_a
        wet = self.S_excess > 0
        drain = np.where(wet, np.maximum(self.S_excess, 0.0) / self.tau_r, 0.0)
        influx = np.where(wet, drain, f["rain"])
        self.infiltration = np.minimum(influx, K_top) * (sat_top < 1.0)
        self.runoff = f["rain"] + drain - self.infiltration
        Tg = self.T[-1]
        for _ in range(2):                                 # land_model.jl:85-86
            for update_skin in (True, False):              # fluxes, skin temperature, fluxes again (implicit skin temperature)
                SW_up = self.albedo * f["SW"]
                LW_up = self.emissivity * self.sigma * (self.Ts + self.Tref) ** 4 + (1 - self.emissivity) * f["LW"]
                R_net = SW_up - f["SW"] + LW_up - f["LW"]
                H_s = self.c_a * self.rho_a * ((self.Ts - f["Ta"]) / r_a)
                H_l = self.Llg * self.rho_a * self.E
                self.G = R_net - H_s - H_l
                if update_skin:
                    self.Ts = Tg - self.G * self.dz[-1, 0] / (2 * self.kappa_skin)
        self.R_net, self.H_s, self.H_l = R_net, H_s, H_l

    def land_step(self, dt, f):
        """ForwardEuler: update_state! (auxiliaries, tendencies) and explicit_step! with the coupling Flux BCs."""
        self.surface(f)
        dsat, dU = self.tendencies(self.T[-1])             # default (zero-flux) temperature halo: 2 T_top - T_top
        dU[-1] -= self.G / self.dz[-1, 0]
        dsat[-1] -= (-self.infiltration) / self.dz[-1, 0]
        self.advance(dt, dsat, dU)

    def _coupled(self, dsat, dU):
        """compute_z_bcs! at explicit_step! time: the Flux BCs take the current ground heat flux / infiltration fields."""
        dsat, dU = dsat.copy(), dU.copy()
        dU[-1] -= self.G / self.dz[-1, 0]
        dsat[-1] -= (-self.infiltration) / self.dz[-1, 0]
        return dsat, dU

    def land_heun_step(self, dt, f_now, f_next):
        """Heun (heun.jl:37-71): the boundary fluxes are added when a state is stepped, from that state's own fields, so the
        final step uses the fluxes evaluated at time t (they are not averaged); the skin temperature of the state is the one
        of its own ``compute_auxiliary!`` at time t."""
        import copy
        self.surface(f_now)
        k1 = self.tendencies(self.T[-1])
        stage = copy.copy(self)
        stage.sat, stage.U, stage.S_excess, stage.Ts = self.sat.copy(), self.U.copy(), self.S_excess.copy(), self.Ts.copy()
        stage.advance(dt, *stage._coupled(*k1))
        stage.surface(f_next)
        k2 = stage.tendencies(stage.T[-1])
        self.advance(dt, *self._coupled((k1[0] + k2[0]) / 2, (k1[1] + k2[1]) / 2))
