/*
 * terrarium_b200.h -- C ABI of the B200-native per-column land time-step.
 *
 * This is the drop-in boundary for ONE hot path of Terrarium.jl: everything that runs inside
 *   timestep!(integrator::ModelIntegrator, ::ForwardEuler | ::Heun, dt)
 *   (reference: src/timesteppers/forward_euler.jl:19-31, src/timesteppers/heun.jl:37-71),
 * i.e. update_state! (src/state_variables.jl:72-80), explicit_step!
 * (src/timesteppers/abstract_timestepper.jl:65-77) and closure! (src/models/soil/soil_model.jl:51-54,
 * src/models/coupled/land_model.jl:98-102) for SoilModel and LandModel (bare ground, or with the PALADYN
 * vegetation / canopy processes of VegetationCarbon) on a ColumnGrid / ColumnRingGrid.  The reference has no FFI of its own for this path (it is pure Julia
 * dispatching KernelAbstractions kernels through Oceananigans' launch!, src/grids/grid_utils.jl:2-6);
 * the entry points below are what a `ccall` based Julia method of timestep!/run!/initialize would bind
 * (see INTEGRATION.md and terrarium.jl_b200/julia/TerrariumB200.jl).
 *
 * Conventions
 *  - every function returns an int status (TRM_OK == 0); a human readable message for the last
 *    failure on the calling thread is available from trm_last_error();
 *  - no exception, no torch/ATen type, no C++ type crosses this boundary: plain pointers and sizes;
 *  - a handle owns all of its device memory; pointers handed out by trm_field_ptr are borrowed;
 *  - calls on one handle must be serialised by the caller; work is stream ordered on the handle's
 *    stream, trm_sync() blocks until it is complete;
 *  - one handle drives one GPU and one contiguous column range [col0, col0+ncol) of the domain.
 *    Columns never exchange data (reference: only d/dz operators, src/Terrarium.jl:27), so a
 *    multi-GPU run is N handles (one process per GPU) with no halo exchange;
 *  - memory layout of every 3-D field is SoA [layer][column], column fastest, layer 0 = BOTTOM
 *    cell (reference convention, docs/src/introduction/numerical_core.md:21-22), leading dimension
 *    `ld` >= ncol padded to 256 bytes so that rows are 128-bit vector aligned;
 *  - all floating point parameters are passed as double and converted once to the handle's
 *    number format NF (float or double), the way Julia constructs `Struct{NF}` from literals.
 *
 * The identical ABI (prefix orc_ instead of trm_) is exported by the CPU oracle in oracle/, which
 * is test infrastructure only and is never loaded by the product path.
 */
#ifndef TERRARIUM_B200_H
#define TERRARIUM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRM_ABI_VERSION 4
#define TRM_MAX_NZ 128       /* per-column layers supported by the fused kernels            */
#define TRM_NUM_USER_INPUTS 8

/* ---- status codes ---------------------------------------------------------------------- */
enum trm_status {
    TRM_OK = 0,
    TRM_ERR_INVALID = 1,      /* bad argument / configuration                                */
    TRM_ERR_CUDA = 2,         /* CUDA runtime failure (message has the CUDA error string)    */
    TRM_ERR_STATE = 3,        /* call sequence error (e.g. step before initialize)           */
    TRM_ERR_UNSUPPORTED = 4,  /* valid in the reference but not built here                   */
    TRM_ERR_NO_DEVICE = 5     /* no usable CUDA device: the product path has NO CPU fallback */
};

/* ---- enumerations ---------------------------------------------------------------------- */
enum trm_dtype       { TRM_F32 = 0, TRM_F64 = 1 };
/* SoilModel (src/models/soil/soil_model.jl:9-59), LandModel with vegetation = nothing
 * (src/models/coupled/land_model.jl:111-125: BareGroundEvaporation + NoCanopyInterception) on either soil
 * hydrology (its default soil has immobile water, land_model.jl:111-112). */
enum trm_model       { TRM_MODEL_SOIL = 0, TRM_MODEL_LAND = 1 };
enum trm_timestepper { TRM_EULER = 0, TRM_HEUN = 1 };
/* src/processes/soil/hydrology/soil_hydrology.jl:13 (NoFlow), soil_hydrology_rre.jl:18 (RichardsEq) */
enum trm_hydrology   { TRM_NOFLOW = 0, TRM_RICHARDS = 1 };
/* FreezeCurves.jl SWRC types used at soil_hydraulic_closures.jl:95-97,115-118 */
enum trm_swrc        { TRM_SWRC_VANGENUCHTEN = 0, TRM_SWRC_BROOKSCOREY = 1 };
/* src/processes/soil/hydrology/soil_hydraulic_properties.jl:163-221 */
enum trm_unsat_k     { TRM_UNSATK_LINEAR = 0, TRM_UNSATK_VANGENUCHTEN = 1 };
/* What the never-filled z-halo of the auxiliary saturation field holds under NoFlow hydrology
 * (unpinned Oceananigans `set!` semantics). Richards always copies. */
enum trm_halo        { TRM_HALO_ZERO = 0, TRM_HALO_COPY = 1 };
/* Skin temperature scheme, src/processes/surface_energy/skin_temperature.jl:12,52 */
enum trm_skin        { TRM_SKIN_IMPLICIT = 0, TRM_SKIN_PRESCRIBED = 1 };
/* LandModel(vegetation = ...): nothing -> BareGroundEvaporation + NoCanopyInterception (land_model.jl:114-125);
 * VegetationCarbon (src/processes/vegetation/vegetation_carbon.jl:6-66: LUEPhotosynthesis, MedlynStomatalConductance,
 * PALADYNAutotrophicRespiration, PALADYNPhenology, PALADYNCarbonDynamics, PALADYNVegetationDynamics,
 * StaticExponentialRootDistribution, FieldCapacityLimitedPAW) with the default SurfaceHydrology
 * (PALADYNCanopyInterception + PALADYNCanopyEvapotranspiration + DirectSurfaceRunoff, surface_hydrology.jl:22-29). */
enum trm_vegetation  { TRM_VEG_NONE = 0, TRM_VEG_CARBON = 1 };
/* Ground evaporation resistance factor, src/processes/surface_hydrology/evapotranspiration/ground_resistance_factor.jl:
 * ConstantEvaporationResistanceFactor (params.evap_beta, :6-11) or SoilMoistureResistanceFactor (Lee & Pielke 1992, :32-57:
 * (1 - cos(pi theta_w / theta_fc))^2 / 4 below field capacity, 1 above; theta_w of the top soil layer). */
enum trm_ground_resistance { TRM_GROUND_RES_CONSTANT = 0, TRM_GROUND_RES_SOIL_MOISTURE = 1 };
/* Schemes of the surface energy balance (src/processes/surface_energy/): ConstantAlbedo (params.albedo / emissivity) or
 * PrescribedAlbedo (inputs TRM_IN_ALBEDO / TRM_IN_EMISSIVITY, albedo.jl:7-44); DiagnosedRadiativeFluxes (radiative_fluxes.jl:
 * 72-209) or PrescribedRadiativeFluxes (outgoing short / longwave radiation are inputs, :13-67); DiagnosedTurbulentFluxes
 * (turbulent_fluxes.jl:20-150) or PrescribedTurbulentFluxes (sensible / latent heat flux are inputs, :9-16). With a prescribed
 * scheme the fields of the same name hold the input values; net radiation and ground heat flux are still diagnosed
 * (R_net = SW_up - SW_down + LW_up - LW_down, G = R_net - H_s - H_l). */
enum trm_albedo_kind    { TRM_ALBEDO_CONSTANT = 0, TRM_ALBEDO_PRESCRIBED = 1 };
enum trm_radiative_kind { TRM_RADIATIVE_DIAGNOSED = 0, TRM_RADIATIVE_PRESCRIBED = 1 };
enum trm_turbulent_kind { TRM_TURBULENT_DIAGNOSED = 0, TRM_TURBULENT_PRESCRIBED = 1 };
/* Arithmetic contract of the CUDA kernels.
 *  FAITHFUL: same operations in the same order as the reference/oracle (true divisions,
 *            pow where the reference calls ^), no FMA contraction.
 *  FAST:     algebraically identical but uses reciprocal metrics, cbrt/sqrt/rsqrt special cases
 *            for van Genuchten n = 2 and FMA contraction. Results differ from FAITHFUL by
 *            rounding only (tests pin the tolerance).                                        */
enum trm_math        { TRM_MATH_FAITHFUL = 0, TRM_MATH_FAST = 1 };

/* Boundary conditions (Oceananigans Value/Gradient/Flux semantics). */
enum trm_bc_kind { TRM_BC_DEFAULT = 0 /* zero flux */, TRM_BC_VALUE = 1, TRM_BC_GRADIENT = 2, TRM_BC_FLUX = 3 };
enum trm_bc_slot {
    TRM_BC_TEMPERATURE_TOP = 0,   /* PrescribedSurfaceTemperature, src/models/soil/soil_model_bcs.jl:17 */
    TRM_BC_TEMPERATURE_BOTTOM = 1,/* PrescribedBottomTemperature, :22                                    */
    TRM_BC_ENERGY_TOP = 2,        /* GroundHeatFlux (Flux on internal_energy), :6                        */
    TRM_BC_ENERGY_BOTTOM = 3,     /* GeothermalHeatFlux, :12                                             */
    TRM_BC_SATURATION_TOP = 4,    /* InfiltrationFlux, :29 (value is the flux itself, positive upward)   */
    TRM_BC_SATURATION_BOTTOM = 5, /* ImpermeableBoundary, :34                                            */
    TRM_BC_PRESSURE_TOP = 6,
    TRM_BC_PRESSURE_BOTTOM = 7,   /* FreeDrainage = Gradient(0), :40                                     */
    TRM_BC_NSLOTS = 8
};

/* Per-column 2-D input variables (src/input_output/input_sources.jl; atmosphere inputs of
 * src/processes/atmosphere/prescribed_atmosphere.jl:89-99,147-149,192-195,220-224). */
enum trm_input_id {
    TRM_IN_USER0 = 0,  /* .. TRM_IN_USER0 + TRM_NUM_USER_INPUTS-1: BC inputs such as T_ub */
    TRM_IN_AIR_TEMPERATURE = 8,
    TRM_IN_AIR_PRESSURE = 9,
    TRM_IN_WINDSPEED = 10,
    TRM_IN_SPECIFIC_HUMIDITY = 11,
    TRM_IN_RAINFALL = 12,
    TRM_IN_SNOWFALL = 13,
    TRM_IN_SHORTWAVE_DOWN = 14,
    TRM_IN_LONGWAVE_DOWN = 15,
    TRM_IN_DAYTIME_LENGTH = 16,
    TRM_IN_CO2 = 17,
    TRM_IN_SKIN_TEMPERATURE = 18, /* only read with TRM_SKIN_PRESCRIBED */
    TRM_IN_SAI = 19,                    /* stem area index, canopy_interception.jl:54 (no default: 0)          */
    TRM_IN_DAILY_LEAF_RESPIRATION = 20, /* autotrophic_respiration.jl:33 (no default: 0)                       */
    /* prescribed variants of the surface energy balance (read only when the corresponding scheme is selected) */
    TRM_IN_ALBEDO = 21,                 /* PrescribedAlbedo, src/processes/surface_energy/albedo.jl:7-14        */
    TRM_IN_EMISSIVITY = 22,
    TRM_IN_SHORTWAVE_UP = 23,           /* PrescribedRadiativeFluxes, radiative_fluxes.jl:13-23                 */
    TRM_IN_LONGWAVE_UP = 24,
    TRM_IN_SENSIBLE_HEAT_FLUX = 25,     /* PrescribedTurbulentFluxes, turbulent_fluxes.jl:9-16                  */
    TRM_IN_LATENT_HEAT_FLUX = 26,
    TRM_IN_COUNT = 27
};
/* How an input is produced at clock time t (device resident, evaluated inside the stage kernel). */
enum trm_source {
    TRM_SRC_CONST = 0,     /* one scalar for all columns                                        */
    TRM_SRC_FIELD = 1,     /* per-column vector owned by the handle, set with trm_set_input_field */
    TRM_SRC_SINUSOID = 2,  /* clamp(mean[c] + amp[c]*sin(2*pi*t/period - phase[c]), lo, hi)      */
    TRM_SRC_TABLE = 3,     /* snapshots values[nt][ncol] at times[nt]; linear in time, flat outside
                              (Oceananigans FieldTimeSeries[Time(t)]: v2*n + v1*(1-n))                  */
    TRM_SRC_RASTER = 4,    /* same data, update rule of RasterInputSource (ext/TerrariumRastersExt/
                              TerrariumRastersExt.jl:96-121): x1 + eps*(x2-x1)/dt in Float64, the node value
                              on a node, flat outside the time axis                                      */
    TRM_SRC_FIELD_PAIR = 5 /* two per-column vectors for ONE step: the values at the step's start time t and at
                              t + dt (trm_set_input_field_pair). A function valued boundary condition f(x, t)
                              evaluated on the host: ForwardEuler and Heun stage 1 read the first vector, Heun stage
                              2 reads the second one where the reference re-evaluates the function at the stage
                              clock t + dt (Value / Gradient halos, heun.jl:53) and the first one for Flux BCs
                              (added with the time-n state, heun.jl:63-66)                               */
};

/* Fields that can be read / written / borrowed. 3-D fields are [nz][ld]; the hydraulic
 * conductivity is a z-face field [nz+1][ld]; 2-D fields are [ld]. */
enum trm_field_id {
    TRM_F_INTERNAL_ENERGY = 0,        /* prognostic, src/processes/soil/energy/soil_energy.jl:47       */
    TRM_F_TEMPERATURE = 1,            /* closure, soil_energy_closures.jl:22-25                        */
    TRM_F_LIQUID_WATER_FRACTION = 2,
    TRM_F_SATURATION_WATER_ICE = 3,   /* auxiliary (NoFlow) / prognostic (Richards)                    */
    TRM_F_PRESSURE_HEAD = 4,          /* closure, soil_hydraulic_closures.jl:14-16 (Richards only)     */
    TRM_F_HYDRAULIC_CONDUCTIVITY = 5, /* z-face auxiliary, soil_hydrology.jl:81                        */
    TRM_F_SURFACE_EXCESS_WATER = 6,   /* prognostic 2-D (Richards)                                     */
    TRM_F_WATER_TABLE = 7,
    TRM_F_GROUND_TEMPERATURE = 8,     /* read-only alias of the top temperature layer                  */
    TRM_F_SKIN_TEMPERATURE = 9,       /* LandModel 2-D fields from here                                */
    TRM_F_GROUND_HEAT_FLUX = 10,
    TRM_F_SHORTWAVE_UP = 11,
    TRM_F_LONGWAVE_UP = 12,
    TRM_F_NET_RADIATION = 13,
    TRM_F_SENSIBLE_HEAT_FLUX = 14,
    TRM_F_LATENT_HEAT_FLUX = 15,
    TRM_F_EVAPORATION_GROUND = 16,
    TRM_F_INFILTRATION = 17,
    TRM_F_SURFACE_RUNOFF = 18,
    TRM_F_TEND_INTERNAL_ENERGY = 19,  /* materialised only by trm_compute_tendencies (debug/tests)     */
    TRM_F_TEND_SATURATION = 20,
    /* vegetated LandModel (TRM_VEG_CARBON) ; all 2-D unless noted */
    TRM_F_CARBON_VEGETATION = 21,         /* prognostic, carbon_dynamics.jl:43                                 */
    TRM_F_VEGETATION_AREA_FRACTION = 22,  /* prognostic, vegetation_dynamics.jl:23                             */
    TRM_F_CANOPY_WATER = 23,              /* prognostic, canopy_interception.jl:48                             */
    TRM_F_BALANCED_LEAF_AREA_INDEX = 24,  /* auxiliaries from here */
    TRM_F_LEAF_AREA_INDEX = 25,
    TRM_F_PHENOLOGY_FACTOR = 26,
    TRM_F_CANOPY_WATER_CONDUCTANCE = 27,
    TRM_F_LEAF_TO_AIR_CO2_RATIO = 28,
    TRM_F_NET_ASSIMILATION = 29,          /* read back by the next evaluation of the stomatal conductance       */
    TRM_F_LEAF_RESPIRATION = 30,
    TRM_F_GROSS_PRIMARY_PRODUCTION = 31,
    TRM_F_AUTOTROPHIC_RESPIRATION = 32,
    TRM_F_NET_PRIMARY_PRODUCTION = 33,
    TRM_F_SOIL_MOISTURE_LIMITING_FACTOR = 34,
    TRM_F_CANOPY_WATER_INTERCEPTION = 35,
    TRM_F_CANOPY_WATER_REMOVAL = 36,
    TRM_F_SATURATION_CANOPY_WATER = 37,
    TRM_F_RAINFALL_GROUND = 38,
    TRM_F_EVAPORATION_CANOPY = 39,
    TRM_F_TRANSPIRATION = 40,
    TRM_F_PLANT_AVAILABLE_WATER = 41,     /* 3-D [nz][ld], materialised by trm_compute_auxiliary                */
    TRM_F_ROOT_FRACTION = 42,             /* 3-D static function of depth (root_distribution.jl:47-56): get only */
    TRM_F_COUNT = 43
};

/* ---- parameters ---------------------------------------------------------------------------
 * Defaults (trm_default_params) are the reference's package defaults. */
typedef struct trm_params {
    /* ConstantSoilPorosity, src/processes/soil/stratigraphy/soil_porosity.jl:7-13 */
    double mineral_porosity;      /* 0.49 */
    double organic_porosity;      /* 0.9  */
    /* ConstantSoilCarbonDensity, src/processes/soil/biogeochem/constant_soil_carbon.jl:10-16 */
    double rho_soc;               /* 0.0  */
    double rho_org;               /* 1300 */
    /* SoilThermalConductivities / SoilHeatCapacities, soil_thermal_properties.jl:13-45 ;
       order: water, ice, air, mineral, organic */
    double kappa[5];              /* 0.57 2.2 0.025 3.8 0.25 */
    double heatcap[5];            /* 4.2e6 1.9e6 1.25e3 2.0e6 2.5e6 */
    /* PhysicalConstants, src/processes/physical_constants.jl:9-51 */
    double rho_w;                 /* 1000  */
    double Lsl;                   /* 3.34e5 */
    double Llg;                   /* 2.257e6 */
    double rho_a;                 /* 1.293 */
    double c_a;                   /* 1005.7 */
    double Tref;                  /* 273.15 */
    double sigma;                 /* 5.6704e-8 */
    double eps_mw;                /* 0.622 (ratio of molecular weights) */
    /* Soil hydraulics, soil_hydraulic_properties.jl:62-76 and FreezeCurves SWRC parameters */
    double K_sat;                 /* 1e-5 */
    double vg_alpha;              /* VanGenuchten alpha [1/m] (FreezeCurves default 1.0) */
    double vg_n;                  /* VanGenuchten n (FreezeCurves default 2.0) */
    double bc_psis;               /* BrooksCorey air entry head [m] (0.01) */
    double bc_lambda;             /* BrooksCorey pore size index (0.2) */
    double theta_res;             /* residual water content (0.0) */
    double impedance;             /* UnsatKVanGenuchten ice impedance Omega (7) */
    double vwc_forcing;           /* constant user VWC forcing [1/s] added in every cell
                                     (soil_hydrology.jl:39, test/soil/soil_hydrology_tests.jl:191-233) */
    /* Surface energy balance */
    double albedo;                /* ConstantAlbedo 0.3, src/processes/surface_energy/albedo.jl:22 */
    double emissivity;            /* 0.97 */
    double kappa_skin;            /* ImplicitSkinTemperature kappa_s 2.0, skin_temperature.jl:54 */
    double C_h;                   /* ConstantAerodynamics 1.2e-3, atmosphere/aerodynamics.jl:8 */
    double min_windspeed;         /* 0.01, prescribed_atmosphere.jl:80 */
    /* Surface hydrology */
    double tau_r;                 /* DirectSurfaceRunoff 3600 s, runoff/direct_surface_runoff.jl:17 */
    double evap_beta;             /* ConstantEvaporationResistanceFactor 1.0 */
    /* ---- vegetated LandModel (TRM_VEG_CARBON) ---- */
    double field_capacity;        /* ConstantSoilHydraulics 0.25, soil_hydraulic_properties.jl:77 */
    double wilting_point;         /* 0.05, :80 */
    double C_mass;                /* PhysicalConstants 12.0 gC/mol, physical_constants.jl:50 */
    /* LUEPhotosynthesis, src/processes/vegetation/photosynthesis.jl:19-67 */
    double tau25, Kc25, Ko25, q10_tau, q10_Kc, q10_Ko;       /* 2600 30 3e4 0.57 2.1 1.2 */
    double alpha_leaf, alpha_a, alpha_C3, cq, k_ext;         /* 0.17 0.5 0.08 4.6e-6 0.5 */
    double T_CO2_high, T_CO2_low, T_photos_high, T_photos_low, theta_r;   /* 42 -4 30 15 0.7 */
    /* MedlynStomatalConductance, stomatal_conductance.jl:18-25 */
    double g1, g_min;             /* 2.3 0.5 */
    /* PALADYNAutotrophicRespiration, autotrophic_respiration.jl:16-25 */
    double cn_sapwood, cn_root, aws;                         /* 330 29 10 */
    /* PALADYNCarbonDynamics, carbon_dynamics.jl:18-39 */
    double SLA, awl, LAI_min, LAI_max, gamma_L, gamma_R, gamma_S;   /* 10 2 1 6 0.3 0.3 0.05 */
    /* PALADYNVegetationDynamics, vegetation_dynamics.jl:14-20 */
    double nu_seed, gamma_v_min;  /* 0.001 0.002 */
    /* StaticExponentialRootDistribution, root_distribution.jl:25-31 */
    double root_a, root_b;        /* 7 2 */
    /* PALADYNCanopyInterception, canopy_interception.jl:33-45 */
    double alpha_int, k_ext_can, w_can_max, tau_w;           /* 0.2 0.5 2e-4 86400 */
    /* PALADYNCanopyEvapotranspiration, canopy_evapotranspiration.jl:32-44 */
    double C_can;                 /* 0.006 */
} trm_params;

typedef struct trm_bc {
    int32_t kind;       /* trm_bc_kind */
    int32_t input;      /* trm_input_id that provides the value / gradient / flux per column */
} trm_bc;

typedef struct trm_config {
    int32_t abi_version;      /* must be TRM_ABI_VERSION */
    int32_t dtype;            /* trm_dtype */
    int64_t ncol;             /* columns owned by this handle */
    int64_t col0;             /* global index of the first owned column (informational) */
    int32_t nz;               /* layers, 1..TRM_MAX_NZ */
    int32_t device;           /* CUDA device ordinal */
    int32_t model;            /* trm_model */
    int32_t timestepper;      /* trm_timestepper */
    int32_t hydrology;        /* trm_hydrology */
    int32_t swrc;             /* trm_swrc */
    int32_t unsat_k;          /* trm_unsat_k */
    int32_t sat_halo;         /* trm_halo (NoFlow only) */
    int32_t skin;             /* trm_skin (LandModel only) */
    int32_t math;             /* trm_math */
    int32_t vegetation;       /* trm_vegetation (LandModel only) */
    int32_t ground_resistance;/* trm_ground_resistance (LandModel only) */
    int32_t albedo_kind;      /* trm_albedo_kind (LandModel only) */
    int32_t radiative;        /* trm_radiative_kind (LandModel only) */
    int32_t turbulent;        /* trm_turbulent_kind (LandModel only) */
    int32_t reserved0;        /* must be 0 */
    const double* z_faces;    /* nz+1 face elevations, bottom .. 0 (column_grid.jl:30-31); copied */
    trm_params params;
    trm_bc bc[TRM_BC_NSLOTS];
} trm_config;

/* Local (per handle) diagnostics; a multi-GPU caller reduces them with NCCL:
 * sums with ncclSum, minima with ncclMin, maxima with ncclMax. */
typedef struct trm_diag {
    double energy;        /* sum_c sum_k U*dz                 [J/m^2 summed over columns]  */
    double water;         /* sum_c (sum_k sat*por*dz + S_excess)  [m summed over columns]   */
    double t_min, t_max;  /* extrema of temperature */
    double sat_min, sat_max;
    double nan_count;     /* non-finite values in U, T, sat */
    double ncol;          /* columns contributing */
} trm_diag;

typedef struct trm_handle trm_handle;

/* ---- lifecycle ------------------------------------------------------------------------- */
void        trm_default_params(trm_params* p);
void        trm_default_config(trm_config* c);   /* SoilModel, Euler, NoFlow, f64, defaults */
int         trm_create(const trm_config* cfg, trm_handle** out);
int         trm_destroy(trm_handle* h);
const char* trm_last_error(void);
int         trm_abi_version(void);
int         trm_sync(trm_handle* h);

/* ---- layout / zero-copy access ---------------------------------------------------------
 * devptr: device pointer to element [0][0]; ld: leading dimension in elements;
 * nrows: nz, nz+1 or 1. The Julia wrapper unsafe_wrap()s these as CuArrays (interior(field)). */
int trm_field_ptr(trm_handle* h, int field_id, void** devptr, int64_t* ld, int32_t* nrows);
/* Same borrow for READING only (output gathers over NCCL, plotting): the library keeps treating its stored closure
 * fields as current, so the next step does not pay for re-reading them. Valid after trm_sync / a synchronous call. */
int trm_field_view(trm_handle* h, int field_id, const void** devptr, int64_t* ld, int32_t* nrows);
/* Host <-> device copies of whole fields. Host layout is dense [nrows][ncol] in the handle's
 * dtype (replaces set!(field, ...) / interior(field), src/initializers.jl:23-27). Writing
 * TEMPERATURE / SATURATION before trm_initialize sets the initial condition. */
int trm_set_field(trm_handle* h, int field_id, const void* host, int64_t count);
int trm_get_field(trm_handle* h, int field_id, void* host, int64_t count);

/* ---- inputs / forcing (device resident; replaces InputSource / FieldTimeSeriesInputSource,
 *      src/input_output/input_sources.jl:81-171, and function valued BCs) ----------------- */
int trm_set_input_const(trm_handle* h, int input_id, double value);
int trm_set_input_field(trm_handle* h, int input_id, const void* host_values /* [ncol] NF */);
/* Function valued boundary conditions (examples/simulations/soil_heat_global.jl:72-93) cannot run inside a CUDA kernel: the
 * host evaluates f(x, t) and f(x, t + dt) before each step and hands both over (TRM_SRC_FIELD_PAIR). */
int trm_set_input_field_pair(trm_handle* h, int input_id, const void* values_t /* [ncol] NF */, const void* values_t_plus_dt /* [ncol] NF */);
int trm_set_input_sinusoid(trm_handle* h, int input_id, const void* mean, const void* amp,
                           const void* phase /* each [ncol] NF */, double period,
                           double lo, double hi /* clamp; use -INFINITY/INFINITY for none */);
int trm_set_input_table(trm_handle* h, int input_id, int32_t nt, const double* times,
                        const void* values /* [nt][ncol] NF */);
/* Time-varying raster already gathered to the owned columns (idxmap = findall(mask), TerrariumRastersExt.jl:44):
 * same arguments as trm_set_input_table, evaluated with the RasterInputSource rule (TRM_SRC_RASTER). */
int trm_set_input_raster(trm_handle* h, int input_id, int32_t nt, const double* times,
                         const void* values /* [nt][ncol] NF */);
/* Borrow the per-column device vector of a TRM_SRC_FIELD input so that a coupled model
 * (e.g. SpeedyWeather, examples/simulations/speedy_dry_land.jl) can write forcing in place. */
int trm_input_ptr(trm_handle* h, int input_id, void** devptr);
/* Values of an input variable as of the last update_inputs! (start of the last step / trm_initialize): what reading the
 * input Field `state.<name>` returns in the reference (state_variables.jl:154-162). host: [ncol] NF. */
int trm_get_input(trm_handle* h, int input_id, void* host, int64_t count);

/* ---- model ------------------------------------------------------------------------------ */
/* initialize!(state, model) after the user initializers ran: hydrology closure + hydraulics,
 * then the inverse energy closure T -> U (src/processes/soil/soil_coupled.jl:45-54). */
int trm_initialize(trm_handle* h);
/* reset!(integrator.state) at the top of initialize!(integrator) (model_integrator.jl:98): zeroes every state, auxiliary,
 * stage and accumulator field and the clock. Call it BEFORE writing the initial conditions of a re-initialisation
 * (a fresh handle is already zero). Inputs / boundary sources and the grid are kept. */
int trm_reset(trm_handle* h);
/* nsteps x timestep!(integrator, dt; finalize = false) (forward_euler.jl:19-31, heun.jl:37-71):
 * one fused kernel launch per timestepper stage (ForwardEuler: one per step, Heun: two per step). */
int trm_step(trm_handle* h, double dt, int64_t nsteps);
/* compute_auxiliary!(state, model) (soil_model.jl:39-42, land_model.jl:79-88): what
 * timestep!(...; finalize = true) and run! do after stepping (model_integrator.jl:81-87,125-131). */
int trm_compute_auxiliary(trm_handle* h);
/* update_state!(...; compute_tendencies = true) with the tendencies (incl. flux BCs)
 * materialised in TRM_F_TEND_* (tests / debugging only, not on the hot path). */
int trm_compute_tendencies(trm_handle* h);

int trm_get_clock(trm_handle* h, double* time, int64_t* iteration);
/* (restart from a snapshot: set the fields, then the clock; the time is rounded to NF like the reference's clock and
 * also becomes the time of the last update_inputs!, so that diagnostics before the first step see the restart time) */
int trm_set_clock(trm_handle* h, double time, int64_t iteration);

/* ---- diagnostics ------------------------------------------------------------------------ */
int trm_diagnostics(trm_handle* h, trm_diag* out);
/* Same numbers left in device memory as 8 doubles in trm_diag order (for ncclAllReduce). */
int trm_diagnostics_device(trm_handle* h, double** dev_out);

/* ---- global diagnostics over NCCL (one process per GPU, one handle per process) ----------------------------------
 * The columns never exchange data; NCCL only serves the global diagnostic reductions (and output gathers, for which
 * trm_field_view hands the device buffers to the caller's own ncclAllGather / NCCL.jl). The library binds to the NCCL
 * the host process already has loaded (NCCL.jl's or torch's libnccl.so.2; dlsym first, dlopen("libnccl.so.2") otherwise),
 * so a communicator created by the host's own NCCL binding can be adopted, and a host without a binding can create one here:
 *   trm_nccl_get_unique_id  rank 0 fills 128 bytes (ncclUniqueId); the caller distributes them to the other ranks
 *                           (a file, MPI, Julia's Distributed, torch.distributed ...)
 *   trm_nccl_comm_init      ncclCommInitRank on the handle's device; the handle owns (and destroys) the communicator
 *   trm_nccl_comm_adopt     use an ncclComm_t the caller owns (must come from the same libnccl instance)
 *   trm_diagnostics_allreduce  local reduction (as trm_diagnostics) + two ncclAllReduce calls on the handle's stream:
 *                           energy, water, nan_count, ncol are summed, minima / maxima are combined; every rank gets the
 *                           global trm_diag. TRM_ERR_UNSUPPORTED when no NCCL library can be found. */
#define TRM_NCCL_UNIQUE_ID_BYTES 128
int trm_nccl_get_unique_id(void* id_bytes);
int trm_nccl_comm_init(trm_handle* h, int32_t nranks, int32_t rank, const void* id_bytes);
int trm_nccl_comm_adopt(trm_handle* h, void* nccl_comm);
int trm_diagnostics_allreduce(trm_handle* h, trm_diag* out);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t trm_launch_count(trm_handle* h);
/* Elapsed device time [ms] of the most recent trm_step call measured with CUDA events on the
 * handle's stream (the events bracket exactly the fused stage kernels of that call). */
int trm_last_step_ms(trm_handle* h, float* ms);

/* ---- asynchronous variants (coupled-model usage: forcing in, surface state out, every step) ------------
 * All three only enqueue work; trm_sync() waits for everything. Host buffers must be pinned and must stay
 * valid / untouched until the next trm_sync().
 *  trm_set_input_field_async: upload of a per-column input on a copy stream into the back half of a double
 *      buffer; steps enqueued AFTER the call read the new values, steps already enqueued keep the old ones.
 *  trm_step_async: trm_step without the final synchronisation (trm_last_step_ms is valid after trm_sync).
 *  trm_get_field_async: snapshot of the field after all steps enqueued so far (device-to-device on the compute
 *      stream), downloaded on a second copy stream while later steps run.                                  */
int trm_set_input_field_async(trm_handle* h, int input_id, const void* pinned_host_values /* [ncol] NF */);
int trm_step_async(trm_handle* h, double dt, int64_t nsteps);
int trm_get_field_async(trm_handle* h, int field_id, void* pinned_host, int64_t count);
/* Time-averaged output (Oceananigans AveragedTimeInterval / WindowedTimeAverage): trm_accumulate adds weight * field to
 * a per-field accumulator owned by the handle (allocated and zeroed on first use; weight = the step's dt);
 * trm_get_accumulated returns scale * accumulator (scale = 1 / window length) in the layout of trm_get_field and, if
 * `reset` is non-zero, zeroes it for the next window. */
int trm_accumulate(trm_handle* h, int field_id, double weight);
int trm_get_accumulated(trm_handle* h, int field_id, void* host, int64_t count, double scale, int32_t reset);
/* Page-locked host memory for the asynchronous entry points (cudaHostAlloc / cudaFreeHost), so that a caller without
 * a CUDA binding of its own (the Julia / Python host side) can own pinned buffers. */
int trm_host_alloc(int64_t bytes, void** host);
int trm_host_free(void* host);
/* Same with flags: TRM_HOST_WRITE_COMBINED for buffers the host only writes and the GPU only reads (forcing rings). */
enum trm_host_flags { TRM_HOST_DEFAULT = 0, TRM_HOST_WRITE_COMBINED = 1 };
int trm_host_alloc_ex(int64_t bytes, int32_t flags, void** host);

/* ---- per-step exchange with a host-side coupler through mapped host memory -------------------------------
 * Coupled-model usage (examples/simulations/speedy_dry_land.jl:45-68: every coupling step the atmosphere hands the land
 * model its forcing and reads the surface state back) without a copy per step: the stage kernel itself reads the
 * per-column input from, and writes the 2-D result field to, page-locked host memory that is mapped into the device's
 * address space (trm_host_alloc / trm_host_alloc_ex). No copy engine, no staging buffer, no event chain: a coupled
 * step is ONE call, trm_step_async(h, dt, 1).
 *   host_in  : nslots x ncol values (NF); the step that advances the clock from iteration k to k+1 reads the values of
 *              input `input_id` from slot k % nslots (a TRM_SRC_FIELD input: constant over the step, both Heun stages);
 *   host_out : nslots x ncol values (NF); the same step stores field `field_id` (any 2-D field of the model) of the
 *              NEW state in slot k % nslots. TRM_F_GROUND_TEMPERATURE is written by the stage kernel itself.
 * Either pointer may be NULL (one direction only); both NULL removes the binding. trm_host_io_wait(h, k) returns once
 * the step that produced iteration k has completed: its output slot may be read and its input slot refilled. The
 * binding starts at the handle's current iteration (trm_get_clock). */
int trm_bind_host_io(trm_handle* h, int input_id, const void* host_in, int field_id, void* host_out, int32_t nslots);
int trm_host_io_wait(trm_handle* h, int64_t iteration);

/* ---- ColumnRingGrid conversions (src/grids/column_ring_grid.jl:102-149) --------------------------------
 * The columns of a ColumnRingGrid are the `true` points of a mask over a ring grid of `nring` points, in ring
 * order. trm_set_ring_index hands over, for every owned column, its position in the ring grid (host array of
 * ncol int64). trm_get_field_ring scatters a field into a dense host array [nrows][nring] with `fill_value` at
 * the unmasked points (RingGrids.Field(field, grid; fill_value)); trm_set_field_ring gathers the masked points
 * of such an array into the field (Field(ring_field, grid)). Scatter / gather run on the device.           */
int trm_set_ring_index(trm_handle* h, const int64_t* ring_index, int64_t nring);
int trm_get_field_ring(trm_handle* h, int field_id, void* host_ring, int64_t count, double fill_value);
int trm_set_field_ring(trm_handle* h, int field_id, const void* host_ring, int64_t count);

/* Tuning knob: threads per block of the register-streaming stage kernel (multiple of 32 in [32, 128], default
 * 128). The shared-memory staged ForwardEuler kernel has a compile-time block size. */
int trm_set_block_size(trm_handle* h, int block);

#ifdef __cplusplus
}
#endif
#endif /* TERRARIUM_B200_H */
