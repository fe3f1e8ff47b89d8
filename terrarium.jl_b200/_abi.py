"""ctypes mirror of ``include/terrarium_b200.h`` (the C ABI of the per-column land time-step).

Only declarations live here: struct layouts, enum values and a ``bind`` helper that attaches
argument/return types to the entry points of a loaded shared library.  The same declarations
serve the product library (prefix ``trm_``) and -- in tests only -- the CPU checker under ``oracle/``, which
exports the identical ABI under its own prefix.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

TRM_ABI_VERSION = 4
TRM_MAX_NZ = 128
TRM_NUM_USER_INPUTS = 8

# status
TRM_OK, TRM_ERR_INVALID, TRM_ERR_CUDA, TRM_ERR_STATE, TRM_ERR_UNSUPPORTED, TRM_ERR_NO_DEVICE = range(6)
# enums
TRM_F32, TRM_F64 = 0, 1
TRM_MODEL_SOIL, TRM_MODEL_LAND = 0, 1
TRM_EULER, TRM_HEUN = 0, 1
TRM_NOFLOW, TRM_RICHARDS = 0, 1
TRM_SWRC_VANGENUCHTEN, TRM_SWRC_BROOKSCOREY = 0, 1
TRM_UNSATK_LINEAR, TRM_UNSATK_VANGENUCHTEN = 0, 1
TRM_HALO_ZERO, TRM_HALO_COPY = 0, 1
TRM_SKIN_IMPLICIT, TRM_SKIN_PRESCRIBED = 0, 1
TRM_MATH_FAITHFUL, TRM_MATH_FAST = 0, 1
TRM_VEG_NONE, TRM_VEG_CARBON = 0, 1
TRM_GROUND_RES_CONSTANT, TRM_GROUND_RES_SOIL_MOISTURE = 0, 1
TRM_BC_DEFAULT, TRM_BC_VALUE, TRM_BC_GRADIENT, TRM_BC_FLUX = 0, 1, 2, 3
(TRM_BC_TEMPERATURE_TOP, TRM_BC_TEMPERATURE_BOTTOM, TRM_BC_ENERGY_TOP, TRM_BC_ENERGY_BOTTOM,
 TRM_BC_SATURATION_TOP, TRM_BC_SATURATION_BOTTOM, TRM_BC_PRESSURE_TOP, TRM_BC_PRESSURE_BOTTOM) = range(8)
TRM_BC_NSLOTS = 8
TRM_IN_USER0 = 0
(TRM_IN_AIR_TEMPERATURE, TRM_IN_AIR_PRESSURE, TRM_IN_WINDSPEED, TRM_IN_SPECIFIC_HUMIDITY, TRM_IN_RAINFALL,
 TRM_IN_SNOWFALL, TRM_IN_SHORTWAVE_DOWN, TRM_IN_LONGWAVE_DOWN, TRM_IN_DAYTIME_LENGTH, TRM_IN_CO2,
 TRM_IN_SKIN_TEMPERATURE, TRM_IN_SAI, TRM_IN_DAILY_LEAF_RESPIRATION) = range(8, 21)
TRM_IN_ALBEDO, TRM_IN_EMISSIVITY, TRM_IN_SHORTWAVE_UP, TRM_IN_LONGWAVE_UP, TRM_IN_SENSIBLE_HEAT_FLUX, TRM_IN_LATENT_HEAT_FLUX = 21, 22, 23, 24, 25, 26
TRM_IN_COUNT = 27
TRM_SRC_CONST, TRM_SRC_FIELD, TRM_SRC_SINUSOID, TRM_SRC_TABLE, TRM_SRC_RASTER = 0, 1, 2, 3, 4

FIELD_IDS = {
    "internal_energy": 0, "temperature": 1, "liquid_water_fraction": 2, "saturation_water_ice": 3,
    "pressure_head": 4, "hydraulic_conductivity": 5, "surface_excess_water": 6, "water_table": 7,
    "ground_temperature": 8, "skin_temperature": 9, "ground_heat_flux": 10, "surface_shortwave_up": 11,
    "surface_longwave_up": 12, "surface_net_radiation": 13, "sensible_heat_flux": 14, "latent_heat_flux": 15,
    "evaporation_ground": 16, "infiltration": 17, "surface_runoff": 18,
    "tendency_internal_energy": 19, "tendency_saturation_water_ice": 20,
    # vegetated LandModel
    "carbon_vegetation": 21, "vegetation_area_fraction": 22, "canopy_water": 23, "balanced_leaf_area_index": 24,
    "leaf_area_index": 25, "phenology_factor": 26, "canopy_water_conductance": 27, "leaf_to_air_co2_ratio": 28,
    "net_assimilation": 29, "leaf_respiration": 30, "gross_primary_production": 31, "autotrophic_respiration": 32,
    "net_primary_production": 33, "soil_moisture_limiting_factor": 34, "canopy_water_interception": 35,
    "canopy_water_removal": 36, "saturation_canopy_water": 37, "rainfall_ground": 38, "evaporation_canopy": 39,
    "transpiration": 40, "plant_available_water": 41, "root_fraction": 42,
}
VEGETATION_FIELDS = tuple(n for n, i in FIELD_IDS.items() if i >= 21)
FIELDS_3D = ("internal_energy", "temperature", "liquid_water_fraction", "saturation_water_ice", "pressure_head",
             "tendency_internal_energy", "tendency_saturation_water_ice", "plant_available_water", "root_fraction")
FIELDS_FACE = ("hydraulic_conductivity",)
INPUT_IDS = {
    "air_temperature": TRM_IN_AIR_TEMPERATURE, "air_pressure": TRM_IN_AIR_PRESSURE, "windspeed": TRM_IN_WINDSPEED,
    "specific_humidity": TRM_IN_SPECIFIC_HUMIDITY, "rainfall": TRM_IN_RAINFALL, "snowfall": TRM_IN_SNOWFALL,
    "surface_shortwave_down": TRM_IN_SHORTWAVE_DOWN, "surface_longwave_down": TRM_IN_LONGWAVE_DOWN,
    "daytime_length": TRM_IN_DAYTIME_LENGTH, "CO2": TRM_IN_CO2, "skin_temperature": TRM_IN_SKIN_TEMPERATURE,
    "SAI": TRM_IN_SAI, "daily_leaf_respiration": TRM_IN_DAILY_LEAF_RESPIRATION,
    "albedo": TRM_IN_ALBEDO, "emissivity": TRM_IN_EMISSIVITY,
}
# inputs of the prescribed flux schemes: they share their names with the fields the diagnosed schemes write
PRESCRIBED_FLUX_INPUTS = {"surface_shortwave_up": TRM_IN_SHORTWAVE_UP, "surface_longwave_up": TRM_IN_LONGWAVE_UP,
                          "sensible_heat_flux": TRM_IN_SENSIBLE_HEAT_FLUX, "latent_heat_flux": TRM_IN_LATENT_HEAT_FLUX}
TRM_ALBEDO_CONSTANT, TRM_ALBEDO_PRESCRIBED = 0, 1
TRM_RADIATIVE_DIAGNOSED, TRM_RADIATIVE_PRESCRIBED = 0, 1
TRM_TURBULENT_DIAGNOSED, TRM_TURBULENT_PRESCRIBED = 0, 1


# vegetated LandModel parameters, in header order
VEGETATION_PARAMS = (
    "field_capacity", "wilting_point", "C_mass",
    "tau25", "Kc25", "Ko25", "q10_tau", "q10_Kc", "q10_Ko", "alpha_leaf", "alpha_a", "alpha_C3", "cq", "k_ext",
    "T_CO2_high", "T_CO2_low", "T_photos_high", "T_photos_low", "theta_r",
    "g1", "g_min", "cn_sapwood", "cn_root", "aws",
    "SLA", "awl", "LAI_min", "LAI_max", "gamma_L", "gamma_R", "gamma_S",
    "nu_seed", "gamma_v_min", "root_a", "root_b",
    "alpha_int", "k_ext_can", "w_can_max", "tau_w", "C_can",
)


class trm_params(C.Structure):
    _fields_ = [
        ("mineral_porosity", C.c_double), ("organic_porosity", C.c_double), ("rho_soc", C.c_double), ("rho_org", C.c_double),
        ("kappa", C.c_double * 5), ("heatcap", C.c_double * 5),
        ("rho_w", C.c_double), ("Lsl", C.c_double), ("Llg", C.c_double), ("rho_a", C.c_double), ("c_a", C.c_double),
        ("Tref", C.c_double), ("sigma", C.c_double), ("eps_mw", C.c_double),
        ("K_sat", C.c_double), ("vg_alpha", C.c_double), ("vg_n", C.c_double), ("bc_psis", C.c_double),
        ("bc_lambda", C.c_double), ("theta_res", C.c_double), ("impedance", C.c_double), ("vwc_forcing", C.c_double),
        ("albedo", C.c_double), ("emissivity", C.c_double), ("kappa_skin", C.c_double), ("C_h", C.c_double),
        ("min_windspeed", C.c_double), ("tau_r", C.c_double), ("evap_beta", C.c_double),
    ] + [(n, C.c_double) for n in VEGETATION_PARAMS]


class trm_bc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("input", C.c_int32)]


class trm_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("dtype", C.c_int32), ("ncol", C.c_int64), ("col0", C.c_int64),
        ("nz", C.c_int32), ("device", C.c_int32), ("model", C.c_int32), ("timestepper", C.c_int32),
        ("hydrology", C.c_int32), ("swrc", C.c_int32), ("unsat_k", C.c_int32), ("sat_halo", C.c_int32),
        ("skin", C.c_int32), ("math", C.c_int32), ("vegetation", C.c_int32), ("ground_resistance", C.c_int32),
        ("albedo_kind", C.c_int32), ("radiative", C.c_int32), ("turbulent", C.c_int32), ("reserved0", C.c_int32),
        ("z_faces", C.POINTER(C.c_double)),
        ("params", trm_params),
        ("bc", trm_bc * TRM_BC_NSLOTS),
    ]


class trm_diag(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("energy", "water", "t_min", "t_max", "sat_min", "sat_max", "nan_count", "ncol")]


# name -> (restype, argtypes); every symbol the header declares
_H = C.c_void_p
SIGNATURES = {
    "default_params": (None, [C.POINTER(trm_params)]),
    "default_config": (None, [C.POINTER(trm_config)]),
    "create": (C.c_int, [C.POINTER(trm_config), C.POINTER(_H)]),
    "destroy": (C.c_int, [_H]),
    "last_error": (C.c_char_p, []),
    "abi_version": (C.c_int, []),
    "sync": (C.c_int, [_H]),
    "field_ptr": (C.c_int, [_H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "field_view": (C.c_int, [_H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "set_field": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64]),
    "get_field": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64]),
    "set_input_const": (C.c_int, [_H, C.c_int, C.c_double]),
    "set_input_field": (C.c_int, [_H, C.c_int, C.c_void_p]),
    "set_input_field_pair": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p]),
    "set_input_sinusoid": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double]),
    "set_input_table": (C.c_int, [_H, C.c_int, C.c_int32, C.POINTER(C.c_double), C.c_void_p]),
    "set_input_raster": (C.c_int, [_H, C.c_int, C.c_int32, C.POINTER(C.c_double), C.c_void_p]),
    "get_input": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64]),
    "input_ptr": (C.c_int, [_H, C.c_int, C.POINTER(C.c_void_p)]),
    "initialize": (C.c_int, [_H]),
    "step": (C.c_int, [_H, C.c_double, C.c_int64]),
    "compute_auxiliary": (C.c_int, [_H]),
    "compute_tendencies": (C.c_int, [_H]),
    "get_clock": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "set_clock": (C.c_int, [_H, C.c_double, C.c_int64]),
    "diagnostics": (C.c_int, [_H, C.POINTER(trm_diag)]),
    "diagnostics_device": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "nccl_get_unique_id": (C.c_int, [C.c_void_p]),
    "nccl_comm_init": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p]),
    "nccl_comm_adopt": (C.c_int, [_H, C.c_void_p]),
    "diagnostics_allreduce": (C.c_int, [_H, C.POINTER(trm_diag)]),
    "launch_count": (C.c_int64, [_H]),
    "last_step_ms": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "set_block_size": (C.c_int, [_H, C.c_int]),
    "set_input_field_async": (C.c_int, [_H, C.c_int, C.c_void_p]),
    "step_async": (C.c_int, [_H, C.c_double, C.c_int64]),
    "get_field_async": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64]),
    "accumulate": (C.c_int, [_H, C.c_int, C.c_double]),
    "get_accumulated": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_int32]),
    "host_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "host_free": (C.c_int, [C.c_void_p]),
    "host_alloc_ex": (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]),
    "bind_host_io": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int32]),
    "host_io_wait": (C.c_int, [_H, C.c_int64]),
    "reset": (C.c_int, [_H]),
    "set_ring_index": (C.c_int, [_H, C.POINTER(C.c_int64), C.c_int64]),
    "get_field_ring": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64, C.c_double]),
    "set_field_ring": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64]),
}
# entry points that only make sense on a device and that the CPU oracle does not export
DEVICE_ONLY = ("field_ptr", "field_view", "host_alloc", "host_free", "host_alloc_ex", "bind_host_io", "host_io_wait", "input_ptr", "diagnostics_device", "nccl_get_unique_id", "nccl_comm_init", "nccl_comm_adopt", "diagnostics_allreduce", "launch_count", "last_step_ms", "set_block_size",
               "set_input_field_async", "step_async", "get_field_async", "set_ring_index", "get_field_ring",
               "set_field_ring")


class BoundLibrary:
    """A loaded shared library whose ``<prefix><name>`` symbols are exposed as attributes ``name``."""

    def __init__(self, cdll: C.CDLL, prefix: str, skip=()):
        self.cdll, self.prefix = cdll, prefix
        for name, (res, args) in SIGNATURES.items():
            if name in skip:
                continue
            fn = getattr(cdll, prefix + name)  # AttributeError if the symbol is missing: fail loudly
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def check(self, status: int, what: str = ""):
        if status != TRM_OK:
            msg = self.last_error().decode("utf-8", "replace")
            raise TerrariumError(status, f"{self.prefix}{what} failed with status {status}: {msg}")


class TerrariumError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


def np_dtype(dtype_code: int):
    return np.float32 if dtype_code == TRM_F32 else np.float64


def dtype_code(nf) -> int:
    nf = np.dtype(nf)
    if nf == np.float32:
        return TRM_F32
    if nf == np.float64:
        return TRM_F64
    raise ValueError(f"unsupported number format {nf}; Terrarium supports Float32 and Float64")
