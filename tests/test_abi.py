"""The C ABI: every symbol `include/terrarium_b200.h` declares is exported by the built library, the ctypes
mirror has the same struct layout as the C header, and the product path fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from common import trm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "terrarium_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 25
    lib = C.CDLL(trm.LIB_PATH)   # loading must work without a GPU: no device call at load time
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # ... and the ctypes mirror binds all of them
    bound = {"trm_" + n for n in trm.abi.SIGNATURES}
    assert set(names) <= bound, sorted(set(names) - bound)
    assert lib.trm_abi_version() == trm.abi.TRM_ABI_VERSION


def test_ctypes_layout_matches_header():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "terrarium_b200.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(trm_params), sizeof(trm_bc), sizeof(trm_config), sizeof(trm_diag),
           offsetof(trm_config, z_faces), offsetof(trm_config, params), offsetof(trm_config, bc));
    return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "layout.c"), os.path.join(d, "layout")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    a = trm.abi
    want = [C.sizeof(a.trm_params), C.sizeof(a.trm_bc), C.sizeof(a.trm_config), C.sizeof(a.trm_diag),
            a.trm_config.z_faces.offset, a.trm_config.params.offset, a.trm_config.bc.offset]
    assert got == want


def test_defaults_match_the_reference_package_defaults():
    lib = trm.cuda_library()
    p = trm.abi.trm_params()
    lib.default_params(C.byref(p))
    assert (p.mineral_porosity, p.K_sat, p.tau_r, p.albedo, p.emissivity) == (0.49, 1.0e-5, 3600.0, 0.3, 0.97)
    assert list(p.kappa) == [0.57, 2.2, 0.025, 3.8, 0.25]
    assert list(p.heatcap) == [4.2e6, 1.9e6, 1.25e3, 2.0e6, 2.5e6]
    # vegetation / canopy defaults of the library equal the host mirror's dataclass defaults (= the reference's)
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.ExponentialSpacing(dz_max=1.0, N=4), 1)
    q = trm.build_params(trm.LandModel(grid))
    for name in trm.abi.VEGETATION_PARAMS:
        if name in ("field_capacity", "wilting_point"):
            continue   # the library default is ConstantSoilHydraulics'; LandModel(grid) has SoilHydraulicsSURFEX (clay = 0 -> 0)
        assert getattr(p, name) == getattr(q, name), name
    assert (q.field_capacity, q.wilting_point) == (0.0, 0.0)
    assert (p.tau25, p.g1, p.SLA, p.w_can_max, p.C_can, p.field_capacity) == (2600.0, 2.3, 10.0, 2.0e-4, 0.006, 0.25)


def test_argument_validation_needs_no_device():
    lib = trm.cuda_library()
    cfg = trm.abi.trm_config()
    lib.default_config(C.byref(cfg))
    h = C.c_void_p()
    cfg.ncol, cfg.nz = 4, 1          # nz < 2
    assert lib.create(C.byref(cfg), C.byref(h)) == trm.abi.TRM_ERR_INVALID
    assert b"nz" in lib.last_error()
    cfg.nz, cfg.abi_version = 4, 99
    assert lib.create(C.byref(cfg), C.byref(h)) == trm.abi.TRM_ERR_INVALID
    assert lib.step(None, 1.0, 1) == trm.abi.TRM_ERR_INVALID


def _cuda_device_present():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0
    except FileNotFoundError:
        return False


@pytest.mark.skipif(_cuda_device_present(), reason="a GPU is present: the no-device error path cannot be exercised")
def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (TRM_ERR_NO_DEVICE), never compute on the CPU."""
    grid = trm.ColumnGrid(trm.B200(), np.float64, trm.UniformSpacing(dz=0.1, N=4), 2)
    with pytest.raises(trm.TerrariumError) as err:
        trm.initialize(trm.SoilModel(grid), trm.ForwardEuler())
    assert err.value.status == trm.abi.TRM_ERR_NO_DEVICE
    assert "no CPU fallback" in str(err.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "terrarium.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".jl")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_integrator" not in text and "libterrarium_oracle" not in text and "orc_" not in text, f


def test_enum_values_match_header():
    """Field / input / source / option codes of the ctypes mirror equal the header's enumerators (compiled, not parsed)."""
    a = trm.abi
    pairs = {
        "TRM_F_INTERNAL_ENERGY": a.FIELD_IDS["internal_energy"], "TRM_F_SURFACE_RUNOFF": a.FIELD_IDS["surface_runoff"],
        "TRM_F_TEND_SATURATION": a.FIELD_IDS["tendency_saturation_water_ice"], "TRM_F_CARBON_VEGETATION": a.FIELD_IDS["carbon_vegetation"],
        "TRM_F_NET_ASSIMILATION": a.FIELD_IDS["net_assimilation"], "TRM_F_TRANSPIRATION": a.FIELD_IDS["transpiration"],
        "TRM_F_PLANT_AVAILABLE_WATER": a.FIELD_IDS["plant_available_water"], "TRM_F_ROOT_FRACTION": a.FIELD_IDS["root_fraction"],
        "TRM_F_COUNT": max(a.FIELD_IDS.values()) + 1,
        "TRM_IN_AIR_TEMPERATURE": a.TRM_IN_AIR_TEMPERATURE, "TRM_IN_CO2": a.TRM_IN_CO2, "TRM_IN_SKIN_TEMPERATURE": a.TRM_IN_SKIN_TEMPERATURE,
        "TRM_IN_SAI": a.TRM_IN_SAI, "TRM_IN_DAILY_LEAF_RESPIRATION": a.TRM_IN_DAILY_LEAF_RESPIRATION, "TRM_IN_COUNT": a.TRM_IN_COUNT,
        "TRM_SRC_TABLE": a.TRM_SRC_TABLE, "TRM_SRC_RASTER": a.TRM_SRC_RASTER, "TRM_BC_FLUX": a.TRM_BC_FLUX, "TRM_BC_NSLOTS": a.TRM_BC_NSLOTS,
        "TRM_VEG_CARBON": a.TRM_VEG_CARBON, "TRM_GROUND_RES_SOIL_MOISTURE": a.TRM_GROUND_RES_SOIL_MOISTURE, "TRM_MATH_FAST": a.TRM_MATH_FAST,
        "TRM_SKIN_PRESCRIBED": a.TRM_SKIN_PRESCRIBED, "TRM_ABI_VERSION": a.TRM_ABI_VERSION, "TRM_MAX_NZ": a.TRM_MAX_NZ,
    }
    names = sorted(pairs)
    prog = '#include <stdio.h>\n#include "terrarium_b200.h"\nint main(void) {\n' + "".join(f'    printf("%d\\n", (int){n});\n' for n in names) + "    return 0;\n}\n"
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "enums.c"), os.path.join(d, "enums")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    assert dict(zip(names, got)) == pairs
    # the field-id table is dense and in header order for the vegetation block
    veg = [a.FIELD_IDS[n] for n in a.VEGETATION_FIELDS]
    assert veg == list(range(21, 43))


def test_julia_glue_matches_the_abi():
    """``julia/TerrariumB200.jl`` cannot be executed here (no Julia), so its tables are checked statically against the ctypes
    mirror (itself pinned to the compiled header above): field ids, input ids, boundary slots, the size of ``TrmParams`` /
    ``TrmConfig`` / ``TrmDiag`` counted from the Julia field declarations, and that every ``ccall`` names an exported
    symbol with the argument count the header declares."""
    import re
    a = trm.abi
    text = open(os.path.join(ROOT, "terrarium.jl_b200", "julia", "TerrariumB200.jl")).read()
    # @enum FieldId
    enum = re.search(r"@enum FieldId::Cint (.*)", text).group(1)
    ids = {k: int(v) for k, v in re.findall(r"(\w+)=(\d+)", enum)}
    assert ids and all(a.FIELD_IDS[k] == v for k, v in ids.items())
    assert set(a.VEGETATION_FIELDS) <= set(ids)
    # input ids and BC slots
    inputs = dict(re.findall(r":(\w+) => (\d+)", re.search(r"const INPUT_ID = Dict\((.*?)\)\n", text, re.S).group(1)))
    assert inputs and all(a.INPUT_IDS[k] == int(v) for k, v in inputs.items())
    slots = re.findall(r"\(:(\w+), :(\w+)\) => (\d+)", re.search(r"const BC_SLOT = Dict\((.*?)\)\n", text, re.S).group(1))
    want = {("temperature", "top"): a.TRM_BC_TEMPERATURE_TOP, ("temperature", "bottom"): a.TRM_BC_TEMPERATURE_BOTTOM,
            ("internal_energy", "top"): a.TRM_BC_ENERGY_TOP, ("internal_energy", "bottom"): a.TRM_BC_ENERGY_BOTTOM,
            ("saturation_water_ice", "top"): a.TRM_BC_SATURATION_TOP, ("saturation_water_ice", "bottom"): a.TRM_BC_SATURATION_BOTTOM,
            ("pressure_head", "top"): a.TRM_BC_PRESSURE_TOP, ("pressure_head", "bottom"): a.TRM_BC_PRESSURE_BOTTOM}
    assert {(f, s): int(v) for f, s, v in slots} == want
    assert int(re.search(r"const TRM_ABI_VERSION = Int32\((\d+)\)", text).group(1)) == a.TRM_ABI_VERSION

    # struct sizes from the Julia declarations (all members are naturally aligned: 4-byte members come in pairs)
    def struct_bytes(name):
        body = re.search(rf"struct {name}\b(.*?)\nend", text, re.S).group(1)
        body = re.sub(r"#.*", "", body)
        size = 0
        for typ in re.findall(r"::\s*([\w{}, ]+?)\s*(?:;|\n|$)", body):
            m = re.match(r"NTuple\{(\w+), (\w+)\}", typ)
            n, base = (m.group(1), m.group(2)) if m else (1, typ)
            n = a.TRM_BC_NSLOTS if n == "TRM_BC_NSLOTS" else int(n)
            size += n * {"Cdouble": 8, "Int64": 8, "Int32": 4, "Ptr{Cdouble}": 8, "TrmBC": 8, "TrmParams": C.sizeof(a.trm_params)}[base]
        return size
    def field_names(name):
        body = re.sub(r"#.*", "", re.search(rf"struct {name}\b(.*?)\nend", text, re.S).group(1))
        return re.findall(r"(\w+)::", body)
    assert field_names("TrmConfig") == [f[0] for f in a.trm_config._fields_]
    julia_params, c_params = field_names("TrmParams"), [f[0] for f in a.trm_params._fields_]
    assert julia_params[-1] == "vegetation" and julia_params[:-1] == c_params[:len(julia_params) - 1]
    assert len(c_params) - (len(julia_params) - 1) == 40          # the NTuple{40} block = the vegetation / canopy parameters
    assert struct_bytes("TrmParams") == C.sizeof(a.trm_params)
    assert struct_bytes("TrmConfig") == C.sizeof(a.trm_config)
    assert struct_bytes("TrmDiag") == C.sizeof(a.trm_diag)

    # every ccall: exported symbol, argument tuple as long as the header's parameter list
    header = open(os.path.join(ROOT, "include", "terrarium_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    nargs = {}
    for name, params in re.findall(r"\b(trm_\w+)\s*\(([^;{]*?)\)\s*;", header):
        params = params.strip()
        nargs[name] = 0 if params in ("", "void") else params.count(",") + 1
    calls = re.findall(r"ccall\(\(:(trm_\w+), LIB\), \w+,\s*\(([^()]*(?:\{[^{}]*\}[^()]*)*)\)", text)
    assert len(calls) >= 25
    for name, argtypes in calls:
        assert name in nargs, name
        depth, n = 0, (1 if argtypes.strip() else 0)
        for ch in argtypes.strip().rstrip(","):
            depth += ch == "{"
            depth -= ch == "}"
            n += (ch == "," and depth == 0)
        assert n == nargs[name], (name, argtypes, nargs[name])
