"""Generates the golden fixtures of tests/golden/*.npz.

The reference (Terrarium.jl, pure Julia) cannot be executed in this environment, so these vectors are NOT outputs of
the reference: they are outputs of the CPU oracle (oracle/terrarium_oracle.cpp), frozen at the revision in which
the oracle passed every known-answer test of the reference (tests/test_golden.py).  They pin the oracle against
silent drift and give the CUDA path a fixed target that does not depend on rebuilding the oracle.

    python tests/golden/make_golden.py          # rewrites the fixtures (only after a deliberate oracle change)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle")):
    sys.path.insert(0, p)

from common import synthetic_land_case, synthetic_soil_case  # noqa: E402


def synthetic_vegetated_case(*a, **kw):
    from test_vegetation import synthetic_vegetated_case as f
    return f(*a, **kw)


NCOL = 16
CASES = {
    # name: (builder kwargs, dt, steps, fields)
    "soil_richards_euler": (lambda e: synthetic_soil_case(e, NCOL), 60.0, 200,
                            ("internal_energy", "temperature", "liquid_water_fraction", "saturation_water_ice", "pressure_head", "water_table")),
    "soil_richards_heun": (lambda e: synthetic_soil_case(e, NCOL, heun=True), 60.0, 200,
                           ("internal_energy", "temperature", "liquid_water_fraction", "saturation_water_ice", "pressure_head", "water_table")),
    "soil_heat_only_euler": (lambda e: synthetic_soil_case(e, NCOL, richards=False), 300.0, 200,
                             ("internal_energy", "temperature", "liquid_water_fraction")),
    "land_bare_ground_euler": (lambda e: synthetic_land_case(e, NCOL, windspeed=0.5), 60.0, 200,
                               ("internal_energy", "temperature", "saturation_water_ice", "pressure_head", "skin_temperature", "ground_heat_flux",
                                "latent_heat_flux", "sensible_heat_flux", "infiltration", "surface_runoff", "surface_excess_water")),
    "land_default_soil_heun": (lambda e: synthetic_land_case(e, NCOL, windspeed=0.5, richards=False, heun=True), 60.0, 200,
                               ("internal_energy", "temperature", "liquid_water_fraction", "skin_temperature", "ground_heat_flux")),
    "land_vegetation_heun": (lambda e: synthetic_vegetated_case(e, NCOL, heun=True), 60.0, 200,
                             ("internal_energy", "temperature", "saturation_water_ice", "skin_temperature", "ground_heat_flux", "latent_heat_flux",
                              "carbon_vegetation", "vegetation_area_fraction", "canopy_water", "net_assimilation", "net_primary_production",
                              "canopy_water_conductance", "transpiration", "evaporation_canopy", "evaporation_ground", "rainfall_ground",
                              "soil_moisture_limiting_factor")),
}


def run(name, engine):
    build, dt, steps, fields = CASES[name]
    integ = build(engine)
    integ.step(dt, steps)
    return {f: getattr(integ.state, f).numpy() for f in fields}


if __name__ == "__main__":
    for name in (sys.argv[1:] or CASES):   # optional: only the named cases
        out = run(name, "oracle")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
