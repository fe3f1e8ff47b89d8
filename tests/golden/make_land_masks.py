"""Freeze the land masks of BASELINE configs 2-4 as a small fixture (run in the build container, where the reference
checkout is mounted; the GPU box has no /root/reference):

    python tests/golden/make_land_masks.py

``era5_land_masks.npz`` holds, for N72 and N145, the boolean mask ``land_sea_frac .> 0.5`` in Float32
(``examples/simulations/soil_heat_global.jl:29-38``) packed to bits in ring order, and the Gaussian latitudes /
longitudes (degrees) of the full grid. Decoded with ``terrarium_jl_b200.netcdf4`` (no netCDF4 / h5py in the image)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import terrarium_jl_b200 as trm  # noqa: E402

out = {}
for name in ("N72", "N145"):
    with trm.netcdf4.File(f"/root/reference/inputs/era5-land_land_sea_mask_{name}.nc") as f:
        frac = f.variables["lsm"].scaled(np.float32)[0]
        out[f"{name}_bits"] = np.packbits(frac.reshape(-1) > np.float32(0.5))
        out[f"{name}_lat"] = f.variables["lat"][:]
        out[f"{name}_lon"] = f.variables["lon"][:]
        print(name, frac.shape, int((frac > np.float32(0.5)).sum()))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "era5_land_masks.npz"), **out)
