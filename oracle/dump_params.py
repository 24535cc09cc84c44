#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle).  Dump the variables of the reference's test/data/clm_params_c180524.nc as text files
(one per variable, doubles as hex floats, character arrays raw) into oracle/_ref/clm_params/, for the NetCDF stand-in
oracle/shim_serial/netcdf.h that the serial builds of the reference's unit tests read them through."""
import os, pathlib, sys
import numpy as np
from scipy.io import netcdf_file
HERE = pathlib.Path(__file__).resolve().parent
REF = pathlib.Path(os.environ.get("ELMK_REFERENCE", "/root/reference"))
OUT = HERE / "_ref" / "clm_params"

def dump():
    OUT.mkdir(parents=True, exist_ok=True)
    f = netcdf_file(str(REF / "test/data/clm_params_c180524.nc"), "r", mmap=False)
    n = 0
    for name, var in f.variables.items():
        a = np.array(var.data)
        p = OUT / f"{name}.txt"
        if a.dtype.kind == "S":
            p.write_bytes(a.tobytes())
        elif a.dtype.kind in "fi":
            p.write_text("\n".join(float(x).hex() for x in a.astype(np.float64).ravel()) + "\n")
        else:
            continue
        n += 1
    return n

if __name__ == "__main__":
    print(dump(), "variables ->", OUT)
