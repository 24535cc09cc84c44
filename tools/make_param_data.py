#!/usr/bin/env python3
"""Convert the parameter files that ship with the reference's tests into one .npz for the product.

Inputs (data, not code): /root/reference/test/data/clm_params_c180524.nc (E3SM PFT parameter file,
NetCDF-3 classic, read with scipy) and test/data/SnowOptics_IN.txt (text dump of the SNICAR 5-band
optics file, the same tables test/test_SurfAlb.cc:301-336 loads).  Output:
elmkernels_b200/data/elm_params.npz, committed, so nothing reads /root/reference at run time.
"""
import pathlib, sys
import numpy as np
from scipy.io import netcdf_file

R = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
PFT_VARS = ("fnr act25 kcha koha cpha vcmaxha jmaxha tpuha lmrha vcmaxhd jmaxhd tpuhd lmrhd lmrse qe theta_cj "
            "bbbopt mbbopt c3psn slatop leafcn flnr fnitr dleaf smpso smpsc tc_stress z0mr displar xl roota_par "
            "rootb_par rholvis rholnir rhosvis rhosnir taulvis taulnir tausvis tausnir").split()
out = {}
nc = netcdf_file(str(R / "test/data/clm_params_c180524.nc"), "r", mmap=False)
for v in PFT_VARS:
    a = np.array(nc.variables[v][:], dtype=np.float64).reshape(-1)
    out["pft_" + v] = a.copy()
names = nc.variables["pftname"][:]
out["pftname"] = np.array([b"".join(r).decode().strip() for r in names])
nc.close()
shapes = {"snw": (5, 1471), "bc": (10, 5), "bcenh": (8, 10, 5), "band": (5,)}
for line in open(R / "test/data/SnowOptics_IN.txt"):
    tok = line.split()
    if len(tok) < 3 or tok[0] in ("NSTEP", "dtime", "!!!"):
        continue
    a = np.array(tok[1:], dtype=np.float64)
    n = tok[0]
    shp = shapes["bcenh"] if n == "bcenh" else shapes["snw"] if "_snw_" in n else shapes["bc"] if n.endswith(("bc1", "bc2")) else shapes["band"]
    out["snicar_" + n] = a.reshape(shp)
dst = pathlib.Path("elmkernels_b200/data/elm_params.npz")
np.savez_compressed(dst, **out)
print(dst, dst.stat().st_size, "bytes;", len(out), "arrays")
print({k: v.shape for k, v in out.items() if k.startswith("snicar_")})
print(out["pftname"][:17], out["pft_c3psn"][:17], out["pft_tc_stress"])
