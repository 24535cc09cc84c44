#!/usr/bin/env python3
"""Generate tests/golden/chain_64col.npz: inputs and outputs of the REFERENCE ITSELF (oracle/_ref =
/root/reference compiled unmodified behind the C ABI) for 64 synthetic columns x 6 full timesteps.
Run in the container where /root/reference is mounted; the .npz is committed."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elmkernels_b200 import abi, ensemble, params

N, STEPS = 64, 6
lib = abi.Library(os.path.join(ROOT, "oracle/_ref/libelmref.so"))
P = params.load_params()
cfg = ensemble.EnsembleConfig(ncols=N, seed=424242, h2osfc_fraction=0.15, soil_temp_spread=6.0)
S0 = ensemble.make_state(cfg, P, lib.fields)
c = lib.columns(N)
c.set_tables(P)
c.upload_state(S0)
F = ensemble.Forcing(N, seed=99)
out = {"ncols": N, "nsteps": STEPS}
out.update({"s0_" + k: v for k, v in S0.items()})
for s in range(STEPS):
    st = {k: c.download(k) for k in ("vtype", "snow_depth", "frac_sno", "htop", "hbot", "tlai", "tsai")}
    f = F.at(s, st)
    out.update({f"f{s}_{k}": v for k, v in f.items()})
    c.upload_state(f)
    c.init_timestep(True)
    c.step()
assert c.errors() == (0, -1)
out.update({"out_" + k: v for k, v in c.download_state().items()})
dst = os.path.join(ROOT, "tests/golden/chain_64col.npz")
np.savez_compressed(dst, **out)
print(dst, os.path.getsize(dst), "bytes; snl", np.bincount(out["out_snl"], minlength=6))
