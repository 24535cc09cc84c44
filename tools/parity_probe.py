#!/usr/bin/env python3
"""Development probe (GPU box): how far from BIT-identical the CUDA library is to the checker.
  1. every kernel group in isolation on identical inputs: columns with any field element whose bits differ;
  2. the full chain free-running: per step, columns that differ in any bit and columns outside 1e-8.
Usage: parity_probe.py [ncols] [steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("ELMREF_SCRUB_STACK", "1")
import elmkernels_b200
from elmkernels_b200 import abi, ensemble, params
import parity

ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 48
lib = elmkernels_b200.load()
chk = abi.Library(os.path.join(ROOT, "oracle", "_ref", "libelmref.so"))
P = params.load_params()


def bitdiff(pair, names=None):
    cols = np.zeros(pair.n, bool); per = {}
    for k in (names or pair.A.field_names):
        ra, rb = pair.a.download(k), pair.b.download(k)
        if ra.dtype.kind == "f":
            m = (ra.view(np.uint64) != rb.view(np.uint64)) & ~(np.isnan(ra) & np.isnan(rb))
        else:
            m = ra != rb
        if m.any():
            mc = m if m.ndim == 1 else m.any(axis=1)
            cols |= mc
            with np.errstate(all="ignore"):
                rel = np.abs(ra.astype(float) - rb.astype(float)) / np.maximum(np.abs(ra.astype(float)), 1e-300)
            per[k] = (int(mc.sum()), float(np.nanmax(np.where(m, rel, 0))))
    return cols, per


for seed, h2osfc, tspread in [(20240005, 0.0, 0.0), (20240003, 0.2, 8.0)]:
    cfg = ensemble.EnsembleConfig(ncols=ncols, seed=seed, h2osfc_fraction=h2osfc, soil_temp_spread=tspread)
    pair = parity.Pair(chk, lib, P, cfg)
    tot = {}
    for step in range(4):
        pair.begin_step()
        for g in range(abi.G_ALL.bit_length()):
            pair.resync(); pair.run(groups=1 << g)
            cols, per = bitdiff(pair)
            t = tot.setdefault(abi.GROUP_NAMES[g], [0, {}])
            t[0] += int(cols.sum())
            for k, v in per.items():
                o = t[1].get(k, (0, 0.0)); t[1][k] = (o[0] + v[0], max(o[1], v[1]))
    print(f"== isolation, seed {seed}: columns (of {4 * ncols}) with any bit difference, per group")
    for g, (n, per) in tot.items():
        worst = sorted(per.items(), key=lambda kv: -kv[1][1])[:4]
        print(f"  {g:22s} {n:6d}  " + "  ".join(f"{k}:{v[0]}@{v[1]:.1e}" for k, v in worst))

cfg = ensemble.EnsembleConfig(ncols=2 * ncols, seed=20240005, soil_temp_spread=6.0)
pair = parity.Pair(chk, lib, P, cfg)
print(f"== free run, {2 * ncols} columns")
for step in range(steps):
    pair.begin_step(); pair.run()
    if step < 4 or step % 8 == 7:
        cols, per = bitdiff(pair)
        out = parity.differing_columns(pair, 1e-8)
        worst = sorted(per.items(), key=lambda kv: -kv[1][1])[:3]
        print(f"  step {step:3d}: bit-different columns {int(cols.sum()):6d}, outside 1e-8: {len(out):5d}   "
              + "  ".join(f"{k}:{v[0]}@{v[1]:.1e}" for k, v in worst))
print("errors", pair.a.errors(), pair.b.errors())
