#!/usr/bin/env python3
"""Development tool: push the same synthetic ensemble through two libraries that export the elmk C ABI
(default: oracle/_ref = the reference itself, and oracle/port = the host build of the physics core) and
report, per kernel group and per field, the largest relative difference.  With --resync the second
library's state is overwritten by the first one's after every group, so each group is checked in
isolation on identical inputs."""
import argparse, os, sys
os.environ.setdefault("ELMREF_SCRUB_STACK", "1")
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elmkernels_b200 import abi, params, ensemble

ap = argparse.ArgumentParser()
ap.add_argument("--a", default=os.path.join(ROOT, "oracle/_ref/libelmref.so"))
ap.add_argument("--b", default=os.path.join(ROOT, "oracle/port/_build/libelmport.so"))
ap.add_argument("--ncols", type=int, default=2048)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--resync", action="store_true")
ap.add_argument("--tol", type=float, default=1e-13)
ap.add_argument("--h2osfc", type=float, default=0.0)
ap.add_argument("--tspread", type=float, default=0.0)
ap.add_argument("--seed", type=int, default=20240005)
ap.add_argument("--groups", type=lambda s: int(s, 0), default=abi.G_ALL)
ap.add_argument("-v", action="store_true")
a = ap.parse_args()

P = params.load_params()
A, B = abi.Library(a.a), abi.Library(a.b)
print("A =", A.backend, " B =", B.backend)
cfg = ensemble.EnsembleConfig(ncols=a.ncols, seed=a.seed, h2osfc_fraction=a.h2osfc, soil_temp_spread=a.tspread)
S0 = ensemble.make_state(cfg, P, A.fields)
ca, cb = A.columns(a.ncols), B.columns(a.ncols)
for c in (ca, cb):
    c.set_tables(P)
    c.upload_state(S0)
F = ensemble.Forcing(a.ncols)
names = A.field_names

def reldiff(x, y):
    x = x.astype(np.float64); y = y.astype(np.float64)
    d = np.abs(x - y)
    s = np.maximum(np.abs(x), np.abs(y))
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(d == 0, 0.0, d / np.where(s > 0, s, 1.0))
    r = np.where(np.isnan(x) != np.isnan(y), np.inf, r)
    r = np.where(np.isnan(x) & np.isnan(y), 0.0, r)
    return r

worst = {}
for step in range(a.steps):
    st = {k: ca.download(k) for k in ("vtype", "snow_depth", "frac_sno", "htop", "hbot", "tlai", "tsai")}
    f = F.at(step, st)
    for c in (ca, cb):
        c.upload_state(f)
        c.init_timestep(True)
    for g in range(abi.G_ALL.bit_length()):
        if not (a.groups >> g) & 1:
            continue
        ca.step(groups=1 << g); cb.step(groups=1 << g)
        sa = ca.download_state(); sb = cb.download_state()
        bad = []
        for k in names:
            r = reldiff(sa[k], sb[k])
            m = float(r.max()) if r.size else 0.0
            key = (abi.GROUP_NAMES[g], k)
            if m > worst.get(key, 0.0):
                worst[key] = m
            if m > a.tol:
                idx = np.unravel_index(np.argmax(r), r.shape)
                bad.append((k, m, idx, sa[k][idx], sb[k][idx], int((r > a.tol).sum())))
        if bad and a.v:
            print(f"step {step} group {abi.GROUP_NAMES[g]}:")
            for k, m, idx, va, vb, cnt in bad:
                print(f"   {k:22s} rel {m:.3e} at {idx} A={va!r} B={vb!r}  ({cnt} elems)")
        if a.resync:
            cb.upload_state(sa)
    ea, eb = ca.errors(), cb.errors()
    snl = ca.download("snl")
    print(f"step {step}: errors A={ea} B={eb}  snl hist {np.bincount(snl, minlength=6).tolist()}")
print("worst relative differences above tol:")
for (g, k), m in sorted(worst.items(), key=lambda kv: -kv[1]):
    if m > a.tol:
        print(f"  {g:20s} {k:22s} {m:.3e}")
