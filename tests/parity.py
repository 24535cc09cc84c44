"""Shared machinery of the parity tests: drive two libraries exporting the elmk C ABI with the same
seeded ensemble and compare every per-column field."""
from __future__ import annotations

import numpy as np

from elmkernels_b200 import abi, ensemble

# THE BAR: the CUDA library reproduces the reference BIT FOR BIT - every element of every per-column field, closed-form
# and iterative kernel groups alike, isolated or free-running.  (BASELINE.json asks for 1e-12 / 1e-8 relative; both are
# implied.)  That is possible because every operation of the chain is defined to the last bit on both sides: IEEE
# add / multiply / divide / sqrt with FMA contraction off, and transcendentals that return the bits of the libm the
# reference calls (csrc/elmk_libm.h).  There is no tolerance, no per-field scale and no allowance for outlier columns.
RTOL_CLOSED = 0.0
RTOL_ITER = 0.0

FORCING_STATE = ("vtype", "snow_depth", "frac_sno", "htop", "hbot", "tlai", "tsai")


def mismatch(ref: np.ndarray, got: np.ndarray, rtol: float = 0.0) -> np.ndarray:
    """Boolean mask of elements that differ.  rtol == 0 (the default, and what every CUDA-versus-oracle test uses):
    the 64 bits of every double must be the same (the sign of a zero included; any NaN matches any NaN).  rtol > 0 is
    only for comparisons against the ELM Fortran fixtures, whose tolerance the test states:
    |ref - got| <= rtol * max(|ref|, |got|).  Integers must be equal."""
    if ref.dtype.kind in "iu":
        return ref != got
    r = np.ascontiguousarray(ref, dtype=np.float64)
    g = np.ascontiguousarray(got, dtype=np.float64)
    ok = (r.view(np.uint64) == g.view(np.uint64)) | (np.isnan(r) & np.isnan(g))
    if rtol > 0.0:
        with np.errstate(invalid="ignore", over="ignore"):
            ok |= np.abs(r - g) <= rtol * np.maximum(np.abs(r), np.abs(g))
    return ~ok


class Pair:
    """Two handles (a = checker, b = under test) fed with the same ensemble."""

    def __init__(self, lib_a, lib_b, params, cfg: ensemble.EnsembleConfig, forcing_seed: int = 7,
                 night_fraction=None):
        self.A, self.B = lib_a, lib_b
        self.n = cfg.ncols
        self.state0 = ensemble.make_state(cfg, params, lib_a.fields)
        self.a, self.b = lib_a.columns(self.n), lib_b.columns(self.n)
        for c in (self.a, self.b):
            c.set_tables(params)
            c.upload_state(self.state0)
        self.forcing = ensemble.Forcing(self.n, seed=forcing_seed, night_fraction=night_fraction)
        self.step_no = 0

    def begin_step(self):
        st = {k: self.a.download(k) for k in FORCING_STATE}
        f = self.forcing.at(self.step_no, st)
        for c in (self.a, self.b):
            c.upload_state(f)
            c.init_timestep(True)
        self.step_no += 1

    def run(self, groups=abi.G_ALL, dtime=1800.0):
        self.a.step(dtime=dtime, groups=groups)
        self.b.step(dtime=dtime, groups=groups)

    def resync(self):
        """Overwrite b's state with a's (isolates one group on identical inputs)."""
        self.b.upload_state(self.a.download_state())

    def compare(self, rtol: float = 0.0, names=None):
        """Returns {field: (n_bad, worst_rel, index, ref, got)} for the fields with differing elements."""
        bad = {}
        for k in (names or self.A.field_names):
            ra, rb = self.a.download(k), self.b.download(k)
            m = mismatch(ra, rb, rtol)
            if m.any():
                r = ra.astype(np.float64)
                g = rb.astype(np.float64)
                with np.errstate(invalid="ignore", divide="ignore"):
                    rel = np.abs(r - g) / np.maximum(np.maximum(np.abs(r), np.abs(g)), 1e-300)
                rel = np.where(m, np.nan_to_num(rel, nan=np.inf), 0.0)
                idx = np.unravel_index(int(np.argmax(rel)), rel.shape)
                bad[k] = (int(m.sum()), float(rel[idx]), tuple(int(i) for i in idx), float(r[idx]), float(g[idx]))
        return bad


def differing_columns(pair: "Pair", rtol: float = 0.0, names=None) -> np.ndarray:
    """Indices of the columns that have at least one differing field element."""
    bad = np.zeros(pair.n, dtype=bool)
    for k in (names or pair.A.field_names):
        m = mismatch(pair.a.download(k), pair.b.download(k), rtol)
        bad |= m if m.ndim == 1 else m.any(axis=1)
    return np.nonzero(bad)[0]


def fmt(bad) -> str:
    return "\n".join(f"  {k}: {v[0]} elements, worst rel {v[1]:.3e} at {v[2]} ref={v[3]!r} got={v[4]!r}"
                     for k, v in sorted(bad.items(), key=lambda kv: -kv[1][1]))
