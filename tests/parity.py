"""Shared machinery of the parity tests: drive two libraries exporting the elmk C ABI with the same
seeded ensemble and compare every per-column field."""
from __future__ import annotations

import numpy as np

from elmkernels_b200 import abi, ensemble

# Fields produced by (or downstream of) the iterative solvers: the CanopyFluxes stability /
# photosynthesis iteration and the SoilTemperature solve + phase change.  BASELINE.json: 1e-8 relative
# on converged fluxes and temperatures, 1e-12 for closed-form kernels.
RTOL_CLOSED = 1e-12
RTOL_ITER = 1e-8

FORCING_STATE = ("vtype", "snow_depth", "frac_sno", "htop", "hbot", "tlai", "tsai")


def mismatch(ref: np.ndarray, got: np.ndarray, rtol: float, scale: float | None = None) -> np.ndarray:
    """Boolean mask of elements outside tolerance.

    An element passes when |ref-got| <= rtol*max(|ref|,|got|), or when both are tiny relative to the
    field: |ref-got| <= rtol*scale with scale = largest finite |ref| of the field (differences of large
    cancelling terms, e.g. balance residuals, carry the absolute rounding error of their operands).
    Integers must be equal.  NaN must match NaN."""
    if ref.dtype.kind in "iu":
        return ref != got
    r = ref.astype(np.float64)
    g = got.astype(np.float64)
    fin = np.isfinite(r)
    if scale is None:
        scale = float(np.max(np.abs(r[fin]))) if fin.any() else 0.0
    with np.errstate(invalid="ignore", over="ignore"):
        d = np.abs(r - g)
        ok = d <= rtol * np.maximum(np.abs(r), np.abs(g))
        ok |= d <= rtol * scale
    ok |= (np.isnan(r) & np.isnan(g)) | ((r == g))
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

    def compare(self, rtol: float, names=None, exclude_cols=None):
        """Returns {field: (n_bad, worst_rel, column)} for fields outside tolerance."""
        bad = {}
        for k in (names or self.A.field_names):
            ra, rb = self.a.download(k), self.b.download(k)
            m = mismatch(ra, rb, rtol)
            if exclude_cols is not None:
                m[exclude_cols] = False
            if m.any():
                r = ra.astype(np.float64)
                g = rb.astype(np.float64)
                with np.errstate(invalid="ignore", divide="ignore"):
                    rel = np.abs(r - g) / np.maximum(np.maximum(np.abs(r), np.abs(g)), 1e-300)
                rel = np.where(m, np.nan_to_num(rel, nan=np.inf), 0.0)
                idx = np.unravel_index(int(np.argmax(rel)), rel.shape)
                bad[k] = (int(m.sum()), float(rel[idx]), tuple(int(i) for i in idx), float(r[idx]), float(g[idx]))
        return bad


def fmt(bad) -> str:
    return "\n".join(f"  {k}: {v[0]} elements, worst rel {v[1]:.3e} at {v[2]} ref={v[3]!r} got={v[4]!r}"
                     for k, v in sorted(bad.items(), key=lambda kv: -kv[1][1]))
