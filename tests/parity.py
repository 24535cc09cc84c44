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

# Balance residuals are differences of large, nearly cancelling terms: their rounding error is absolute,
# set by the magnitude of the operands (W/m2 of radiation and heat fluxes, kg/m2 of column water), not by
# the (near-zero) residual itself.  They are compared with atol = rtol * operand scale.
RESIDUAL_SCALE = {"errsol": 1.0e3, "errlon": 1.0e3, "errseb": 1.0e3, "soil_e_balance": 1.0e3, "errh2o": 1.0e4,
                  "errh2osno": 1.0e3, "dwb": 10.0}
# Natural magnitudes of fields that are themselves sums/differences of larger terms (net fluxes, two-stream
# and SNICAR fractions built from exponentials of opposite sign): an element passes when it is within
# rtol relative OR within rtol * (natural magnitude of its operands).
import re as _re
_SCALES = [(_re.compile(r"^(eflx_|sabg|sabv$|fsa$|fsr$|dlrad$|ulrad$|netrad$|xmf|parsun_z$|parsha_z$)"), 100.0),  # W/m2
           (_re.compile(r"^(qflx_|mflx_)"), 1.0e-4),                                                            # kg/m2/s
           (_re.compile(r"^(alb|fab|ftdd$|ftid$|ftii$|flx_abs|fsun_z$|fwet$|fdry$)"), 1.0)]                      # fractions


def field_scale(name: str):
    if name in RESIDUAL_SCALE:
        return RESIDUAL_SCALE[name]
    for rx, v in _SCALES:
        if rx.match(name):
            return v
    return None

FORCING_STATE = ("vtype", "snow_depth", "frac_sno", "htop", "hbot", "tlai", "tsai")


def mismatch(ref: np.ndarray, got: np.ndarray, rtol: float, scale: float | None = None) -> np.ndarray:
    """Boolean mask of elements outside tolerance.

    An element passes when |ref-got| <= rtol*max(|ref|,|got|), or when both are tiny relative to the
    field: |ref-got| <= rtol*scale with scale = the median non-zero finite |ref| of the field, i.e. its
    typical magnitude (robust against the 1e97 outliers of reference quirk 5), or the operand scale given
    for balance residuals (differences of large cancelling terms carry the absolute rounding error of
    their operands).
    Integers must be equal.  NaN must match NaN."""
    if ref.dtype.kind in "iu":
        return ref != got
    r = ref.astype(np.float64)
    g = got.astype(np.float64)
    fin = np.isfinite(r)
    if scale is None:
        nz = np.abs(r[fin & (r != 0.0)])
        scale = float(np.median(nz)) if nz.size else 0.0
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
            m = mismatch(ra, rb, rtol, field_scale(k))
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


def outlier_columns(pair: "Pair", rtol: float, names=None) -> np.ndarray:
    """Indices of the columns that have at least one field element outside tolerance."""
    bad = np.zeros(pair.n, dtype=bool)
    for k in (names or pair.A.field_names):
        m = mismatch(pair.a.download(k), pair.b.download(k), rtol, field_scale(k))
        bad |= m if m.ndim == 1 else m.any(axis=1)
    return np.nonzero(bad)[0]


def check_with_rare_flips(pair: "Pair", rtol: float, max_outliers: int, loose: float = 2e-2, what: str = ""):
    """The iterative solvers stop on thresholds (|dT_leaf| < 0.01 K, |dLE| < 0.1 W/m2, secant/Brent brackets):
    a last-bit difference between CUDA's and glibc's exp/log/pow can, rarely, change an iteration count, and
    the column then differs by up to the convergence tolerance instead of by rounding.  All columns must be
    within `rtol` except at most `max_outliers`, which must still agree within `loose`."""
    out = outlier_columns(pair, rtol)
    assert len(out) <= max_outliers, (f"{what}: {len(out)} columns outside rtol={rtol:g} (allowed {max_outliers})\n"
                                      + fmt(pair.compare(rtol)))
    if len(out):
        worse = pair.compare(loose, exclude_cols=None)
        worse = {k: v for k, v in worse.items() if k not in ("imelt", "snl")}
        assert not worse, f"{what}: outlier columns {out.tolist()} exceed the loose bound {loose:g}\n{fmt(worse)}"
    return out


def fmt(bad) -> str:
    return "\n".join(f"  {k}: {v[0]} elements, worst rel {v[1]:.3e} at {v[2]} ref={v[3]!r} got={v[4]!r}"
                     for k, v in sorted(bad.items(), key=lambda kv: -kv[1][1]))
