"""Synthetic land-column ensembles for parity tests and benchmarks.

Everything here is host-side numpy: it produces the *inputs* of the column timestep - the
time-invariant column properties, a cold-start state and the per-step atmospheric forcing and
phenology - in the reference's host layout (column outer).  The one-time initialisation follows
the reference's init functions, restated vectorised over columns:

  vertical grid          elm_kokkos_interface.cc:137-184 (the standard CLM exponential grid)
  init_topo_slope, init_melt_factor, init_micro_sigma      src/physics/init_topography_impl.hh:7-37
  init_snow_layers, init_snow_state                        src/physics/init_snow_state_impl.hh:12-150
  pedotransfer, soil_hydraulic_params, init_soil_hydraulics src/physics/soil_texture_hydraulic_model_impl.hh:6-124
  init_vegrootfr, init_soil_temp, init_soilh2o_state       src/physics/init_soil_state_impl.hh:10-222
  snow-cover restoration after the cold start              SURVEY.md section 8(d) (Niu-Yang form of init_snow_state)
  phenology burial of LAI/SAI by snow                      src/physics/phenology_physics_impl.hh:36-61
  VIS/NIR split of incident shortwave                      src/physics/atm_physics_impl.hh:126-141

tests/test_init_cpu.py checks the one-time initialisation against the reference's own functions (through the
oracle library).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from .abi import F64, I32, U8
from .params import psn_rows

NLEVSNO, NLEVGRND, NLEVSOI, NLEVBED = 5, 15, 10, 15
TFRZ, DENICE, DENH2O, BDSNO, ZLND, SNW_RDS_MIN = 273.15, 917.0, 1000.0, 250.0, 0.01, 54.526
SECSPDAY, STEBOL = 86400.0, 5.67e-8

_NP = {F64: np.float64, I32: np.int32, U8: np.uint8}


def vertical_grid():
    """CLM soil grid: node depth z_j = 0.025 (exp(0.5 (j - 0.5)) - 1), 15 layers, below 5 snow slots."""
    j = np.arange(1, NLEVGRND + 1)
    z = 0.025 * (np.exp(0.5 * (j - 0.5)) - 1.0)
    dzs = np.empty(NLEVGRND)
    dzs[0] = 0.5 * (z[0] + z[1])
    dzs[1:-1] = 0.5 * (z[2:] - z[:-2])
    dzs[-1] = z[-1] - z[-2]
    zi = np.empty(NLEVGRND + 1)
    zi[0] = 0.0
    zi[1:-1] = 0.5 * (z[:-1] + z[1:])
    zi[-1] = z[-1] + 0.5 * dzs[-1]
    dz = np.concatenate([np.zeros(NLEVSNO), dzs])
    zsoi = np.concatenate([np.zeros(NLEVSNO), z])
    zisoi = np.concatenate([np.zeros(NLEVSNO), zi])
    return dz, zsoi, zisoi


def soil_hydraulics(pct_sand, pct_clay, organic, organic_max, zsoi_soil):
    """Vectorised init_soil_hydraulics: inputs [n,15] (zsoi_soil [15]); returns dict of [n,15] (+csol [n,20])."""
    n = pct_sand.shape[0]
    sand = pct_sand.copy()
    clay = pct_clay.copy()
    sand[:, NLEVSOI:] = pct_sand[:, NLEVSOI - 1:NLEVSOI]
    clay[:, NLEVSOI:] = pct_clay[:, NLEVSOI - 1:NLEVSOI]
    om_frac = (organic / organic_max) ** 2.0
    om_frac[:, NLEVSOI:] = 0.0
    z = zsoi_soil[None, :]
    zsapric, pcalpha, pcbeta, om_tkd, om_tkm, om_csol = 0.5, 0.5, 0.139, 0.05, 0.25, 2.5
    watsat = 0.489 - 0.00126 * sand
    sucsat = 10.0 * 10.0 ** (1.88 - 0.0131 * sand)
    xksat = 0.0070556 * 10.0 ** (-0.884 + 0.0153 * sand)
    om_watsat = np.maximum(0.93 - 0.1 * (z / zsapric), 0.83)
    om_b = np.minimum(2.7 + 9.3 * (z / zsapric), 12.0)
    om_sucsat = np.minimum(10.3 - 0.2 * (z / zsapric), 10.1)
    om_hksat = np.maximum(0.28 - 0.2799 * (z / zsapric), 0.0001)
    bulk_den = (1.0 - watsat) * 2.7e3
    tkm = (1.0 - om_frac) * (8.8 * sand + 2.92 * clay) / (sand + clay) + om_tkm * om_frac
    watsat = (1.0 - om_frac) * watsat + om_watsat * om_frac
    bsw = (1.0 - om_frac) * (2.91 + 0.159 * clay) + om_frac * om_b
    sucsat = (1.0 - om_frac) * sucsat + om_sucsat * om_frac
    perc_norm = (1.0 - pcalpha) ** (-pcbeta)
    with np.errstate(invalid="ignore"):
        perc_frac = np.where(om_frac > pcalpha, perc_norm * np.abs(om_frac - pcalpha) ** pcbeta, 0.0)
    uncon_frac = (1.0 - om_frac) + (1.0 - perc_frac) * om_frac
    uncon_hksat = np.where(om_frac < 1.0, uncon_frac / ((1.0 - om_frac) / xksat + ((1.0 - perc_frac) * om_frac) / om_hksat), 0.0)
    hksat = uncon_frac * uncon_hksat + (perc_frac * om_frac) * om_hksat
    tkmg = tkm ** (1.0 - watsat)
    tkdry = ((0.135 * bulk_den + 64.7) / (2.7e3 - 0.947 * bulk_den)) * (1.0 - om_frac) + om_tkd * om_frac
    csol15 = ((1.0 - om_frac) * (2.128 * sand + 2.385 * clay) / (sand + clay) + om_csol * om_frac) * 1.0e6
    csol15[:, NLEVSOI:] = 2.0e6
    watdry = watsat * (316230.0 / sucsat) ** (-1.0 / bsw)
    watopt = watsat * (158490.0 / sucsat) ** (-1.0 / bsw)
    watfc = watsat * (0.1 / (hksat * SECSPDAY)) ** (1.0 / (2.0 * bsw + 3.0))
    # the reference writes csol(i) for i in 0..14 of a 20-wide row and later reads rows 5..19
    # (init_soil_hydraulics vs calc_soil_heat_capacity): keep that placement, rows 15..19 stay zero
    csol = np.zeros((n, NLEVSNO + NLEVGRND))
    csol[:, :NLEVGRND] = csol15
    return dict(watsat=watsat, bsw=bsw, sucsat=sucsat, watdry=watdry, watopt=watopt, watfc=watfc, tkmg=tkmg,
                tkdry=tkdry, csol=csol)


def snow_layers(snow_depth, dz, zsoi, zisoi):
    """Vectorised init_snow_layers: fills snow rows of dz/zsoi/zisoi ([n,20],[n,20],[n,21]) in place; returns snl."""
    d = snow_depth
    n = d.shape[0]
    snl = np.zeros(n, dtype=np.int32)
    dzs = np.zeros((n, NLEVSNO))
    m = (d >= 0.01) & (d <= 0.03); snl[m] = 1; dzs[m, 4] = d[m]
    m = (d > 0.03) & (d <= 0.04); snl[m] = 2; dzs[m, 3] = d[m] / 2.0; dzs[m, 4] = dzs[m, 3]
    m = (d > 0.04) & (d <= 0.07); snl[m] = 2; dzs[m, 3] = 0.02; dzs[m, 4] = d[m] - dzs[m, 3]
    m = (d > 0.07) & (d <= 0.12); snl[m] = 3; dzs[m, 2] = 0.02; dzs[m, 3] = (d[m] - 0.02) / 2.0; dzs[m, 4] = dzs[m, 3]
    m = (d > 0.12) & (d <= 0.18); snl[m] = 3; dzs[m, 2] = 0.02; dzs[m, 3] = 0.05; dzs[m, 4] = d[m] - dzs[m, 2] - dzs[m, 3]
    m = (d > 0.18) & (d <= 0.29); snl[m] = 4; dzs[m, 1] = 0.02; dzs[m, 2] = 0.05
    dzs[m, 3] = (d[m] - dzs[m, 1] - dzs[m, 2]) / 2.0; dzs[m, 4] = dzs[m, 3]
    m = (d > 0.29) & (d <= 0.41); snl[m] = 4; dzs[m, 1] = 0.02; dzs[m, 2] = 0.05; dzs[m, 3] = 0.11
    dzs[m, 4] = d[m] - dzs[m, 1] - dzs[m, 2] - dzs[m, 3]
    m = (d > 0.41) & (d <= 0.64); snl[m] = 5; dzs[m, 0] = 0.02; dzs[m, 1] = 0.05; dzs[m, 2] = 0.11
    dzs[m, 3] = (d[m] - dzs[m, 0] - dzs[m, 1] - dzs[m, 2]) / 2.0; dzs[m, 4] = dzs[m, 3]
    m = d > 0.64; snl[m] = 5; dzs[m, 0] = 0.02; dzs[m, 1] = 0.05; dzs[m, 2] = 0.11; dzs[m, 3] = 0.23
    dzs[m, 4] = d[m] - dzs[m, 0] - dzs[m, 1] - dzs[m, 2] - dzs[m, 3]
    dz[:, :NLEVSNO] = dzs
    zsoi[:, :NLEVSNO] = 0.0
    zisoi[:, :NLEVSNO] = 0.0
    for j in range(NLEVSNO - 1, -1, -1):
        act = j >= NLEVSNO - snl
        zsoi[act, j] = zisoi[act, j + 1] - 0.5 * dz[act, j]
        zisoi[act, j] = zisoi[act, j + 1] - dz[act, j]
    return snl


def root_fractions(vtype, roota, rootb, zisoi):
    n = vtype.shape[0]
    rootfr = np.zeros((n, NLEVGRND))
    a = roota[vtype][:, None]
    b = rootb[vtype][:, None]
    zi = zisoi[:, NLEVSNO:]
    up, lo = zi[:, :NLEVSOI - 1], zi[:, 1:NLEVSOI]
    rootfr[:, :NLEVSOI - 1] = 0.5 * (np.exp(-a * up) + np.exp(-b * up) - np.exp(-a * lo) - np.exp(-b * lo))
    last = zi[:, NLEVSOI - 1:NLEVSOI]
    rootfr[:, NLEVSOI - 1:NLEVSOI] = 0.5 * (np.exp(-a * last) + np.exp(-b * last))
    rootfr[vtype == 0] = 0.0
    return rootfr


@dataclass
class EnsembleConfig:
    ncols: int
    seed: int = 20240005
    mixed_pft: bool = True          # vtype ~ U{1..16}; otherwise 12 (c3 arctic grass, the fixture PFT)
    snow_fraction: float = 0.5      # share of columns that start with a snow pack
    bare_fraction: float = 0.2      # share of columns with tlai == 0 (bare-ground flux path)
    h2osfc_fraction: float = 0.0    # share of columns that start with standing surface water
    soil_temp_spread: float = 0.0   # +- K of per-column perturbation on the 274 K cold start
    aerosol: bool = True
    organic_max: float = 130.0


def make_state(cfg: EnsembleConfig, params: Dict[str, np.ndarray], fields: Dict[str, tuple]) -> Dict[str, np.ndarray]:
    """All per-column fields of include/elmk_fields.def for a cold-started ensemble (host layout)."""
    n = cfg.ncols
    rng = np.random.default_rng(cfg.seed)
    S = {k: np.zeros((n,) if nl == 1 else (n, nl), dtype=_NP[dt]) for k, (_, dt, nl) in fields.items()}
    S["vtype"][:] = rng.integers(1, 17, n) if cfg.mixed_pft else 12
    S["isoicol"][:] = rng.integers(0, 20, n)
    pct_sand = rng.uniform(5.0, 60.0, (n, NLEVGRND))
    pct_clay = rng.uniform(5.0, 40.0, (n, NLEVGRND))
    organic = rng.uniform(0.0, 60.0, (n, NLEVGRND))
    organic[:, NLEVSOI:] = 0.0
    dz1, z1, zi1 = vertical_grid()
    S["dz"][:] = dz1
    S["zsoi"][:] = z1
    S["zisoi"][:] = zi1
    S["topo_slope"][:] = max(0.070044865858546, 0.2)
    S["topo_std"][:] = 3.96141847422387
    S["n_melt"][:] = 200.0 / np.maximum(10.0, S["topo_std"])
    slope0 = 0.4 ** (-1.0 / 3.0)
    S["micro_sigma"][:] = (S["topo_slope"] + slope0) ** (-3.0)
    depth = np.where(rng.uniform(size=n) < cfg.snow_fraction, rng.uniform(0.02, 1.0, n), 0.0)
    snl = snow_layers(depth, S["dz"], S["zsoi"], S["zisoi"])
    S["snl"][:] = snl
    S.update(soil_hydraulics(pct_sand, pct_clay, organic, cfg.organic_max, z1[NLEVSNO:]))
    S["rootfr"][:] = root_fractions(S["vtype"], params["pft_roota_par"], params["pft_rootb_par"], S["zisoi"])
    active = np.arange(NLEVSNO)[None, :] >= (NLEVSNO - snl)[:, None]
    # init_soil_temp: snow 250 K, soil 274 K (optionally perturbed per column, layer-coherent)
    tsoil = 274.0 + (cfg.soil_temp_spread * rng.uniform(-1.0, 1.0, n))[:, None] * np.linspace(1.0, 0.2, NLEVGRND)[None, :]
    S["t_soisno"][:, NLEVSNO:] = tsoil
    S["t_soisno"][:, :NLEVSNO] = np.where(active, 250.0, 0.0)
    # init_snow_state (cold start) + restoration of the snow cover, SURVEY.md section 8(d)
    S["snw_rds"][:] = np.where(active, SNW_RDS_MIN, 0.0)
    # init_soilh2o_state
    vol = np.minimum(0.15, S["watsat"])
    vol[:, NLEVBED:] = 0.0
    S["h2osoi_vol"][:] = vol
    frozen = S["t_soisno"][:, NLEVSNO:] <= TFRZ
    S["h2osoi_ice"][:, NLEVSNO:] = np.where(frozen, S["dz"][:, NLEVSNO:] * DENICE * vol, 0.0)
    S["h2osoi_liq"][:, NLEVSNO:] = np.where(frozen, 0.0, S["dz"][:, NLEVSNO:] * DENH2O * vol)
    S["h2osoi_ice"][:, :NLEVSNO] = np.where(active, S["dz"][:, :NLEVSNO] * 250.0, 0.0)
    swe = S["h2osoi_ice"][:, :NLEVSNO].sum(axis=1)
    has = snl > 0
    S["snow_depth"][:] = np.where(has, depth, 0.0)
    S["h2osno"][:] = np.where(has, swe, 0.0)
    S["int_snow"][:] = S["h2osno"]
    with np.errstate(divide="ignore", invalid="ignore"):
        fs = np.tanh(depth / (2.5 * ZLND * np.minimum(400.0, swe / depth) / 100.0))
    S["frac_sno"][:] = np.where(has, fs, 0.0)
    S["frac_sno_eff"][:] = S["frac_sno"]
    S["t_grnd"][:] = S["t_soisno"][np.arange(n), NLEVSNO - snl]
    # hard-wired values of ELMInterface::setup (elm_kokkos_interface.cc:98-135)
    S["veg_active"][:] = 1
    S["t_h2osfc"][:] = 274.0
    S["altmax_indx"][:] = 5
    S["t10"][:] = 276.0
    S["t_veg"][:] = 283.0
    S["forc_hgt"][:] = 30.0
    for k in ("forc_hgt_u_patch", "forc_hgt_t_patch", "forc_hgt_q_patch"):
        S[k][:] = 30.0
    if cfg.h2osfc_fraction > 0.0:
        wet = rng.uniform(size=n) < cfg.h2osfc_fraction
        S["h2osfc"][:] = np.where(wet, rng.uniform(0.0, 5.0, n), 0.0)
    S["psn_pft"][:] = psn_rows(params)[S["vtype"]]
    # phenology inputs (time-invariant here) and per-column forcing phases
    tlai = np.where(rng.uniform(size=n) < cfg.bare_fraction, 0.0, rng.uniform(0.1, 4.1, n))
    S["tlai"][:] = tlai
    S["tsai"][:] = np.where(tlai > 0.0, 0.25 * tlai + 0.1, 0.0)
    S["htop"][:] = rng.uniform(0.2, 2.2, n)
    S["hbot"][:] = 0.1 * S["htop"]
    if cfg.aerosol:
        for k, scale in (("bcphi", 2e-13), ("bcpho", 1e-13), ("bcdep", 5e-14), ("dst1_1", 3e-12), ("dst1_2", 1e-12),
                         ("dst2_1", 4e-12), ("dst2_2", 1e-12), ("dst3_1", 2e-12), ("dst3_2", 1e-12),
                         ("dst4_1", 1e-12), ("dst4_2", 5e-13)):
            S["aer_" + k][:] = scale * rng.uniform(0.5, 1.5, n)
    return S


class Forcing:
    """Per-step atmospheric forcing + phenology, a deterministic function of (column, step)."""

    NAMES = ("coszen forc_tbot forc_thbot forc_pbot forc_qbot forc_lwrad forc_u forc_v forc_rain forc_snow "
             "forc_solad forc_solai elai esai frac_veg_nosno_alb").split()

    def __init__(self, ncols: int, seed: int = 7, dtime: float = 1800.0, night_fraction: Optional[float] = None):
        rng = np.random.default_rng(seed)
        self.n, self.dtime = ncols, dtime
        self.phase = rng.uniform(size=ncols)
        self.tbias = rng.uniform(-12.0, 12.0, ncols)
        self.pbot = rng.uniform(95000.0, 103000.0, ncols)
        self.rh = rng.uniform(0.4, 0.9, ncols)
        self.wind = rng.uniform(0.5, 6.5, ncols)
        self.pphase = rng.uniform(size=ncols)
        self.pamp = rng.uniform(size=ncols)
        self.night_fraction = night_fraction

    def at(self, step: int, state: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
        """Forcing for `step`; needs the current snow_depth/frac_sno and tlai/tsai/htop/hbot/vtype for burial."""
        hr = step * self.dtime / 3600.0
        h = np.mod(hr / 24.0 + self.phase, 1.0)
        if self.night_fraction is not None:
            # place exactly the requested share of columns at night at every step
            h = np.where(self.phase < self.night_fraction, 0.05 + 0.15 * self.phase / max(self.night_fraction, 1e-9),
                         0.30 + 0.40 * (self.phase - self.night_fraction) / max(1.0 - self.night_fraction, 1e-9))
        cz = np.maximum(0.0, np.sin(2.0 * np.pi * (h - 0.25)))
        tb = 271.0 + self.tbias + 6.0 * np.sin(2.0 * np.pi * (h - 0.3))
        e = 611.0 * np.exp(17.27 * (tb - 273.15) / (tb - 35.85))
        q = np.maximum(1e-9, self.rh * 0.622 * e / (self.pbot - 0.378 * e))
        event = np.mod(self.pphase * 13.0 + step * 0.07, 1.0) < 0.35
        prec = np.where(event, 2.0e-4 * self.pamp, 0.0)
        frain = np.clip((tb - 273.15) * 0.5, 0.0, 1.0)
        sw = np.maximum(600.0 * cz * 0.5, 0.0)
        rv = np.minimum(0.99, np.maximum(0.17639 + 0.00380 * sw - 9.0039e-06 * sw ** 2 + 8.1351e-09 * sw ** 3, 0.01))
        rn = np.minimum(0.99, np.maximum(0.29548 + 0.00504 * sw - 1.4957e-05 * sw ** 2 + 1.4881e-08 * sw ** 3, 0.01))
        vt, sd, fsno = state["vtype"], state["snow_depth"], state["frac_sno"]
        htop, hbot = state["htop"], state["hbot"]
        woody = (vt > 0) & (vt <= 11)
        ol = np.minimum(np.maximum(sd - hbot, 0.0), htop - hbot)
        fb = np.where(woody, 1.0 - ol / np.maximum(1.0e-06, htop - hbot), 1.0 - np.maximum(np.minimum(sd, 0.2), 0.0) / 0.2)
        elai = np.maximum(state["tlai"] * (1.0 - fsno) + state["tlai"] * fb * fsno, 0.0)
        esai = np.maximum(state["tsai"] * (1.0 - fsno) + state["tsai"] * fb * fsno, 0.0)
        elai = np.where(elai < 0.05, 0.0, elai)
        esai = np.where(esai < 0.05, 0.0, esai)
        return dict(
            coszen=cz, forc_tbot=tb, forc_thbot=tb.copy(), forc_pbot=self.pbot, forc_qbot=q,
            forc_lwrad=0.8 * STEBOL * tb ** 4, forc_u=self.wind, forc_v=np.zeros(self.n),
            forc_rain=frain * prec, forc_snow=(1.0 - frain) * prec,
            forc_solad=np.stack([rv * sw, rn * sw], axis=1), forc_solai=np.stack([(1.0 - rv) * sw, (1.0 - rn) * sw], axis=1),
            elai=elai, esai=esai, frac_veg_nosno_alb=((elai + esai) >= 0.05).astype(np.int32))
