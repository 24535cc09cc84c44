"""Shared by the CPU and GPU tests of the forcing / phenology producers: synthetic raw series that reach every
branch of the reference's functors (temperature cap, pressure floor, RH vs specific humidity, longwave fallback on
both sides of the plausible range, night, rain / snow ramp, burial of trees vs grasses, bare PFT)."""
import numpy as np

from elmkernels_b200 import ensemble

OUT_ATM = ("forc_tbot forc_thbot forc_pbot forc_qbot forc_lwrad forc_solad forc_solai forc_rain forc_snow forc_u forc_v "
           "forc_hgt forc_hgt_u_patch forc_hgt_t_patch forc_hgt_q_patch").split()
OUT_PHEN = "tlai tsai htop hbot elai esai frac_veg_nosno_alb".split()


def atm_series(n, ntimes=6, seed=3):
    r = np.random.default_rng(seed)
    s = {
        "TBOT": r.uniform(235.0, 312.0, (ntimes, n)),
        "PBOT": r.uniform(60000.0, 103000.0, (ntimes, n)),
        "QBOT": r.uniform(5.0, 100.0, (ntimes, n)),          # relative humidity in percent
        "FLDS": r.uniform(20.0, 700.0, (ntimes, n)),          # both sides of the 50..600 window
        "FSDS": r.uniform(0.0, 900.0, (ntimes, n)),
        "PREC": np.where(r.uniform(size=(ntimes, n)) < 0.4, r.uniform(-1e-5, 2e-3, (ntimes, n)), 0.0),
        "WIND": r.uniform(0.0, 12.0, (ntimes, n)),
    }
    s["TBOT"][:, :5] = 330.0     # above the 323 K cap
    s["PBOT"][:, 5:9] = 3.0e4    # below the 4e4 Pa floor
    s["TBOT"][:, 9:14] = r.uniform(272.0, 276.0, (ntimes, 5))   # inside the rain / snow ramp
    return s


def phen_series(n, seed=4):
    r = np.random.default_rng(seed)
    lai = r.uniform(0.0, 5.0, (3, n))
    lai[:, ::7] = r.uniform(0.0, 0.06, (3, len(range(0, n, 7))))   # around the 0.05 cut
    return {"MLAI": lai, "MSAI": 0.25 * lai + r.uniform(0.0, 0.2, (3, n)),
            "MHTOP": r.uniform(0.1, 25.0, (3, n)), "MHBOT": r.uniform(0.0, 0.1, (3, n))}


def prepare(lib, params, n, seed=21):
    cols = lib.columns(n)
    cols.set_tables(params)
    st = ensemble.make_state(ensemble.EnsembleConfig(ncols=n, seed=seed), params, lib.fields)
    r = np.random.default_rng(seed)
    st["vtype"] = r.integers(0, 17, n).astype(np.int32)            # includes the bare PFT 0
    st["coszen"] = np.where(r.uniform(size=n) < 0.3, 0.0, r.uniform(0.0, 1.0, n))
    st["snow_depth"] = np.where(r.uniform(size=n) < 0.5, 0.0, r.uniform(0.0, 1.5, n))
    st["frac_sno"] = np.where(st["snow_depth"] > 0, r.uniform(0.05, 1.0, n), 0.0)
    cols.upload_state(st)
    for k, v in atm_series(n).items():
        cols.atm_series(k, v)
    for k, v in phen_series(n).items():
        cols.phen_series(k, v)
    return cols


def run(lib, params, n, rh=True):
    cols = prepare(lib, params, n)
    out = []
    for t_idx, wt1 in ((0, 1.0), (2, 0.25), (4, 0.0)):
        cols.atm_forcing(t_idx, wt1, 1.0 - wt1, rh)
        cols.phenology(0 if t_idx < 3 else 1, 0.5 + 0.1 * t_idx, 0.5 - 0.1 * t_idx)
        out.append({k: cols.download(k) for k in OUT_ATM + OUT_PHEN})
    cols.close()
    return out


def solar(lib, params, n=4000, seed=5):
    """elmk_solar_step for columns spread over the globe (poles and the date line included) at several times of the
    year and of the day: [(coszen[n], dayl, max_dayl)]."""
    r = np.random.default_rng(seed)
    lat = np.deg2rad(r.uniform(-90.0, 90.0, n))
    lon = np.deg2rad(r.uniform(-180.0, 180.0, n))
    lat[:6] = np.array([np.pi / 2, -np.pi / 2, 0.0, np.deg2rad(66.6), np.deg2rad(-66.6), 1.5])
    lon[:6] = np.array([0.0, np.pi, -np.pi, 0.0, 3.0, -3.0])
    cols = lib.columns(n)
    cols.set_tables(params)
    cols.set_coordinates(lat, lon)
    out = []
    for doy in (0, 79, 171, 264, 354):
        for frac in (0.0, 0.21, 0.5, 0.77, 0.98):
            for dt in (1800.0, 3600.0, 10800.0):
                dayl, mx = cols.solar_step(dt, doy + frac + 1.0, doy + 1)
                out.append((cols.download("coszen"), dayl, mx))
    # one site for all columns (the reference's own use)
    cols.set_coordinates(0.6, -1.9)
    dayl, mx = cols.solar_step(1800.0, 100.3, 100)
    out.append((cols.download("coszen"), dayl, mx))
    cols.close()
    return out
