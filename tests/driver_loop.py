"""The reference driver's whole loop through the C ABI, device-side producers included:

    setup:     soil grid + texture + initial snow depth  ->  elmk_init_columns           (initialize_kokkos_elm)
    each step: elmk_solar_step: coszen of every column (own latitude / longitude), day lengths (kokkos_init_timestep :27-35)
               elmk_phenology, elmk_atm_forcing                                           (update_phenology, get_forcing)
               elmk_init_timestep, elmk_step(all groups)                                  (init_step_kernel, advance)

Shared by the CPU test (host port against the reference) and the GPU test (CUDA library against the reference)."""
import numpy as np

from elmkernels_b200 import ensemble, forcing

DT = 1800.0
FORC_DT_DAYS = 3.0 / 24.0      # three-hourly forcing records


def series(n, ntimes, seed=11):
    r = np.random.default_rng(seed)
    hours = np.arange(ntimes)[:, None] * 3.0
    tb = 268.0 + r.uniform(-10, 10, n)[None, :] + 5.0 * np.sin(2 * np.pi * (hours - 9.0) / 24.0)
    return {
        "TBOT": tb,
        "PBOT": np.broadcast_to(r.uniform(95000.0, 103000.0, n)[None, :], (ntimes, n)).copy(),
        "QBOT": r.uniform(40.0, 95.0, (ntimes, n)),                      # relative humidity, percent
        "FLDS": r.uniform(180.0, 380.0, (ntimes, n)),
        "FSDS": np.maximum(0.0, 700.0 * np.sin(2 * np.pi * (hours - 6.0) / 24.0)) * r.uniform(0.5, 1.0, (ntimes, n)),
        "PREC": np.where(r.uniform(size=(ntimes, n)) < 0.3, r.uniform(0.0, 4e-4, (ntimes, n)), 0.0),
        "WIND": r.uniform(0.5, 7.0, (ntimes, n)),
    }


def monthly(n, seed=12):
    r = np.random.default_rng(seed)
    lai = np.where(r.uniform(size=n) < 0.2, 0.0, r.uniform(0.3, 4.0, n))[None, :] * np.array([0.8, 1.0, 1.2])[:, None]
    return {"MLAI": lai, "MSAI": 0.25 * lai + 0.1 * (lai > 0), "MHTOP": np.broadcast_to(r.uniform(0.2, 18.0, n), (3, n)).copy(),
            "MHBOT": np.broadcast_to(r.uniform(0.01, 0.15, n), (3, n)).copy()}


def run(lib, params, n, nsteps, collect=None):
    cols = lib.columns(n)
    cols.set_tables(params)
    r = np.random.default_rng(3)
    st = {k: cols.host_array(k) for k in ("vtype", "isoicol", "topo_slope", "topo_std", "dz", "zsoi", "zisoi", "veg_active",
                                          "t_h2osfc", "t10", "t_veg", "altmax_indx")}
    st["vtype"][:] = r.integers(1, 17, n)
    st["isoicol"][:] = r.integers(0, 20, n)
    st["topo_slope"][:] = 0.070044865858546
    st["topo_std"][:] = 3.96141847422387
    dz, z, zi = ensemble.vertical_grid()
    st["dz"][:] = dz; st["zsoi"][:] = z; st["zisoi"][:] = zi
    st["veg_active"][:] = 1
    st["t_h2osfc"][:] = 274.0; st["t10"][:] = 276.0; st["t_veg"][:] = 283.0; st["altmax_indx"][:] = 5
    cols.upload_state(st)
    sand, clay = r.uniform(5.0, 60.0, (n, 15)), r.uniform(5.0, 40.0, (n, 15))
    org = r.uniform(0.0, 60.0, (n, 15)); org[:, 10:] = 0.0
    cols.init_columns(sand, clay, org, 130.0, np.zeros(n))          # cold start without snow, as the reference driver
    ntimes = 2 + int(nsteps * DT / 86400.0 / FORC_DT_DAYS) + 1
    for k, v in series(n, ntimes).items():
        cols.atm_series(k, v)
    for k, v in monthly(n).items():
        cols.phen_series(k, v)
    # every column at its own place on the globe (the reference driver has one site)
    cols.set_coordinates(np.deg2rad(r.uniform(-70.0, 75.0, n)), np.deg2rad(r.uniform(-180.0, 180.0, n)))
    out = []
    for step in range(nsteps):
        sec = step * DT
        centred_days = (sec + DT / 2.0) / 86400.0
        doy = 195                                                   # 14 July
        dayl, max_dayl = cols.solar_step(DT, doy + sec / 86400.0 + 1.0, doy + 1)
        m1 = forcing.first_month_idx(7, 14, sec % 86400.0) - 5     # series holds June, July, August
        pw1, pw2 = forcing.monthly_data_weights(7, 14, sec % 86400.0)
        cols.phenology(m1, pw1, pw2)
        t_idx, w1, w2 = forcing.forcing_time_weights(centred_days, FORC_DT_DAYS)
        cols.atm_forcing(t_idx, w1, w2, True)
        cols.init_timestep(False)                                   # forcing heights were just set by atm_forcing
        cols.step(dtime=DT, dayl=dayl, max_dayl=max_dayl)
        if collect and step in collect:
            out.append(cols.download_state())
    err = cols.errors()
    final = cols.download_state()
    cols.close()
    return out, final, err
