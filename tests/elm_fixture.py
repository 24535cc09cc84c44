"""Replay of the reference's ELM (Fortran) golden dump of test_CanHydro (BASELINE.json config 1) through a
library exporting the elmk C ABI: the 48 records are 48 columns; the five kernels test/test_CanHydro.cc:203-218
calls are group a3 (interception, ground_flux, snow_init, fraction_h2osfc) followed by group a1 (fraction_wet,
which the test evaluates with the post-interception canopy water)."""
import os

import numpy as np

from elmkernels_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RENAME = {"forc_t": "forc_tbot", "z": "zsoi", "zi": "zisoi"}
# the 38 variables test_CanHydro.cc:221-261 compares (those that are per-column fields of the ABI)
COMPARED = ("frac_veg_nosno elai esai h2ocan qflx_snwcp_liq qflx_snwcp_ice qflx_rain_grnd qflx_snow_grnd do_capsnow "
            "t_grnd qflx_snow_melt n_melt snow_depth h2osno int_snow frac_sno_eff frac_sno snl micro_sigma h2osfc "
            "frac_h2osfc forc_rain forc_snow forc_t fwet fdry dz z zi swe_old h2osoi_liq h2osoi_ice t_soisno "
            "frac_iceold snw_rds").split()


def replay(lib, params):
    z = np.load(os.path.join(ROOT, "tests", "golden", "elm_canopy_hydrology.npz"))
    n = len(z["steps"])
    assert np.all(z["in_oldfflag"] == z["in_oldfflag"][0]) and np.all(z["in_dewmx"] == z["in_dewmx"][0])
    cols = lib.columns(n)
    cols.set_tables(params, dewmx=float(z["in_dewmx"][0, 0]), oldfflag=int(z["in_oldfflag"][0, 0]))
    for key in z.files:
        if not key.startswith("in_"):
            continue
        name = RENAME.get(key[3:], key[3:])
        if name in lib.fields:
            a = z[key]
            _, dt, nl = lib.fields[name]
            a = np.nan_to_num(a, nan=0.0) if dt != abi.F64 else a
            cols.upload(name, a.reshape(n) if nl == 1 else a)
    cols.step(dtime=1800.0, groups=abi.G_CANOPY_HYDROLOGY)
    cols.step(dtime=1800.0, groups=abi.G_FRAC_WET)
    worst, got_all = {}, {}
    for v in COMPARED:
        ref = z["out_" + v]
        got = cols.download(RENAME.get(v, v)).astype(np.float64).reshape(ref.shape)
        got_all[v] = got
        d = np.abs(got - ref)
        s = np.maximum(np.abs(ref), 1e-300)
        worst[v] = float(np.max(np.where(d == 0, 0.0, d / s)))
    replay.last = got_all   # the replayed values themselves, for library-against-library comparisons
    return worst


CANFLUX_RENAME = {"forc_t": "forc_tbot", "forc_q": "forc_qbot", "forc_th": "forc_thbot"}
CANFLUX_COMPARED = ("btran displa z0mv z0hv z0qv t_veg qflx_tran_veg qflx_evap_veg eflx_sh_veg eflx_sh_grnd eflx_sh_snow "
                    "eflx_sh_soil eflx_sh_h2osfc qflx_evap_soi qflx_ev_snow qflx_ev_soil qflx_ev_h2osfc dlrad ulrad cgrnds "
                    "cgrndl cgrnd t_ref2m q_ref2m rh_ref2m h2ocan rootr eff_porosity").split()


def replay_canopy_fluxes(lib, params, day=False):
    """The ELM Fortran dump of test_CanFlux (test/test_CanFlux.cc:328-453) through group a7.  The kernel-group
    wrapper derives the CO2 partial pressure from a constant 355 ppmv (canopy_fluxes_kokkos.cc:49-51) while the
    dump carries ELM's time-varying value.  day=False: the 47 night records - where the stomatal root-find and hence
    CO2 do not enter - replayed with the wrapper's constants; they exercise moisture stress, the Monin-Obukhov
    iteration, the leaf energy balance and compute_flux.  day=True: the 50 daytime records with the dump's partial
    pressures handed in through elmk_set_gas_pressures, as test_CanFlux.cc hands them to the library functions; they
    exercise photosynthesis (C3, vtype 12) on top."""
    from elmkernels_b200.params import psn_rows
    z = np.load(os.path.join(ROOT, "tests", "golden", "elm_canopy_fluxes.npz"))
    night = (z["in_parsun_z"][:, 0] <= 0.0) & (z["in_parsha_z"][:, 0] <= 0.0)
    pick = ~night if day else night
    n = int(pick.sum())
    cols = lib.columns(n)
    cols.set_tables(params)
    for key in z.files:
        if not key.startswith("in_"):
            continue
        name = CANFLUX_RENAME.get(key[3:], key[3:])
        if name in lib.fields:
            a = z[key][pick]
            _, dt, nl = lib.fields[name]
            cols.upload(name, a.reshape(n) if nl == 1 else a)
    cols.upload("psn_pft", np.repeat(psn_rows(params)[12][None, :], n, axis=0))
    cols.fill("veg_active", 1)
    if day:
        cols.set_gas_pressures(z["in_forc_pco2"][pick].reshape(n), z["in_forc_po2"][pick].reshape(n))
    cols.step(dtime=1800.0, dayl=float(z["in_dayl"][0, 0]), max_dayl=float(z["in_max_dayl"][0, 0]), groups=abi.G_CANOPY_FLUXES)
    assert cols.errors() == (0, -1)
    worst, got_all = {}, {}
    for v in CANFLUX_COMPARED:
        ref = z["out_" + v][pick]
        got = cols.download(v).astype(np.float64).reshape(ref.shape)
        got_all[v] = got
        d = np.abs(got - ref)
        s = np.maximum(np.abs(ref), 1e-300)
        # canopy water is ~1e-19..1e-6 kg/m2 in these records: differences below 1e-15 kg/m2 are rounding residue
        worst[v] = float(np.max(np.where(d <= 1e-15, 0.0, d / s)))
    replay_canopy_fluxes.last = got_all
    return n, worst


SURFALB_RENAME = {"mss_cnc_bcphi": "cnc_bcphi", "mss_cnc_bcpho": "cnc_bcpho", "mss_cnc_dst1": "cnc_dst1",
                  "mss_cnc_dst2": "cnc_dst2", "mss_cnc_dst3": "cnc_dst3", "mss_cnc_dst4": "cnc_dst4"}
# what test_SurfAlb.cc:540-595 compares, as far as it is a per-column field of the ABI (fabd_sun / fabd_sha are
# wrapper-local in the reference)
SURFALB_COMPARED = ("albsod albsoi albsnd albsni albgrd albgri flx_absdv flx_absdn flx_absiv flx_absin tlai_z fsun_z "
                    "fabd_sun_z fabd_sha_z fabi_sun_z fabi_sha_z albd ftid ftdd fabd albi ftii fabi fabi_sun fabi_sha").split()


def replay_surface_albedo(lib, params):
    """The ELM Fortran dump of test_SurfAlb (test/test_SurfAlb.cc:391-537; 95 records, 47 of them sunlit, all with a
    thin snow cover h2osno = 0.015 m without snow layers - the flg_nosnl branch of SNICAR) through group a2: the test
    calls surface_albedo::init_timestep, soil_albedo, the four SNICAR functions for direct and for diffuse light,
    ground_albedo, flux_absorption_factor, canopy_layer_lai and two_stream_solver - the body of kokkos_albedo_snicar.
    The soil-colour albedos arrive per record in the dump; they become rows of the colour table."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "elm_surface_albedo.npz"))
    n = len(z["steps"])
    pairs = np.concatenate([z["in_albsat"], z["in_albdry"]], axis=1)
    uniq, colour = np.unique(pairs, axis=0, return_inverse=True)
    assert len(uniq) <= 20
    p = dict(params)
    p["albsat"] = np.array(params["albsat"], dtype=np.float64)
    p["albdry"] = np.array(params["albdry"], dtype=np.float64)
    p["albsat"][:len(uniq)] = uniq[:, :2]
    p["albdry"][:len(uniq)] = uniq[:, 2:]
    cols = lib.columns(n)
    cols.set_tables(p)
    for key in z.files:
        if not key.startswith("in_"):
            continue
        name = SURFALB_RENAME.get(key[3:], key[3:])
        if name in lib.fields and name not in ("albsat", "albdry"):
            a = z[key]
            _, dt, nl = lib.fields[name]
            a = np.nan_to_num(a, nan=0.0, posinf=0.0, neginf=0.0) if dt != abi.F64 else a
            cols.upload(name, a.reshape(n) if nl == 1 else a)
    cols.upload("isoicol", colour.reshape(n).astype(np.int32))
    cols.fill("vtype", 12)
    cols.step(dtime=1800.0, groups=abi.G_ALBEDO)
    assert cols.errors() == (0, -1)
    worst, got_all = {}, {}
    for v in SURFALB_COMPARED:
        ref = z["out_" + v]
        got = cols.download(v).astype(np.float64).reshape(ref.shape)
        got_all[v] = got
        d = np.abs(got - ref)
        s = np.maximum(np.abs(ref), 1e-300)
        worst[v] = float(np.max(np.where(d == 0, 0.0, d / s)))
    replay_surface_albedo.last = got_all
    return n, worst
