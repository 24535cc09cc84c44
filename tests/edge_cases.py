"""Edge-case ensembles shared by the CPU (port vs reference) and GPU (CUDA vs reference) parity tests."""
import numpy as np

from elmkernels_b200 import ensemble


def capped_snow(state, rng):
    """Columns above the snow-capping threshold (h2osno > 1000 kg/m2 => do_capsnow) with five thick layers."""
    n = state["snl"].shape[0]
    sel = rng.uniform(size=n) < 0.5
    dzs = np.array([0.02, 0.05, 0.11, 0.23, 3.0])
    for k in ("dz",):
        state[k][sel, :5] = dzs
    state["snl"][sel] = 5
    state["h2osoi_ice"][sel, :5] = dzs * 350.0
    state["h2osoi_liq"][sel, :5] = 0.0
    state["t_soisno"][sel, :5] = 255.0
    state["snw_rds"][sel, :5] = 54.526
    zi = state["zisoi"]
    z = state["zsoi"]
    for j in range(4, -1, -1):
        z[sel, j] = zi[sel, j + 1] - 0.5 * state["dz"][sel, j]
        zi[sel, j] = zi[sel, j + 1] - state["dz"][sel, j]
    swe = state["h2osoi_ice"][:, :5].sum(axis=1)
    depth = state["dz"][:, :5].sum(axis=1)
    state["h2osno"][sel] = swe[sel]
    state["int_snow"][sel] = swe[sel]
    state["snow_depth"][sel] = depth[sel]
    state["frac_sno"][sel] = 1.0
    state["frac_sno_eff"][sel] = 1.0
    state["t_grnd"][sel] = 255.0
    return state


def all_bare(state, rng):
    state["tlai"][:] = 0.0
    state["tsai"][:] = 0.0
    return state


def hot_and_wet(state, rng):
    """Warm soil, ponded surface water, wet canopy: exercises the surface-water Newton iteration and its
    phase change, quirk 5 (1e97 ground heat flux) included."""
    n = state["snl"].shape[0]
    state["h2osfc"][:] = rng.uniform(0.0, 8.0, n)
    state["t_h2osfc"][:] = rng.uniform(268.0, 285.0, n)
    state["h2ocan"][:] = rng.uniform(0.0, 0.3, n)
    state["t_soisno"][:, 5:] += rng.uniform(-6.0, 12.0, (n, 1))
    return state


CASES = {"capped_snow": capped_snow, "all_bare": all_bare, "hot_and_wet": hot_and_wet}


def build(name, n, params, fields, seed=99):
    rng = np.random.default_rng(seed)
    st = ensemble.make_state(ensemble.EnsembleConfig(ncols=n, seed=seed, soil_temp_spread=4.0), params, fields)
    return CASES[name](st, rng)
