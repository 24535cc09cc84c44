"""Shared by the CPU and GPU tests of the one-time column initialisation (elmk_init_columns)."""
import numpy as np

from elmkernels_b200 import ensemble

# every field the reference's initialisation lambda writes (initialize_elm_kokkos.cc:374-431)
WRITTEN = ("psn_pft topo_slope n_melt micro_sigma snl dz zsoi zisoi watsat bsw sucsat watdry watopt watfc tkmg tkdry csol "
           "rootfr t_soisno t_grnd h2osno int_snow snow_depth h2osfc h2ocan frac_h2osfc fwet fdry frac_sno snw_rds "
           "h2osoi_vol h2osoi_liq h2osoi_ice").split()


def inputs(n, seed=31):
    r = np.random.default_rng(seed)
    sand = r.uniform(5.0, 60.0, (n, 15))
    clay = r.uniform(5.0, 40.0, (n, 15))
    org = r.uniform(0.0, 130.0, (n, 15))       # up to organic_max: both sides of the percolation threshold
    org[:, 10:] = 0.0
    org[::9, :] = 130.0                          # om_frac == 1: the all-organic branch
    # snow depths on every branch of init_snow_layers, boundaries included
    edges = np.array([0.0, 0.005, 0.01, 0.02, 0.03, 0.035, 0.04, 0.05, 0.07, 0.1, 0.12, 0.15, 0.18, 0.25, 0.29, 0.35, 0.41,
                      0.5, 0.64, 0.8, 1.5])
    depth = np.where(np.arange(n) < len(edges), np.resize(edges, n), r.uniform(0.0, 1.0, n))
    return sand, clay, org, depth


def run(lib, params, n):
    cols = lib.columns(n)
    cols.set_tables(params)
    r = np.random.default_rng(5)
    st = {k: cols.host_array(k) for k in ("vtype", "topo_slope", "topo_std", "dz", "zsoi", "zisoi", "t_soisno")}
    st["vtype"][:] = r.integers(0, 17, n)                       # includes the bare PFT
    st["topo_slope"][:] = r.uniform(0.0, 0.6, n)                # both sides of the 0.2 floor
    st["topo_std"][:] = r.uniform(1.0, 40.0, n)                 # both sides of the 10 m floor
    dz, z, zi = ensemble.vertical_grid()
    st["dz"][:] = dz; st["zsoi"][:] = z; st["zisoi"][:] = zi
    st["t_soisno"][:] = 123.0                                    # rows the initialisation leaves alone must stay
    cols.upload_state(st)
    sand, clay, org, depth = inputs(n)
    cols.init_columns(sand, clay, org, 130.0, depth)
    out = {k: cols.download(k) for k in WRITTEN}
    cols.close()
    return out
