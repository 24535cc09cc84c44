"""Global parameter tables of the column timestep (PFT constants, SNICAR optics, snow-age, soil colour).

The PFT and SNICAR numbers come from the parameter files that ship with the reference's tests
(converted once by tools/make_param_data.py into data/elm_params.npz).  The reference's snow-age
file and soil-colour file are not in its repository, so those tables are synthetic:
snow-age values are inert for the results (reference snow_hydrology_impl.hh:217-223 clamps the grain
radius to SNW_RDS_MIN from both sides) but must be finite; the soil-colour albedo tables are
monotone in colour class with NIR = 2 x VIS, as in SURVEY.md section 8(d).
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np

_DATA = os.path.join(os.path.dirname(__file__), "data", "elm_params.npz")


def load_params() -> Dict[str, np.ndarray]:
    z = np.load(_DATA)
    p = {k: np.array(z[k]) for k in z.files if k != "pftname"}
    p["pftname"] = z["pftname"]
    # synthetic soil-colour albedo tables [20][2] (VIS, NIR)
    sat_vis = np.linspace(0.25, 0.04, 20)
    dry_vis = np.linspace(0.36, 0.08, 20)
    p["albsat"] = np.stack([sat_vis, 2.0 * sat_vis], axis=1)
    p["albdry"] = np.stack([dry_vis, 2.0 * dry_vis], axis=1)
    # synthetic (finite) snow-age fit parameters [11][31][8]
    shape = (11, 31, 8)
    p["snowage_tau"] = np.full(shape, 10.0)
    p["snowage_kappa"] = np.full(shape, 1.0)
    p["snowage_drdt0"] = np.full(shape, 0.5)
    return p


def psn_rows(params: Dict[str, np.ndarray]) -> np.ndarray:
    """[17][27] table: row v = the reference's PFTData::get_pft_psn(v) (pft_data_impl.hh:64-98)."""
    from .abi import PSN_ORDER
    rows = np.zeros((17, 27))
    for j, n in enumerate(PSN_ORDER):
        a = params["pft_" + n]
        rows[:, j] = a[0] if n == "tc_stress" else a[:17]
    return rows
