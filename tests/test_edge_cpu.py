"""Edge cases, port vs the reference itself, bit for bit: snow capping, all-bare ensembles, ponded surface water,
tiny column counts, and the error convention (a reference `throw` becomes the same error bit on the same column)."""
import numpy as np
import pytest

import edge_cases
import parity
from elmkernels_b200 import abi, ensemble


def _run(a_lib, b_lib, params, state, steps, night_fraction=None):
    n = state["snl"].shape[0]
    a, b = a_lib.columns(n), b_lib.columns(n)
    for c in (a, b):
        c.set_tables(params)
        c.upload_state(state)
    F = ensemble.Forcing(n, seed=3, night_fraction=night_fraction)
    for s in range(steps):
        f = F.at(s, {k: a.download(k) for k in parity.FORCING_STATE})
        for c in (a, b):
            c.upload_state(f)
            c.init_timestep(True)
            c.step()
    return a, b


@pytest.mark.parametrize("case", sorted(edge_cases.CASES))
def test_edge_ensembles_bit_exact(case, ref_lib, port_lib, params):
    st = edge_cases.build(case, 384, params, ref_lib.fields)
    a, b = _run(ref_lib, port_lib, params, st, steps=6)
    if case == "capped_snow":
        assert a.download("do_capsnow").sum() > 50
    for k in ref_lib.field_names:
        assert np.array_equal(a.download(k), b.download(k), equal_nan=True), (case, k)
    assert a.errors() == b.errors()


@pytest.mark.parametrize("n", [1, 2, 33])
def test_tiny_column_counts(n, ref_lib, port_lib, params):
    st = ensemble.make_state(ensemble.EnsembleConfig(ncols=n, seed=n), params, ref_lib.fields)
    a, b = _run(ref_lib, port_lib, params, st, steps=3)
    for k in ref_lib.field_names:
        assert np.array_equal(a.download(k), b.download(k), equal_nan=True), k


@pytest.mark.parametrize("night", [0.0, 1.0])
def test_all_day_and_all_night(night, ref_lib, port_lib, params):
    st = ensemble.make_state(ensemble.EnsembleConfig(ncols=256, seed=8), params, ref_lib.fields)
    a, b = _run(ref_lib, port_lib, params, st, steps=3, night_fraction=night)
    cz = a.download("coszen")
    assert (cz > 0).all() if night == 0.0 else (cz == 0).all()
    for k in ref_lib.field_names:
        assert np.array_equal(a.download(k), b.download(k), equal_nan=True), k


def test_reference_throw_becomes_error_bit(ref_lib, port_lib, params):
    """A snow grain radius outside the SNICAR table makes the reference throw (snow_snicar_impl.hh:76); both
    sides must flag exactly those columns with ELMK_ERR_SNICAR_RADIUS and leave the others untouched."""
    n = 64
    st = ensemble.make_state(ensemble.EnsembleConfig(ncols=n, seed=4, snow_fraction=1.0), params, ref_lib.fields)
    bad_cols = np.array([3, 17, 40])
    st["snw_rds"][bad_cols, 4] = 10.0
    res = []
    for lib in (ref_lib, port_lib):
        c = lib.columns(n)
        c.set_tables(params)
        c.upload_state(st)
        F = ensemble.Forcing(n, seed=3, night_fraction=0.0)
        c.upload_state(F.at(0, st))
        c.init_timestep(True)
        c.step(groups=abi.G_FRAC_WET | abi.G_ALBEDO)
        res.append((c.errors(), c.download("errmask")))
    for (any_err, first), mask in res:
        assert any_err == 1 << 1 and first == 3
        assert set(np.nonzero(mask)[0]) == set(bad_cols)
