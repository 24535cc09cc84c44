"""Forcing functors and phenology: the host build of the device code (oracle/port) against
the reference's own functors (oracle/_ref), bit for bit; host time bookkeeping against the reference's formulas."""
import numpy as np
import pytest

import forcing_cases as F
from elmkernels_b200 import forcing


@pytest.mark.parametrize("rh", [True, False])
def test_port_matches_reference_functors_bit_for_bit(ref_lib, port_lib, params, rh):
    a, b = F.run(ref_lib, params, 1500, rh), F.run(port_lib, params, 1500, rh)
    for sa, sb in zip(a, b):
        for k in sa:
            np.testing.assert_array_equal(sa[k], sb[k], err_msg=k)
    # the cases reach the branches they were built for
    last = a[-1]
    assert (last["forc_tbot"] == 323.0).any() and (last["forc_pbot"] == 4.0e4).any()
    assert (last["forc_rain"] > 0).any() and (last["forc_snow"] > 0).any()
    assert (last["frac_veg_nosno_alb"] == 0).any() and (last["frac_veg_nosno_alb"] == 1).any()


def test_series_bounds_are_checked(port_lib, params):
    from elmkernels_b200 import abi
    cols = F.prepare(port_lib, params, 32)
    with pytest.raises(abi.ElmkError):
        cols.atm_forcing(5, 0.5, 0.5)      # needs records 5 and 6 of a 6-record series
    with pytest.raises(abi.ElmkError):
        cols.phenology(2, 0.5, 0.5)
    cols.close()


def test_time_weights_follow_the_reference_formulas():
    # a step centred 1.25 forcing intervals after the start of the data (atm_data_impl.hh:191-199)
    t, w1, w2 = forcing.forcing_time_weights(1.25 * 0.125, 0.125)
    assert (t, w1, w2) == (1, 0.75, 0.25)
    # a model time a hair before a record boundary belongs to that record (forc_t_idx_aligned), not to the one before
    assert forcing.forcing_time_index(2.9999999999 * 0.125, 0.125) == 3 and forcing.forcing_time_index(3.0 * 0.125, 0.125) == 3
    assert forcing.forcing_time_index(2.999 * 0.125, 0.125) == 2
    assert forcing.forcing_time_weights(2.9999999999 * 0.125, 0.125) == (3, 1.0, 0.0)
    # monthly_data.cc:7-61: mid-January sits exactly on January's value; early January interpolates from December
    assert forcing.first_month_idx(1, 16, 43200.0) == 0 and forcing.monthly_data_weights(1, 16, 43200.0) == (1.0, 0.0)
    assert forcing.first_month_idx(1, 1, 0.0) == 11
    w1, w2 = forcing.monthly_data_weights(1, 1, 0.0)
    assert abs(w1 - 0.5) < 1e-15 and abs(w1 + w2 - 1.0) < 1e-15
    assert forcing.first_month_idx(12, 31, 0.0) == 11


def test_solar_geometry_port_matches_reference_bit_for_bit(ref_lib, port_lib, params):
    """average_cosz / daylength / max_daylength (init_timestep_kokkos.cc:27-35): the reference's functions per column
    against the split evaluation of the product (latitude terms once on the host, arc cosine and sines per column)."""
    a, b = F.solar(ref_lib, params), F.solar(port_lib, params)
    lit = 0
    for (ca, da, ma), (cb, db, mb) in zip(a, b):
        assert np.array_equal(ca.view(np.uint64), cb.view(np.uint64))
        assert da == db and ma == mb
        lit += int((ca > 0).sum())
    assert lit > 1000 and np.all(a[-1][0] == a[-1][0][0])
