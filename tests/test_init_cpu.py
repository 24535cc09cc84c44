"""One-time column initialisation: host build of the device code (oracle/port) against the
reference's own init functions (oracle/_ref), bit for bit, on every branch of init_snow_layers and of the organic
soil mixing."""
import numpy as np

import init_cases as I


def test_port_matches_reference_init_bit_for_bit(ref_lib, port_lib, params):
    a, b = I.run(ref_lib, params, 700), I.run(port_lib, params, 700)
    for k in I.WRITTEN:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert set(np.unique(a["snl"])) == {0, 1, 2, 3, 4, 5}
    assert (a["t_soisno"][:, :5] == 123.0).any() and (a["t_soisno"][:, 5:] == 274.0).all()
