"""The whole driver loop (initialisation, phenology, forcing functors, bookkeeping, eleven kernel groups) on the host
port against the reference, bit for bit, 24 steps."""
import numpy as np

import driver_loop as D


def test_port_runs_the_driver_loop_like_the_reference(ref_lib, port_lib, params):
    _, fa, ea = D.run(ref_lib, params, 600, 24)
    _, fb, eb = D.run(port_lib, params, 600, 24)
    assert ea == eb == (0, -1)
    for k in fa:
        np.testing.assert_array_equal(fa[k], fb[k], err_msg=k)
    # the loop went through day and night, rain and (after the cold soil froze some water) both soil phases
    assert fa["h2osoi_ice"][:, 5:].max() >= 0.0 and np.isfinite(fa["t_veg"]).all()
