"""elmk_init_columns on the CUDA library against the reference's init functions: closed form, 1e-12 relative (pow / exp /
tanh differ from glibc in the last bit); integers, layer geometry and the 1e36 markers equal."""
import numpy as np
import pytest

import init_cases as I
import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [21, 5000])
def test_cuda_init_matches_the_reference(cuda_lib, checker, params, n):
    a, b = I.run(checker, params, n), I.run(cuda_lib, params, n)
    for k in I.WRITTEN:
        if k in ("snl", "dz", "zsoi", "zisoi", "t_soisno", "snw_rds", "h2osoi_vol"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
        else:
            bad = parity.mismatch(a[k], b[k])
            assert not bad.any(), f"{k}: {int(bad.sum())} elements differ, e.g. {a[k][bad][:3]} vs {b[k][bad][:3]}"
