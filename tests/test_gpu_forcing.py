"""Forcing functors and phenology on the CUDA library against the reference's functors (oracle/_ref): closed-form
arithmetic, so every element within 1e-12 relative (the longwave fallback goes through exp); integer flags equal."""
import numpy as np
import pytest

import forcing_cases as F
import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rh", [True, False])
@pytest.mark.parametrize("n", [130, 40000])
def test_cuda_forcing_and_phenology_match_the_reference(cuda_lib, checker, params, n, rh):
    a, b = F.run(checker, params, n, rh), F.run(cuda_lib, params, n, rh)
    for sa, sb in zip(a, b):
        for k in sa:
            bad = parity.mismatch(sa[k], sb[k])
            assert not bad.any(), f"{k}: {int(bad.sum())} elements differ, e.g. {sa[k][bad][:3]} vs {sb[k][bad][:3]}"


def test_cuda_solar_geometry_matches_the_reference(cuda_lib, checker, params):
    """coszen of every column (device: arc cosine + four sines per column from the libm restatements) and the two day
    lengths, against the reference's functions, bit for bit."""
    a, b = F.solar(checker, params, n=20000), F.solar(cuda_lib, params, n=20000)
    for (ca, da, ma), (cb, db, mb) in zip(a, b):
        bad = parity.mismatch(ca, cb)
        assert not bad.any(), f"coszen: {int(bad.sum())} columns differ, e.g. {ca[bad][:3]} vs {cb[bad][:3]}"
        assert da == db and ma == mb
