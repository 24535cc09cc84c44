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
