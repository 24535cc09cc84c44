"""The whole driver loop on the CUDA library against the reference: cold-start initialisation, then 24 steps of
phenology + forcing functors + bookkeeping + the eleven kernel groups, everything produced on the device from
resident series.  Every field of the final state must equal the reference's bit for bit."""
import numpy as np
import pytest

import driver_loop as D
import parity

pytestmark = pytest.mark.gpu


def test_cuda_runs_the_driver_loop_like_the_reference(cuda_lib, checker, params):
    n, steps = 4096, 24
    _, fa, ea = D.run(checker, params, n, steps)
    _, fb, eb = D.run(cuda_lib, params, n, steps)
    assert ea == eb == (0, -1)
    worst = {k: int(parity.mismatch(fa[k], fb[k]).sum()) for k in fa}
    worst = {k: v for k, v in worst.items() if v}
    assert not worst, f"fields with differing elements after {steps} steps: {worst}"
