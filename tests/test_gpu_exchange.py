"""The overlapped exchange on the CUDA library: posting step k+1's inputs while step k computes and fetching
step k's outputs while step k+1 computes must give exactly the state and outputs of the serial
upload -> step -> download sequence (double-buffered staging, event ordering between the three streams)."""
import numpy as np
import pytest

import test_exchange_cpu as X

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 777, 20000])
def test_exchange_equals_serial_movement(cuda_lib, params, n):
    X.check_same(X.run_serial(cuda_lib, params, n, 5), X.run_exchange(cuda_lib, params, n, 5))


def test_exchange_with_pinned_buffers_overlaps_and_matches(cuda_lib, params):
    import torch
    n = 65536
    serial = X.run_serial(cuda_lib, params, n, 3)
    # the same run with pinned host buffers (the asynchronous path of cudaMemcpyAsync)
    orig = type(cuda_lib.columns(1)).host_array

    def pinned(self, name, m=None):
        a = orig(self, name, m)
        return torch.from_numpy(a).pin_memory().numpy()

    import elmkernels_b200.abi as abi
    abi.Columns.host_array = pinned
    try:
        piped = X.run_exchange(cuda_lib, params, n, 3)
    finally:
        abi.Columns.host_array = orig
    X.check_same(serial, piped)
