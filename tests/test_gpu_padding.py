"""Handles whose column count is not a multiple of the 128-column padding: the threads of the padding columns run the
soil-temperature body (its block barriers need whole blocks) on zero-filled state.  After ten steps the padding must
not have raised an error bit, must hold only finite values, and the valid columns must equal those of a handle that
carries the same columns without padding."""
import numpy as np
import pytest
import torch

import parity
from elmkernels_b200 import ensemble

pytestmark = pytest.mark.gpu


class Raw:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def test_padding_columns_stay_clean(cuda_lib, params):
    n = 1000   # padded to 1024
    cfg = ensemble.EnsembleConfig(ncols=1024, seed=5, h2osfc_fraction=0.1, soil_temp_spread=5.0)
    st = ensemble.make_state(cfg, params, cuda_lib.fields)
    F = ensemble.Forcing(1024, seed=9)
    a, b = cuda_lib.columns(n), cuda_lib.columns(1024)
    for c, m in ((a, n), (b, 1024)):
        c.set_tables(params)
        c.upload_state({k: np.ascontiguousarray(v[:m]) for k, v in st.items()})
    for step in range(10):
        f = F.at(step, {k: b.download(k) for k in parity.FORCING_STATE})
        for c, m in ((a, n), (b, 1024)):
            c.upload_state({k: np.ascontiguousarray(v[:m]) for k, v in f.items()})
            c.init_timestep(True)
            c.step()
    assert a.errors() == (0, -1)
    for k in cuda_lib.field_names:
        assert np.array_equal(a.download(k), b.download(k)[:n], equal_nan=True), k
    a.sync()
    dev = torch.device("cuda", 0)
    for k in cuda_lib.field_names:
        _, dt, nl = cuda_lib.fields[k]
        ptr, stride = a.device_ptr(k)
        assert stride == 1024
        t = torch.as_tensor(Raw(ptr, nl * stride, {0: "<f8", 1: "<i4", 2: "|u1"}[dt]), device=dev).reshape(nl, stride)[:, n:]
        if dt == 0:
            assert bool(torch.isfinite(t).all()), f"{k}: non-finite values in the padding columns"
        if k == "errmask":
            assert int(t.abs().max()) == 0, "padding columns raised error bits"
