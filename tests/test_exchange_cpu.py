"""The overlapped per-step exchange (elmk_exchange_*) on the CPU checker libraries: same call sequence as the
product library, results identical to the plain upload / download calls."""
import numpy as np
import pytest

from elmkernels_b200 import ensemble

IN = ("coszen forc_tbot forc_thbot forc_pbot forc_qbot forc_lwrad forc_u forc_v forc_rain forc_snow "
      "forc_solad forc_solai elai esai frac_veg_nosno_alb").split()
OUT = "dtend_column_h2o errh2o errh2osno dwb errsol errlon errseb netrad errmask".split()


def run_serial(lib, params, n, steps):
    cols = lib.columns(n)
    cols.set_tables(params)
    st = ensemble.make_state(ensemble.EnsembleConfig(ncols=n, seed=5), params, lib.fields)
    cols.upload_state(st)
    forcing = ensemble.Forcing(n, seed=9)
    outs = []
    for k in range(steps):
        f = forcing.at(k, st)
        cols.upload_state({name: f[name] for name in IN})
        cols.init_timestep(True)
        cols.step()
        outs.append({name: cols.download(name) for name in OUT})
    final = cols.download_state()
    cols.close()
    return outs, final


def run_exchange(lib, params, n, steps):
    cols = lib.columns(n)
    cols.set_tables(params)
    st = ensemble.make_state(ensemble.EnsembleConfig(ncols=n, seed=5), params, lib.fields)
    cols.upload_state(st)
    forcing = ensemble.Forcing(n, seed=9)
    x = cols.exchange(IN, OUT)

    def inputs(k):
        f = forcing.at(k, st)
        return [np.ascontiguousarray(f[name], dtype=cols.host_array(name).dtype) for name in IN]

    bufs = [[cols.host_array(name) for name in OUT] for _ in range(2)]
    outs = []
    x.post(inputs(0))
    for k in range(steps):
        x.commit()
        if k + 1 < steps:
            x.post(inputs(k + 1))       # in flight while step k runs
        cols.init_timestep(True)
        cols.step()
        x.fetch(bufs[k & 1])
        if k >= 1:
            x.wait()
            outs.append({name: a.copy() for name, a in zip(OUT, bufs[(k - 1) & 1])})
    x.wait()
    outs.append({name: a.copy() for name, a in zip(OUT, bufs[(steps - 1) & 1])})
    final = cols.download_state()
    cols.close()
    return outs, final


def check_same(a, b):
    outs_a, fin_a = a
    outs_b, fin_b = b
    assert len(outs_a) == len(outs_b)
    for oa, ob in zip(outs_a, outs_b):
        for k in OUT:
            np.testing.assert_array_equal(oa[k], ob[k], err_msg=k)
    for k in fin_a:
        np.testing.assert_array_equal(fin_a[k], fin_b[k], err_msg=k)


def test_exchange_equals_upload_download_on_the_port(port_lib, params):
    check_same(run_serial(port_lib, params, 96, 4), run_exchange(port_lib, params, 96, 4))


def test_exchange_call_order_is_enforced(port_lib, params):
    from elmkernels_b200 import abi
    cols = port_lib.columns(8)
    x = cols.exchange(IN[:2], OUT[:1])
    with pytest.raises(abi.ElmkError):
        x.commit()                       # nothing posted
    a = [cols.host_array(n) for n in IN[:2]]
    x.post(a); x.post(a)
    with pytest.raises(abi.ElmkError):
        x.post(a)                        # two posts already in flight
    cols.close()
