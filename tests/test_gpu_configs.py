"""BASELINE.json configs 2-4 as parity cases, at sizes the checker finishes in seconds:
  config 2  SurfaceAlbedo + SurfaceRadiation two-stream       groups a1 + a2 + a4, 25 % night columns, after two
                                                                spin-up steps of the whole chain
  config 3  CanopyHydrology + CanopyTemperature + BareGround   groups a3 + a5 + a6, half of the columns bare, standing
                                                                surface water on a fifth
  config 4  CanopyFluxes stability + photosynthesis iteration  group a7, 40 % night, all PFTs incl. C4
Every field must equal the checker's bit for bit (tests/parity.py).
Config 1 (test_CanHydro on the ELM forcing of test/data) is tests/test_gpu_parity.py::test_elm_fortran_dump_of_test_canhydro;
config 5 (full step, persistent state) is test_full_chain_free_running and bench.py."""
import numpy as np
import pytest

import parity
from elmkernels_b200 import abi, ensemble

pytestmark = pytest.mark.gpu
N = 16384


def spun_up(checker, cuda_lib, params, cfg, night_fraction, steps=2):
    pair = parity.Pair(checker, cuda_lib, params, cfg, night_fraction=night_fraction)
    for _ in range(steps):
        pair.begin_step()
        pair.run()
        pair.resync()
    pair.begin_step()
    return pair


def run_groups(pair, groups):
    for g in groups:
        pair.run(groups=g)


def test_config2_albedo_and_surface_radiation(cuda_lib, checker, params):
    pair = spun_up(checker, cuda_lib, params, ensemble.EnsembleConfig(ncols=N, seed=20240002), night_fraction=0.25)
    night = pair.a.download("coszen") <= 0.0
    assert 0.2 < night.mean() < 0.3
    run_groups(pair, [abi.G_FRAC_WET, abi.G_ALBEDO, abi.G_SURFACE_RADIATION])
    bad = pair.compare()
    assert not bad, parity.fmt(bad)
    assert pair.b.errors() == pair.a.errors() == (0, -1)


def test_config3_hydrology_temperature_bareground(cuda_lib, checker, params):
    cfg = ensemble.EnsembleConfig(ncols=N, seed=20240003, bare_fraction=0.5, h2osfc_fraction=0.2)
    pair = spun_up(checker, cuda_lib, params, cfg, night_fraction=None)
    bare = pair.a.download("frac_veg_nosno") == 0
    assert 0.4 < bare.mean() < 0.7
    run_groups(pair, [abi.G_FRAC_WET, abi.G_ALBEDO, abi.G_CANOPY_HYDROLOGY, abi.G_SURFACE_RADIATION])
    pair.resync()
    run_groups(pair, [abi.G_CANOPY_TEMPERATURE, abi.G_BAREGROUND_FLUXES])
    bad = pair.compare()
    assert not bad, parity.fmt(bad)


def test_config4_canopy_fluxes_mixed_pft(cuda_lib, checker, params):
    cfg = ensemble.EnsembleConfig(ncols=N, seed=20240004, bare_fraction=0.0)
    pair = spun_up(checker, cuda_lib, params, cfg, night_fraction=0.4)
    vt = pair.a.download("vtype")
    assert set(np.unique(vt)) == set(range(1, 17)), "every PFT, C4 grass (14) included"
    run_groups(pair, [abi.G_FRAC_WET, abi.G_ALBEDO, abi.G_CANOPY_HYDROLOGY, abi.G_SURFACE_RADIATION,
                      abi.G_CANOPY_TEMPERATURE, abi.G_BAREGROUND_FLUXES])
    pair.resync()
    pair.run(groups=abi.G_CANOPY_FLUXES)
    bad = pair.compare()
    assert not bad, parity.fmt(bad)
    night = (pair.a.download("parsun_z").reshape(N, -1)[:, 0] <= 0.0) & (pair.a.download("parsha_z").reshape(N, -1)[:, 0] <= 0.0)
    assert 0.3 < night.mean() < 0.6
