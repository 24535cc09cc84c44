"""GPU parity tests: the CUDA library (through the C ABI) against the oracle on the same seeded inputs.

Checker = oracle/_ref (the reference's own code) when its prebuilt library is present, else oracle/port.
The bar is bit-for-bit equality of every field (tests/parity.py): BASELINE.json's 1e-12 for closed-form kernel groups
and 1e-8 for the iterative CanopyFluxes and SoilTemperature paths are implied, with no allowance of any kind."""
import os

import numpy as np
import pytest

import parity
from elmkernels_b200 import abi, ensemble

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_backend_is_cuda(cuda_lib):
    assert cuda_lib.backend == "cuda-sm100a"


def test_upload_download_round_trip(cuda_lib):
    """Layout conversion (reference host layout <-> column-innermost HBM layout), ragged sizes."""
    rng = np.random.default_rng(1)
    for n in (1, 63, 64, 65, 1000, 4097):
        cols = cuda_lib.columns(n)
        for name in ("t_soisno", "zisoi", "psn_pft", "forc_tbot", "snl", "imelt", "veg_active", "forc_solad"):
            _, dt, nl = cuda_lib.fields[name]
            shape = (n,) if nl == 1 else (n, nl)
            a = (rng.uniform(-1e3, 1e3, shape) if dt == abi.F64 else rng.integers(0, 100, shape)).astype(abi._NP[dt])
            cols.upload(name, a)
            assert np.array_equal(cols.download(name), a), (n, name)
        # partial ranges
        if n > 10:
            a = rng.uniform(size=(5, 20))
            cols.upload("dz", a, col0=3)
            assert np.array_equal(cols.download("dz", col0=3, n=5), a)
        cols.close()


def test_fill_and_errors(cuda_lib):
    cols = cuda_lib.columns(777)
    cols.fill("t_grnd", 271.5)
    assert np.all(cols.download("t_grnd") == 271.5)
    assert cols.errors() == (0, -1)
    e = np.zeros(777, np.int32)
    e[500] = 4
    e[123] = 2
    cols.upload("errmask", e)
    assert cols.errors() == (6, 123)
    cols.clear_errors()
    assert cols.errors() == (0, -1)


def test_step_requires_tables(cuda_lib):
    cols = cuda_lib.columns(64)
    with pytest.raises(abi.ElmkError):
        cols.step()


def test_unsupported_land_unit_is_rejected(cuda_lib, params):
    cols = cuda_lib.columns(64)
    with pytest.raises(abi.ElmkError):
        cols.set_tables(params, land=dict(ltype=3))


@pytest.mark.parametrize("seed,h2osfc,tspread", [(20240005, 0.0, 0.0), (20240003, 0.2, 8.0)])
def test_each_group_in_isolation(cuda_lib, checker, params, seed, h2osfc, tspread):
    """Every kernel group on inputs identical to the checker's (state re-synchronised before each group)."""
    cfg = ensemble.EnsembleConfig(ncols=4096, seed=seed, h2osfc_fraction=h2osfc, soil_temp_spread=tspread)
    pair = parity.Pair(checker, cuda_lib, params, cfg)
    for step in range(4):
        pair.begin_step()
        bad = pair.compare(0.0, names=["h2osno_old", "dtbegin_column_h2o", "do_capsnow", "frac_veg_nosno", "frac_iceold"])
        assert not bad, f"init_timestep step {step}\n{parity.fmt(bad)}"
        for g in range(abi.G_ALL.bit_length()):
            pair.resync()
            pair.run(groups=1 << g)
            bad = pair.compare()
            assert not bad, f"step {step} group {abi.GROUP_NAMES[g]}: bits differ\n{parity.fmt(bad)}"
    assert pair.b.errors() == pair.a.errors()


def test_full_chain_free_running(cuda_lib, checker, params):
    """48 steps (one day) of the full chain with state persistent on the device, never re-synchronised."""
    cfg = ensemble.EnsembleConfig(ncols=8192, seed=20240005, soil_temp_spread=6.0)
    pair = parity.Pair(checker, cuda_lib, params, cfg)
    for step in range(48):
        pair.begin_step()
        pair.run()
        if step % 8 == 7 or step < 2:
            bad = pair.compare()   # every field, the eight balance diagnostics of a11 included
            assert not bad, f"step {step}: bits differ\n{parity.fmt(bad)}"
    assert pair.b.errors() == pair.a.errors() == (0, -1)
    # the optional global diagnostic (sum/min/max over columns): min and max exactly, sums up to their association
    ra, rb = pair.a.diag_reduce(), pair.b.diag_reduce()
    assert np.array_equal(ra[8:], rb[8:])
    assert np.allclose(ra[:8], rb[:8], rtol=1e-12, atol=1e-12 * np.max(np.abs(ra[:8])))


def test_split_and_fused_plans_agree(cuda_lib, params):
    """The production launch plan (SNICAR as warp tasks, re-packed CanopyFluxes, compacted bare-ground columns, fused
    snow + flux + diagnostics launch) and one plain launch per kernel group with one thread per column give identical
    bits."""
    for cfg in (ensemble.EnsembleConfig(ncols=2048, seed=3, h2osfc_fraction=0.1, soil_temp_spread=5.0),
                ensemble.EnsembleConfig(ncols=3000, seed=17, snow_fraction=1.0, soil_temp_spread=5.0)):
        pair = parity.Pair(cuda_lib, cuda_lib, params, cfg, night_fraction=0.1)
        pair.a.set_plan("split")
        for _ in range(6):
            pair.begin_step()
            pair.run()
        assert not pair.compare()
        assert pair.b.launch_count > 0 and pair.a.launch_count > 0


def test_golden_vectors(cuda_lib, params):
    """The committed vectors produced by the reference itself (tests/golden/chain_64col.npz)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "chain_64col.npz"))
    n = int(z["ncols"])
    cols = cuda_lib.columns(n)
    cols.set_tables(params)
    cols.upload_state({k[3:]: z[k] for k in z.files if k.startswith("s0_")})
    for s in range(int(z["nsteps"])):
        cols.upload_state({k[len(f"f{s}_"):]: z[k] for k in z.files if k.startswith(f"f{s}_")})
        cols.init_timestep(True)
        cols.step()
    assert cols.errors() == (0, -1)
    bad = {}
    for k in z.files:
        if k.startswith("out_"):
            m = parity.mismatch(z[k], cols.download(k[4:]))
            if m.any():
                bad[k[4:]] = int(m.sum())
    assert not bad, bad


def test_large_ensemble_properties(cuda_lib, params):
    """Size-independent properties at a bench-like size (no oracle): column independence (a column's result
    does not depend on which other columns share the launch) and the shortwave/longwave identities."""
    n = 1 << 18
    cfg = ensemble.EnsembleConfig(ncols=n, seed=77, soil_temp_spread=6.0)
    st = ensemble.make_state(cfg, params, cuda_lib.fields)
    F = ensemble.Forcing(n, seed=5)
    big = cuda_lib.columns(n)
    big.set_tables(params)
    big.upload_state(st)
    lo, m = 100_003, 4096
    sub = cuda_lib.columns(m)
    sub.set_tables(params)
    sub.upload_state({k: v[lo:lo + m] for k, v in st.items()})
    for step in range(4):
        f = F.at(step, {k: big.download(k) for k in parity.FORCING_STATE})
        big.upload_state(f)
        sub.upload_state({k: v[lo:lo + m] for k, v in f.items()})
        for c in (big, sub):
            c.init_timestep(True)
            c.step()
    assert big.errors() == (0, -1)
    for k in cuda_lib.field_names:
        assert np.array_equal(big.download(k, col0=lo, n=m), sub.download(k), equal_nan=True), k
    assert np.max(np.abs(big.download("errsol"))) < 1e-9
    assert np.max(np.abs(big.download("errlon"))) < 1e-9
    snl = big.download("snl")
    assert snl.min() >= 0 and snl.max() <= 5
    t = big.download("t_soisno")
    assert np.isfinite(t).all() and t[:, 5:].min() > 150.0 and t.max() < 350.0


def test_full_size_handle_is_periodic_in_its_tiles(cuda_lib, params):
    """BASELINE.json config 5 at its per-GPU size, ragged: 2,097,152 + 77 columns built as eight copies (and a head) of
    one 262,144-column ensemble.  Columns are independent, so after three full steps of the production plan every tile
    of the big handle must hold, bit for bit, what a 262,144-column handle of its own holds - whatever block, SNICAR /
    bare-ground window or queue position a column lands in at the large size.  (The 262,144-column results themselves
    are tied to the oracle by the tests above and, at full size, by bench.py's verification.)"""
    m, tiles, tail = 1 << 18, 8, 77
    n = m * tiles + tail
    cfg = ensemble.EnsembleConfig(ncols=m, seed=123, soil_temp_spread=6.0, h2osfc_fraction=0.1)
    st = ensemble.make_state(cfg, params, cuda_lib.fields)
    F = ensemble.Forcing(m, seed=9, night_fraction=0.3)
    small = cuda_lib.columns(m)
    small.set_tables(params)
    small.upload_state(st)
    big = cuda_lib.columns(n)
    big.set_tables(params)
    for t in range(tiles):
        big.upload_state(st, col0=t * m)
    big.upload_state({k: v[:tail] for k, v in st.items()}, col0=tiles * m)
    for step in range(3):
        f = F.at(step, {k: small.download(k) for k in parity.FORCING_STATE})
        small.upload_state(f)
        for t in range(tiles):
            big.upload_state(f, col0=t * m)
        big.upload_state({k: v[:tail] for k, v in f.items()}, col0=tiles * m)
        for c in (small, big):
            c.init_timestep(True)
            c.step()
    assert small.errors() == (0, -1) and big.errors() == (0, -1)
    for k in cuda_lib.field_names:
        want = small.download(k)
        for t in (0, 3, tiles - 1):
            assert not parity.mismatch(want, big.download(k, col0=t * m, n=m)).any(), (k, t)
        assert not parity.mismatch(want[:tail], big.download(k, col0=tiles * m, n=tail)).any(), (k, "tail")


# ELM's Fortran is another implementation (different compiler, different libm): the reference's own tests accept
# it at 1e-15 .. 1e-10 depending on the variable.  Two comparisons: the CUDA replay against the checker's replay of the
# same records, bit for bit; and against the Fortran values at the tolerances of BASELINE.json.
ELM_RTOL_CLOSED = 1e-12
ELM_RTOL_ITER = 1e-8


def test_elm_fortran_dump_of_test_canhydro(cuda_lib, checker, params):
    """BASELINE.json config 1 on the GPU: the ELM Fortran golden vectors of test_CanHydro, 38 variables x 48 records."""
    import elm_fixture
    worst = elm_fixture.replay(cuda_lib, params)
    got = elm_fixture.replay.last
    elm_fixture.replay(checker, params)
    ref = elm_fixture.replay.last
    diff = [k for k in ref if parity.mismatch(ref[k], got[k]).any()]
    assert not diff, f"CUDA and checker replays differ in {diff}"
    bad = {k: v for k, v in worst.items() if v > ELM_RTOL_CLOSED}
    assert not bad, bad


def test_elm_fortran_dump_of_test_canflux_night_records(cuda_lib, checker, params):
    """ELM Fortran golden vectors of test_CanFlux (night records) on the GPU."""
    import elm_fixture
    n, worst = elm_fixture.replay_canopy_fluxes(cuda_lib, params)
    got = elm_fixture.replay_canopy_fluxes.last
    assert n == 47
    elm_fixture.replay_canopy_fluxes(checker, params)
    ref = elm_fixture.replay_canopy_fluxes.last
    diff = [k for k in ref if parity.mismatch(ref[k], got[k]).any()]
    assert not diff, f"CUDA and checker replays differ in {diff}"
    bad = {k: v for k, v in worst.items() if v > ELM_RTOL_ITER}
    assert not bad, bad


def test_elm_fortran_dump_of_test_surfalb(cuda_lib, checker, params):
    """ELM Fortran golden vectors of test_SurfAlb (95 records, SNICAR without snow layers + two-stream) on the GPU:
    bit-identical to the checker's replay and within 1e-15 of the Fortran values."""
    import elm_fixture
    n, worst = elm_fixture.replay_surface_albedo(cuda_lib, params)
    got = elm_fixture.replay_surface_albedo.last
    assert n == 95
    elm_fixture.replay_surface_albedo(checker, params)
    ref = elm_fixture.replay_surface_albedo.last
    diff = [k for k in ref if parity.mismatch(ref[k], got[k]).any()]
    assert not diff, f"CUDA and checker replays differ in {diff}"
    bad = {k: v for k, v in worst.items() if v > 1e-15}
    assert not bad, bad


def test_elm_fortran_dump_of_test_canflux_day_records(cuda_lib, port_lib, params):
    """The 50 daytime records (photosynthesis active, ELM's own CO2 / O2 partial pressures through
    elmk_set_gas_pressures) on the GPU: bit-identical to the host port, which tests/test_oracle_cpu.py pins to the
    Fortran values, and within the same bound of the Fortran values itself."""
    import elm_fixture
    n, worst = elm_fixture.replay_canopy_fluxes(cuda_lib, params, day=True)
    got = elm_fixture.replay_canopy_fluxes.last
    assert n == 50
    elm_fixture.replay_canopy_fluxes(port_lib, params, day=True)
    ref = elm_fixture.replay_canopy_fluxes.last
    diff = [k for k in ref if parity.mismatch(ref[k], got[k]).any()]
    assert not diff, f"CUDA and port replays differ in {diff}"
    bad = {k: v for k, v in worst.items() if v > 3.5e-12}
    assert not bad, bad


@pytest.mark.parametrize("case", ["capped_snow", "all_bare", "hot_and_wet"])
def test_edge_ensembles(case, cuda_lib, checker, params):
    """Snow capping (do_capsnow), all-bare ensembles and ponded surface water (incl. the 1e97 ground heat flux of
    reference quirk 5), free-running for six steps."""
    import edge_cases
    st = edge_cases.build(case, 2048, params, cuda_lib.fields)
    pair = parity.Pair(checker, cuda_lib, params, ensemble.EnsembleConfig(ncols=2048, seed=99, soil_temp_spread=4.0))
    for c in (pair.a, pair.b):
        c.upload_state(st)
    for step in range(6):
        pair.begin_step()
        pair.run()
    if case == "capped_snow":
        assert pair.b.download("do_capsnow").sum() > 200
    bad = pair.compare()
    assert not bad, f"{case}: bits differ\n{parity.fmt(bad)}"
    assert pair.a.errors() == pair.b.errors()


@pytest.mark.parametrize("n", [1, 2, 33, 129])
def test_tiny_column_counts(n, cuda_lib, checker, params):
    pair = parity.Pair(checker, cuda_lib, params, ensemble.EnsembleConfig(ncols=n, seed=n))
    for _ in range(3):
        pair.begin_step()
        pair.run()
    assert not pair.compare()


def test_reference_throw_becomes_error_bit(cuda_lib, checker, params):
    n = 4096
    st = ensemble.make_state(ensemble.EnsembleConfig(ncols=n, seed=4, snow_fraction=1.0), params, cuda_lib.fields)
    bad_cols = np.array([3, 1000, 4095])
    st["snw_rds"][bad_cols, 4] = 10.0
    out = []
    for lib in (checker, cuda_lib):
        c = lib.columns(n)
        c.set_tables(params)
        c.upload_state(st)
        c.upload_state(ensemble.Forcing(n, seed=3, night_fraction=0.0).at(0, st))
        c.init_timestep(True)
        c.step(groups=abi.G_FRAC_WET | abi.G_ALBEDO)
        out.append((c.errors(), set(np.nonzero(c.download("errmask"))[0])))
    assert out[0] == out[1] == ((1 << 1, 3), set(bad_cols))
    assert cuda_lib.dll.elmk_error_text(1 << 1).decode().startswith("ELM ERROR: SNICAR")
