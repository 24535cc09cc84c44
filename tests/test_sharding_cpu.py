"""The N>1 path on CPU: two gloo ranks, each owning a contiguous column range (host port as the backend),
must reproduce the single-handle run column for column, and the all-reduced balance diagnostic must equal the
single-handle reduction."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, STEPS = 1000, 3   # deliberately not divisible into equal halves of 128


def test_shard_ranges_cover_without_overlap():
    from elmkernels_b200.sharding import shard_range
    for total in (0, 1, 7, 1000, 16 * 2**20):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _full_run(lib_path):
    sys.path.insert(0, ROOT)
    from elmkernels_b200 import abi, ensemble, params
    lib = abi.Library(lib_path)
    P = params.load_params()
    st = ensemble.make_state(ensemble.EnsembleConfig(ncols=N, seed=31, soil_temp_spread=5.0), P, lib.fields)
    F = ensemble.Forcing(N, seed=32)
    forc = []
    cols = lib.columns(N)
    cols.set_tables(P)
    cols.upload_state(st)
    for s in range(STEPS):
        f = F.at(s, {k: cols.download(k) for k in ("vtype", "snow_depth", "frac_sno", "htop", "hbot", "tlai", "tsai")})
        forc.append(f)
        cols.upload_state(f)
        cols.init_timestep(True)
        cols.step()
    return lib, P, st, forc, cols


def _worker(rank, world, port, lib_path, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from elmkernels_b200 import abi, params
    from elmkernels_b200.sharding import reduce_diagnostics, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(out_dir, "inputs.npz"))
    lo, hi = shard_range(N, rank, world)
    lib = abi.Library(lib_path)
    cols = lib.columns(hi - lo)
    cols.set_tables(params.load_params())
    cols.upload_state({k[3:]: z[k][lo:hi] for k in z.files if k.startswith("s0_")})
    for s in range(STEPS):
        cols.upload_state({k[len(f"f{s}_"):]: z[k][lo:hi] for k in z.files if k.startswith(f"f{s}_")})
        cols.init_timestep(True)
        cols.step()                      # no communication on the step
    red = reduce_diagnostics(cols.diag_reduce())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, red=red,
             **{k: cols.download(k) for k in ("t_grnd", "t_soisno", "snl", "errseb", "h2osno", "eflx_lh_tot")})
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_single_handle(port_lib, tmp_path):
    lib, P, st, forc, cols = _full_run(port_lib.path)
    inputs = {"s0_" + k: v for k, v in st.items()}
    for s, f in enumerate(forc):
        inputs.update({f"f{s}_{k}": v for k, v in f.items()})
    np.savez(tmp_path / "inputs.npz", **inputs)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, port_lib.path, str(tmp_path)), nprocs=2, join=True)
    full_red = cols.diag_reduce()
    for rank in range(2):
        z = np.load(tmp_path / f"rank{rank}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        for k in ("t_grnd", "t_soisno", "snl", "errseb", "h2osno", "eflx_lh_tot"):
            assert np.array_equal(z[k], cols.download(k, col0=lo, n=hi - lo), equal_nan=True), (rank, k)
        red = z["red"]
        assert np.array_equal(red[8:], full_red[8:])                       # min / max exact
        assert np.allclose(red[:8], full_red[:8], rtol=1e-12, atol=1e-9)   # sums up to association
