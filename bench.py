#!/usr/bin/env python3
"""Benchmark of the ELM column timestep on B200 (contract: one JSON line on stdout from rank 0).

Default workload = BASELINE.json config 5, the configuration the metric "column-steps/sec, full ELM step" is quoted
on: the full chain of kokkos_driver's timestep over 2,097,152 synthetic land columns per GPU (16M columns on 8 GPUs:
weak scaling, contiguous column ranges per GPU, no collective on the step), state persistent in HBM across steps.
One step = what ELMInterface::advance does for one dt (driver/kokkos/elm_kokkos_interface.cc:269-322), everything
on the device:
    elmk_solar_step   coszen of every column (its own latitude / longitude), day lengths
    elmk_phenology    LAI / SAI / canopy heights from resident monthly values
    elmk_atm_forcing  the eight forcing functors from resident raw series (forcing refreshed per step on device)
    elmk_init_timestep, elmk_step(all groups a1..a11)

--config 2 | 3 | 4 run the sub-chains of BASELINE.json configs 2-4 through the same entry point (group mask):
    2  frac_wet + albedo/SNICAR + surface radiation       1,048,576 columns per GPU, 25 % night, 1112 B / column-step
    3  canopy hydrology + canopy temperature + bare-ground 4,194,304 columns per GPU, half bare, ponds on a fifth, 1249 B
    4  CanopyFluxes                                        4,194,304 vegetated columns per GPU, 40 % night, all PFTs, 2020 B;
                                                           the line carries the histogram of stability-iteration passes

  value     column-steps/s, device-timed (CUDA events on the handle's stream, max over ranks), inputs resident in HBM
  e2e       the same through the C ABI with HOST buffers inside the timed region: config 5 uploads one new record of
            each of the seven raw forcing series per step from pinned host memory (elmk_atm_series_row) and reads the
            eight per-column balance diagnostics + the error word back every step (overlapped exchange); configs 2-4
            upload the 17 forcing / phenology fields and read back a result field + the error word
  roofline  the dominant launch: algorithmic bytes per launch / its mean CUDA-event duration, against the measured HBM
            copy bandwidth; `step` is the same for the whole step with the frozen byte counts of BASELINE.md section 4
  verify    65,536 columns of this rank's ensemble stepped by oracle/_ref (the reference's own code) beside the GPU,
            outside the timed region: every field compared bit for bit
  cpu_baseline  oracle/_ref with OpenMP on all host cores on a bounded sample of the same workload

`--impl reference` times that CPU path alone, as the reference arm (same --config).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGORITHMIC_BYTES_PER_COLUMN_STEP = 6325   # SURVEY.md section 8(d) / BASELINE.md section 4, element level, frozen (config 5)
DECLARED_BYTES_PER_COLUMN_STEP = 7201      # declared-extent upper bound IN + 2*PROG + OUT (7117 of SURVEY.md + imelt carried across steps + errmask)
HBM_FALLBACK_GBS = 6650.0
CHUNK = 1 << 18                            # columns generated / uploaded at a time
DT = 1800.0
FORC_DT_DAYS = 3.0 / 24.0                  # three-hourly forcing records
RING = 8                                   # time levels of the resident forcing window (e2e refreshes one per step)
DOY0 = 195                                 # the run starts on 14 July

FORCING_FIELDS = ("coszen forc_tbot forc_thbot forc_pbot forc_qbot forc_lwrad forc_u forc_v forc_rain forc_snow "
                  "forc_solad forc_solai elai esai frac_veg_nosno_alb").split()
DIAG_FIELDS = "dtend_column_h2o errh2o errh2osno dwb errsol errlon errseb netrad errmask".split()


def configs():
    from elmkernels_b200 import abi
    return {
        5: dict(name="full ELM column timestep (solar geometry + phenology + forcing functors + init_timestep + groups "
                     "a1..a11, BASELINE.json config 5)", metric="column-steps/sec, full ELM step", mask=abi.G_ALL,
                ncols=1 << 21, bytes=ALGORITHMIC_BYTES_PER_COLUMN_STEP, ens=dict(soil_temp_spread=6.0), night=None, out=DIAG_FIELDS),
        2: dict(name="SurfaceAlbedo + SurfaceRadiation two-stream (groups a1 + a2 + a4, BASELINE.json config 2)",
                metric="column-steps/sec, albedo + surface radiation", mask=abi.G_FRAC_WET | abi.G_ALBEDO | abi.G_SURFACE_RADIATION,
                ncols=1 << 20, bytes=1112, ens=dict(), night=0.25, out=["fsa", "fsr", "errmask"]),
        3: dict(name="CanopyHydrology + CanopyTemperature + BareGroundFluxes (groups a3 + a5 + a6, BASELINE.json config 3)",
                metric="column-steps/sec, hydrology + temperature + bare-ground fluxes",
                mask=abi.G_CANOPY_HYDROLOGY | abi.G_CANOPY_TEMPERATURE | abi.G_BAREGROUND_FLUXES, ncols=1 << 22, bytes=1249,
                ens=dict(bare_fraction=0.5, h2osfc_fraction=0.2), night=None, out=["t_grnd", "eflx_sh_grnd", "errmask"]),
        4: dict(name="CanopyFluxes stability + photosynthesis iteration, mixed PFTs (group a7, BASELINE.json config 4)",
                metric="column-steps/sec, CanopyFluxes", mask=abi.G_CANOPY_FLUXES, ncols=1 << 22, bytes=2020,
                ens=dict(bare_fraction=0.0), night=0.4, out=["t_veg", "eflx_sh_veg", "errmask"]),
    }


def static_json(*names):
    for n in names:
        try:
            return json.load(open(os.path.join(ROOT, "profiles", n))), n
        except Exception:
            pass
    return None, None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---- the synthetic workload: a deterministic function of (rank, chunk), the same on the GPU and on the CPU checker ----
def chunk_seed(rank, chunk):
    return 20240000 + 5 + 7919 * (rank * 64 + chunk)


def make_chunk(P, fields, ncols, seed, cfg):
    from elmkernels_b200 import ensemble
    st = ensemble.make_state(ensemble.EnsembleConfig(ncols=ncols, seed=seed, **cfg["ens"]), P, fields)
    return st, ensemble.Forcing(ncols, seed=seed + 1, night_fraction=cfg["night"])


def chunk_series(ncols, seed, ntimes):
    """Raw forcing series (ntimes, ncols) and monthly phenology (3, ncols) + coordinates of one chunk."""
    r = np.random.default_rng(seed + 2)
    hours = np.arange(ntimes)[:, None] * 3.0
    tb = 270.0 + r.uniform(-12, 12, ncols)[None, :] + 5.0 * np.sin(2 * np.pi * (hours - 9.0) / 24.0)
    atm = {
        "TBOT": tb,
        "PBOT": np.broadcast_to(r.uniform(95000.0, 103000.0, ncols)[None, :], (ntimes, ncols)).copy(),
        "QBOT": r.uniform(40.0, 95.0, (ntimes, ncols)),                      # relative humidity, percent
        "FLDS": r.uniform(180.0, 380.0, (ntimes, ncols)),
        "FSDS": r.uniform(200.0, 800.0, (ntimes, ncols)),                    # scaled by the cosine of the zenith angle on the device
        "PREC": np.where(r.uniform(size=(ntimes, ncols)) < 0.3, r.uniform(0.0, 4e-4, (ntimes, ncols)), 0.0),
        "WIND": r.uniform(0.5, 7.0, (ntimes, ncols)),
    }
    lai = np.where(r.uniform(size=ncols) < 0.2, 0.0, r.uniform(0.3, 4.0, ncols))[None, :] * np.array([0.8, 1.0, 1.2])[:, None]
    phen = {"MLAI": lai, "MSAI": 0.25 * lai + 0.1 * (lai > 0),
            "MHTOP": np.broadcast_to(r.uniform(0.2, 2.2, ncols), (3, ncols)).copy(),
            "MHBOT": np.broadcast_to(r.uniform(0.01, 0.15, ncols), (3, ncols)).copy()}
    lat, lon = np.deg2rad(r.uniform(-65.0, 72.0, ncols)), np.deg2rad(r.uniform(-180.0, 180.0, ncols))
    return atm, phen, lat, lon


class Workload:
    """One handle (GPU or CPU checker) carrying columns [0, n) of this rank's ensemble, stepped through the C ABI."""

    def __init__(self, lib, P, cfg, key, n, rank, device=0, pinned=None, gen_n=None):
        """gen_n: column count of the handle whose first n columns this one mirrors (the generator draws whole chunks:
        a checker that carries a subsample has to draw the same chunk and keep its head)."""
        from elmkernels_b200 import forcing
        self.F, self.cfg, self.key, self.n = forcing, cfg, key, n
        self.cols = lib.columns(n, device=device)
        self.cols.set_tables(P)
        if os.environ.get("ELMK_PLAN"):   # development: "split" = one launch per kernel group (per-group times)
            self.cols.set_plan(os.environ["ELMK_PLAN"])
        self.step_no = 0
        full = key == 5
        ntimes = RING
        atm = {k: np.empty((ntimes, n)) for k in self.cols.ATM_VARS} if full else None
        phen = {k: np.empty((3, n)) for k in self.cols.PHEN_VARS} if full else None
        lat, lon = (np.empty(n), np.empty(n)) if full else (None, None)
        self.forc_host = pinned    # {field: pinned array [n(,nlev)]} filled here when given (configs 2-4 e2e)
        gen_n = gen_n or n
        for c0 in range(0, n, CHUNK):
            m = min(CHUNK, n - c0)
            mg = min(CHUNK, gen_n - c0)      # columns the generator draws for this chunk
            seed = chunk_seed(rank, c0 // CHUNK)
            st, fg = make_chunk(P, lib.fields, mg, seed, cfg)
            f = fg.at(12 + (c0 // CHUNK) % 24, st)
            if m < mg:
                st = {k: np.ascontiguousarray(v[:m]) for k, v in st.items()}
                f = {k: np.ascontiguousarray(v[:m]) for k, v in f.items()}
            self.cols.upload_state(st, col0=c0)
            self.cols.upload_state(f, col0=c0)
            if pinned is not None:
                for k in FORCING_FIELDS:
                    pinned[k][c0:c0 + m] = f[k]
            if full:
                a, p, la, lo = chunk_series(mg, seed, ntimes)
                for k in atm:
                    atm[k][:, c0:c0 + m] = a[k][:, :m]
                for k in phen:
                    phen[k][:, c0:c0 + m] = p[k][:, :m]
                lat[c0:c0 + m], lon[c0:c0 + m] = la[:m], lo[:m]
            del st, f
        self.atm_host = atm   # (RING, n) per variable: the e2e leg re-sends these records one per step
        if full:
            for k, v in atm.items():
                self.cols.atm_series(k, v)
            for k, v in phen.items():
                self.cols.phen_series(k, v)
            self.cols.set_coordinates(lat, lon)
        self.saved = None
        if key == 4:
            # CanopyFluxes starts every step from the same leaf temperature and canopy water (else the second step finds
            # the first one's converged answer and the iteration is over at once): its four read-write scalars are
            # restored before every step, inside the timed region
            self.rw = ("t_veg", "h2ocan", "displa", "z0mv")
        if key in (2, 3, 4):
            self.spin_up()

    def spin_up(self):
        """Two steps of the whole chain so that the sub-chain sees realistic inputs; for configs 3
        and 4 the groups that precede theirs in the chain run once more on the forcing of the timed steps."""
        from elmkernels_b200 import abi
        for _ in range(2):
            self.cols.init_timestep(True)
            self.cols.step()
        self.cols.init_timestep(True)
        before = {2: 0, 3: abi.G_FRAC_WET | abi.G_ALBEDO,
                  4: abi.G_FRAC_WET | abi.G_ALBEDO | abi.G_CANOPY_HYDROLOGY | abi.G_SURFACE_RADIATION |
                  abi.G_CANOPY_TEMPERATURE | abi.G_BAREGROUND_FLUXES}[self.key]
        if before:
            self.cols.step(groups=before)
        if self.key == 4:
            self.saved = {k: self.cols.download(k) for k in self.rw}
            self.restore = None
            if self.cols.lib.backend.startswith("cuda"):
                # device-resident copies, restored with device-to-device copies on the handle's stream
                import torch
                dev = torch.device("cuda", torch.cuda.current_device())
                stream = torch.cuda.ExternalStream(self.cols.stream, device=dev)

                class Raw:   # a field of the handle as a CUDA array
                    def __init__(self, ptr, n):
                        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
                views = {k: torch.as_tensor(Raw(self.cols.device_ptr(k)[0], self.n), device=dev) for k in self.rw}
                keep = {k: v.clone() for k, v in views.items()}

                def restore():
                    with torch.cuda.stream(stream):
                        for k in self.rw:
                            views[k].copy_(keep[k], non_blocking=True)
                self.restore = restore

    def step(self):
        c = self.cols
        if self.key == 5:
            sec = self.step_no * DT
            dayl, max_dayl = c.solar_step(DT, DOY0 + sec / 86400.0 + 1.0, DOY0 + 1 + int(sec // 86400))
            m1 = self.F.first_month_idx(7, 14, sec % 86400.0) - 5     # the series hold June, July, August
            pw1, pw2 = self.F.monthly_data_weights(7, 14, sec % 86400.0)
            c.phenology(m1, pw1, pw2)
            t_idx, w1, w2 = self.F.forcing_time_weights((sec + DT / 2.0) / 86400.0, FORC_DT_DAYS)
            c.atm_forcing(t_idx % (RING - 1), w1, w2, True)           # the resident window is a ring of records
            c.init_timestep(False)
            c.step(dtime=DT, dayl=dayl, max_dayl=max_dayl)
        else:
            if self.key == 4:
                if self.restore:
                    self.restore()
                else:
                    c.upload_state(self.saved)
            if self.key == 3:
                c.init_timestep(True)
            c.step(dtime=DT, groups=self.cfg["mask"])
        self.step_no += 1


def use_all_host_threads(share=1):
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arms are meant to use every host core
    (bound threads, one per core)."""
    n = max(1, (os.cpu_count() or 1) // share)
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def checker_library():
    from elmkernels_b200 import abi
    os.environ.setdefault("ELMREF_SCRUB_STACK", "1")   # deterministic column order in the oracle (oracle/shim/Kokkos_Core.hpp)
    path, kind = os.path.join(ROOT, "oracle", "_ref", "libelmref.so"), "reference"
    if not os.path.exists(path):
        path, kind = os.path.join(ROOT, "oracle", "port", "_build", "libelmport.so"), "port"
    if not os.path.exists(path):
        return None, None
    return abi.Library(path), kind


def time_cpu(P, cfg, key, ncols, steps, warmup):
    """The reference's own code on a bounded sample: `warmup` untimed steps, then `steps` steps timed one by one;
    the value is columns / median step time."""
    cores = use_all_host_threads()
    lib, kind = checker_library()
    if lib is None:
        return None
    w = Workload(lib, P, cfg, key, ncols, rank=0)
    for _ in range(warmup):
        w.step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        w.step()
        times.append(time.perf_counter() - t0)
    w.cols.close()
    med = float(np.median(times))
    return {"value": ncols / med, "unit": "column-steps/s", "cores": cores, "kind": kind,
            "sample": f"{ncols} columns of the same synthetic workload, {warmup} warm-up + {steps} timed steps "
                      f"(median step time), OpenMP bound to all {cores} host cores ({lib.backend})",
            "ms_per_step": 1e3 * med, "total_s": float(np.sum(times))}


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    from elmkernels_b200 import params as prm
    P = prm.load_params()
    b = time_cpu(P, cfg, args.config, args.cpu_cols, args.steps, max(args.warmup, 2))
    if b is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref nor oracle/port is built"}), flush=True)
        return
    line = {"impl": "reference", "metric": cfg["metric"], "value": b["value"], "unit": "column-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 2), "ms_per_step": b["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "columns_per_gpu": args.ncols, "dtime_s": DT},
            "cpu_baseline": {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": b["value"], "unit": "column-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def launch_bytes(group_bytes, mask):
    """Algorithmic HBM bytes per column of one launch covering the groups in `mask` (chain order)."""
    if mask == (1 << 6):
        return 2020   # CanopyFluxes alone: the frozen figure of BASELINE.md section 4
    produced, rd, wr = {}, 0, {}
    for g, name in enumerate(group_bytes["order"]):
        if not (mask >> g) & 1:
            continue
        t = group_bytes["groups"][name]
        for f, b in t["read_elems"].items():
            need = b - produced.get(f, 0)
            if need > 0:
                rd += need
                produced[f] = b   # now on chip
        for f, b in t["write_elems"].items():
            wr[f] = max(wr.get(f, 0), b)
            produced[f] = max(produced.get(f, 0), b)
    return rd + sum(wr.values())


def verify(lib, P, cfg, key, rank, gpu: Workload, nv, steps):
    """Columns [0, nv) of this rank on the CPU checker beside the GPU handle, `steps` steps from the same state: every
    field bit for bit.  Runs before the timed region (the GPU state simply is `steps` steps older afterwards)."""
    chk, kind = checker_library()
    if chk is None:
        return {"skipped": "no checker library"}
    use_all_host_threads()
    ref = Workload(chk, P, cfg, key, nv, rank, gen_n=gpu.n)
    ref.step_no = gpu.step_no
    bad, elements = {}, 0
    for _ in range(steps):
        ref.step()
        gpu.step()
    for k in lib.field_names:
        a, b = ref.cols.download(k), gpu.cols.download(k, col0=0, n=nv)
        elements += a.size
        if a.dtype.kind == "f":
            m = (a.view(np.uint64) != b.view(np.uint64)) & ~(np.isnan(a) & np.isnan(b))
        else:
            m = a != b
        if m.any():
            bad[k] = int(m.sum())
    ea, eb = ref.cols.errors(), gpu.cols.errors()
    ref.cols.close()
    return {"columns": nv, "steps": steps, "fields": len(lib.field_names), "elements_compared": elements,
            "mismatching_elements": int(sum(bad.values())), "mismatching_fields": bad, "bit_identical": not bad,
            "checker": chk.backend, "kind": kind, "error_words_equal": ea[0] == eb[0]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[2, 3, 4, 5], help="BASELINE.json config (default 5: the full step)")
    ap.add_argument("--ncols", type=int, default=0, help="columns per GPU (default: the size SURVEY.md 8(d) gives the config)")
    ap.add_argument("--cpu-cols", type=int, default=1 << 17, help="columns of the CPU-baseline sample")
    ap.add_argument("--verify-cols", type=int, default=1 << 16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    cfg = configs()[args.config]
    args.ncols = args.ncols or cfg["ncols"]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    import elmkernels_b200
    from elmkernels_b200 import params as prm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = elmkernels_b200.load()
    P = prm.load_params()
    n, key = args.ncols, args.config

    def pinned(name, rows=None):
        _, dt, nl = lib.fields[name]
        shape = (n,) if nl == 1 else (n, nl)
        return torch.empty(shape, dtype={0: torch.float64, 1: torch.int32, 2: torch.uint8}[dt], pin_memory=True).numpy()

    forc_host = {k: pinned(k) for k in FORCING_FIELDS} if key != 5 else None
    W = Workload(lib, P, cfg, key, n, rank, device=local, pinned=forc_host)
    cols = W.cols
    cols.sync()
    stream = torch.cuda.ExternalStream(cols.stream, device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity beside the measurement: a subsample of this rank's columns on the reference's own code ----
    check = None
    if not args.no_verify:
        check = verify(lib, P, cfg, key, rank, W, min(args.verify_cols, n), steps=2)

    # ---- warm-up ----
    for _ in range(args.warmup):
        W.step()
    cols.sync()
    any_err, first = cols.errors()
    if any_err:
        raise SystemExit(f"column error bits {any_err:#x} (first column {first}) during warm-up")

    # ---- timed region: K steps, inputs resident in HBM ----
    sampler = ClockSampler(local)
    sampler.start()
    cols.timing(2 if os.environ.get("ELMK_TIMING_DETAIL") else True)
    l0 = cols.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        W.step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = cols.launch_count - l0
    kern = cols.timing_read()
    cols.timing(False)
    hist = cols.canflux_pass_histogram() if key in (4, 5) and not os.environ.get("ELMK_PLAN") else None

    # ---- end to end: host buffers in, per-column results out, every step, through the C ABI ----
    out_fields = cfg["out"]
    res = [[pinned(k) for k in out_fields] for _ in range(2)]
    d2h = sum(a.nbytes for a in res[0])
    if key == 5:
        # one new record of every raw series per step: pinned rows -> the ring slot that the forcing functors read
        # two records later (elmk_atm_series_row copies on its own stream while the step computes)
        # (the host keeps the window of records in pinned memory and re-sends the record of the slot it refreshes, so
        #  that the forcing - hence the work per step - has the same statistics as in the device-timed region)
        rows = {v: torch.from_numpy(W.atm_host[v]).pin_memory().numpy() for v in cols.ATM_VARS}
        h2d = sum(a[0].nbytes for a in rows.values())
        xch = cols.exchange([], out_fields)
    else:
        h2d = sum(a.nbytes for a in forc_host.values())
        xch = cols.exchange(FORCING_FIELDS, out_fields)
        fin = [forc_host[k] for k in FORCING_FIELDS]
    host_checks = []

    def e2e_steps(k_steps):
        if key != 5:
            xch.post(fin)
        for k in range(k_steps):
            if key == 5:
                t_next = (W.F.forcing_time_index((W.step_no * DT + DT / 2.0) / 86400.0, FORC_DT_DAYS) + 3) % (RING - 1)
                for v in cols.ATM_VARS:
                    cols.atm_series_row(v, t_next, rows[v][t_next])
            else:
                xch.commit()
                if k + 1 < k_steps:
                    xch.post_wait()                             # the previous post has left the host buffers:
                    fin[1][:8] += 1.0e-9 * ((k & 1) * 2 - 1)  # the host may write the next step's forcing into them
                    xch.post(fin)
            W.step()
            xch.fetch(res[k & 1])
            if k >= 1:
                xch.wait()                                   # step k-1's results are on the host
                host_checks.append(int(res[(k - 1) & 1][-1].max()))
        xch.wait()
        host_checks.append(int(res[(k_steps - 1) & 1][-1].max()))

    e2e_steps(2)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    g0.record(stream)
    t_host0 = time.perf_counter()
    e2e_steps(args.steps)
    cols.sync()
    t_host1 = time.perf_counter()
    g1.record(stream)
    barrier()
    # the device->host copies finish on the exchange's copy stream, after the last event on the step stream:
    # the host clock around the same region (which ends with every result on the host) is the e2e time
    ms_e2e = max(g0.elapsed_time(g1), 1e3 * (t_host1 - t_host0))
    clocks = sampler.stop()
    any_err, first = cols.errors()

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=f"cuda:{local}")
    ok = torch.tensor([1.0 if (check is None or check.get("bit_identical", True)) else 0.0], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if key == 5:
            # optional global balance diagnostic: sum/min/max of the eight a11 fields over all GPUs (NCCL),
            # outside the timed region - the step itself has no collective
            from elmkernels_b200.sharding import reduce_diagnostics
            reduce_diagnostics(cols.diag_reduce(), device=f"cuda:{local}")
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = measured_peak()
        gb = json.load(open(os.path.join(ROOT, "elmkernels_b200", "data", "group_bytes.json")))
        other = {k[0]: k[2] / k[3] for k in kern if not k[1]}   # launches outside elmk_step (and sub-launches with ELMK_TIMING_DETAIL)
        kern = [k for k in kern if k[1]]
        total_kernel_ms = sum(k[2] for k in kern) or 1.0
        top = max(kern, key=lambda k: k[2])
        per_launch_ms = top[2] / top[3]
        bytes_per_col = launch_bytes(gb, top[1])
        achieved = bytes_per_col * n / (per_launch_ms * 1e-3) / 1e9
        value = n * world * args.steps / (ms * 1e-3)
        e2e = n * world * args.steps / (ms_e2e * 1e-3)
        counters, csrc = static_json("r2_kernel_counters.json", "r2a_kernel_counters.json", "r1_kernel_counters.json")
        fpk, _ = static_json("r1_fp64_peak.json")
        cg = (counters or {}).get("groups", {}) if key == 5 else {}
        traffic = cg[top[0]]["dram_bytes_per_column"] * n if top[0] in cg else None
        fp64 = None
        if top[0] in cg and fpk:
            # FP64 pipe: thread-level DADD+DMUL+DFMA instructions per launch / live duration, against the measured
            # instruction rate of the pipe (a DFMA counts as one instruction)
            inst = cg[top[0]]["fp64_inst_per_column"] * n
            peak_inst = fpk["dfma_tflops"] / 2.0
            fp64 = {"achieved": inst / (per_launch_ms * 1e-3) / 1e12, "peak": peak_inst, "unit": "T FP64 inst/s",
                    "frac": inst / (per_launch_ms * 1e-3) / 1e12 / peak_inst,
                    "inst_per_column": cg[top[0]]["fp64_inst_per_column"],
                    "peak_source": "measured DFMA microbenchmark (profiles/r1_fp64_peak.json)",
                    "counts_source": "ncu, profiles/%s (%d columns)" % (csrc, counters["ncols"])}
        step_gbs = cfg["bytes"] * n * args.steps / (ms * 1e-3) / 1e9
        line = {
            "metric": cfg["metric"], "value": value, "unit": "column-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"] + ", synthetic mixed-PFT ensemble with snl 0..5, state persistent in HBM, "
                                   "inputs of every step produced on the device" if key == 5 else cfg["name"] + ", synthetic ensemble",
                       "baseline_config": key, "columns_per_gpu": n, "columns_total": n * world, "dtime_s": DT,
                       "parallelism": f"columns sharded over {world} GPU(s), no collective on the step",
                       "l2": "inputs larger than L2: %.1f GB of column state per GPU vs 126 MB" % (n * 5765 / 1e9)},
            "roofline": {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "fp64": fp64,
                         "note": "launches of this path are bounded by FP64 dependent-issue latency and instruction fetch, "
                                 "not by HBM: DESIGN.md section 4",
                         "bytes_per_column": bytes_per_col, "ms_per_launch": per_launch_ms,
                         "share_of_step": top[2] / total_kernel_ms,
                         "step": {"bytes_per_column": cfg["bytes"], "achieved": step_gbs, "frac": step_gbs / peak},
                         "kernels": {k[0]: {"ms_per_launch": k[2] / k[3], "share": k[2] / total_kernel_ms,
                                            "bytes_per_column": launch_bytes(gb, k[1]),
                                            "GBps": launch_bytes(gb, k[1]) * n / (k[2] / k[3] * 1e-3) / 1e9,
                                            "traffic_bytes_per_column": cg.get(k[0], {}).get("dram_bytes_per_column"),
                                            "fp64_inst_per_column": cg.get(k[0], {}).get("fp64_inst_per_column")}
                                     for k in kern}},
            "e2e": {"value": e2e, "unit": "column-steps/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e / args.steps,
                    "what": ("one new record of the seven raw forcing series in, eight balance diagnostics + error word out"
                             if key == 5 else "17 forcing / phenology fields in, %s out" % " + ".join(out_fields))},
            "other_launches_ms": other, "gpu_launches": launches, "clocks": clocks,
            "errors": {"any": any_err, "first_column": first},
            "verify": dict(check, all_ranks_bit_identical=bool(float(ok[0]) > 0.5)) if check else None,
        }
        if hist is not None:
            nz = np.nonzero(hist)[0]
            line["canflux_passes"] = {"histogram": {int(k): int(hist[k]) for k in nz},
                                      "mean": float((hist * np.arange(42)).sum() / max(hist.sum(), 1)),
                                      "columns": int(hist.sum())}
        if world == 1 and not args.no_cpu_baseline:
            b = time_cpu(P, cfg, key, args.cpu_cols, steps=5, warmup=2)
            line["cpu_baseline"] = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")} if b else None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
