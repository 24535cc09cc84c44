#!/usr/bin/env python3
"""Benchmark of the ELM column timestep on B200 (contract: one JSON line on stdout from rank 0).

Workload (BASELINE.json config 5, the configuration the metric "column-steps/sec, full ELM step" is quoted
on): the full kernel chain of kokkos_driver's timestep - init_timestep bookkeeping + groups a1..a11 - over
2,097,152 synthetic land columns per GPU (16M columns on 8 GPUs: weak scaling, contiguous column ranges
per GPU, no collective on the step), state persistent in HBM across steps.

  value   column-steps/s with the forcing already resident in HBM (device-timed, max over ranks)
  e2e     the same through the C ABI with HOST buffers: every step uploads that step's atmospheric forcing
          and phenology (17 fields) from pinned host memory and reads back the eight per-column balance
          diagnostics + the error word
  roofline  the dominant kernel of the step: algorithmic bytes per launch (tools/group_bytes.py, element
          level) / its mean CUDA-event duration, against the measured HBM copy bandwidth
  cpu_baseline  the reference's own code (oracle/_ref, OpenMP on all host cores) on a bounded sample

`--impl reference` times that CPU path alone, as the reference arm.
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

ALGORITHMIC_BYTES_PER_COLUMN_STEP = 6325   # SURVEY.md section 8(d) / BASELINE.md section 4, element level, frozen
DECLARED_BYTES_PER_COLUMN_STEP = 7201      # declared-extent upper bound IN + 2*PROG + OUT (7117 of SURVEY.md + imelt carried across steps + errmask)
HBM_FALLBACK_GBS = 6650.0
METRIC = "column-steps/sec, full ELM step"
CHUNK = 1 << 18                            # columns generated / uploaded at a time

FORCING_FIELDS = ("coszen forc_tbot forc_thbot forc_pbot forc_qbot forc_lwrad forc_u forc_v forc_rain forc_snow "
                  "forc_solad forc_solai elai esai frac_veg_nosno_alb").split()
RESULT_FIELDS = "dtend_column_h2o errh2o errh2osno dwb errsol errlon errseb netrad errmask".split()


def kernel_counters():
    """ncu counters of one step of this workload (profiles/r1_kernel_counters.json, written by tools/ncu_counters.py
    from a capture of `bench.py --ncols 524288 --steps 1`): DRAM bytes and FP64 instructions per column and launch
    group.  Static evidence committed with the repo - nothing is profiled while the benchmark runs."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r1_kernel_counters.json")))
    except Exception:
        return None


def fp64_peak():
    """Measured FP64 pipe peak of this pool's B200 (tools/fp64_peak.cu -> profiles/r1_fp64_peak.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r1_fp64_peak.json")))
    except Exception:
        return None


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


def make_chunk(params, fields, ncols, seed):
    from elmkernels_b200 import ensemble
    cfg = ensemble.EnsembleConfig(ncols=ncols, seed=seed, soil_temp_spread=6.0)
    st = ensemble.make_state(cfg, params, fields)
    forcing = ensemble.Forcing(ncols, seed=seed + 1)
    return st, forcing


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arms are meant to use every host core."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref), all host threads."""
    if rank != 0:
        return
    use_all_host_threads()
    from elmkernels_b200 import abi, params as prm
    path = os.path.join(ROOT, "oracle", "_ref", "libelmref.so")
    kind = "reference"
    if not os.path.exists(path):
        path, kind = os.path.join(ROOT, "oracle", "port", "_build", "libelmport.so"), "port"
    lib = abi.Library(path)
    P = prm.load_params()
    n = args.cpu_cols
    st, forcing = make_chunk(P, lib.fields, n, 20240000 + 5)
    cols = lib.columns(n)
    cols.set_tables(P)
    cols.upload_state(st)
    f = forcing.at(12, st)
    cols.upload_state(f)
    for _ in range(args.warmup):
        cols.init_timestep(True)
        cols.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cols.init_timestep(True)
        cols.step()
    dt = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    value = n * args.steps / dt
    sample = f"{n} columns x {args.steps} steps (of the {args.ncols}-column/GPU workload), {args.warmup} warm-up steps"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "column-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "full ELM column timestep (init_timestep + groups a1..a11), synthetic mixed-PFT "
                                   "ensemble with snl 0..5", "columns_per_gpu": args.ncols, "dtime_s": 1800,
                       "backend": lib.backend},
            "cpu_baseline": {"value": value, "unit": "column-steps/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "column-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(P, ncols, steps=3, warmup=1):
    from elmkernels_b200 import abi
    path = os.path.join(ROOT, "oracle", "_ref", "libelmref.so")
    kind = "reference"
    if not os.path.exists(path):
        path, kind = os.path.join(ROOT, "oracle", "port", "_build", "libelmport.so"), "port"
    if not os.path.exists(path):
        return None
    use_all_host_threads()
    lib = abi.Library(path)
    st, forcing = make_chunk(P, lib.fields, ncols, 20240000 + 5)
    cols = lib.columns(ncols)
    cols.set_tables(P)
    cols.upload_state(st)
    cols.upload_state(forcing.at(12, st))
    for _ in range(warmup):
        cols.init_timestep(True)
        cols.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        cols.init_timestep(True)
        cols.step()
    dt = time.perf_counter() - t0
    cols.close()
    return {"value": ncols * steps / dt, "unit": "column-steps/s", "cores": os.cpu_count() or 1, "kind": kind,
            "sample": f"{ncols} columns x {steps} steps of the same synthetic ensemble, {warmup} warm-up, "
                      f"OpenMP over all host cores ({lib.backend})"}


def launch_bytes(group_bytes, mask):
    """Algorithmic HBM bytes per column of one launch covering the groups in `mask` (chain order)."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncols", type=int, default=1 << 21, help="columns per GPU")
    ap.add_argument("--cpu-cols", type=int, default=1 << 17, help="columns of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import elmkernels_b200
    from elmkernels_b200 import abi, params as prm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = elmkernels_b200.load()
    P = prm.load_params()
    n = args.ncols
    cols = lib.columns(n, device=local)
    cols.set_tables(P)

    # ---- synthetic ensemble: unique columns, generated and uploaded chunk by chunk ----
    forc_host = {}
    for k in FORCING_FIELDS:
        _, dt, nl = lib.fields[k]
        t = torch.empty((n,) if nl == 1 else (n, nl), dtype={0: torch.float64, 1: torch.int32, 2: torch.uint8}[dt],
                        pin_memory=True)
        forc_host[k] = t.numpy()
    for c0 in range(0, n, CHUNK):
        m = min(CHUNK, n - c0)
        st, forcing = make_chunk(P, lib.fields, m, 20240000 + 5 + 7919 * (rank * 64 + c0 // CHUNK))
        cols.upload_state(st, col0=c0)
        f = forcing.at(12 + (c0 // CHUNK) % 24, st)
        for k in FORCING_FIELDS:
            forc_host[k][c0:c0 + m] = f[k]
        del st, f
    up_plan = cols.plan(FORCING_FIELDS, [forc_host[k] for k in FORCING_FIELDS])
    res_host = {}
    for k in RESULT_FIELDS:
        _, dt, nl = lib.fields[k]
        res_host[k] = torch.empty((n,), dtype={0: torch.float64, 1: torch.int32}[dt], pin_memory=True).numpy()
    down_plan = cols.plan(RESULT_FIELDS, [res_host[k] for k in RESULT_FIELDS])
    h2d = sum(a.nbytes for a in forc_host.values())
    d2h = sum(a.nbytes for a in res_host.values())
    cols.upload_many(up_plan)
    cols.sync()

    stream = torch.cuda.ExternalStream(cols.stream, device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        cols.init_timestep(True)
        cols.step()

    # ---- warm-up ----
    for _ in range(args.warmup):
        one_step()
    cols.sync()
    any_err, first = cols.errors()
    if any_err:
        raise SystemExit(f"column error bits {any_err:#x} (first column {first}) during warm-up")

    # ---- timed region: K steps, forcing resident in HBM ----
    sampler = ClockSampler(local)
    sampler.start()
    cols.timing(2 if os.environ.get("ELMK_TIMING_DETAIL") else True)
    l0 = cols.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        one_step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = cols.launch_count - l0
    kern = cols.timing_read()
    cols.timing(False)

    # ---- end to end: host forcing in, per-column diagnostics out, every step ----
    # Through the overlapped exchange of the C ABI (elmk_exchange_*): every step's forcing is copied from pinned
    # host memory and every step's diagnostics are copied back to pinned host memory inside the timed region; the
    # copies of step k+1's inputs and of step k's results run on two copy streams while step k+1 computes.  The
    # host reads step k's result (the error word) before it issues step k+2.
    xch = cols.exchange(FORCING_FIELDS, RESULT_FIELDS)
    fin = [forc_host[k] for k in FORCING_FIELDS]
    fout = [[res_host[k] for k in RESULT_FIELDS],
            [torch.empty_like(torch.from_numpy(res_host[k]), pin_memory=True).numpy() for k in RESULT_FIELDS]]
    host_checks = []

    def e2e_steps(k_steps):
        xch.post(fin)
        for k in range(k_steps):
            xch.commit()
            if k + 1 < k_steps:
                xch.post(fin)
            one_step()
            xch.fetch(fout[k & 1])
            if k >= 1:
                xch.wait()                                   # step k-1's diagnostics are on the host
                host_checks.append(int(fout[(k - 1) & 1][-1].max()))
        xch.wait()
        host_checks.append(int(fout[(k_steps - 1) & 1][-1].max()))

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
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # optional global balance diagnostic: sum/min/max of the eight a11 fields over all GPUs (NCCL),
        # outside the timed region - the step itself has no collective
        from elmkernels_b200.sharding import reduce_diagnostics
        global_diag = reduce_diagnostics(cols.diag_reduce(), device=f"cuda:{local}")
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = measured_peak()
        gb = json.load(open(os.path.join(ROOT, "elmkernels_b200", "data", "group_bytes.json")))
        other = {k[0]: k[2] / k[3] for k in kern if not k[1]}   # init_timestep (and, with ELMK_TIMING_DETAIL, sub-launches)
        kern = [k for k in kern if k[1]]   # launches of elmk_step (init_timestep has mask 0)
        total_kernel_ms = sum(k[2] for k in kern) or 1.0
        top = max(kern, key=lambda k: k[2])
        per_launch_ms = top[2] / top[3]
        bytes_per_col = launch_bytes(gb, top[1])
        achieved = bytes_per_col * n / (per_launch_ms * 1e-3) / 1e9
        value = n * world * args.steps / (ms * 1e-3)
        e2e = n * world * args.steps / (ms_e2e * 1e-3)
        counters, fpk = kernel_counters(), fp64_peak()
        cg = (counters or {}).get("groups", {})
        traffic = cg[top[0]]["dram_bytes_per_column"] * n if top[0] in cg else None
        fp64 = None
        if top[0] in cg and fpk:
            # FP64 pipe: thread-level DADD+DMUL+DFMA instructions per launch / live duration, against the measured
            # instruction rate of the pipe (a DFMA counts as one instruction; the physics is built without FMA
            # contraction, so its flop rate is bounded by the DMUL/DADD figure)
            inst = cg[top[0]]["fp64_inst_per_column"] * n
            peak_inst = fpk["dfma_tflops"] / 2.0
            fp64 = {"achieved": inst / (per_launch_ms * 1e-3) / 1e12, "peak": peak_inst, "unit": "T FP64 inst/s",
                    "frac": inst / (per_launch_ms * 1e-3) / 1e12 / peak_inst,
                    "flop_per_column": cg[top[0]]["flop_per_column"], "inst_per_column": cg[top[0]]["fp64_inst_per_column"],
                    "peak_source": "measured DFMA microbenchmark (profiles/r1_fp64_peak.json)",
                    "counts_source": "ncu, profiles/r1_kernel_counters.json (%d columns)" % counters["ncols"]}
        line = {
            "metric": METRIC, "value": value, "unit": "column-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "full ELM column timestep (init_timestep + groups a1..a11, BASELINE.json config 5), "
                                   "synthetic mixed-PFT ensemble with snl 0..5, state persistent in HBM",
                       "columns_per_gpu": n, "columns_total": n * world, "dtime_s": 1800,
                       "parallelism": f"columns sharded over {world} GPU(s), no collective on the step",
                       "l2": "inputs larger than L2: %.1f GB of column state per GPU vs 126 MB" % (n * 5765 / 1e9)},
            "roofline": {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "fp64": fp64,
                         "note": "the dominant launch is bounded by FP64 latency / issue, not by HBM: see fp64 and DESIGN.md section 4",
                         "bytes_per_column": bytes_per_col, "ms_per_launch": per_launch_ms,
                         "share_of_step": top[2] / total_kernel_ms,
                         "step": {"bytes_per_column": ALGORITHMIC_BYTES_PER_COLUMN_STEP,
                                  "achieved": ALGORITHMIC_BYTES_PER_COLUMN_STEP * n * args.steps / (ms * 1e-3) / 1e9,
                                  "frac": ALGORITHMIC_BYTES_PER_COLUMN_STEP * n * args.steps / (ms * 1e-3) / 1e9 / peak},
                         "kernels": {k[0]: {"ms_per_launch": k[2] / k[3], "share": k[2] / total_kernel_ms,
                                            "bytes_per_column": launch_bytes(gb, k[1]),
                                            "GBps": launch_bytes(gb, k[1]) * n / (k[2] / k[3] * 1e-3) / 1e9,
                                            "traffic_bytes_per_column": cg.get(k[0], {}).get("dram_bytes_per_column"),
                                            "fp64_inst_per_column": cg.get(k[0], {}).get("fp64_inst_per_column")}
                                     for k in kern}},
            "e2e": {"value": e2e, "unit": "column-steps/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e / args.steps},
            "other_launches_ms": other, "gpu_launches": launches, "clocks": clocks, "errors": {"any": any_err, "first_column": first},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(P, args.cpu_cols)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
