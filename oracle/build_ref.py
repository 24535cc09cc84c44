#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle).  Build oracle/_ref/libelmref.so from the reference sources.

The reference's own files are compiled where they lie under /root/reference (nothing is copied):
its ten kernel-group wrappers driver/kokkos/*_kokkos.cc with -DENABLE_KOKKOS against the
host Kokkos stand-in in oracle/shim, four of its utility .cc files, and oracle/ref_capi.cc
(the C ABI of include/elmk_b200.h on top of the reference's ELMState).  Flags follow
BASELINE.md section 3: g++ -std=c++17 -O2 -fopenmp, no -march, no -ffast-math (so no FMA contraction);
-DNDEBUG turns the reference's two in-kernel assert()s into no-ops instead of process aborts.

Also builds the reference's seven unit tests (test/test_*.cc) into oracle/_ref/ so that the
golden vectors in test/data can be replayed through the reference itself (oracle/run_ref_tests.py).

Outputs go to oracle/_ref/ only (git-ignored, shipped to the GPU box by gpurun).
"""
import os, pathlib, subprocess, sys, concurrent.futures as cf

HERE = pathlib.Path(__file__).resolve().parent
REF = pathlib.Path(os.environ.get("ELMK_REFERENCE", "/root/reference"))
OUT = HERE / "_ref"
WRAPPERS = ["albedo", "bareground_fluxes", "canopy_fluxes", "canopy_hydrology", "canopy_temperature",
            "conserved_quantity", "snow_hydrology", "soil_temperature", "surface_fluxes", "surface_radiation"]
EXTRA = ["src/utils/utils.cc", "src/utils/read_input.cc", "src/data/monthly_data.cc",
         "src/physics/day_length.cc", "src/physics/incident_shortwave.cc"]
TESTS = ["CanHydro", "CanSunShade", "SurfRad", "CanTemp", "BGFlux", "SurfAlb", "CanFlux"]
CXX = os.environ.get("ELMK_CXX", "/usr/bin/g++")
BASE = ["-std=c++17", "-O2", "-fPIC", "-w"]
INC = ["-I" + str(HERE / "shim"), "-I" + str(REF / "driver/kokkos"), "-I" + str(REF / "src/physics"),
       "-I" + str(REF / "src/data"), "-I" + str(REF / "src/utils")]


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(map(str, cmd)) + "\n" + r.stdout + r.stderr)
        raise SystemExit("oracle/_ref build failed")


def newer(target, sources):
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(pathlib.Path(s).stat().st_mtime <= t for s in sources)


def build_lib(force=False):
    lib = OUT / "libelmref.so"
    srcs = [REF / f"driver/kokkos/{w}_kokkos.cc" for w in WRAPPERS] + [REF / e for e in EXTRA] + [HERE / "ref_capi.cc"]
    deps = srcs + [HERE / "shim/Kokkos_Core.hpp", HERE / "shim/netcdf.h", HERE / "exchange_host.h", HERE.parent / "include/elmk_b200.h",
                   HERE.parent / "include/elmk_fields.def"]
    if not force and newer(lib, deps):
        return lib
    OUT.mkdir(exist_ok=True)
    flags = BASE + ["-fopenmp", "-DENABLE_KOKKOS", "-DNDEBUG", '-DINPUT_DATA_DIR="/nonexistent/"'] + INC
    objs = []
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        futs = []
        for s in srcs:
            o = OUT / (pathlib.Path(s).stem + ".o")
            objs.append(o)
            futs.append(ex.submit(run, [CXX] + flags + ["-c", str(s), "-o", str(o)]))
        for f in futs:
            f.result()
    run([CXX, "-shared", "-fopenmp", "-o", str(lib)] + [str(o) for o in objs])
    for o in objs:
        o.unlink()
    return lib


def build_tests(force=False):
    """The reference's own unit tests, serial ELM::Array backend (no Kokkos), as test/CMakeLists.txt:5-31."""
    shim = HERE / "shim_serial"
    flags = BASE + ['-DTEST_DATA_DIR="%s/"' % (REF / "test/data"), "-I" + str(shim), "-I" + str(HERE / "shim"),
                    "-I" + str(REF / "src/physics"), "-I" + str(REF / "src/data"), "-I" + str(REF / "src/utils")]
    OUT.mkdir(exist_ok=True)
    jobs = []
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for t in TESTS:
            exe = OUT / f"test_{t}"
            src = REF / f"test/test_{t}.cc"
            if not force and newer(exe, [src, shim / "zero_new.cc", shim / "netcdf.h"]):
                continue
            extra = [str(REF / "src/utils/read_test_input.cc")]
            opt = []
            if t in ("SurfAlb", "CanFlux"):
                # these two read the PFT constants of clm_params_c180524.nc (served as text dumps by shim_serial/netcdf.h)
                # and scratch arrays of ELM::Array that nothing initialises (shim_serial/zero_new.cc)
                extra += [str(REF / "src/utils/read_input.cc"), str(REF / "src/utils/utils.cc"), str(shim / "zero_new.cc")]
                opt = ['-DELMK_NC_DUMP_DEFAULT="%s"' % (OUT / "clm_params")]
                if not (OUT / "clm_params" / "pftname.txt").exists():
                    import dump_params
                    dump_params.dump()
            jobs.append(ex.submit(run, [CXX] + flags + opt + [str(src)] + extra + ["-o", str(exe)]))
        for j in jobs:
            j.result()


if __name__ == "__main__":
    if not REF.exists():
        raise SystemExit(f"{REF} not present: oracle/_ref can only be (re)built where the reference is mounted")
    force = "--force" in sys.argv
    print(build_lib(force))
    if "--tests" in sys.argv:
        build_tests(force)
        print("reference unit tests built")
