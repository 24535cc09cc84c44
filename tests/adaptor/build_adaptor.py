#!/usr/bin/env python3
"""Build tests/adaptor/_build/adaptor_steps: the C++ host adaptor on the reference's own ELMState, linked against the
product library.  Needs the reference headers (this container); the binary travels to the GPU box with the snapshot
(_build/ is git-ignored, not gpurun-ignored).  Called by __graft_entry__.build()."""
import os, pathlib, subprocess, sys
HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = pathlib.Path(os.environ.get("ELMK_REFERENCE", "/root/reference"))


def build(force=False):
    exe = HERE / "_build" / "adaptor_steps"
    srcs = [HERE / "adaptor_steps.cc", ROOT / "include/elm_b200.hh", ROOT / "include/elmk_b200.h", ROOT / "include/elmk_members.h"]
    if not force and exe.exists() and all(s.stat().st_mtime <= exe.stat().st_mtime for s in srcs):
        return exe
    if not REF.is_dir():
        return None
    exe.parent.mkdir(exist_ok=True)
    libdir = ROOT / "elmkernels_b200"
    cmd = ["g++", "-std=c++17", "-O1", "-w", "-fopenmp", "-DENABLE_KOKKOS", '-DINPUT_DATA_DIR="/nonexistent/"',
           f"-I{ROOT}/oracle/shim", f"-I{ROOT}/include", f"-I{REF}/driver/kokkos", f"-I{REF}/src/physics",
           f"-I{REF}/src/data", f"-I{REF}/src/utils", str(HERE / "adaptor_steps.cc"), f"{REF}/src/utils/utils.cc",
           f"{REF}/src/utils/read_input.cc", f"-L{libdir}", "-lelmk_b200", "-Wl,-rpath,$ORIGIN/../../../elmkernels_b200",
           "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-6000:])
        raise SystemExit("adaptor_steps build failed")
    return exe


# The reference's own unit tests, compiled UNCHANGED from where they lie against the library-level API of include/elm/
# (which takes the place of the reference's src/physics on the include path) and linked against the product library:
# every ELM::<namespace>::<function> call of the test then runs on the GPU.  The ELM Fortran fixture files the tests
# read at run time are copied next to the binaries (build output: git-ignored, travels with the snapshot).
REF_TESTS = {"CanHydro": ["CanopyHydrology_IN.txt", "CanopyHydrology_OUT.txt"],
             "SurfRad": ["SurfaceRadiation_IN.txt", "SurfaceRadiation_OUT.txt"],
             "CanSunShade": ["CanopySunShadeFractions_IN.txt", "CanopySunShadeFractions_OUT.txt"],
             "CanTemp": ["CanopyTemperature_IN.txt", "CanopyTemperature_OUT.txt"],
             "BGFlux": ["BareGroundFluxes_IN.txt", "BareGroundFluxes_OUT.txt"],
             "CanFlux": ["CanopyFluxes_IN.txt", "CanopyFluxes_OUT.txt"]}
# test_CanFlux also reads the PFT constants of clm_params_c180524.nc through ELM::IO::read_pft_var: served, as in the
# oracle's build of the same test, by oracle/shim_serial/netcdf.h from the text dumps of oracle/dump_params.py (copied
# next to the binary), with the zero-filling operator new of oracle/shim_serial/zero_new.cc (the test reads scratch
# arrays it never initialises)
NEEDS_PARAMS = {"CanFlux"}


def build_reference_tests(force=False):
    import shutil
    if not REF.is_dir():
        return []
    out = HERE / "_build"
    (out / "data").mkdir(parents=True, exist_ok=True)
    built = []
    for t, data in REF_TESTS.items():
        exe = out / f"ref_test_{t}"
        for d in data:
            shutil.copyfile(REF / "test/data" / d, out / "data" / d)
        deps = list((ROOT / "include/elm").glob("*.h*")) + [ROOT / "include/elmk_b200.h", REF / f"test/test_{t}.cc"]
        if not force and exe.exists() and all(s.stat().st_mtime <= exe.stat().st_mtime for s in deps):
            built.append(exe)
            continue
        extra, inc, tail_inc = [], [], []
        if t in NEEDS_PARAMS:
            sys.path.insert(0, str(ROOT / "oracle"))
            import dump_params
            if not (dump_params.OUT / "pftname.txt").exists():
                dump_params.dump()
            shutil.copytree(dump_params.OUT, out / "data" / "clm_params", dirs_exist_ok=True)
            inc = [f"-I{ROOT}/oracle/shim_serial", '-DELMK_NC_DUMP_DEFAULT="data/clm_params"']
            # (test_CanFlux includes data_types.hh -> elm_state.h, whose data headers pull snow_snicar.h, atm_physics.h
            #  and phenology_physics.h for constants and data-manager functors: those resolve to the reference's
            #  src/physics, placed AFTER include/elm - every function the test calls resolves to include/elm)
            tail_inc = [f"-I{REF}/src/physics"]
            extra = [str(REF / "src/utils/read_input.cc"), str(REF / "src/utils/utils.cc"), str(ROOT / "oracle/shim_serial/zero_new.cc")]
        cmd = ["g++", "-std=c++17", "-O1", "-w", '-DTEST_DATA_DIR="data/"'] + inc + [f"-I{ROOT}/include/elm", f"-I{ROOT}/include",
               f"-I{REF}/src/data", f"-I{REF}/src/utils"] + tail_inc + [str(REF / f"test/test_{t}.cc"), str(REF / "src/utils/read_test_input.cc")] + extra + [
               f"-L{ROOT}/elmkernels_b200", "-lelmk_b200", "-Wl,-rpath,$ORIGIN/../../../elmkernels_b200", "-o", str(exe)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr[-6000:])
            raise SystemExit(f"reference test_{t} against include/elm: build failed")
        built.append(exe)
    return built


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_reference_tests(force="--force" in sys.argv))
