"""The C++ host adaptor include/elm_b200.hh compiles against the reference's own ELMState (every drop-in
wrapper instantiates for ELMStateType) and moves data through the index-map interface correctly.  Needs the
reference headers, so it runs only where /root/reference is mounted (this container, not the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ELMK_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference headers not present")
def test_adaptor_compiles_against_reference_state_and_round_trips(port_lib, tmp_path):
    exe = tmp_path / "adaptor_check"
    libdir = os.path.dirname(port_lib.path)
    cmd = ["g++", "-std=c++17", "-O1", "-w", "-fopenmp", "-DENABLE_KOKKOS", '-DINPUT_DATA_DIR="/nonexistent/"',
           f"-I{ROOT}/oracle/shim", f"-I{ROOT}/include", f"-I{REF}/driver/kokkos", f"-I{REF}/src/physics",
           f"-I{REF}/src/data", f"-I{REF}/src/utils", f"{ROOT}/tests/adaptor/adaptor_check.cc", f"{REF}/src/utils/utils.cc",
           f"{REF}/src/utils/read_input.cc", f"-L{libdir}", "-lelmport", f"-Wl,-rpath,{libdir}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")
def test_reference_unit_tests_compile_unchanged_against_the_function_api():
    """Six of the reference's seven test/test_*.cc build from where they lie with include/elm in place of the
    reference's src/physics (tests/adaptor/build_adaptor.py); they run on the GPU in tests/test_gpu_adaptor.py."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "adaptor"))
    import build_adaptor
    import __graft_entry__ as ge
    ge.build_product()
    built = [os.path.basename(str(p)) for p in build_adaptor.build_reference_tests()]
    assert sorted(built) == ["ref_test_BGFlux", "ref_test_CanFlux", "ref_test_CanHydro", "ref_test_CanSunShade",
                             "ref_test_CanTemp", "ref_test_SurfRad"]
