import os
import sys

import pytest

# make the oracle deterministic where the reference reads uninitialised stack (oracle/shim/Kokkos_Core.hpp)
os.environ.setdefault("ELMREF_SCRUB_STACK", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def params():
    from elmkernels_b200 import params as p
    return p.load_params()


@pytest.fixture(scope="session")
def ref_lib():
    """oracle/_ref/libelmref.so: the reference's own code behind the C ABI (prebuilt; travels to the GPU box)."""
    from elmkernels_b200 import abi
    path = os.path.join(ROOT, "oracle", "_ref", "libelmref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference"):
            import subprocess
            subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "build_ref.py")])
        else:
            pytest.skip("oracle/_ref/libelmref.so not built and /root/reference not present")
    return abi.Library(path)


@pytest.fixture(scope="session")
def port_lib():
    """oracle/port: host build of the physics core (test infrastructure)."""
    from elmkernels_b200 import abi
    sys.path.insert(0, os.path.join(ROOT, "oracle", "port"))
    import build_port
    return abi.Library(str(build_port.build()))


@pytest.fixture(scope="session")
def checker(request):
    """What the CUDA library is compared with: the reference itself when oracle/_ref is present, else the pinned port."""
    from elmkernels_b200 import abi
    path = os.path.join(ROOT, "oracle", "_ref", "libelmref.so")
    if os.path.exists(path):
        return abi.Library(path)
    return request.getfixturevalue("port_lib")


@pytest.fixture(scope="session")
def cuda_lib():
    import elmkernels_b200
    return elmkernels_b200.load()
