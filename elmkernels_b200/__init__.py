"""elmkernels_b200 - B200-native (hand-written FP64 CUDA, sm_100a) implementation of the ELMKernels
per-column land-surface timestep, behind the C ABI of include/elmk_b200.h.

The product library is libelmk_b200.so in this directory.  There is no CPU fallback: `load()` raises
if the CUDA library has not been built (python __graft_entry__.py, or python elmkernels_b200/build.py).
"""
import os

from .abi import Columns, ElmkError, Library  # noqa: F401
from . import abi  # noqa: F401

_LIB = None
# ELMK_LIB: development override used for A/B measurements of differently built copies of the same CUDA library
LIB_PATH = os.environ.get("ELMK_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libelmk_b200.so")


def load() -> Library:
    """The CUDA product library.  Fails loudly when it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ElmkError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                            "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback")
        _LIB = Library(LIB_PATH)
        if not _LIB.backend.startswith("cuda"):
            raise ElmkError(f"{LIB_PATH} reports backend {_LIB.backend!r}, expected the CUDA build")
    return _LIB
