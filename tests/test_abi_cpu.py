"""CPU-side checks of the drop-in boundary: the product library exists, loads, and exports every symbol
that include/elmk_b200.h declares; its field table equals include/elmk_fields.def.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "elmk_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(elmk_[a-z_0-9]+)\s*\(", text)))


def field_rows():
    rows = []
    for line in open(os.path.join(ROOT, "include", "elmk_fields.def")):
        m = re.match(r"ELMK_FIELD\((\w+), (\w+), (\d+), (\w+)\)", line)
        if m:
            rows.append((m.group(1), {"F64": 0, "I32": 1, "U8": 2}[m.group(2)], int(m.group(3)), m.group(4)))
    return rows


@pytest.fixture(scope="module")
def product_path():
    import __graft_entry__ as ge
    ge.build_product()
    from elmkernels_b200 import LIB_PATH
    assert os.path.exists(LIB_PATH), "libelmk_b200.so was not built"
    return LIB_PATH


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for s in ("elmk_create", "elmk_destroy", "elmk_set_tables", "elmk_upload", "elmk_download", "elmk_step",
              "elmk_init_timestep", "elmk_errors", "elmk_diag_reduce", "elmk_sync"):
        assert s in syms
    assert len(syms) >= 24


def test_product_library_exports_every_declared_symbol(product_path):
    dll = ctypes.CDLL(product_path)
    missing = [s for s in declared_symbols() if not hasattr(dll, s)]
    assert not missing, f"libelmk_b200.so does not export {missing}"
    dll.elmk_backend.restype = ctypes.c_char_p
    assert dll.elmk_backend() == b"cuda-sm100a"
    dll.elmk_abi_version.restype = ctypes.c_int
    assert dll.elmk_abi_version() == 6


def test_field_table_matches_def_file(product_path):
    from elmkernels_b200 import abi
    lib = abi.Library(product_path)
    rows = field_rows()
    assert len(rows) == 223 and len(lib.field_names) == len(rows)
    for (name, dt, nlev, _), got in zip(rows, lib.field_names):
        assert name == got
        assert lib.fields[name][1:] == (dt, nlev)


def test_algorithmic_bytes_per_column_step():
    """The roofline denominator of bench.py is derived from the field classes."""
    import bench
    size = {0: 8, 1: 4, 2: 1}
    rows = field_rows()
    declared = sum(size[dt] * n * {"IN": 1, "PROG": 2, "OUT": 1, "DEAD": 0}[cls] for _, dt, n, cls in rows)
    assert declared == bench.DECLARED_BYTES_PER_COLUMN_STEP
    assert bench.ALGORITHMIC_BYTES_PER_COLUMN_STEP == 6325


def test_product_has_no_cpu_fallback(product_path):
    """elmk_create on a machine without a CUDA device must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from elmkernels_b200 import abi
    lib = abi.Library(product_path)
    with pytest.raises(abi.ElmkError):
        lib.columns(16)


def test_load_refuses_a_non_cuda_library(tmp_path, monkeypatch):
    import elmkernels_b200
    monkeypatch.setattr(elmkernels_b200, "_LIB", None)
    monkeypatch.setattr(elmkernels_b200, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(elmkernels_b200.ElmkError):
        elmkernels_b200.load()
