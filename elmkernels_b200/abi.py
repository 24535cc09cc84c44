"""ctypes binding of the C ABI declared in include/elmk_b200.h.

`Library` wraps one shared object that exports the elmk_* symbols; `Columns` is one handle
(one device, one contiguous range of land columns).  The product library is
elmkernels_b200/libelmk_b200.so (hand-written CUDA for sm_100a); the test suite also points
this same binding at the oracle libraries under oracle/ - the ABI is identical on purpose so that
parity tests push the same bytes through both.

Host arrays use the reference's layout (column outer, level inner: numpy shape (ncols,) or
(ncols, nlev), C-contiguous), see reference src/utils/array.hh:176-183.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, Mapping, Optional, Sequence

import numpy as np

F64, I32, U8 = 0, 1, 2
COL_OUTER, COL_INNER = 0, 1
_NP = {F64: np.float64, I32: np.int32, U8: np.uint8}

# kernel groups, include/elmk_b200.h
G_FRAC_WET = 1 << 0
G_ALBEDO = 1 << 1
G_CANOPY_HYDROLOGY = 1 << 2
G_SURFACE_RADIATION = 1 << 3
G_CANOPY_TEMPERATURE = 1 << 4
G_BAREGROUND_FLUXES = 1 << 5
G_CANOPY_FLUXES = 1 << 6
G_SOIL_TEMPERATURE = 1 << 7
G_SNOW_HYDROLOGY = 1 << 8
G_SURFACE_FLUXES = 1 << 9
G_CONSERVATION = 1 << 10
G_ALL = 0x7FF
GROUP_NAMES = ["frac_wet", "albedo_snicar", "canopy_hydrology", "surface_radiation", "canopy_temperature",
               "bareground_fluxes", "canopy_fluxes", "soil_temperature", "snow_hydrology", "surface_fluxes",
               "conservation"]

NPFT_TABLES = 40
PFT_ORDER = ("fnr act25 kcha koha cpha vcmaxha jmaxha tpuha lmrha vcmaxhd jmaxhd tpuhd lmrhd lmrse qe theta_cj "
             "bbbopt mbbopt c3psn slatop leafcn flnr fnitr dleaf smpso smpsc tc_stress z0mr displar xl roota_par "
             "rootb_par rholvis rholnir rhosvis rhosnir taulvis taulnir tausvis tausnir").split()
# the 27 members of the reference's PFTDataPSN, in declaration order (src/data/pft_data.h:20-24)
PSN_ORDER = PFT_ORDER[:27]
SNICAR_BAND = [f"{k}_{s}" for s in ("oc1", "oc2", "dst1", "dst2", "dst3", "dst4")
               for k in ("ss_alb", "asm_prm", "ext_cff_mss")]
SNICAR_SNOW = [f"{k}_snw_{d}" for d in ("drc", "dfs") for k in ("ss_alb", "asm_prm", "ext_cff_mss")]
SNICAR_BC = [f"{k}_{s}" for s in ("bc1", "bc2") for k in ("ss_alb", "asm_prm", "ext_cff_mss")]

_PD = C.POINTER(C.c_double)


class Tables(C.Structure):
    """struct elmk_tables"""
    _fields_ = [("ltype", C.c_int32), ("ctype", C.c_int32), ("vtype", C.c_int32), ("urbpoi", C.c_int32),
                ("lakpoi", C.c_int32), ("oldfflag", C.c_int32), ("dewmx", C.c_double),
                ("pft", _PD * NPFT_TABLES), ("albsat", _PD), ("albdry", _PD), ("snicar_band", _PD * 18),
                ("snicar_snow", _PD * 6), ("snicar_bc", _PD * 6), ("bcenh", _PD), ("snowage", _PD * 3)]


def table_arrays(params):
    """The arrays behind the pointer members of struct elmk_tables, in member order (pft[40], albsat, albdry,
    snicar_band[18], snicar_snow[6], snicar_bc[6], bcenh, snowage[3])."""
    out = [params["pft_" + n] for n in PFT_ORDER] + [params["albsat"], params["albdry"]]
    out += [params["snicar_" + n] for n in SNICAR_BAND] + [params["snicar_" + n] for n in SNICAR_SNOW]
    out += [params["snicar_" + n] for n in SNICAR_BC] + [params["snicar_bcenh"]]
    out += [params["snowage_" + n] for n in ("tau", "kappa", "drdt0")]
    return out


class ElmkError(RuntimeError):
    pass


class Library:
    """One shared object exporting the elmk_* C ABI."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise ElmkError(f"shared library not found: {path}")
        self.path = path
        self.dll = C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))
        d = self.dll
        H = C.c_void_p
        sig = {
            "elmk_abi_version": (C.c_int, []),
            "elmk_backend": (C.c_char_p, []),
            "elmk_field_count": (C.c_int, []),
            "elmk_field_id": (C.c_int, [C.c_char_p]),
            "elmk_field_info": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
            "elmk_create": (C.c_int, [C.POINTER(H), C.c_int, C.c_int64]),
            "elmk_destroy": (C.c_int, [H]),
            "elmk_last_error": (C.c_char_p, [H]),
            "elmk_ncols": (C.c_int64, [H]),
            "elmk_set_tables": (C.c_int, [H, C.POINTER(Tables)]),
            "elmk_upload": (C.c_int, [H, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int]),
            "elmk_download": (C.c_int, [H, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int]),
            "elmk_fill": (C.c_int, [H, C.c_int, C.c_double]),
            "elmk_upload_many": (C.c_int, [H, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_int64,
                                           C.c_int64, C.c_int]),
            "elmk_download_many": (C.c_int, [H, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_int64,
                                             C.c_int64, C.c_int]),
            "elmk_init_timestep": (C.c_int, [H, C.c_int]),
            "elmk_step": (C.c_int, [H, C.c_double, C.c_double, C.c_double, C.c_uint32]),
            "elmk_set_coordinates": (C.c_int, [H, _PD, _PD, C.c_int64]),
            "elmk_solar_step": (C.c_int, [H, C.c_double, C.c_double, C.c_int, _PD, _PD]),
            "elmk_fn_call": (C.c_int, [C.c_int, C.c_int, _PD, C.c_int64]),
            "elmk_set_gas_pressures": (C.c_int, [C.c_void_p, _PD, _PD]),
            "elmk_sync": (C.c_int, [H]),
            "elmk_set_plan": (C.c_int, [H, C.c_int]),
            "elmk_launch_count": (C.c_int64, [H]),
            "elmk_errors": (C.c_int, [H, C.POINTER(C.c_uint32), C.POINTER(C.c_int64)]),
            "elmk_clear_errors": (C.c_int, [H]),
            "elmk_error_text": (C.c_char_p, [C.c_uint32]),
            "elmk_diag_reduce": (C.c_int, [H, _PD]),
            "elmk_device_ptr": (C.c_int, [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
            "elmk_stream": (C.c_int, [H, C.POINTER(C.c_void_p)]),
            "elmk_exchange_create": (C.c_int, [H, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.POINTER(H)]),
            "elmk_exchange_destroy": (C.c_int, [H]),
            "elmk_exchange_post": (C.c_int, [H, C.POINTER(C.c_void_p)]),
            "elmk_exchange_commit": (C.c_int, [H]),
            "elmk_exchange_fetch": (C.c_int, [H, C.POINTER(C.c_void_p)]),
            "elmk_exchange_wait": (C.c_int, [H]),
            "elmk_exchange_post_wait": (C.c_int, [H]),
            "elmk_math_eval": (C.c_int, [H, C.c_int, C.c_int64, _PD, _PD, _PD]),
            "elmk_init_columns": (C.c_int, [H, _PD, _PD, _PD, C.c_double, _PD]),
            "elmk_atm_series": (C.c_int, [H, C.c_int, _PD, C.c_int]),
            "elmk_atm_forcing": (C.c_int, [H, C.c_int, C.c_double, C.c_double, C.c_int]),
            "elmk_atm_series_row": (C.c_int, [H, C.c_int, C.c_int, _PD]),
            "elmk_canflux_pass_histogram": (C.c_int, [H, C.POINTER(C.c_int64)]),
            "elmk_phen_series": (C.c_int, [H, C.c_int, _PD, C.c_int]),
            "elmk_phenology": (C.c_int, [H, C.c_int, C.c_double, C.c_double]),
            "elmk_timing_enable": (C.c_int, [H, C.c_int]),
            "elmk_timing_read": (C.c_int, [H, C.c_int, C.POINTER(C.c_char_p), _PD, C.POINTER(C.c_int64),
                                           C.POINTER(C.c_uint32)]),
        }
        self.symbols = list(sig)
        for name, (res, args) in sig.items():
            fn = getattr(d, name)  # AttributeError here = the library does not export the declared ABI
            fn.restype, fn.argtypes = res, args
        self.backend = d.elmk_backend().decode()
        self.fields: Dict[str, tuple] = {}
        self.field_names = []
        nm, dt, nl = C.c_char_p(), C.c_int(), C.c_int()
        for i in range(d.elmk_field_count()):
            d.elmk_field_info(i, C.byref(nm), C.byref(dt), C.byref(nl))
            self.fields[nm.value.decode()] = (i, dt.value, nl.value)
            self.field_names.append(nm.value.decode())

    def columns(self, ncols: int, device: int = 0) -> "Columns":
        return Columns(self, ncols, device)


class Exchange:
    """Double-buffered host<->device exchange that overlaps with the step (include/elmk_b200.h, elmk_exchange_*).
    Host arrays are in the reference layout (ncols[, nlev]); pass pinned memory for real overlap."""

    def __init__(self, cols: "Columns", in_names: Sequence[str], out_names: Sequence[str]):
        self.cols, self.in_names, self.out_names = cols, list(in_names), list(out_names)
        ids_in = (C.c_int * len(self.in_names))(*[cols._spec(n)[0] for n in self.in_names])
        ids_out = (C.c_int * len(self.out_names))(*[cols._spec(n)[0] for n in self.out_names])
        self._x = C.c_void_p()
        cols._check(cols.lib.dll.elmk_exchange_create(cols._h, len(self.in_names), ids_in, len(self.out_names), ids_out,
                                                     C.byref(self._x)), "elmk_exchange_create")
        self._keep = []

    def _ptrs(self, names, arrays):
        assert len(arrays) == len(names)
        for n, a in zip(names, arrays):
            _, dt, nl = self.cols._spec(n)
            assert a.flags["C_CONTIGUOUS"] and a.shape[0] == self.cols.ncols and a.dtype == _NP[dt], n
        return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])

    def post(self, arrays: Sequence[np.ndarray]):
        self._keep = (self._keep + [list(arrays)])[-4:]   # the copies are asynchronous: keep the buffers alive
        self.cols._check(self.cols.lib.dll.elmk_exchange_post(self._x, self._ptrs(self.in_names, arrays)), "elmk_exchange_post")

    def commit(self):
        self.cols._check(self.cols.lib.dll.elmk_exchange_commit(self._x), "elmk_exchange_commit")

    def fetch(self, arrays: Sequence[np.ndarray]):
        self._keep = (self._keep + [list(arrays)])[-4:]
        self.cols._check(self.cols.lib.dll.elmk_exchange_fetch(self._x, self._ptrs(self.out_names, arrays)), "elmk_exchange_fetch")

    def wait(self):
        self.cols._check(self.cols.lib.dll.elmk_exchange_wait(self._x), "elmk_exchange_wait")

    def post_wait(self):
        """Block until the oldest post has left its host buffers (they may then be refilled)."""
        self.cols._check(self.cols.lib.dll.elmk_exchange_post_wait(self._x), "elmk_exchange_post_wait")

    def close(self):
        if self._x:
            self.cols.lib.dll.elmk_exchange_destroy(self._x)
            self._x = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Columns:
    """One handle: `ncols` land columns resident on one device."""

    def __init__(self, lib: Library, ncols: int, device: int = 0):
        self.lib, self.ncols = lib, int(ncols)
        self._h = C.c_void_p()
        self._keep = None
        self._exchanges = []
        rc = lib.dll.elmk_create(C.byref(self._h), int(device), self.ncols)
        if rc != 0:
            raise ElmkError(f"elmk_create({ncols}) failed with {rc} on {lib.backend}")

    # -- plumbing --
    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.dll.elmk_last_error(self._h)
            raise ElmkError(f"{what} failed with {rc}: {msg.decode() if msg else ''}")

    def close(self):
        for x in getattr(self, "_exchanges", []):
            x.close()
        self._exchanges = []
        if self._h:
            self.lib.dll.elmk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- tables --
    def set_tables(self, params: Mapping[str, np.ndarray], land: Optional[Mapping[str, int]] = None,
                   dewmx: float = 0.1, oldfflag: int = 1):
        """params: arrays keyed pft_<name>, snicar_<name>, albsat, albdry, snowage_{tau,kappa,drdt0}."""
        land = {**dict(ltype=1, ctype=1, vtype=12, urbpoi=0, lakpoi=0), **(land or {})}
        t = Tables()
        keep = []

        def ptr(a, size):
            a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
            if a.size < size:
                raise ElmkError(f"table too small: {a.size} < {size}")
            keep.append(a)
            return a.ctypes.data_as(_PD)

        for k in ("ltype", "ctype", "vtype", "urbpoi", "lakpoi"):
            setattr(t, k, int(land[k]))
        t.oldfflag, t.dewmx = int(oldfflag), float(dewmx)
        for i, n in enumerate(PFT_ORDER):
            t.pft[i] = ptr(params["pft_" + n], 1 if n == "tc_stress" else 17)
        t.albsat, t.albdry = ptr(params["albsat"], 40), ptr(params["albdry"], 40)
        for i, n in enumerate(SNICAR_BAND):
            t.snicar_band[i] = ptr(params["snicar_" + n], 5)
        for i, n in enumerate(SNICAR_SNOW):
            t.snicar_snow[i] = ptr(params["snicar_" + n], 5 * 1471)
        for i, n in enumerate(SNICAR_BC):
            t.snicar_bc[i] = ptr(params["snicar_" + n], 50)
        t.bcenh = ptr(params["snicar_bcenh"], 400)
        for i, n in enumerate(("tau", "kappa", "drdt0")):
            t.snowage[i] = ptr(params["snowage_" + n], 11 * 31 * 8)
        self._check(self.lib.dll.elmk_set_tables(self._h, C.byref(t)), "elmk_set_tables")
        self._keep = keep

    # -- state movement --
    def _spec(self, name: str):
        try:
            return self.lib.fields[name]
        except KeyError:
            raise ElmkError(f"unknown field {name!r}") from None

    def host_array(self, name: str, n: Optional[int] = None) -> np.ndarray:
        _, dt, nl = self._spec(name)
        n = self.ncols if n is None else n
        return np.zeros((n,) if nl == 1 else (n, nl), dtype=_NP[dt])

    def upload(self, name: str, arr: np.ndarray, col0: int = 0):
        fid, dt, nl = self._spec(name)
        a = np.ascontiguousarray(arr, dtype=_NP[dt])
        n = a.shape[0]
        if a.size != n * nl:
            raise ElmkError(f"{name}: expected {nl} elements per column, got shape {a.shape}")
        self._check(self.lib.dll.elmk_upload(self._h, fid, a.ctypes.data, col0, n, COL_OUTER), f"elmk_upload({name})")

    def download(self, name: str, col0: int = 0, n: Optional[int] = None, out: Optional[np.ndarray] = None):
        fid, dt, nl = self._spec(name)
        n = self.ncols - col0 if n is None else n
        a = self.host_array(name, n) if out is None else out
        self._check(self.lib.dll.elmk_download(self._h, fid, a.ctypes.data, col0, n, COL_OUTER),
                    f"elmk_download({name})")
        return a

    def upload_state(self, state: Mapping[str, np.ndarray], col0: int = 0):
        for k, v in state.items():
            self.upload(k, v, col0)

    def download_state(self, names: Optional[Iterable[str]] = None) -> Dict[str, np.ndarray]:
        return {k: self.download(k) for k in (names or self.lib.field_names)}

    def fill(self, name: str, value: float):
        self._check(self.lib.dll.elmk_fill(self._h, self._spec(name)[0], float(value)), f"elmk_fill({name})")

    def plan(self, names: Sequence[str], arrays: Sequence[np.ndarray]):
        """Pre-marshal a (fields, host buffers) pair for upload_many/download_many."""
        ids = (C.c_int * len(names))(*[self._spec(n)[0] for n in names])
        ptrs = (C.c_void_p * len(names))(*[a.ctypes.data for a in arrays])
        return ids, ptrs, len(names), list(arrays)

    def upload_many(self, plan, col0: int = 0, n: Optional[int] = None):
        ids, ptrs, k, _ = plan
        self._check(self.lib.dll.elmk_upload_many(self._h, k, ids, ptrs, col0, self.ncols if n is None else n,
                                                  COL_OUTER), "elmk_upload_many")

    def download_many(self, plan, col0: int = 0, n: Optional[int] = None):
        ids, ptrs, k, _ = plan
        self._check(self.lib.dll.elmk_download_many(self._h, k, ids, ptrs, col0, self.ncols if n is None else n,
                                                    COL_OUTER), "elmk_download_many")

    def exchange(self, in_names: Sequence[str], out_names: Sequence[str]) -> "Exchange":
        """Overlapped per-step exchange of a fixed set of input and output fields (elmk_exchange_*)."""
        x = Exchange(self, in_names, out_names)
        self._exchanges.append(x)
        return x

    def init_columns(self, pct_sand: np.ndarray, pct_clay: np.ndarray, organic: np.ndarray, organic_max: float,
                     snow_depth: np.ndarray):
        """One-time cold start of every column (initialize_elm_kokkos.cc:374-431) from soil texture (ncols, 15) and the
        initial snow depth (ncols,); vtype, topo_slope, topo_std and the soil grid must be in the state already."""
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (pct_sand, pct_clay, organic)]
        d = np.ascontiguousarray(snow_depth, dtype=np.float64)
        for x in a:
            if x.shape != (self.ncols, 15):
                raise ElmkError(f"soil texture arrays must have shape ({self.ncols}, 15), got {x.shape}")
        if d.shape != (self.ncols,):
            raise ElmkError("snow_depth must have one value per column")
        self._check(self.lib.dll.elmk_init_columns(self._h, a[0].ctypes.data_as(_PD), a[1].ctypes.data_as(_PD),
                                                   a[2].ctypes.data_as(_PD), float(organic_max), d.ctypes.data_as(_PD)),
                    "elmk_init_columns")

    # -- per-step input producers on the device (forcing functors, phenology) --
    ATM_VARS = ("TBOT", "PBOT", "QBOT", "FLDS", "FSDS", "PREC", "WIND")
    PHEN_VARS = ("MLAI", "MSAI", "MHTOP", "MHBOT")

    def _series(self, fn, what, var_index: int, arr: np.ndarray):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        if a.ndim != 2 or a.shape[1] != self.ncols:
            raise ElmkError(f"{what}: expected shape (ntimes, {self.ncols}), got {a.shape}")
        self._check(fn(self._h, var_index, a.ctypes.data_as(_PD), a.shape[0]), what)

    def atm_series(self, var: str, arr: np.ndarray):
        """Raw forcing series of one variable, shape (ntimes, ncols) as AtmDataManager::data."""
        self._series(self.lib.dll.elmk_atm_series, f"elmk_atm_series({var})", self.ATM_VARS.index(var), arr)

    def atm_series_row(self, var: str, t: int, row: np.ndarray):
        """Replace time level t of a resident series (asynchronous for pinned `row`, which must stay alive)."""
        assert row.dtype == np.float64 and row.shape == (self.ncols,) and row.flags["C_CONTIGUOUS"]
        self._check(self.lib.dll.elmk_atm_series_row(self._h, self.ATM_VARS.index(var), int(t), row.ctypes.data_as(_PD)),
                    "elmk_atm_series_row")

    def canflux_pass_histogram(self) -> np.ndarray:
        """hist[k] = vegetated columns whose stability iteration took k passes in the last step (CUDA library only)."""
        hist = np.zeros(42, dtype=np.int64)
        self._check(self.lib.dll.elmk_canflux_pass_histogram(self._h, hist.ctypes.data_as(C.POINTER(C.c_int64))),
                    "elmk_canflux_pass_histogram")
        return hist

    def atm_forcing(self, t_idx: int, wt1: float, wt2: float, qbot_is_rh: bool = True):
        self._check(self.lib.dll.elmk_atm_forcing(self._h, int(t_idx), float(wt1), float(wt2), int(qbot_is_rh)),
                    "elmk_atm_forcing")

    def phen_series(self, var: str, arr: np.ndarray):
        self._series(self.lib.dll.elmk_phen_series, f"elmk_phen_series({var})", self.PHEN_VARS.index(var), arr)

    def phenology(self, start_idx: int, wt1: float, wt2: float):
        self._check(self.lib.dll.elmk_phenology(self._h, int(start_idx), float(wt1), float(wt2)), "elmk_phenology")

    # -- stepping --
    def set_coordinates(self, lat_r, lon_r):
        """Latitude / longitude [rad]: scalars (one site for every column, as in the reference) or one per column."""
        lat = np.ascontiguousarray(np.atleast_1d(lat_r), dtype=np.float64)
        lon = np.ascontiguousarray(np.atleast_1d(lon_r), dtype=np.float64)
        assert lat.shape == lon.shape and lat.size in (1, self.ncols)
        self._check(self.lib.dll.elmk_set_coordinates(self._h, lat.ctypes.data_as(_PD), lon.ctypes.data_as(_PD), lat.size),
                    "elmk_set_coordinates")

    def solar_step(self, dtime: float, decday: float, doy1: int):
        """coszen of every column for the step starting at decimal day `decday`; returns (dayl, max_dayl)."""
        d, m = C.c_double(), C.c_double()
        self._check(self.lib.dll.elmk_solar_step(self._h, dtime, decday, int(doy1), C.byref(d), C.byref(m)), "elmk_solar_step")
        return d.value, m.value

    def init_timestep(self, reset_forc_hgt: bool = True):
        self._check(self.lib.dll.elmk_init_timestep(self._h, int(reset_forc_hgt)), "elmk_init_timestep")

    def step(self, dtime: float = 1800.0, dayl: float = 50000.0, max_dayl: float = 86400.0, groups: int = G_ALL):
        self._check(self.lib.dll.elmk_step(self._h, dtime, dayl, max_dayl, groups), "elmk_step")

    def set_gas_pressures(self, forc_pco2=None, forc_po2=None):
        """Per-column CO2 / O2 partial pressures [Pa] for the photosynthesis of group a7 (None, None: back to the
        reference wrapper's constants)."""
        if forc_pco2 is None and forc_po2 is None:
            self._check(self.lib.dll.elmk_set_gas_pressures(self._h, None, None), "elmk_set_gas_pressures")
            return
        a = np.ascontiguousarray(forc_pco2, dtype=np.float64).reshape(self.ncols)
        b = np.ascontiguousarray(forc_po2, dtype=np.float64).reshape(self.ncols)
        self._check(self.lib.dll.elmk_set_gas_pressures(self._h, a.ctypes.data_as(_PD), b.ctypes.data_as(_PD)),
                    "elmk_set_gas_pressures")

    def set_plan(self, plan: str):
        """'fused' (default, production) or 'split' (one launch per kernel group, one thread per column)."""
        self._check(self.lib.dll.elmk_set_plan(self._h, {"fused": 0, "split": 1}[plan]), "elmk_set_plan")

    def sync(self):
        self._check(self.lib.dll.elmk_sync(self._h), "elmk_sync")

    @property
    def launch_count(self) -> int:
        return int(self.lib.dll.elmk_launch_count(self._h))

    def errors(self):
        any_, first = C.c_uint32(), C.c_int64()
        self._check(self.lib.dll.elmk_errors(self._h, C.byref(any_), C.byref(first)), "elmk_errors")
        return any_.value, first.value

    def clear_errors(self):
        self._check(self.lib.dll.elmk_clear_errors(self._h), "elmk_clear_errors")

    MATH = {"exp": 0, "log": 1, "log10": 2, "pow": 3, "atan": 4, "cos": 5, "tanh": 6, "erf": 7, "acos": 8, "div": 9}

    def math_eval(self, fn: str, x: np.ndarray, y: Optional[np.ndarray] = None) -> np.ndarray:
        """The library's own transcendental `fn` at x (and y), evaluated where the library computes."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        yp = None
        if y is not None:
            y = np.ascontiguousarray(y, dtype=np.float64)
            assert y.shape == x.shape
            yp = y.ctypes.data_as(_PD)
        self._check(self.lib.dll.elmk_math_eval(self._h, self.MATH[fn], x.size, x.ctypes.data_as(_PD), yp,
                                                out.ctypes.data_as(_PD)), "elmk_math_eval")
        return out

    def diag_reduce(self) -> np.ndarray:
        out = np.zeros(24)
        self._check(self.lib.dll.elmk_diag_reduce(self._h, out.ctypes.data_as(_PD)), "elmk_diag_reduce")
        return out

    @property
    def stream(self) -> int:
        p = C.c_void_p()
        self._check(self.lib.dll.elmk_stream(self._h, C.byref(p)), "elmk_stream")
        return p.value or 0

    def timing(self, on):
        """False / True / 2 (also the sub-launches of the composite launches)."""
        self._check(self.lib.dll.elmk_timing_enable(self._h, int(on)), "elmk_timing_enable")

    def timing_read(self):
        """[(launch name, group mask, total ms, launches)] since timing(True)."""
        k = 64
        names = (C.c_char_p * k)()
        ms = (C.c_double * k)()
        cnt = (C.c_int64 * k)()
        masks = (C.c_uint32 * k)()
        n = self.lib.dll.elmk_timing_read(self._h, k, names, ms, cnt, masks)
        if n < 0:
            self._check(n, "elmk_timing_read")
        return [(names[i].decode(), int(masks[i]), float(ms[i]), int(cnt[i])) for i in range(min(n, k))]

    def device_ptr(self, name: str):
        p, s = C.c_void_p(), C.c_int64()
        self._check(self.lib.dll.elmk_device_ptr(self._h, self._spec(name)[0], C.byref(p), C.byref(s)),
                    "elmk_device_ptr")
        return p.value, s.value
