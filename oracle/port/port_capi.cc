// TEST INFRASTRUCTURE (oracle) - not part of the product path.
//
// libelmport.so: the CPU restatement of the column timestep.  It compiles the single-source physics
// core elmkernels_b200/csrc/phys_*.h (each function there cites the reference file:line it follows)
// with g++ for the host, behind the same C ABI as the product library (include/elmk_b200.h), using
// the same column-innermost storage.  Built with -DELMK_EXACT_POW and without FMA contraction
// (x86-64 baseline), it is pinned bit-for-bit / to 1e-15 against oracle/_ref (the reference itself)
// and against the reference's golden vectors by tests/test_oracle_*.py.  The GPU parity tests then use
// it (and oracle/_ref) as the checker.  The product library never links or loads this file.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../elmkernels_b200/csrc/elmk_state.h"
#include "../../elmkernels_b200/csrc/phys_albedo.h"
#include "../../elmkernels_b200/csrc/phys_bareground.h"
#include "../../elmkernels_b200/csrc/phys_canflux.h"
#include "../../elmkernels_b200/csrc/phys_cantemp.h"
#include "../../elmkernels_b200/csrc/phys_forcing.h"
#include "../../elmkernels_b200/csrc/phys_hydrology.h"
#include "../../elmkernels_b200/csrc/phys_init.h"
#include "../../elmkernels_b200/csrc/phys_radiation.h"
#include "../../elmkernels_b200/csrc/phys_snow.h"
#include "../../elmkernels_b200/csrc/phys_soiltemp.h"
#include "../../elmkernels_b200/csrc/phys_surfflux.h"
#include "../../include/elmk_b200.h"

namespace {
using namespace elmk;

struct Spec { const char* name; int dtype; int nlev; };
const Spec kSpecs[] = {
#define F64 ELMK_F64
#define I32 ELMK_I32
#define U8 ELMK_U8
#define ELMK_FIELD(name, type, nlev, cls) {#name, type, nlev},
#include "../../include/elmk_fields.def"
#undef ELMK_FIELD
#undef F64
#undef I32
#undef U8
};
constexpr int kNumFields = sizeof(kSpecs) / sizeof(kSpecs[0]);
inline size_t esize(int dt) { return dt == ELMK_F64 ? 8 : dt == ELMK_I32 ? 4 : 1; }

struct PortCtx {
  int64_t ncols = 0, np = 0;
  Cols cols;
  Tables tab;
  std::vector<void*> base;
  std::vector<double> snw[2][3], snowage[3];
  std::vector<double> atm[ATM_NVARS], phen[PHEN_NVARS];   // [ntimes][ncols]
  std::vector<double> coords;   // sin(lat), cos(lat), tan(lat), lon: [4][ncoords]
  std::vector<double> gas[2];   // CO2, O2 partial pressures (elmk_set_gas_pressures)
  int64_t ncoords = 0;
  double lat0 = 0.0;
  bool tables_set = false;
  int64_t launches = 0;
  std::string last_error;
};
PortCtx* ctx(elmk_handle h) { return reinterpret_cast<PortCtx*>(h); }

template <class F> void for_columns(PortCtx& c, F f) {
  const long n = static_cast<long>(c.ncols);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (long i = 0; i < n; ++i) f(static_cast<int>(i));
  c.launches += 1;
}
} // namespace

extern "C" {

int elmk_abi_version(void) { return ELMK_ABI_VERSION; }
const char* elmk_backend(void) { return "port-host"; }
int elmk_field_count(void) { return kNumFields; }
int elmk_field_id(const char* name) {
  for (int i = 0; i < kNumFields; ++i)
    if (std::strcmp(kSpecs[i].name, name) == 0) return i;
  return -1;
}
int elmk_field_info(int field, const char** name, int* dtype, int* nlev) {
  if (field < 0 || field >= kNumFields) return ELMK_EINVAL;
  if (name) *name = kSpecs[field].name;
  if (dtype) *dtype = kSpecs[field].dtype;
  if (nlev) *nlev = kSpecs[field].nlev;
  return ELMK_OK;
}

int elmk_create(elmk_handle* out, int, int64_t ncols) {
  if (!out || ncols <= 0 || ncols > INT32_MAX) return ELMK_EINVAL;
  auto* c = new PortCtx;
  c->ncols = ncols;
  c->np = (ncols + 31) / 32 * 32;
  c->cols.np = c->np;
  c->cols.ncols = static_cast<int>(ncols);
  c->cols.npi = static_cast<int>(c->np);
  c->cols.pco2_in = nullptr;
  c->cols.po2_in = nullptr;
  c->base.resize(kNumFields);
  for (int f = 0; f < kNumFields; ++f)
    c->base[f] = std::calloc(static_cast<size_t>(c->np) * kSpecs[f].nlev, esize(kSpecs[f].dtype));
  int f = 0;
#define ELMK_FIELD(name, type, nlev, cls) c->cols.name = static_cast<elmk_##type*>(c->base[f++]);
#include "../../include/elmk_fields.def"
#undef ELMK_FIELD
  *out = reinterpret_cast<elmk_handle>(c);
  return ELMK_OK;
}
int elmk_destroy(elmk_handle h) {
  if (!h) return ELMK_OK;
  for (void* p : ctx(h)->base) std::free(p);
  delete ctx(h);
  return ELMK_OK;
}
const char* elmk_last_error(elmk_handle h) { return h ? ctx(h)->last_error.c_str() : "null handle"; }
int64_t elmk_ncols(elmk_handle h) { return ctx(h)->ncols; }

int elmk_set_tables(elmk_handle h, const elmk_tables* t) {
  PortCtx& c = *ctx(h);
  if (!(t->ltype == ISTSOIL || t->ltype == ISTCROP) || t->urbpoi || t->lakpoi) {
    c.last_error = "only soil/crop land units without lake or urban points are on the hot path";
    return ELMK_EUNSUPPORTED;
  }
  Tables& T = c.tab;
  T.ltype = t->ltype; T.ctype = t->ctype; T.vtype = t->vtype; T.urbpoi = t->urbpoi; T.lakpoi = t->lakpoi;
  T.oldfflag = t->oldfflag; T.dewmx = t->dewmx;
  for (int v = 0; v < 17; ++v) {
    T.z0mr[v] = t->pft[27][v]; T.displar[v] = t->pft[28][v]; T.xl[v] = t->pft[29][v];
    T.rhol[v][0] = t->pft[32][v]; T.rhol[v][1] = t->pft[33][v];
    T.rhos[v][0] = t->pft[34][v]; T.rhos[v][1] = t->pft[35][v];
    T.taul[v][0] = t->pft[36][v]; T.taul[v][1] = t->pft[37][v];
    T.taus[v][0] = t->pft[38][v]; T.taus[v][1] = t->pft[39][v];
    for (int k = 0; k < 26; ++k) T.psn[k][v] = t->pft[k][v];
    T.psn[26][v] = t->pft[26][0];
    T.roota[v] = t->pft[30][v]; T.rootb[v] = t->pft[31][v];
  }
  std::memcpy(T.albsat, t->albsat, sizeof(T.albsat));
  std::memcpy(T.albdry, t->albdry, sizeof(T.albdry));
  for (int s = 0; s < 6; ++s)
    for (int k = 0; k < 3; ++k) std::memcpy(T.aer_band[s][k], t->snicar_band[s * 3 + k], sizeof(double) * NBND_SNW);
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < 3; ++k) std::memcpy(T.bc[s][k], t->snicar_bc[s * 3 + k], sizeof(double) * 10 * NBND_SNW);
  std::memcpy(T.bcenh, t->bcenh, sizeof(T.bcenh));
  for (int d = 0; d < 2; ++d)
    for (int k = 0; k < 3; ++k) {
      c.snw[d][k].assign(t->snicar_snow[d * 3 + k], t->snicar_snow[d * 3 + k] + NBND_SNW * ELMK_MIE_SNW);
      T.snw[d][k] = c.snw[d][k].data();
    }
  for (int k = 0; k < 3; ++k) {
    c.snowage[k].assign(t->snowage[k], t->snowage[k] + 11 * 31 * 8);
    T.snowage[k] = c.snowage[k].data();
  }
  c.tables_set = true;
  return ELMK_OK;
}

static int move_field(PortCtx& c, int field, void* host, int64_t col0, int64_t n, int layout, bool up) {
  if (field < 0 || field >= kNumFields || col0 < 0 || n < 0 || col0 + n > c.ncols) return ELMK_EINVAL;
  const int nlev = kSpecs[field].nlev;
  const size_t es = esize(kSpecs[field].dtype);
  char* dev = static_cast<char*>(c.base[field]);
  char* hb = static_cast<char*>(host);
  for (int l = 0; l < nlev; ++l)
    for (int64_t col = 0; col < n; ++col) {
      char* d = dev + (static_cast<size_t>(l) * c.np + col0 + col) * es;
      char* s = hb + ((layout == ELMK_COL_OUTER) ? (static_cast<size_t>(col) * nlev + l)
                                                 : (static_cast<size_t>(l) * n + col)) * es;
      if (up) std::memcpy(d, s, es); else std::memcpy(s, d, es);
    }
  return ELMK_OK;
}
int elmk_upload(elmk_handle h, int field, const void* host, int64_t col0, int64_t n, int layout) {
  return move_field(*ctx(h), field, const_cast<void*>(host), col0, n, layout, true);
}
int elmk_download(elmk_handle h, int field, void* host, int64_t col0, int64_t n, int layout) {
  return move_field(*ctx(h), field, host, col0, n, layout, false);
}
int elmk_upload_many(elmk_handle h, int nf, const int* fields, const void* const* hosts, int64_t col0, int64_t n,
                     int layout) {
  for (int i = 0; i < nf; ++i)
    if (int rc = elmk_upload(h, fields[i], hosts[i], col0, n, layout)) return rc;
  return ELMK_OK;
}
int elmk_download_many(elmk_handle h, int nf, const int* fields, void* const* hosts, int64_t col0, int64_t n,
                       int layout) {
  for (int i = 0; i < nf; ++i)
    if (int rc = elmk_download(h, fields[i], hosts[i], col0, n, layout)) return rc;
  return ELMK_OK;
}
int elmk_fill(elmk_handle h, int field, double value) {
  PortCtx& c = *ctx(h);
  if (field < 0 || field >= kNumFields) return ELMK_EINVAL;
  const size_t count = static_cast<size_t>(c.np) * kSpecs[field].nlev;
  if (kSpecs[field].dtype == ELMK_F64) std::fill_n(static_cast<double*>(c.base[field]), count, value);
  else if (kSpecs[field].dtype == ELMK_I32) std::fill_n(static_cast<int*>(c.base[field]), count, static_cast<int>(value));
  else std::fill_n(static_cast<unsigned char*>(c.base[field]), count, static_cast<unsigned char>(value != 0.0));
  return ELMK_OK;
}

int elmk_init_columns(elmk_handle h, const double* pct_sand, const double* pct_clay, const double* organic,
                      double organic_max, const double* snow_depth) {
  PortCtx& c = *ctx(h);
  if (!c.tables_set) return ELMK_ENOTABLES;
  // host[col][lev] -> [lev][ncols], the layout column_init reads
  std::vector<double> t[3];
  const double* src[3] = {pct_sand, pct_clay, organic};
  for (int k = 0; k < 3; ++k) {
    t[k].resize((size_t)NLEVGRND * c.ncols);
    for (int64_t i = 0; i < c.ncols; ++i)
      for (int l = 0; l < NLEVGRND; ++l) t[k][(size_t)l * c.ncols + i] = src[k][(size_t)i * NLEVGRND + l];
  }
  const InitInputs X{t[0].data(), t[1].data(), t[2].data(), snow_depth, (long long)c.ncols, organic_max};
  for_columns(c, [&](int i) { column_init(c.cols, c.tab, X, i); });
  return ELMK_OK;
}

int elmk_atm_series(elmk_handle h, int var, const double* host, int ntimes) {
  PortCtx& c = *ctx(h);
  if (var < 0 || var >= ATM_NVARS || !host || ntimes < 2) return ELMK_EINVAL;
  c.atm[var].assign(host, host + (size_t)ntimes * c.ncols);
  return ELMK_OK;
}
int elmk_atm_series_row(elmk_handle h, int var, int t, const double* host) {
  PortCtx& c = *ctx(h);
  if (var < 0 || var >= ATM_NVARS || !host || t < 0 || (size_t)(t + 1) * c.ncols > c.atm[var].size()) return ELMK_EINVAL;
  std::copy(host, host + c.ncols, c.atm[var].begin() + (size_t)t * c.ncols);
  return ELMK_OK;
}
int elmk_atm_forcing(elmk_handle h, int t_idx, double wt1, double wt2, int qbot_is_rh) {
  PortCtx& c = *ctx(h);
  AtmSeries A;
  A.stride = c.ncols;
  for (int v = 0; v < ATM_NVARS; ++v) {
    if (t_idx < 0 || (size_t)(t_idx + 2) * c.ncols > c.atm[v].size()) return ELMK_EINVAL;
    A.v[v] = c.atm[v].data();
  }
  for_columns(c, [&](int i) { column_atm_forcing(c.cols, A, t_idx, wt1, wt2, qbot_is_rh != 0, i); });
  return ELMK_OK;
}
int elmk_set_gas_pressures(elmk_handle h, const double* forc_pco2, const double* forc_po2) {
  PortCtx& c = *ctx(h);
  if ((forc_pco2 == nullptr) != (forc_po2 == nullptr)) return ELMK_EINVAL;
  if (!forc_pco2) {
    c.cols.pco2_in = nullptr;
    c.cols.po2_in = nullptr;
    return ELMK_OK;
  }
  c.gas[0].assign(forc_pco2, forc_pco2 + c.ncols);
  c.gas[1].assign(forc_po2, forc_po2 + c.ncols);
  c.cols.pco2_in = c.gas[0].data();
  c.cols.po2_in = c.gas[1].data();
  return ELMK_OK;
}
int elmk_set_coordinates(elmk_handle h, const double* lat_r, const double* lon_r, int64_t n) {
  PortCtx& c = *ctx(h);
  if (!lat_r || !lon_r || !(n == 1 || n == c.ncols)) return ELMK_EINVAL;
  c.coords.resize((size_t)4 * n);
  for (int64_t i = 0; i < n; ++i) {
    c.coords[i] = std::sin(lat_r[i]);
    c.coords[n + i] = std::cos(lat_r[i]);
    c.coords[2 * n + i] = std::tan(solar::ensure_tan_defined(lat_r[i]));
    c.coords[3 * n + i] = lon_r[i];
  }
  c.ncoords = n;
  c.lat0 = lat_r[0];
  return ELMK_OK;
}
int elmk_solar_step(elmk_handle h, double dtime, double decday, int doy1, double* dayl, double* max_dayl) {
  PortCtx& c = *ctx(h);
  if (c.coords.empty()) return ELMK_EINVAL;
  const double declin = solar::declination((int)decday);
  SolarStep G;
  G.sin_lat = c.coords.data(); G.cos_lat = G.sin_lat + c.ncoords; G.tan_lat = G.sin_lat + 2 * c.ncoords; G.lon = G.sin_lat + 3 * c.ncoords;
  G.per_column = c.ncoords > 1;
  G.dtrad = dtime * solar::TWO_PI / 86400.0;
  G.frac2pi = (decday - std::floor(decday)) * solar::TWO_PI;
  G.sin_decl = std::sin(declin);
  G.cos_decl = std::cos(declin);
  G.tan_decl = std::tan(solar::ensure_tan_defined(declin));
  for_columns(c, [&](int i) { column_coszen(c.cols, G, i); });
  if (dayl) *dayl = solar::daylength(c.lat0, solar::declination(doy1));
  if (max_dayl) *max_dayl = solar::max_daylength(c.lat0);
  return ELMK_OK;
}
int elmk_phen_series(elmk_handle h, int var, const double* host, int nmonths) {
  PortCtx& c = *ctx(h);
  if (var < 0 || var >= PHEN_NVARS || !host || nmonths < 2) return ELMK_EINVAL;
  c.phen[var].assign(host, host + (size_t)nmonths * c.ncols);
  return ELMK_OK;
}
int elmk_phenology(elmk_handle h, int start_idx, double wt1, double wt2) {
  PortCtx& c = *ctx(h);
  PhenSeries P;
  P.stride = c.ncols;
  for (int v = 0; v < PHEN_NVARS; ++v) {
    if (start_idx < 0 || (size_t)(start_idx + 2) * c.ncols > c.phen[v].size()) return ELMK_EINVAL;
    P.v[v] = c.phen[v].data();
  }
  for_columns(c, [&](int i) { column_phenology(c.cols, P, start_idx, wt1, wt2, i); });
  return ELMK_OK;
}

int elmk_init_timestep(elmk_handle h, int reset_forc_hgt) {
  PortCtx& c = *ctx(h);
  for_columns(c, [&](int i) { column_init_timestep(c.cols, c.tab, reset_forc_hgt, i); });
  return ELMK_OK;
}

int elmk_step(elmk_handle h, double dtime, double dayl, double max_dayl, uint32_t mask) {
  PortCtx& c = *ctx(h);
  if (!c.tables_set) return ELMK_ENOTABLES;
  const Cols& S = c.cols;
  const Tables& T = c.tab;
  StepArgs A{dtime, dayl, max_dayl};
  if (mask & ELMK_G_FRAC_WET) for_columns(c, [&](int i) { column_frac_wet(S, T, i); });
  if (mask & ELMK_G_ALBEDO) for_columns(c, [&](int i) { column_albedo(S, T, i); });
  if (mask & ELMK_G_CANOPY_HYDROLOGY) for_columns(c, [&](int i) { column_canopy_hydrology(S, T, dtime, i); });
  if (mask & ELMK_G_SURFACE_RADIATION) for_columns(c, [&](int i) { column_surface_radiation(S, T, i); });
  if (mask & ELMK_G_CANOPY_TEMPERATURE) for_columns(c, [&](int i) { column_canopy_temperature(S, T, i); });
  if (mask & ELMK_G_BAREGROUND_FLUXES) for_columns(c, [&](int i) { column_bareground_fluxes(S, T, i); });
  if (mask & ELMK_G_CANOPY_FLUXES) for_columns(c, [&](int i) { column_canopy_fluxes(S, T, A, i); });
  if (mask & ELMK_G_SOIL_TEMPERATURE) for_columns(c, [&](int i) { column_soil_temperature(S, T, dtime, i); });
  if (mask & ELMK_G_SNOW_HYDROLOGY) for_columns(c, [&](int i) { column_snow_hydrology(S, T, dtime, i); });
  if (mask & ELMK_G_SURFACE_FLUXES) for_columns(c, [&](int i) { column_surface_fluxes(S, T, dtime, i); });
  if (mask & ELMK_G_CONSERVATION) for_columns(c, [&](int i) { column_conservation(S, T, dtime, i); });
  return ELMK_OK;
}
int elmk_sync(elmk_handle) { return ELMK_OK; }
int64_t elmk_launch_count(elmk_handle h) { return ctx(h)->launches; }

int elmk_errors(elmk_handle h, uint32_t* any, int64_t* first_col) {
  PortCtx& c = *ctx(h);
  uint32_t acc = 0;
  int64_t first = -1;
  for (int64_t i = 0; i < c.ncols; ++i) {
    const uint32_t w = static_cast<uint32_t>(c.cols.errmask[i]);
    if (w && first < 0) first = i;
    acc |= w;
  }
  if (any) *any = acc;
  if (first_col) *first_col = first;
  return ELMK_OK;
}
int elmk_clear_errors(elmk_handle h) {
  PortCtx& c = *ctx(h);
  std::fill_n(c.cols.errmask, c.np, 0);
  return ELMK_OK;
}
const char* elmk_error_text(uint32_t) { return "see include/elmk_b200.h"; }

int elmk_diag_reduce(elmk_handle h, double out[24]) {
  PortCtx& c = *ctx(h);
  const double* d[8] = {c.cols.dtend_column_h2o, c.cols.errh2o, c.cols.errh2osno, c.cols.dwb,
                        c.cols.errsol, c.cols.errlon, c.cols.errseb, c.cols.netrad};
  for (int k = 0; k < 8; ++k) {
    double s = 0.0, lo = d[k][0], hi = d[k][0];
    for (int64_t i = 0; i < c.ncols; ++i) {
      s += d[k][i];
      lo = std::min(lo, d[k][i]);
      hi = std::max(hi, d[k][i]);
    }
    out[k] = s; out[8 + k] = lo; out[16 + k] = hi;
  }
  return ELMK_OK;
}
int elmk_device_ptr(elmk_handle, int, void**, int64_t*) { return ELMK_EUNSUPPORTED; }
int elmk_stream(elmk_handle, void** stream) { if (stream) *stream = nullptr; return ELMK_OK; }
int elmk_timing_enable(elmk_handle, int) { return ELMK_OK; }
int elmk_timing_read(elmk_handle, int, const char**, double*, int64_t*, uint32_t*) { return 0; }

#include "../exchange_host.h"

} // extern "C"
