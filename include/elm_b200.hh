// elm_b200.hh - host-side C++ adaptor: the reference's kernel-group wrapper API on top of the C ABI.
//
// A driver written against the reference calls, per timestep (driver/kokkos/elm_kokkos_interface.cc:289-318),
//   ELM::kokkos_frac_wet(S); ELM::kokkos_albedo_snicar(S); ELM::kokkos_canopy_hydrology(S, dt); ...
// on one mutable `ELMStateType& S` (src/data/elm_state.h:53-225).  This header provides the same eleven
// functions, with the same names, argument order and meaning, in namespace ELM::b200, templated on the
// state type and reaching the device through include/elmk_b200.h only.  Swapping backends is
//   #include "elm_b200.hh"      and      namespace ELMX = ELM::b200;   (instead of ELM)
// plus linking libelmk_b200.so; nothing in the driver's data structures changes, because the adaptor reads
// and writes `S` exclusively through the reference's own element accessors - `S.field(i)`,
// `S.field(i, lev)` - i.e. the templated ArrayType / index-map interface (Kokkos::View or ELM::Array alike).
//
// Two ways to use it:
//   * drop-in, call by call: each ELM::b200::kokkos_<group>(S[, dt]) uploads the per-column arrays from S,
//     runs that one group on the device and downloads the arrays back into S.  Semantically identical to
//     the reference wrapper, convenient for bringing a driver up; the PCIe traffic makes it slow.
//   * resident: ELM::b200::Device keeps the column state in HBM across steps; the driver uploads only the
//     per-step forcing (upload_forcing), calls advance(), and downloads what it wants to look at
//     (download_primary = the PrimaryVars set of elm_state.h:17-48, download_diagnostics, download).
//
// Error convention: the reference throws std::runtime_error from inside kernels
// (e.g. photosynthesis_impl.hh:232,289,439; snow_snicar_impl.hh:76,618,658,664); the device records one bit
// per throw site per column, and check_errors() rethrows std::runtime_error with the reference's text.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "elmk_b200.h"
#include "elmk_members.h"

namespace ELM {
namespace b200 {

class Device {
public:
  Device(int64_t ncols, int device = 0) : ncols_(ncols) {
    if (elmk_create(&h_, device, ncols) != ELMK_OK) throw std::runtime_error("ELM::b200: elmk_create failed (no CUDA device?)");
  }
  ~Device() { if (h_) elmk_destroy(h_); }
  Device(const Device&) = delete;
  Device& operator=(const Device&) = delete;

  elmk_handle handle() const { return h_; }
  int64_t ncols() const { return ncols_; }

  // ---- tables: PFT constants, SNICAR optics, snow-age fit, soil-colour albedos, LandType, dewmx, oldfflag.
  //      Replaces the table part of initialize_kokkos_elm (initialize_elm_kokkos.cc:267-366). ----
  template <class State> void set_tables(const State& S) {
    elmk_tables t{};
    t.ltype = S.Land.ltype; t.ctype = S.Land.ctype; t.vtype = S.Land.vtype;
    t.urbpoi = S.Land.urbpoi ? 1 : 0; t.lakpoi = S.Land.lakpoi ? 1 : 0;
    t.oldfflag = S.oldfflag; t.dewmx = S.dewmx;
    std::vector<std::vector<double>> keep;
    auto flat = [&keep](const auto& v) -> const double* {
      std::vector<double> a(v.size());
      for (size_t i = 0; i < a.size(); ++i) a[i] = v.data()[i];
      keep.push_back(std::move(a));
      return keep.back().data();
    };
    const auto& p = *S.pft_data;
    const double* pft[ELMK_NPFT_TABLES] = {
        flat(p.fnr), flat(p.act25), flat(p.kcha), flat(p.koha), flat(p.cpha), flat(p.vcmaxha), flat(p.jmaxha),
        flat(p.tpuha), flat(p.lmrha), flat(p.vcmaxhd), flat(p.jmaxhd), flat(p.tpuhd), flat(p.lmrhd), flat(p.lmrse),
        flat(p.qe), flat(p.theta_cj), flat(p.bbbopt), flat(p.mbbopt), flat(p.c3psn), flat(p.slatop), flat(p.leafcn),
        flat(p.flnr), flat(p.fnitr), flat(p.dleaf), flat(p.smpso), flat(p.smpsc), flat(p.tc_stress), flat(p.z0mr),
        flat(p.displar), flat(p.xl), flat(p.roota_par), flat(p.rootb_par), flat(p.rholvis), flat(p.rholnir),
        flat(p.rhosvis), flat(p.rhosnir), flat(p.taulvis), flat(p.taulnir), flat(p.tausvis), flat(p.tausnir)};
    for (int k = 0; k < ELMK_NPFT_TABLES; ++k) t.pft[k] = pft[k];
    t.albsat = flat(S.albsat);
    t.albdry = flat(S.albdry);
    const auto& s = *S.snicar_data;
    const double* band[18] = {flat(s.ss_alb_oc1), flat(s.asm_prm_oc1), flat(s.ext_cff_mss_oc1), flat(s.ss_alb_oc2),
        flat(s.asm_prm_oc2), flat(s.ext_cff_mss_oc2), flat(s.ss_alb_dst1), flat(s.asm_prm_dst1), flat(s.ext_cff_mss_dst1),
        flat(s.ss_alb_dst2), flat(s.asm_prm_dst2), flat(s.ext_cff_mss_dst2), flat(s.ss_alb_dst3), flat(s.asm_prm_dst3),
        flat(s.ext_cff_mss_dst3), flat(s.ss_alb_dst4), flat(s.asm_prm_dst4), flat(s.ext_cff_mss_dst4)};
    for (int k = 0; k < 18; ++k) t.snicar_band[k] = band[k];
    const double* snow[6] = {flat(s.ss_alb_snw_drc), flat(s.asm_prm_snw_drc), flat(s.ext_cff_mss_snw_drc),
        flat(s.ss_alb_snw_dfs), flat(s.asm_prm_snw_dfs), flat(s.ext_cff_mss_snw_dfs)};
    for (int k = 0; k < 6; ++k) t.snicar_snow[k] = snow[k];
    const double* bc[6] = {flat(s.ss_alb_bc1), flat(s.asm_prm_bc1), flat(s.ext_cff_mss_bc1), flat(s.ss_alb_bc2),
        flat(s.asm_prm_bc2), flat(s.ext_cff_mss_bc2)};
    for (int k = 0; k < 6; ++k) t.snicar_bc[k] = bc[k];
    t.bcenh = flat(s.bcenh);
    const auto& a = *S.snw_rds_table;
    t.snowage[0] = flat(a.snowage_tau);
    t.snowage[1] = flat(a.snowage_kappa);
    t.snowage[2] = flat(a.snowage_drdt0);
    check(elmk_set_tables(h_, &t), "elmk_set_tables");
    dayl_ = S.dayl;
    max_dayl_ = S.max_dayl;
  }

  // ---- state movement through the element accessors of S ----
  // every per-column array of S -> device
  template <class State> void upload(const State& S) {
#define X(name, member, nlev) put(#name, member, nlev);
    ELMK_STATE_MEMBERS(X)
    ELMK_AEROSOL_MEMBERS(X)
#undef X
    put_psn(S);
    dayl_ = S.dayl;
    max_dayl_ = S.max_dayl;
  }
  // every per-column array of the device -> S
  template <class State> void download(State& S) {
#define X(name, member, nlev) get(#name, member, nlev);
    ELMK_STATE_MEMBERS(X)
    ELMK_AEROSOL_MEMBERS(X)
#undef X
  }
  // the per-step inputs the caller refreshes before the chain: forcing, phenology, coszen
  template <class State> void upload_forcing(const State& S) {
    put("coszen", S.coszen, 1);
    put("forc_tbot", S.forc_tbot, 1); put("forc_thbot", S.forc_thbot, 1); put("forc_pbot", S.forc_pbot, 1);
    put("forc_qbot", S.forc_qbot, 1); put("forc_lwrad", S.forc_lwrad, 1); put("forc_u", S.forc_u, 1);
    put("forc_v", S.forc_v, 1); put("forc_rain", S.forc_rain, 1); put("forc_snow", S.forc_snow, 1);
    put("forc_hgt", S.forc_hgt, 1);
    put("forc_solad", S.forc_solad, 2); put("forc_solai", S.forc_solai, 2);
    put("elai", S.elai, 1); put("esai", S.esai, 1); put("tlai", S.tlai, 1); put("tsai", S.tsai, 1);
    put("htop", S.htop, 1); put("hbot", S.hbot, 1); put("frac_veg_nosno_alb", S.frac_veg_nosno_alb, 1);
    dayl_ = S.dayl;
    max_dayl_ = S.max_dayl;
  }
  // the PrimaryVars set (elm_state.h:17-48; ELMInterface::copyPrimaryVars, elm_kokkos_interface.cc:324-347)
  template <class Primary> void download_primary(Primary& S) {
    get("snl", S.snl, 1); get("snow_depth", S.snow_depth, 1); get("frac_sno", S.frac_sno, 1);
    get("int_snow", S.int_snow, 1); get("snw_rds", S.snw_rds, 5);
    get("h2osoi_liq", S.h2osoi_liq, 20); get("h2osoi_ice", S.h2osoi_ice, 20); get("h2osoi_vol", S.h2osoi_vol, 15);
    get("h2ocan", S.h2ocan, 1); get("h2osno", S.h2osno, 1); get("h2osfc", S.h2osfc, 1);
    get("t_soisno", S.t_soisno, 20); get("t_grnd", S.t_grnd, 1); get("t_h2osfc", S.t_h2osfc, 1);
    get("t_h2osfc_bef", S.t_h2osfc_bef, 1); get("nrad", S.nrad, 1);
    get("dz", S.dz, 20); get("zsoi", S.zsoi, 20); get("zisoi", S.zisoi, 21);
  }
  // the eight balance diagnostics of kokkos_evaluate_conservation (conserved_quantity_kokkos.cc:13-20), which
  // the reference keeps in wrapper-local Views: out[k] has ncols entries, k in the order of ELMK_DIAGNOSTIC_FIELDS
  std::vector<std::vector<double>> download_diagnostics() {
    std::vector<std::vector<double>> out;
    static const char* names[8] = {"dtend_column_h2o", "errh2o", "errh2osno", "dwb", "errsol", "errlon", "errseb", "netrad"};
    for (const char* n : names) {
      out.emplace_back(static_cast<size_t>(ncols_));
      check(elmk_download(h_, field(n), out.back().data(), 0, ncols_, ELMK_COL_OUTER), n);
    }
    return out;
  }

  // ---- one-time cold start of every column: the per-column lambda of initialize_kokkos_elm
  //      (initialize_elm_kokkos.cc:374-431).  Texture arrays are read as a(col, lev), snow depth as a(col). ----
  template <class Arr2, class Arr1>
  void init_columns(const Arr2& pct_sand, const Arr2& pct_clay, const Arr2& organic, double organic_max, const Arr1& snow_depth) {
    auto rows = [this](const Arr2& a) {
      std::vector<double> b(static_cast<size_t>(ncols_) * ELMK_NLEVGRND);
      for (int64_t i = 0; i < ncols_; ++i)
        for (int l = 0; l < ELMK_NLEVGRND; ++l) b[static_cast<size_t>(i) * ELMK_NLEVGRND + l] = a(static_cast<int>(i), l);
      return b;
    };
    const std::vector<double> s = rows(pct_sand), c = rows(pct_clay), o = rows(organic);
    std::vector<double> d(static_cast<size_t>(ncols_));
    for (int64_t i = 0; i < ncols_; ++i) d[static_cast<size_t>(i)] = snow_depth(static_cast<int>(i));
    check(elmk_init_columns(h_, s.data(), c.data(), o.data(), organic_max, d.data()), "elmk_init_columns");
  }

  // ---- per-step input producers on the device (ELM::get_forcing, atm_forcing_kokkos.cc:48-63; ComputePhenology via
  //      PhenologyDataManager::get_data, phenology_data_impl.hh:46-63).  `data` is the manager's (ntimes, ncells)
  //      array - AtmDataManager::data, PhenologyDataManager::mlai/msai/mhtop/mhbot - read through operator()(t, i). ----
  template <class Arr2> void set_atm_series(int var, const Arr2& data, int ntimes) {
    const std::vector<double> b = flatten2(data, ntimes);
    check(elmk_atm_series(h_, var, b.data(), ntimes), "elmk_atm_series");
  }
  template <class Arr2> void set_phen_series(int var, const Arr2& data, int nmonths) {
    const std::vector<double> b = flatten2(data, nmonths);
    check(elmk_phen_series(h_, var, b.data(), nmonths), "elmk_phen_series");
  }
  // t_idx, wt1, wt2: AtmDataManager::forc_t_idx_check_bounds / forcing_time_weights for the centred step time
  void atm_forcing(int t_idx, double wt1, double wt2, bool qbot_is_rh = true) {
    check(elmk_atm_forcing(h_, t_idx, wt1, wt2, qbot_is_rh ? 1 : 0), "elmk_atm_forcing");
  }
  // start_idx, wt1, wt2: monthly_data::first_month_idx / monthly_data_weights
  void phenology(int start_idx, double wt1, double wt2) { check(elmk_phenology(h_, start_idx, wt1, wt2), "elmk_phenology"); }

  // solar geometry of kokkos_init_timestep (init_timestep_kokkos.cc:27-35): coordinates once (n == 1: the reference's
  // single site for all columns), then per step average_cosz of every column on the device; the day lengths it
  // returns are kept for step().  decday = Utils::decimal_doy(current) + 1, doy1 = current.doy + 1.
  void set_coordinates(const double* lat_r, const double* lon_r, int64_t n = 1) {
    check(elmk_set_coordinates(h_, lat_r, lon_r, n), "elmk_set_coordinates");
  }
  void solar_step(double dtime, double decday, int doy1) {
    check(elmk_solar_step(h_, dtime, decday, doy1, &dayl_, &max_dayl_), "elmk_solar_step");
  }

  // CO2 / O2 partial pressures [Pa] per column for the photosynthesis of kokkos_canopy_fluxes, which the reference
  // derives from constants (canopy_fluxes_kokkos.cc:49-51); nullptr, nullptr returns to those
  void set_gas_pressures(const double* forc_pco2, const double* forc_po2) {
    check(elmk_set_gas_pressures(h_, forc_pco2, forc_po2), "elmk_set_gas_pressures");
  }

  // ---- stepping ----
  // per-column part of kokkos_init_timestep (init_timestep_kokkos.cc:53-72) incl. the reset of forc_hgt_*_patch
  void init_timestep(bool reset_forc_hgt = true) { check(elmk_init_timestep(h_, reset_forc_hgt ? 1 : 0), "elmk_init_timestep"); }
  void step(double dtime, uint32_t groups = ELMK_G_ALL) { check(elmk_step(h_, dtime, dayl_, max_dayl_, groups), "elmk_step"); }
  // ELMInterface::advance without the file IO: bookkeeping + the eleven groups; returns false like the reference
  bool advance(double dtime) {
    init_timestep(true);
    step(dtime, ELMK_G_ALL);
    return false;
  }
  void sync() { check(elmk_sync(h_), "elmk_sync"); }
  void set_daylength(double dayl, double max_dayl) { dayl_ = dayl; max_dayl_ = max_dayl; }

  // rethrows the reference's exception for the first column that hit one of its throw/assert sites
  void check_errors() {
    uint32_t any = 0;
    int64_t first = -1;
    check(elmk_errors(h_, &any, &first), "elmk_errors");
    if (!any) return;
    for (uint32_t bit = 1; bit; bit <<= 1)
      if (any & bit) throw std::runtime_error(std::string(elmk_error_text(bit)) + " (first column " + std::to_string(first) + ")");
  }

private:
  int field(const char* name) {
    const int f = elmk_field_id(name);
    if (f < 0) throw std::runtime_error(std::string("ELM::b200: unknown field ") + name);
    return f;
  }
  void check(int rc, const char* what) {
    if (rc != ELMK_OK) throw std::runtime_error(std::string("ELM::b200: ") + what + " failed: " + elmk_last_error(h_));
  }
  template <class Arr2> std::vector<double> flatten2(const Arr2& a, int n0) {
    std::vector<double> b(static_cast<size_t>(n0) * ncols_);
    for (int t = 0; t < n0; ++t)
      for (int64_t i = 0; i < ncols_; ++i) b[static_cast<size_t>(t) * ncols_ + i] = a(t, static_cast<int>(i));
    return b;
  }
  // gather one array of S through its element accessor into a dense host buffer in the reference's
  // layout (column outer) and hand it to the ABI; works for any ArrayType with operator()(i[, lev])
  template <class Arr> void put(const char* name, const Arr& a, int nlev) {
    int dtype = 0;
    const int f = field(name);
    elmk_field_info(f, nullptr, &dtype, nullptr);
    const size_t n = static_cast<size_t>(ncols_);
    if (dtype == ELMK_F64) { std::vector<double> b(n * nlev); fill(b, a, nlev); check(elmk_upload(h_, f, b.data(), 0, ncols_, ELMK_COL_OUTER), name); }
    else if (dtype == ELMK_I32) { std::vector<int32_t> b(n * nlev); fill(b, a, nlev); check(elmk_upload(h_, f, b.data(), 0, ncols_, ELMK_COL_OUTER), name); }
    else { std::vector<uint8_t> b(n * nlev); fill(b, a, nlev); check(elmk_upload(h_, f, b.data(), 0, ncols_, ELMK_COL_OUTER), name); }
  }
  template <class Arr> void get(const char* name, Arr& a, int nlev) {
    int dtype = 0;
    const int f = field(name);
    elmk_field_info(f, nullptr, &dtype, nullptr);
    const size_t n = static_cast<size_t>(ncols_);
    if (dtype == ELMK_F64) { std::vector<double> b(n * nlev); check(elmk_download(h_, f, b.data(), 0, ncols_, ELMK_COL_OUTER), name); drain(b, a, nlev); }
    else if (dtype == ELMK_I32) { std::vector<int32_t> b(n * nlev); check(elmk_download(h_, f, b.data(), 0, ncols_, ELMK_COL_OUTER), name); drain(b, a, nlev); }
    else { std::vector<uint8_t> b(n * nlev); check(elmk_download(h_, f, b.data(), 0, ncols_, ELMK_COL_OUTER), name); drain(b, a, nlev); }
  }
  template <class T, class Arr> void fill(std::vector<T>& b, const Arr& a, int nlev) {
    for (int64_t i = 0; i < ncols_; ++i) {
      if constexpr (Arr::rank == 1) b[i] = static_cast<T>(a(i));
      else for (int l = 0; l < nlev; ++l) b[i * nlev + l] = static_cast<T>(a(i, l));
    }
  }
  template <class T, class Arr> void drain(const std::vector<T>& b, Arr& a, int nlev) {
    using V = typename Arr::value_type;
    for (int64_t i = 0; i < ncols_; ++i) {
      if constexpr (Arr::rank == 1) a(i) = static_cast<V>(b[i]);
      else for (int l = 0; l < nlev; ++l) a(i, l) = static_cast<V>(b[i * nlev + l]);
    }
  }
  // psn_pft(i) is a struct of 27 doubles (PFTDataPSN, pft_data.h:20-24)
  template <class State> void put_psn(const State& S) {
    std::vector<double> b(static_cast<size_t>(ncols_) * 27);
    for (int64_t i = 0; i < ncols_; ++i) {
      const double* p = reinterpret_cast<const double*>(&S.psn_pft(i));
      for (int k = 0; k < 27; ++k) b[i * 27 + k] = p[k];
    }
    check(elmk_upload(h_, field("psn_pft"), b.data(), 0, ncols_, ELMK_COL_OUTER), "psn_pft");
  }

  elmk_handle h_ = nullptr;
  int64_t ncols_;
  double dayl_ = 0.0, max_dayl_ = 1.0;
};

// ---- drop-in wrappers: same names and arguments as driver/kokkos/<group>_kokkos.hh -------------------
namespace detail {
struct Mirror {
  const void* state;
  int64_t ncols;
  int ltype, ctype, vtype, oldfflag;
  bool urbpoi, lakpoi;
  double dewmx;
  std::unique_ptr<Device> dev;
};
inline std::vector<Mirror>& mirrors() {
  static std::vector<Mirror> cache;
  return cache;
}
template <class State> bool same_scalars(const Mirror& m, const State& S) {
  return m.ltype == S.Land.ltype && m.ctype == S.Land.ctype && m.vtype == S.Land.vtype && m.urbpoi == (bool)S.Land.urbpoi &&
         m.lakpoi == (bool)S.Land.lakpoi && m.oldfflag == S.oldfflag && m.dewmx == S.dewmx;
}
// One device mirror per state object, created on first use.  The entry is keyed on the object's address AND its
// column count (a state destroyed and another allocated at the same address does not inherit the mirror), the tables
// are sent again whenever the scalars of the land unit change, and ELM::b200::release(S) / release_all() free the
// device memory; refresh_tables(S) re-sends the PFT / SNICAR tables after the caller changed them in place.
template <class State> Device& device_for(State& S) {
  auto& cache = mirrors();
  for (size_t k = 0; k < cache.size(); ++k) {
    Mirror& m = cache[k];
    if (m.state != static_cast<const void*>(&S)) continue;
    if (m.ncols != (int64_t)S.num_columns) {   // stale: another object lived at this address
      cache.erase(cache.begin() + k);
      break;
    }
    if (!same_scalars(m, S)) {
      m.dev->set_tables(S);
      m.ltype = S.Land.ltype; m.ctype = S.Land.ctype; m.vtype = S.Land.vtype; m.urbpoi = S.Land.urbpoi; m.lakpoi = S.Land.lakpoi;
      m.oldfflag = S.oldfflag; m.dewmx = S.dewmx;
    }
    return *m.dev;
  }
  Mirror m{static_cast<const void*>(&S), (int64_t)S.num_columns, S.Land.ltype, S.Land.ctype, S.Land.vtype, S.oldfflag,
           (bool)S.Land.urbpoi, (bool)S.Land.lakpoi, S.dewmx, std::make_unique<Device>(S.num_columns)};
  m.dev->set_tables(S);
  cache.push_back(std::move(m));
  return *cache.back().dev;
}
template <class State> void run_group(State& S, double dtime, uint32_t group) {
  Device& d = device_for(S);
  d.upload(S);
  d.step(dtime, group);
  d.download(S);
  d.check_errors();
}
} // namespace detail

// frees the device mirror of S (the drop-in wrappers create one on first use and keep it)
template <class State> void release(State& S) {
  auto& cache = detail::mirrors();
  for (size_t k = 0; k < cache.size(); ++k)
    if (cache[k].state == static_cast<const void*>(&S)) { cache.erase(cache.begin() + k); return; }
}
inline void release_all() { detail::mirrors().clear(); }
// after the caller changed PFT / SNICAR / snow-age tables of S in place
template <class State> void refresh_tables(State& S) { detail::device_for(S).set_tables(S); }

template <class State> void kokkos_frac_wet(State& S) { detail::run_group(S, 0.0, ELMK_G_FRAC_WET); }
template <class State> void kokkos_albedo_snicar(State& S) { detail::run_group(S, 0.0, ELMK_G_ALBEDO); }
template <class State> void kokkos_canopy_hydrology(State& S, const double& dtime) { detail::run_group(S, dtime, ELMK_G_CANOPY_HYDROLOGY); }
template <class State> void kokkos_surface_radiation(State& S) { detail::run_group(S, 0.0, ELMK_G_SURFACE_RADIATION); }
template <class State> void kokkos_canopy_temperature(State& S) { detail::run_group(S, 0.0, ELMK_G_CANOPY_TEMPERATURE); }
template <class State> void kokkos_bareground_fluxes(State& S) { detail::run_group(S, 0.0, ELMK_G_BAREGROUND_FLUXES); }
template <class State> void kokkos_canopy_fluxes(State& S, const double& dtime) { detail::run_group(S, dtime, ELMK_G_CANOPY_FLUXES); }
template <class State> void kokkos_soil_temperature(State& S, const double& dtime) { detail::run_group(S, dtime, ELMK_G_SOIL_TEMPERATURE); }
template <class State, class Date> void kokkos_snow_hydrology(State& S, const double& dtime, const Date&) { detail::run_group(S, dtime, ELMK_G_SNOW_HYDROLOGY); }
template <class State> void kokkos_surface_fluxes(State& S, const double& dtime) { detail::run_group(S, dtime, ELMK_G_SURFACE_FLUXES); }
template <class State> void kokkos_evaluate_conservation(State& S, const double& dtime) { detail::run_group(S, dtime, ELMK_G_CONSERVATION); }
// the per-column lambda of kokkos_init_timestep (init_timestep_kokkos.cc:53-72); its file reading (phenology, forcing)
// stays with the driver
template <class State> void kokkos_init_timestep_columns(State& S) {
  Device& d = detail::device_for(S);
  d.upload(S);
  d.init_timestep(true);
  d.download(S);
}

} // namespace b200

// With ELM_B200_DROP_IN defined before this header is included, the eleven wrappers are ELM::kokkos_<group> themselves:
// a driver written against the reference (driver/kokkos/elm_kokkos_interface.cc:289-318) switches backend by including
// this header in place of the eleven driver/kokkos/<group>_kokkos.hh and linking libelmk_b200.so - no call site changes.
#ifdef ELM_B200_DROP_IN
using b200::kokkos_frac_wet;
using b200::kokkos_albedo_snicar;
using b200::kokkos_canopy_hydrology;
using b200::kokkos_surface_radiation;
using b200::kokkos_canopy_temperature;
using b200::kokkos_bareground_fluxes;
using b200::kokkos_canopy_fluxes;
using b200::kokkos_soil_temperature;
using b200::kokkos_snow_hydrology;
using b200::kokkos_surface_fluxes;
using b200::kokkos_evaluate_conservation;
#endif
} // namespace ELM
