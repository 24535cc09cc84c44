// The C++ host adaptor (include/elm_b200.hh) computing through a library that exports the elmk C ABI, on the
// REFERENCE's own state type: built where /root/reference is mounted (tests/adaptor/build_adaptor.py; reference
// headers used where they lie, host Kokkos stand-in of oracle/shim), linked against libelmk_b200.so, shipped to the
// GPU box with the snapshot and run there by tests/test_gpu_adaptor.py, which compares its output with the checker.
//
//   adaptor_steps <dir> <ncols> <nsteps>
//     <dir>/tables.bin    the members of struct elmk_tables as doubles (order below; every array preceded by its length)
//     <dir>/state.bin     every field of include/elmk_fields.def in order, reference host layout, native element types
//     <dir>/forcing_<k>.bin  the per-step inputs of upload_forcing, doubles (frac_veg_nosno_alb as doubles too)
//     -> <dir>/out.bin    every ELMState member of ELMK_STATE_MEMBERS + aerosol members after the steps, as doubles,
//                         read out of the reference state object through its own element accessors
// Path exercised: blob -> C ABI -> Device::download(S) -> ELMStateType S -> Device::set_tables(S), Device::upload(S)
// (a second device mirror) -> per step upload_forcing(S) + advance(dt) -> download(S); then, with ELM_B200_DROP_IN,
// one more step as eleven ELM::kokkos_<group>(S[, dt]) calls exactly as ELMInterface::advance writes them
// (elm_kokkos_interface.cc:289-318).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "compile_options.hh"
#include "data_types.hh"
#include "elm_constants.h"
#include "utils.hh"
#include "date_time.hh"

#define ELM_B200_DROP_IN
#include "elm_b200.hh"

static std::vector<char> slurp(const std::string& p) {
  std::ifstream f(p, std::ios::binary);
  if (!f) { std::fprintf(stderr, "cannot read %s\n", p.c_str()); std::exit(2); }
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
// one table of the blob: its length, then its values (the reference's views may be shorter: 17 of 25 PFTs)
template <class V> static const double* fill_view(V& v, const double* src) {
  const size_t len = static_cast<size_t>(*src++);
  if (len < v.size()) { std::fprintf(stderr, "table blob: %zu values for a view of %zu\n", len, v.size()); std::exit(2); }
  for (size_t i = 0; i < v.size(); ++i) v.data()[i] = src[i];
  return src + len;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const std::string dir = argv[1];
  const int n = std::atoi(argv[2]), nsteps = std::atoi(argv[3]);
  auto dd = ELM::Utils::create_domain_decomposition_2D(ELM::Utils::square_numprocs(1), {1, 1}, {0, 0});
  ELMStateType S(n, dd, std::string(), ELM::Utils::Date(1985, 1, 1), 1);

  // ---- tables into the reference's data managers (as initialize_kokkos_elm fills them from files) ----
  {
    const std::vector<char> blob = slurp(dir + "/tables.bin");
    const double* p = reinterpret_cast<const double*>(blob.data());
    S.Land.ltype = (int)*p++; S.Land.ctype = (int)*p++; S.Land.vtype = (int)*p++;
    S.Land.urbpoi = *p++ != 0.0; S.Land.lakpoi = *p++ != 0.0;
    S.oldfflag = (int)*p++; S.dewmx = *p++; S.dayl = *p++; S.max_dayl = *p++;
    auto& t = *S.pft_data;
    ViewD1* pft[ELMK_NPFT_TABLES] = {&t.fnr, &t.act25, &t.kcha, &t.koha, &t.cpha, &t.vcmaxha, &t.jmaxha, &t.tpuha, &t.lmrha,
        &t.vcmaxhd, &t.jmaxhd, &t.tpuhd, &t.lmrhd, &t.lmrse, &t.qe, &t.theta_cj, &t.bbbopt, &t.mbbopt, &t.c3psn, &t.slatop,
        &t.leafcn, &t.flnr, &t.fnitr, &t.dleaf, &t.smpso, &t.smpsc, &t.tc_stress, &t.z0mr, &t.displar, &t.xl, &t.roota_par,
        &t.rootb_par, &t.rholvis, &t.rholnir, &t.rhosvis, &t.rhosnir, &t.taulvis, &t.taulnir, &t.tausvis, &t.tausnir};
    for (auto* v : pft) p = fill_view(*v, p);
    p = fill_view(S.albsat, p);
    p = fill_view(S.albdry, p);
    auto& s = *S.snicar_data;
    ViewD1* band[18] = {&s.ss_alb_oc1, &s.asm_prm_oc1, &s.ext_cff_mss_oc1, &s.ss_alb_oc2, &s.asm_prm_oc2, &s.ext_cff_mss_oc2,
        &s.ss_alb_dst1, &s.asm_prm_dst1, &s.ext_cff_mss_dst1, &s.ss_alb_dst2, &s.asm_prm_dst2, &s.ext_cff_mss_dst2,
        &s.ss_alb_dst3, &s.asm_prm_dst3, &s.ext_cff_mss_dst3, &s.ss_alb_dst4, &s.asm_prm_dst4, &s.ext_cff_mss_dst4};
    for (auto* v : band) p = fill_view(*v, p);
    ViewD2* snow[6] = {&s.ss_alb_snw_drc, &s.asm_prm_snw_drc, &s.ext_cff_mss_snw_drc, &s.ss_alb_snw_dfs, &s.asm_prm_snw_dfs,
        &s.ext_cff_mss_snw_dfs};
    for (auto* v : snow) p = fill_view(*v, p);
    ViewD2* bc[6] = {&s.ss_alb_bc1, &s.asm_prm_bc1, &s.ext_cff_mss_bc1, &s.ss_alb_bc2, &s.asm_prm_bc2, &s.ext_cff_mss_bc2};
    for (auto* v : bc) p = fill_view(*v, p);
    p = fill_view(s.bcenh, p);
    auto& a = *S.snw_rds_table;
    p = fill_view(a.snowage_tau, p);
    p = fill_view(a.snowage_kappa, p);
    p = fill_view(a.snowage_drdt0, p);
  }

  // ---- column state: blob -> first device mirror -> S, through the adaptor's accessors ----
  {
    ELM::b200::Device in(n);
    const std::vector<char> blob = slurp(dir + "/state.bin");
    const char* p = blob.data();
    for (int f = 0; f < elmk_field_count(); ++f) {
      int dt = 0, nl = 0;
      elmk_field_info(f, nullptr, &dt, &nl);
      const size_t bytes = (size_t)n * nl * (dt == ELMK_F64 ? 8 : dt == ELMK_I32 ? 4 : 1);
      if (elmk_upload(in.handle(), f, p, 0, n, ELMK_COL_OUTER) != ELMK_OK) return 3;
      p += bytes;
    }
    in.sync();
    in.download(S);
    // psn_pft travels up only (an input): fill the reference's struct array from the blob the way the driver does
    std::vector<double> psn((size_t)n * 27);
    elmk_download(in.handle(), elmk_field_id("psn_pft"), psn.data(), 0, n, ELMK_COL_OUTER);
    for (int i = 0; i < n; ++i) {
      double* q = reinterpret_cast<double*>(&S.psn_pft(i));
      for (int k = 0; k < 27; ++k) q[k] = psn[(size_t)i * 27 + k];
    }
  }

  ELM::b200::Device dev(n);
  dev.set_tables(S);
  dev.upload(S);
  auto read_forcing = [&](int k) {
    const std::vector<char> blob = slurp(dir + "/forcing_" + std::to_string(k) + ".bin");
    const double* p = reinterpret_cast<const double*>(blob.data());
    auto col = [&](auto& v) { for (int i = 0; i < n; ++i) v(i) = static_cast<std::decay_t<decltype(v(0))>>(*p++); };
    auto col2 = [&](auto& v) { for (int i = 0; i < n; ++i) for (int l = 0; l < 2; ++l) v(i, l) = *p++; };
    col(S.coszen); col(S.forc_tbot); col(S.forc_thbot); col(S.forc_pbot); col(S.forc_qbot); col(S.forc_lwrad); col(S.forc_u);
    col(S.forc_v); col(S.forc_rain); col(S.forc_snow); col2(S.forc_solad); col2(S.forc_solai); col(S.elai); col(S.esai);
    col(S.frac_veg_nosno_alb);
  };
  for (int k = 0; k < nsteps; ++k) {
    read_forcing(k);
    dev.upload_forcing(S);
    dev.advance(1800.0);
  }
  dev.download(S);
  dev.check_errors();

  // ---- one more step through the drop-in wrappers, written like ELMInterface::advance ----
  if (argc < 5) {   // (a fifth argument skips the drop-in step: development aid)
    read_forcing(nsteps);
    const double dtime = 1800.0;
    const ELM::Utils::Date current(1985, 1, 1);
    ELM::b200::kokkos_init_timestep_columns(S);
    ELM::kokkos_frac_wet(S);
    ELM::kokkos_albedo_snicar(S);
    ELM::kokkos_canopy_hydrology(S, dtime);
    ELM::kokkos_surface_radiation(S);
    ELM::kokkos_canopy_temperature(S);
    ELM::kokkos_bareground_fluxes(S);
    ELM::kokkos_canopy_fluxes(S, dtime);
    ELM::kokkos_soil_temperature(S, dtime);
    ELM::kokkos_snow_hydrology(S, dtime, current);
    ELM::kokkos_surface_fluxes(S, dtime);
    ELM::kokkos_evaluate_conservation(S, dtime);
  }
  ELM::b200::release(S);

  // ---- every member of the reference state, read through its own accessors ----
  std::ofstream out(dir + "/out.bin", std::ios::binary);
  auto dump = [&](auto& v, int nl) {
    for (int i = 0; i < n; ++i) {
      if constexpr (std::decay_t<decltype(v)>::rank == 1) { const double x = static_cast<double>(v(i)); out.write(reinterpret_cast<const char*>(&x), 8); }
      else for (int l = 0; l < nl; ++l) { const double x = static_cast<double>(v(i, l)); out.write(reinterpret_cast<const char*>(&x), 8); }
    }
  };
#define X(name, member, nlev) dump(member, nlev);
  ELMK_STATE_MEMBERS(X)
  ELMK_AEROSOL_MEMBERS(X)
#undef X
  std::printf("adaptor_steps: %d columns, %d resident steps + 1 step through the drop-in wrappers\n", n, nsteps);
  return 0;
}
