// Compile-and-run check of include/elm_b200.hh against the REFERENCE's own state type.
// Built by tests/test_adaptor_cpu.py where /root/reference is mounted: the reference headers are used
// where they lie (nothing copied), with the host Kokkos stand-in of oracle/shim, and the program is linked
// against a library exporting the elmk C ABI (the host port on a CPU-only machine).
#include <cstdio>
#include <cstdlib>

#include "compile_options.hh"
#include "data_types.hh"
#include "elm_constants.h"
#include "utils.hh"
#include "date_time.hh"

#include "elm_b200.hh"

// every drop-in wrapper must instantiate for the reference's ELMStateType
template void ELM::b200::kokkos_frac_wet<ELMStateType>(ELMStateType&);
template void ELM::b200::kokkos_albedo_snicar<ELMStateType>(ELMStateType&);
template void ELM::b200::kokkos_canopy_hydrology<ELMStateType>(ELMStateType&, const double&);
template void ELM::b200::kokkos_surface_radiation<ELMStateType>(ELMStateType&);
template void ELM::b200::kokkos_canopy_temperature<ELMStateType>(ELMStateType&);
template void ELM::b200::kokkos_bareground_fluxes<ELMStateType>(ELMStateType&);
template void ELM::b200::kokkos_canopy_fluxes<ELMStateType>(ELMStateType&, const double&);
template void ELM::b200::kokkos_soil_temperature<ELMStateType>(ELMStateType&, const double&);
template void ELM::b200::kokkos_snow_hydrology<ELMStateType, ELM::Utils::Date>(ELMStateType&, const double&, const ELM::Utils::Date&);
template void ELM::b200::kokkos_surface_fluxes<ELMStateType>(ELMStateType&, const double&);
template void ELM::b200::kokkos_evaluate_conservation<ELMStateType>(ELMStateType&, const double&);

int main() {
  const int n = 37;
  auto dd = ELM::Utils::create_domain_decomposition_2D(ELM::Utils::square_numprocs(1), {1, 1}, {0, 0});
  ELMStateType A(n, dd, std::string(), ELM::Utils::Date(1985, 1, 1), 1), B(n, dd, std::string(), ELM::Utils::Date(1985, 1, 1), 1);
  // a recognisable pattern in arrays of every rank / element type
  for (int i = 0; i < n; ++i) {
    A.snl(i) = i % 6;
    A.veg_active(i) = (i % 3) != 0;
    A.forc_tbot(i) = 270.0 + i;
    for (int l = 0; l < 20; ++l) { A.t_soisno(i, l) = 250.0 + i + 0.01 * l; A.imelt(i, l) = (i + l) % 3; }
    for (int l = 0; l < 21; ++l) A.zisoi(i, l) = -1.0 * i + l;
    for (int l = 0; l < 5; ++l) A.aero_mass->mss_dst3(i, l) = 1e-9 * (i * 5 + l);
    A.aero_input->bcdep(i) = 1e-13 * i;
    double* p = reinterpret_cast<double*>(&A.psn_pft(i));
    for (int k = 0; k < 27; ++k) p[k] = i + 0.001 * k;
  }
  ELM::b200::Device dev(n);
  dev.upload(A);
  dev.download(B);
  int bad = 0;
  for (int i = 0; i < n; ++i) {
    bad += A.snl(i) != B.snl(i);
    bad += A.veg_active(i) != B.veg_active(i);
    bad += A.forc_tbot(i) != B.forc_tbot(i);
    for (int l = 0; l < 20; ++l) bad += (A.t_soisno(i, l) != B.t_soisno(i, l)) + (A.imelt(i, l) != B.imelt(i, l));
    for (int l = 0; l < 21; ++l) bad += A.zisoi(i, l) != B.zisoi(i, l);
    for (int l = 0; l < 5; ++l) bad += A.aero_mass->mss_dst3(i, l) != B.aero_mass->mss_dst3(i, l);
    bad += A.aero_input->bcdep(i) != B.aero_input->bcdep(i);
  }
  // psn_pft goes up only (it is an input); read it back through the ABI
  std::vector<double> psn(n * 27);
  elmk_download(dev.handle(), elmk_field_id("psn_pft"), psn.data(), 0, n, ELMK_COL_OUTER);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 27; ++k) bad += psn[i * 27 + k] != i + 0.001 * k;
  std::printf("adaptor round trip: %d mismatches\n", bad);
  return bad ? 1 : 0;
}
