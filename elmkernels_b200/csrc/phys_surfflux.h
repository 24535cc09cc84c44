// phys_surfflux.h - surface-flux update (group a10), balance diagnostics (group a11) and the
// per-column part of the begin-of-step bookkeeping.
//
// Parity targets:
//   kokkos_surface_fluxes          reference driver/kokkos/surface_fluxes_kokkos.cc:6-107 ->
//     prev_tgrnd :10, delta_t :24, evap_ratio :32, initial_flux_calc :74, update_surface_fluxes :147,
//     lwrad_outgoing :240, soil_energy_balance :264        (src/physics/surface_fluxes_impl.hh)
//   kokkos_evaluate_conservation   driver/kokkos/conserved_quantity_kokkos.cc:8-81 ->
//     column_water_mass :7, dh2o_dt :19, column_water_balance_error :26, snow_water_balance_error :37,
//     solar_shortwave_balance_error :73, solar_longwave_balance_error :84,
//     surface_energy_balance_error :95, net_radiation :105
//                                                  (src/physics/conserved_quantity_evaluators_impl.hh)
//   per-column body of kokkos_init_timestep  driver/kokkos/init_timestep_kokkos.cc:53-72 ->
//     ELM::init_timestep                           (src/physics/init_timestep_impl.hh:9-39)
// All closed form: bandwidth-bound.
#pragma once
#include "elmk_state.h"

namespace elmk {

// ---- a10 ---------------------------------------------------------------------------------------
ELMK_HD void column_surface_fluxes(const Cols& S, const Tables& T, const double dtime, const int c)
{
  // the number of snow layers is the one AFTER snow hydrology, tssbef was saved before it (quirk 10)
  const int snl = C1(snl);
  const int snotop = NLEVSNO - snl;
  const double fse = C1(frac_sno_eff), fsfc = C1(frac_h2osfc);
  const double t_sfc_bef = C1(t_h2osfc_bef);
  const double tss_snotop = C2(tssbef, snotop), tss_soitop = C2(tssbef, NLEVSNO);
  const double tg = C1(t_grnd);
  const double cgrnds = C1(cgrnds), cgrndl = C1(cgrndl);
  const double htvp = C1(htvp);

  // ground temperature the fluxes were computed with, and its change over the step
  const double t_grnd0 = (snl > 0) ? fse * tss_snotop + (1.0 - fse - fsfc) * tss_soitop + fsfc * t_sfc_bef
                                   : (1.0 - fsfc) * tss_soitop + fsfc * t_sfc_bef;
  const double tinc = tg - t_grnd0;

  // -- initial_flux_calc --
  double sh_grnd = C1(eflx_sh_grnd) + tinc * cgrnds;
  double evap_soi = C1(qflx_evap_soi) + tinc * cgrndl;
  double ev_snow = C1(qflx_ev_snow) + tinc * cgrndl;
  double ev_soil = C1(qflx_ev_soil) + tinc * cgrndl;
  double ev_sfc = C1(qflx_ev_h2osfc) + tinc * cgrndl;

  // -- update_surface_fluxes --
  // the wrapper pairs the ICE of the top snow layer with the LIQUID of the top soil layer (quirk 10)
  const double ice_top = C2(h2osoi_ice, snotop);
  const double liq_top = C2(h2osoi_liq, NLEVSNO);
  double egsmax = (ice_top + liq_top) / dtime;
  if (egsmax < 0.0) egsmax = 0.0;
  const double egirat = (evap_soi > egsmax) ? egsmax / evap_soi : 1.0;
  if (egirat < 1.0) {
    const double save = evap_soi;
    evap_soi *= egirat;
    sh_grnd += (save - evap_soi) * htvp;
    ev_snow *= egirat;
    ev_soil *= egirat;
    ev_sfc *= egirat;
  }
  const int veg = C1(frac_veg_nosno);
  const double emg = C1(emg), lwrad = C1(forc_lwrad), dlrad = C1(dlrad);
  {
    // quirk 5: exponent 40 on the surface-water term and the cube taken of the whole product
    const double lw_grnd = (fse * pow4(tss_snotop) + (1.0 - fse - fsfc) * pow4(tss_soitop) + fsfc * m_pow(t_sfc_bef, 40.0));
    const double p3 = emg * STEBOL * t_grnd0;
    C1(eflx_soil_grnd) = ((1.0 - fse) * C1(sabg_soil) + fse * C1(sabg_snow)) + dlrad +
                         (1.0 - (double)veg) * emg * lwrad - emg * STEBOL * lw_grnd - cube(p3) * (4.0 * tinc) -
                         (sh_grnd + evap_soi * htvp);
  }
  const double evap_veg = C1(qflx_evap_veg);
  C1(eflx_sh_tot) = C1(eflx_sh_veg) + sh_grnd;
  C1(qflx_evap_tot) = evap_veg + evap_soi;
  C1(eflx_lh_tot) = HVAP * evap_veg + htvp * evap_soi;

  double evap_grnd = 0.0, sub_snow = 0.0, dew_snow = 0.0, dew_grnd = 0.0;
  if (ev_snow >= 0.0) {
    if ((liq_top + ice_top) > 0.0) {
      evap_grnd = dmax(ev_snow * (liq_top / (liq_top + ice_top)), 0.0);
    } else {
      evap_grnd = 0.0;
    }
    sub_snow = ev_snow - evap_grnd;
  } else {
    if (tg < TFRZ) {
      dew_snow = fabs(ev_snow);
    } else {
      dew_grnd = fabs(ev_snow);
    }
  }
  if (snl > 0 && C1(do_capsnow)) {
    C1(qflx_snwcp_liq) = C1(qflx_snwcp_liq) + fse * dew_grnd;
    C1(qflx_snwcp_ice) = C1(qflx_snwcp_ice) + fse * dew_snow;
  }
  C1(eflx_sh_grnd) = sh_grnd;
  C1(qflx_evap_soi) = evap_soi;
  C1(qflx_ev_snow) = ev_snow;
  C1(qflx_ev_soil) = ev_soil;
  C1(qflx_ev_h2osfc) = ev_sfc;
  C1(qflx_evap_grnd) = evap_grnd;
  C1(qflx_sub_snow) = sub_snow;
  C1(qflx_dew_snow) = dew_snow;
  C1(qflx_dew_grnd) = dew_grnd;

  // -- lwrad_outgoing --
  {
    const double lw_grnd = (fse * pow4(tss_snotop) + (1.0 - fse - fsfc) * pow4(tss_soitop) + fsfc * pow4(t_sfc_bef));
    const double out = C1(ulrad) + (1 - veg) * (1.0 - emg) * lwrad + (1 - veg) * emg * STEBOL * lw_grnd +
                       4.0 * emg * STEBOL * cube(t_grnd0) * tinc;
    C1(eflx_lwrad_out) = out;
    C1(eflx_lwrad_net) = out - lwrad;
  }

  // -- soil_energy_balance --
  {
    const double t_sfc = C1(t_h2osfc);
    double errsoi = C1(eflx_soil_grnd) - C1(xmf) - C1(xmf_h2osfc) - fsfc * (t_sfc - t_sfc_bef) * (t_sfc / dtime);
    errsoi += C1(eflx_h2osfc_snow);
#pragma unroll
    for (int j = 0; j < NLEVTOT; ++j) {
      if (j >= NLEVSNO - snl && j < NLEVSNO) errsoi -= fse * (C2(t_soisno, j) - C2(tssbef, j)) / C2(fact, j);
      if (j >= NLEVSNO) errsoi -= (C2(t_soisno, j) - C2(tssbef, j)) / C2(fact, j);
    }
    C1(soil_e_balance) = errsoi;
  }
  (void)T;
}

// total water of the column [kg/m2]: canopy + snow + surface water + every layer's ice and liquid,
// summed in the reference's order
ELMK_HD double column_water_mass(const Cols& S, const int c)
{
  double water = C1(h2ocan) + C1(h2osno) + C1(h2osfc);
#pragma unroll
  for (int i = 0; i < NLEVTOT; ++i) water += C2(h2osoi_ice, i) + C2(h2osoi_liq, i);
  return water;
}

// ---- a11 ---------------------------------------------------------------------------------------
ELMK_HD void column_conservation(const Cols& S, const Tables&, const double dtime, const int c)
{
  constexpr double hydrology_source_sink = 0.0;
  const double begwb = C1(dtbegin_column_h2o);
  const double endwb = column_water_mass(S, c);
  C1(dtend_column_h2o) = endwb;
  C1(dwb) = (endwb - begwb) / dtime;
  const double q_snwcp_ice = C1(qflx_snwcp_ice);
  C1(errh2o) = (endwb - begwb) -
               (C1(forc_rain) + C1(forc_snow) - hydrology_source_sink - C1(qflx_evap_tot) - q_snwcp_ice) * dtime;

  double errsno = 0.0;
  if (C1(snl) > 0) {
    const double dew_snow = C1(qflx_dew_snow), dew_grnd = C1(qflx_dew_grnd), sub_snow = C1(qflx_sub_snow);
    const double evap_grnd = C1(qflx_evap_grnd), snow_melt = C1(qflx_snow_melt), sl_top = C1(qflx_sl_top_soil);
    const double fse = C1(frac_sno_eff), rain_grnd = C1(qflx_rain_grnd), snow_grnd = C1(qflx_snow_grnd);
    const double sfc_ice = C1(qflx_h2osfc_ice);
    double sources, sinks;
    if (C1(do_capsnow)) {
      sources = fse * (dew_snow + dew_grnd) + sfc_ice + snow_grnd + rain_grnd;
      sinks = fse * (sub_snow + evap_grnd) + q_snwcp_ice + C1(qflx_snwcp_liq) + snow_melt + sl_top;
    } else {
      constexpr double snow_h2osfc = 0.0;
      sources = (snow_grnd - snow_h2osfc) + fse * (rain_grnd + dew_snow + dew_grnd) + sfc_ice;
      sinks = fse * (sub_snow + evap_grnd) + snow_melt + sl_top;
    }
    errsno = (C1(h2osno) - C1(h2osno_old)) - (sources - sinks) * dtime;
  }
  C1(errh2osno) = errsno;

  const double fsa = C1(fsa), lw_out = C1(eflx_lwrad_out), lw_net = C1(eflx_lwrad_net), lwrad = C1(forc_lwrad);
  C1(errsol) = fsa + C1(fsr) - (C2(forc_solad, 0) + C2(forc_solad, 1) + C2(forc_solai, 0) + C2(forc_solai, 1));
  C1(errlon) = lw_out - lw_net - lwrad;
  C1(errseb) = C1(sabv) + C1(sabg_chk) + lwrad - lw_out - C1(eflx_sh_tot) - C1(eflx_lh_tot) - C1(eflx_soil_grnd);
  C1(netrad) = fsa - lw_net;
}

// ---- begin-of-step bookkeeping -----------------------------------------------------------------
ELMK_HD void column_init_timestep(const Cols& S, const Tables&, const int reset_forc_hgt, const int c)
{
  if (reset_forc_hgt) {
    // ProcessZBOT, reference src/physics/atm_physics_impl.hh:197-202
    const double h = C1(forc_hgt);
    C1(forc_hgt_u_patch) = h;
    C1(forc_hgt_t_patch) = h;
    C1(forc_hgt_q_patch) = h;
  }
  const double h2osno = C1(h2osno);
  C1(h2osno_old) = h2osno;
  C1(dtbegin_column_h2o) = column_water_mass(S, c);
  C1(do_capsnow) = (h2osno > H2OSNO_MAX) ? 1 : 0;
  C1(frac_veg_nosno) = C1(veg_active) ? C1(frac_veg_nosno_alb) : 0;
  const int snl = C1(snl);
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i) {
    if (i >= NLEVSNO - snl) {
      const double ice = C2(h2osoi_ice, i);
      C2(frac_iceold, i) = ice / (C2(h2osoi_liq, i) + ice);
    }
  }
}

} // namespace elmk
