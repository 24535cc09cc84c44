// phys_cantemp.h - canopy temperature group (a5): saves the previous ground temperatures, ground
// temperature, soil-surface humidity factors, saturated humidities, emissivities, roughness lengths,
// forcing heights; zeroes the flux accumulators.  Also the shared saturation-vapour-pressure routine.
//
// Parity target: kokkos_canopy_temperature, reference
// driver/kokkos/canopy_temperature_kokkos.cc:6-131 ->
//   old_ground_temp :9, ground_temp :32, calc_soilalpha :51, calc_soilbeta :133, humidities :143,
//   ground_properties :205, forcing_height :260, init_energy_fluxes :299
//                                                   (src/physics/canopy_temperature_impl.hh)
//   surface_resistance::calc_soilevap_stress        (src/physics/surface_resistance_impl.hh:9-46)
//   qsat                                            (src/physics/qsat_impl.hh:7-79)
#pragma once
#include "elmk_state.h"

namespace elmk {

// Saturation vapour pressure es [Pa], specific humidity qs [kg/kg] and their temperature
// derivatives: 8th-order polynomials over water (0..100 C) and over ice (-75..0 C), Horner form.
// The called copy returns its four results by value: reference parameters of a non-inlined function would force
// the caller's variables - in CanopyFluxes, members of the iteration state - into local memory.
struct QSat { double es, esdT, qs, qsdT; };
ELMK_HD_NOINLINE QSat qsat_values(const double T, const double p)
{
  double es, esdT, qs, qsdT;
  double td = T - TFRZ;
  if (td > 100.0) td = 100.0;
  if (td < -75.0) td = -75.0;
  if (td >= 0.0) {
    es = 6.11213476 + td * (0.444007856 + td * (0.143064234e-01 + td * (0.264461437e-03 + td * (0.305903558e-05 +
         td * (0.196237241e-07 + td * (0.892344772e-10 + td * (-0.373208410e-12 + td * 0.209339997e-15)))))));
    esdT = 0.444017302 + td * (0.286064092e-01 + td * (0.794683137e-03 + td * (0.121211669e-04 + td * (0.103354611e-06 +
           td * (0.404125005e-09 + td * (-0.788037859e-12 + td * (-0.114596802e-13 + td * 0.381294516e-16)))))));
  } else {
    es = 6.11123516 + td * (0.503109514 + td * (0.188369801e-01 + td * (0.420547422e-03 + td * (0.614396778e-05 +
         td * (0.602780717e-07 + td * (0.387940929e-09 + td * (0.149436277e-11 + td * 0.262655803e-14)))))));
    esdT = 0.503277922 + td * (0.377289173e-01 + td * (0.126801703e-02 + td * (0.249468427e-04 + td * (0.313703411e-06 +
           td * (0.257180651e-08 + td * (0.133268878e-10 + td * (0.394116744e-13 + td * 0.498070196e-16)))))));
  }
  es = es * 100.0;
  esdT = esdT * 100.0;
  const double vp = 1.0 / (p - 0.378 * es);
  const double vp1 = 0.622 * vp;
  const double vp2 = vp1 * vp;
  qs = es * vp1;
  qsdT = esdT * vp2 * p;
  return {es, esdT, qs, qsdT};
}
ELMK_HD void qsat(const double T, const double p, double& es, double& esdT, double& qs, double& qsdT)
{
  const QSat q = qsat_values(T, p);
  es = q.es; esdT = q.esdT; qs = q.qs; qsdT = q.qsdT;
}

ELMK_HD void column_canopy_temperature(const Cols& S, const Tables& T, const int c)
{
  const int snl = C1(snl);
  const int top = NLEVSNO - snl;
  const double fsno = C1(frac_sno), fsno_eff = C1(frac_sno_eff), fsfc = C1(frac_h2osfc);
  const double t_sfc = C1(t_h2osfc);

  // -- old_ground_temp: remember the temperatures the solver starts from --
#pragma unroll
  for (int i = 0; i < NLEVTOT; ++i) C2(tssbef, i) = C2(t_soisno, i);
  C1(t_h2osfc_bef) = t_sfc;

  const double t_soil1 = C2(t_soisno, NLEVSNO);
  const double t_top = C2(t_soisno, top);   // == t_soil1 when snl == 0

  // -- ground_temp --
  double tg;
  if (snl > 0) {
    tg = fsno_eff * t_top + (1.0 - fsno_eff - fsfc) * t_soil1 + fsfc * t_sfc;
  } else {
    tg = (1.0 - fsfc) * t_soil1 + fsfc * t_sfc;
  }
  C1(t_grnd) = tg;

  // -- calc_soilalpha (qred, hr) and calc_soilbeta (Lee & Pielke beta) --
  const double liq1 = C2(h2osoi_liq, NLEVSNO), ice1 = C2(h2osoi_ice, NLEVSNO), dz1 = C2(dz, NLEVSNO);
  const double watsat1 = C2(watsat, 0);
  const double wx = (liq1 / DENH2O + ice1 / DENICE) / dz1;
  double fac = dmin(1.0, wx / watsat1);
  fac = dmax(fac, 0.01);
  double psit = -C2(sucsat, 0) * m_pow(fac, (-C2(bsw, 0)));
  psit = dmax(-1.e8, psit);
  const double hr = m_exp(psit / ROVERG / t_soil1);
  // qred = (1 - fsno - fsfc) hr + fsno + fsfc is computed by the reference but only feeds soilalpha (unused)

  const double watfc1 = C2(watfc, 0);
  double soilbeta;
  if (wx < watfc1) {
    double fac_fc = dmin(1.0, wx / watfc1);
    fac_fc = dmax(fac_fc, 0.01);
    soilbeta = (1.0 - fsno - fsfc) * 0.25 * sq(1.0 - m_cos(PI * fac_fc)) + fsno + fsfc;
  } else {
    soilbeta = 1.0;
  }
  C1(soilbeta) = soilbeta;

  // -- humidities --
  const double forc_q = C1(forc_qbot), pbot = C1(forc_pbot);
  double eg, degdT, qsatg, qsatgdT;
  qsat(t_top, pbot, eg, degdT, qsatg, qsatgdT);
  // (the reference's guard "qsatg > forc_q && forc_q > qsatg" can never hold: SURVEY.md quirk 6)
  double qg_snow = qsatg;
  double dqgdT = fsno * qsatgdT;
  qsat(t_soil1, pbot, eg, degdT, qsatg, qsatgdT);
  if (qsatg > forc_q && forc_q > hr * qsatg) {
    qsatg = forc_q;
    qsatgdT = 0.0;
  }
  const double qg_soil = hr * qsatg;
  dqgdT = dqgdT + (1.0 - fsno - fsfc) * hr * qsatgdT;
  if (snl == 0) {
    qg_snow = qg_soil;
    dqgdT = (1.0 - fsfc) * hr * dqgdT;
  }
  qsat(t_sfc, pbot, eg, degdT, qsatg, qsatgdT);
  const double qg_h2osfc = qsatg;
  dqgdT = dqgdT + fsfc * qsatgdT;
  C1(qg_snow) = qg_snow;
  C1(qg_soil) = qg_soil;
  C1(qg_h2osfc) = qg_h2osfc;
  C1(dqgdT) = dqgdT;
  C1(qg) = fsno_eff * qg_snow + (1.0 - fsno_eff - fsfc) * qg_soil + fsfc * qg_h2osfc;

  // -- ground_properties --
  const double lsai = C1(elai) + C1(esai);
  C1(emg) = (1.0 - fsno) * 0.96 + fsno * 0.97;
  C1(emv) = 1.0 - m_exp(-lsai / 1.0);
  double htvp = HVAP;
  if (C2(h2osoi_liq, top) <= 0 && C2(h2osoi_ice, top) > 0.0) htvp = HSUB;
  C1(htvp) = htvp;
  const double z0mg = (fsno > 0.0) ? ZSNO : ZLND;
  C1(z0mg) = z0mg;
  C1(z0hg) = z0mg;
  C1(z0qg) = z0mg;
  // the reference indexes the PFT tables with the GLOBAL Land.vtype here
  const double htop = C1(htop);
  const double z0m = T.z0mr[T.vtype] * htop;
  const double displa = T.displar[T.vtype] * htop;
  C1(z0m) = z0m;
  C1(displa) = displa;
  C1(z0mv) = z0m;
  C1(z0hv) = z0m;
  C1(z0qv) = z0m;
  C1(thv) = C1(forc_thbot) * (1.0 + 0.61 * forc_q);

  // -- forcing_height: the patch heights accumulate (+=) on top of the per-step reset to forc_hgt --
  double hu = C1(forc_hgt_u_patch), ht = C1(forc_hgt_t_patch), hq = C1(forc_hgt_q_patch);
  if (C1(veg_active)) {
    const double add = (C1(frac_veg_nosno) == 0) ? z0mg + displa : z0m + displa;
    hu += add;
    ht += add;
    hq += add;
  }
  C1(forc_hgt_u_patch) = hu;
  C1(forc_hgt_t_patch) = ht;
  C1(forc_hgt_q_patch) = hq;
  C1(thm) = C1(forc_tbot) + 0.0098 * ht;

  // -- init_energy_fluxes --
  C1(eflx_sh_tot) = 0.0;
  C1(eflx_lh_tot) = 0.0;
  C1(eflx_sh_veg) = 0.0;
  C1(qflx_evap_tot) = 0.0;
  C1(qflx_evap_veg) = 0.0;
  C1(qflx_tran_veg) = 0.0;
}

} // namespace elmk
