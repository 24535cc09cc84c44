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

// The eight functions of the group, one per reference function (the library-level API of include/elm/
// canopy_temperature.h calls them one by one through elmk_fn_call; the column body below composes them).  Layer
// values arrive as scalars: the caller picks the top snow / first soil layer out of its rows.
namespace tmp {

// old_ground_temp :9-29 - the temperatures the solver starts from
ELMK_HD void old_ground_temp(const double t_h2osfc, const ColRow t_soisno, double& t_h2osfc_bef, const ColRow tssbef)
{
#pragma unroll
  for (int i = 0; i < NLEVTOT; ++i) tssbef[i] = t_soisno[i];
  t_h2osfc_bef = t_h2osfc;
}

// ground_temp :32-48; t_top = t_soisno(nlevsno - snl), t_soil1 = t_soisno(nlevsno)
ELMK_HD double ground_temp(const int snl, const double fsno_eff, const double fsfc, const double t_sfc, const double t_top,
                           const double t_soil1)
{
  if (snl > 0) return fsno_eff * t_top + (1.0 - fsno_eff - fsfc) * t_soil1 + fsfc * t_sfc;
  return (1.0 - fsfc) * t_soil1 + fsfc * t_sfc;
}

// partial volume of ice and water of the first soil layer (calc_soilalpha :82, calc_soilevap_stress :27)
ELMK_HD double surface_wetness(const double liq1, const double ice1, const double dz1)
{
  return (liq1 / DENH2O + ice1 / DENICE) / dz1;
}

// calc_soilalpha :51-130, soil / crop branch
ELMK_HD void calc_soilalpha(const double fsno, const double fsfc, const double liq1, const double ice1, const double dz1,
                            const double t_soil1, const double watsat1, const double sucsat1, const double bsw1, double& qred,
                            double& hr, double& soilalpha)
{
  const double wx = surface_wetness(liq1, ice1, dz1);
  double fac = dmin(1.0, wx / watsat1);
  fac = dmax(fac, 0.01);
  double psit = -sucsat1 * m_pow(fac, (-bsw1));
  psit = dmax(-1.e8, psit);
  hr = m_exp(psit / ROVERG / t_soil1);
  qred = (1.0 - fsno - fsfc) * hr + fsno + fsfc;
  soilalpha = qred;
}

// calc_soilbeta :133-140 -> surface_resistance::calc_soilevap_stress (surface_resistance_impl.hh:9-46), Lee & Pielke beta
ELMK_HD double calc_soilbeta(const double fsno, const double fsfc, const double watfc1, const double liq1, const double ice1,
                             const double dz1)
{
  const double wx = surface_wetness(liq1, ice1, dz1);
  if (wx < watfc1) {
    double fac_fc = dmin(1.0, wx / watfc1);
    fac_fc = dmax(fac_fc, 0.01);
    return (1.0 - fsno - fsfc) * 0.25 * sq(1.0 - m_cos(PI * fac_fc)) + fsno + fsfc;
  }
  return 1.0;
}

// humidities :143-202, soil / crop branch
ELMK_HD void humidities(const int snl, const double forc_q, const double pbot, const double t_sfc, const double fsno,
                        const double fsno_eff, const double fsfc, const double hr, const double t_top, const double t_soil1,
                        double& qg_snow, double& qg_soil, double& qg, double& qg_h2osfc, double& dqgdT)
{
  double eg, degdT, qsatg, qsatgdT;
  qsat(t_top, pbot, eg, degdT, qsatg, qsatgdT);
  // (the reference's guard "qsatg > forc_q && forc_q > qsatg" can never hold: SURVEY.md quirk 6)
  qg_snow = qsatg;
  dqgdT = fsno * qsatgdT;
  qsat(t_soil1, pbot, eg, degdT, qsatg, qsatgdT);
  if (qsatg > forc_q && forc_q > hr * qsatg) {
    qsatg = forc_q;
    qsatgdT = 0.0;
  }
  qg_soil = hr * qsatg;
  dqgdT = dqgdT + (1.0 - fsno - fsfc) * hr * qsatgdT;
  if (snl == 0) {
    qg_snow = qg_soil;
    dqgdT = (1.0 - fsfc) * hr * dqgdT;
  }
  qsat(t_sfc, pbot, eg, degdT, qsatg, qsatgdT);
  qg_h2osfc = qsatg;
  dqgdT = dqgdT + fsfc * qsatgdT;
  qg = fsno_eff * qg_snow + (1.0 - fsno_eff - fsfc) * qg_soil + fsfc * qg_h2osfc;
}

// ground_properties :205-257; displar_v / z0mr_v = the PFT-table entries at Land.vtype, liq_top / ice_top = the
// water contents of layer nlevsno - snl
ELMK_HD void ground_properties(const double fsno, const double forc_th, const double forc_q, const double elai,
                               const double esai, const double htop, const double displar_v, const double z0mr_v,
                               const double liq_top, const double ice_top, double& emg, double& emv, double& htvp,
                               double& z0mg, double& z0hg, double& z0qg, double& z0mv, double& z0hv, double& z0qv, double& thv,
                               double& z0m, double& displa)
{
  emg = (1.0 - fsno) * 0.96 + fsno * 0.97;
  emv = 1.0 - m_exp(-(elai + esai) / 1.0);
  htvp = HVAP;
  if (liq_top <= 0 && ice_top > 0.0) htvp = HSUB;
  z0mg = (fsno > 0.0) ? ZSNO : ZLND;
  z0hg = z0mg;
  z0qg = z0mg;
  z0m = z0mr_v * htop;
  displa = displar_v * htop;
  z0mv = z0m;
  z0hv = z0m;
  z0qv = z0m;
  thv = forc_th * (1.0 + 0.61 * forc_q);
}

// forcing_height :260-296, soil / crop branch: the patch heights accumulate (+=) on top of the per-step reset
ELMK_HD void forcing_height(const bool veg_active, const int frac_veg_nosno, const double z0m, const double z0mg,
                            const double forc_t, const double displa, double& hu, double& ht, double& hq, double& thm)
{
  if (veg_active) {
    const double add = (frac_veg_nosno == 0) ? z0mg + displa : z0m + displa;
    hu += add;
    ht += add;
    hq += add;
  }
  thm = forc_t + 0.0098 * ht;
}

// init_energy_fluxes :299-332
ELMK_HD void init_energy_fluxes(double& eflx_sh_tot, double& eflx_lh_tot, double& eflx_sh_veg, double& qflx_evap_tot,
                                double& qflx_evap_veg, double& qflx_tran_veg)
{
  eflx_sh_tot = 0.0;
  eflx_lh_tot = 0.0;
  eflx_sh_veg = 0.0;
  qflx_evap_tot = 0.0;
  qflx_evap_veg = 0.0;
  qflx_tran_veg = 0.0;
}

} // namespace tmp

ELMK_HD void column_canopy_temperature(const Cols& S, const Tables& T, const int c)
{
  const int snl = C1(snl);
  const int top = NLEVSNO - snl;
  const double fsno = C1(frac_sno), fsno_eff = C1(frac_sno_eff), fsfc = C1(frac_h2osfc);
  const double t_sfc = C1(t_h2osfc);

  double t_sfc_bef;
  tmp::old_ground_temp(t_sfc, ELMK_ROW(t_soisno), t_sfc_bef, ELMK_ROW(tssbef));
  C1(t_h2osfc_bef) = t_sfc_bef;
  const double t_soil1 = C2(t_soisno, NLEVSNO);
  const double t_top = C2(t_soisno, top);   // == t_soil1 when snl == 0
  C1(t_grnd) = tmp::ground_temp(snl, fsno_eff, fsfc, t_sfc, t_top, t_soil1);

  const double liq1 = C2(h2osoi_liq, NLEVSNO), ice1 = C2(h2osoi_ice, NLEVSNO), dz1 = C2(dz, NLEVSNO);
  double qred, hr, soilalpha;   // (qred only feeds soilalpha, which nothing on the chain reads)
  tmp::calc_soilalpha(fsno, fsfc, liq1, ice1, dz1, t_soil1, C2(watsat, 0), C2(sucsat, 0), C2(bsw, 0), qred, hr, soilalpha);
  C1(soilbeta) = tmp::calc_soilbeta(fsno, fsfc, C2(watfc, 0), liq1, ice1, dz1);

  const double forc_q = C1(forc_qbot);
  double qg_snow, qg_soil, qg, qg_h2osfc, dqgdT;
  tmp::humidities(snl, forc_q, C1(forc_pbot), t_sfc, fsno, fsno_eff, fsfc, hr, t_top, t_soil1, qg_snow, qg_soil, qg, qg_h2osfc,
                  dqgdT);
  C1(qg_snow) = qg_snow;
  C1(qg_soil) = qg_soil;
  C1(qg_h2osfc) = qg_h2osfc;
  C1(dqgdT) = dqgdT;
  C1(qg) = qg;

  // the reference indexes the PFT tables with the GLOBAL Land.vtype here
  double emg, emv, htvp, z0mg, z0hg, z0qg, z0mv, z0hv, z0qv, thv, z0m, displa;
  tmp::ground_properties(fsno, C1(forc_thbot), forc_q, C1(elai), C1(esai), C1(htop), T.displar[T.vtype], T.z0mr[T.vtype],
                         C2(h2osoi_liq, top), C2(h2osoi_ice, top), emg, emv, htvp, z0mg, z0hg, z0qg, z0mv, z0hv, z0qv, thv, z0m,
                         displa);
  C1(emg) = emg;
  C1(emv) = emv;
  C1(htvp) = htvp;
  C1(z0mg) = z0mg;
  C1(z0hg) = z0hg;
  C1(z0qg) = z0qg;
  C1(z0m) = z0m;
  C1(displa) = displa;
  C1(z0mv) = z0mv;
  C1(z0hv) = z0hv;
  C1(z0qv) = z0qv;
  C1(thv) = thv;

  double hu = C1(forc_hgt_u_patch), ht = C1(forc_hgt_t_patch), hq = C1(forc_hgt_q_patch), thm;
  tmp::forcing_height(C1(veg_active) != 0, C1(frac_veg_nosno), z0m, z0mg, C1(forc_tbot), displa, hu, ht, hq, thm);
  C1(thm) = thm;
  C1(forc_hgt_u_patch) = hu;
  C1(forc_hgt_t_patch) = ht;
  C1(forc_hgt_q_patch) = hq;

  tmp::init_energy_fluxes(C1(eflx_sh_tot), C1(eflx_lh_tot), C1(eflx_sh_veg), C1(qflx_evap_tot), C1(qflx_evap_veg),
                          C1(qflx_tran_veg));
}

} // namespace elmk
