// canopy_temperature.h - ELM::canopy_temperature::* of the reference (src/physics/canopy_temperature.h,
// canopy_temperature_impl.hh:9-332) on the B200 backend: identical names, namespace, argument order and meaning; each
// call runs the function's device code (elmkernels_b200/csrc/phys_cantemp.h, namespace tmp) through elmk_fn_call.
// The reference's test/test_CanTemp.cc compiles unchanged with -I<repo>/include/elm in place of
// -I<reference>/src/physics.  Layer rows have nlevsno + nlevgrnd = 20 elements, soil-property rows nlevgrnd = 15.
#pragma once
#include "elm_constants.h"   // the reference's data / constants headers (src/data)
#include "land_data.h"

#include "elm_b200_fn.hh"

namespace ELM::canopy_temperature {

template <typename ArrayD1>
void old_ground_temp(const LandType& Land, const double& t_h2osfc, const ArrayD1 t_soisno, double& t_h2osfc_bef,
                     ArrayD1 tssbef)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_TMP_OLD_GROUND_TEMP).in(t_h2osfc).row(t_soisno, 20, false).io(t_h2osfc_bef).row(tssbef, 20, true).call();
}

template <typename ArrayD1>
void ground_temp(const LandType& Land, const int& snl, const double& frac_sno_eff, const double& frac_h2osfc,
                 const double& t_h2osfc, const ArrayD1 t_soisno, double& t_grnd)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_TMP_GROUND_TEMP).in(snl).in(frac_sno_eff).in(frac_h2osfc).in(t_h2osfc).row(t_soisno, 20, false)
      .io(t_grnd).call();
}

template <typename ArrayD1>
void calc_soilalpha(const LandType& Land, const double& frac_sno, const double& frac_h2osfc, const ArrayD1 h2osoi_liq,
                    const ArrayD1 h2osoi_ice, const ArrayD1 dz, const ArrayD1 t_soisno, const ArrayD1 watsat,
                    const ArrayD1 sucsat, const ArrayD1 bsw, const ArrayD1 watdry, const ArrayD1 watopt, double& qred,
                    double& hr, double& soilalpha)
{
  b200::fn::require_soil(Land);
  qred = 1.0;   // (canopy_temperature_impl.hh:62, before the land-unit branches)
  b200::fn::Args(ELMK_FN_TMP_CALC_SOILALPHA).in(frac_sno).in(frac_h2osfc).row(h2osoi_liq, 20, false).row(h2osoi_ice, 20, false)
      .row(dz, 20, false).row(t_soisno, 20, false).row(watsat, 15, false).row(sucsat, 15, false).row(bsw, 15, false)
      .row(watdry, 15, false).row(watopt, 15, false).io(qred).io(hr).io(soilalpha).call();
}

template <typename ArrayD1>
void calc_soilbeta(const LandType& Land, const double& frac_sno, const double& frac_h2osfc, const ArrayD1 watsat,
                   const ArrayD1 watfc, const ArrayD1 h2osoi_liq, const ArrayD1 h2osoi_ice, const ArrayD1 dz,
                   double& soilbeta)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_TMP_CALC_SOILBETA).in(frac_sno).in(frac_h2osfc).row(watsat, 15, false).row(watfc, 15, false)
      .row(h2osoi_liq, 20, false).row(h2osoi_ice, 20, false).row(dz, 20, false).io(soilbeta).call();
}

template <typename ArrayD1>
void humidities(const LandType& Land, const int& snl, const double& forc_q, const double& forc_pbot, const double& t_h2osfc,
                const double& t_grnd, const double& frac_sno, const double& frac_sno_eff, const double& frac_h2osfc,
                const double& qred, const double& hr, const ArrayD1 t_soisno, double& qg_snow, double& qg_soil, double& qg,
                double& qg_h2osfc, double& dqgdT)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_TMP_HUMIDITIES).in(snl).in(forc_q).in(forc_pbot).in(t_h2osfc).in(t_grnd).in(frac_sno).in(frac_sno_eff)
      .in(frac_h2osfc).in(qred).in(hr).row(t_soisno, 20, false).io(qg_snow).io(qg_soil).io(qg).io(qg_h2osfc).io(dqgdT).call();
}

// displar / z0mr: the PFT tables; the reference reads their entry at Land.vtype (canopy_temperature_impl.hh:245-246)
template <typename ArrayD1, typename SubviewD1>
void ground_properties(const LandType& Land, const int& snl, const double& frac_sno, const double& forc_th,
                       const double& forc_q, const double& elai, const double& esai, const double& htop,
                       const SubviewD1 displar, const SubviewD1 z0mr, const ArrayD1 h2osoi_liq, const ArrayD1 h2osoi_ice,
                       double& emg, double& emv, double& htvp, double& z0mg, double& z0hg, double& z0qg, double& z0mv,
                       double& z0hv, double& z0qv, double& thv, double& z0m, double& displa)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_TMP_GROUND_PROPERTIES).in(snl).in(frac_sno).in(forc_th).in(forc_q).in(elai).in(esai).in(htop)
      .in(static_cast<double>(displar(Land.vtype))).in(static_cast<double>(z0mr(Land.vtype))).row(h2osoi_liq, 20, false)
      .row(h2osoi_ice, 20, false).io(emg).io(emv).io(htvp).io(z0mg).io(z0hg).io(z0qg).io(z0mv).io(z0hv).io(z0qv).io(thv).io(z0m)
      .io(displa).call();
}

inline void forcing_height(const LandType& Land, const bool& veg_active, const int& frac_veg_nosno, const double& z0m,
                           const double& z0mg, const double& forc_t, const double& displa, double& forc_hgt_u_patch,
                           double& forc_hgt_t_patch, double& forc_hgt_q_patch, double& thm)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_TMP_FORCING_HEIGHT).in(veg_active).in(frac_veg_nosno).in(z0m).in(z0mg).in(forc_t).in(displa)
      .io(forc_hgt_u_patch).io(forc_hgt_t_patch).io(forc_hgt_q_patch).io(thm).call();
}

inline void init_energy_fluxes(const LandType&, double& eflx_sh_tot, double& eflx_lh_tot, double& eflx_sh_veg,
                               double& qflx_evap_tot, double& qflx_evap_veg, double& qflx_tran_veg)
{
  b200::fn::Args(ELMK_FN_TMP_INIT_ENERGY_FLUXES).io(eflx_sh_tot).io(eflx_lh_tot).io(eflx_sh_veg).io(qflx_evap_tot)
      .io(qflx_evap_veg).io(qflx_tran_veg).call();
}

} // namespace ELM::canopy_temperature
