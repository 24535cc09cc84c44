// canopy_fluxes.h - ELM::canopy_fluxes::* of the reference (src/physics/canopy_fluxes.h, canopy_fluxes_impl.hh:95-539) on
// the B200 backend: identical names, namespace, argument order and meaning; each call runs the function's device code
// (elmkernels_b200/csrc/phys_canflux.h: canflux_begin, the loop over canflux_iterate, canflux_end - the pieces of the
// re-packed CanopyFluxes kernels) through elmk_fn_call.  The reference's test/test_CanFlux.cc compiles unchanged with
// -I<repo>/include/elm in place of -I<reference>/src/physics.  Layer rows have nlevsno + nlevgrnd = 20 elements, soil
// rows nlevgrnd = 15, canopy-layer rows nlevcan = 1.
#pragma once
#include "elm_constants.h"   // the reference's data / constants headers (src/data)
#include "land_data.h"
#include "pft_data.h"

#include "elm_b200_fn.hh"

namespace ELM::canopy_fluxes {

template <typename ArrayD1>
void initialize_flux(const LandType& Land, const int& snl, const int& frac_veg_nosno, const double& frac_sno,
                     const double& forc_hgt_u_patch, const double& thm, const double& thv, const double& max_dayl,
                     const double& dayl, const int& altmax_indx, const int& altmax_lastyear_indx,
                     const ArrayD1 t_soisno, const ArrayD1 h2osoi_ice, const ArrayD1 h2osoi_liq, const ArrayD1 dz,
                     const ArrayD1 rootfr, const double& tc_stress, const ArrayD1 sucsat, const ArrayD1 watsat,
                     const ArrayD1 bsw, const double& smpso, const double& smpsc, const double& elai,
                     const double& esai, const double& emv, const double& emg, const double& qg, const double& t_grnd,
                     const double& forc_t, const double& forc_pbot, const double& forc_lwrad, const double& forc_u,
                     const double& forc_v, const double& forc_q, const double& forc_th, const double& z0mg,
                     double& btran, double& displa, double& z0mv, double& z0hv, double& z0qv, ArrayD1 rootr,
                     ArrayD1 eff_porosity, double& dayl_factor, double& air, double& bir, double& cir, double& el,
                     double& qsatl, double& qsatldT, double& taf, double& qaf, double& um, double& ur, double& obu,
                     double& zldis, double& delq, double& t_veg)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_CF_INITIALIZE_FLUX).in(snl).in(frac_veg_nosno).in(frac_sno).in(forc_hgt_u_patch).in(thm).in(thv)
      .in(max_dayl).in(dayl).in(altmax_indx).in(altmax_lastyear_indx).row(t_soisno, 20, false).row(h2osoi_ice, 20, false)
      .row(h2osoi_liq, 20, false).row(dz, 20, false).row(rootfr, 15, false).in(tc_stress).row(sucsat, 15, false)
      .row(watsat, 15, false).row(bsw, 15, false).in(smpso).in(smpsc).in(elai).in(esai).in(emv).in(emg).in(qg).in(t_grnd)
      .in(forc_t).in(forc_pbot).in(forc_lwrad).in(forc_u).in(forc_v).in(forc_q).in(forc_th).in(z0mg)
      .io(btran).io(displa).io(z0mv).io(z0hv).io(z0qv).row(rootr, 15, true).row(eff_porosity, 15, true).io(dayl_factor).io(air)
      .io(bir).io(cir).io(el).io(qsatl).io(qsatldT).io(taf).io(qaf).io(um).io(ur).io(obu).io(zldis).io(delq).io(t_veg).call();
}

template <typename ArrayD1>
void stability_iteration(
    const LandType& Land, const double& dtime, const int& snl, const int& frac_veg_nosno, const double& frac_sno,
    const double& forc_hgt_u_patch, const double& forc_hgt_t_patch, const double& forc_hgt_q_patch, const double& fwet,
    const double& fdry, const double& laisun, const double& laisha, const double& forc_rho, const double& snow_depth,
    const double& soilbeta, const double& frac_h2osfc, const double& t_h2osfc, const double& sabv, const double& h2ocan,
    const double& htop, const ArrayD1 t_soisno, const double& air, const double& bir, const double& cir,
    const double& ur, const double& zldis, const double& displa, const double& elai, const double& esai,
    const double& t_grnd, const double& forc_pbot, const double& forc_q, const double& forc_th, const double& z0mg,
    const double& z0mv, const double& z0hv, const double& z0qv, const double& thm, const double& thv, const double& qg,
    const PFTDataPSN& psn_pft, const int& nrad, const double& t10, const ArrayD1 tlai_z, const double& vcmaxcintsha,
    const double& vcmaxcintsun, const ArrayD1 parsha_z, const ArrayD1 parsun_z, const ArrayD1 laisha_z,
    const ArrayD1 laisun_z, const double& forc_pco2, const double& forc_po2, const double& dayl_factor, double& btran,
    double& qflx_tran_veg, double& qflx_evap_veg, double& eflx_sh_veg, double& wtg, double& wtl0, double& wta0,
    double& wtal, double& el, double& qsatl, double& qsatldT, double& taf, double& qaf, double& um, double& dth,
    double& dqh, double& obu, double& temp1, double& temp2, double& temp12m, double& temp22m, double& tlbef,
    double& delq, double& dt_veg, double& t_veg, double& wtgq, double& wtalq, double& wtlq0, double& wtaq0)
{
  b200::fn::require_soil(Land);
  b200::fn::Args args(ELMK_FN_CF_STABILITY_ITERATION);
  args.in(Land.vtype).in(dtime).in(snl).in(frac_veg_nosno).in(frac_sno).in(forc_hgt_u_patch).in(forc_hgt_t_patch)
      .in(forc_hgt_q_patch).in(fwet).in(fdry).in(laisun).in(laisha).in(forc_rho).in(snow_depth).in(soilbeta).in(frac_h2osfc)
      .in(t_h2osfc).in(sabv).in(h2ocan).in(htop).row(t_soisno, 20, false).in(air).in(bir).in(cir).in(ur).in(zldis).in(displa)
      .in(elai).in(esai).in(t_grnd).in(forc_pbot).in(forc_q).in(forc_th).in(z0mg).in(z0mv).in(z0hv).in(z0qv).in(thm).in(thv)
      .in(qg);
  // the 27 members of PFTDataPSN in their order (pft_data.h:20-24)
  const double psn[27] = {psn_pft.fnr, psn_pft.act25, psn_pft.kcha, psn_pft.koha, psn_pft.cpha, psn_pft.vcmaxha, psn_pft.jmaxha,
                          psn_pft.tpuha, psn_pft.lmrha, psn_pft.vcmaxhd, psn_pft.jmaxhd, psn_pft.tpuhd, psn_pft.lmrhd,
                          psn_pft.lmrse, psn_pft.qe, psn_pft.theta_cj, psn_pft.bbbopt, psn_pft.mbbopt, psn_pft.c3psn,
                          psn_pft.slatop, psn_pft.leafcn, psn_pft.flnr, psn_pft.fnitr, psn_pft.dleaf, psn_pft.smpso,
                          psn_pft.smpsc, psn_pft.tc_stress};
  for (double v : psn) args.in(v);
  args.in(nrad).in(t10).row(tlai_z, 1, false).in(vcmaxcintsha).in(vcmaxcintsun).row(parsha_z, 1, false).row(parsun_z, 1, false)
      .row(laisha_z, 1, false).row(laisun_z, 1, false).in(forc_pco2).in(forc_po2).in(dayl_factor)
      .io(btran).io(qflx_tran_veg).io(qflx_evap_veg).io(eflx_sh_veg).io(wtg).io(wtl0).io(wta0).io(wtal).io(el).io(qsatl)
      .io(qsatldT).io(taf).io(qaf).io(um).io(dth).io(dqh).io(obu).io(temp1).io(temp2).io(temp12m).io(temp22m).io(tlbef).io(delq)
      .io(dt_veg).io(t_veg).io(wtgq).io(wtalq).io(wtlq0).io(wtaq0).call();
}

template <typename ArrayD1>
void compute_flux(const LandType& Land, const double& dtime, const int& snl, const int& frac_veg_nosno,
                  const double& frac_sno, const ArrayD1 t_soisno, const double& frac_h2osfc, const double& t_h2osfc,
                  const double& sabv, const double& qg_snow, const double& qg_soil, const double& qg_h2osfc,
                  const double& dqgdT, const double& htvp, const double& wtg, const double& wtl0, const double& wta0,
                  const double& wtal, const double& air, const double& bir, const double& cir, const double& qsatl,
                  const double& qsatldT, const double& dth, const double& dqh, const double& temp1, const double& temp2,
                  const double& temp12m, const double& temp22m, const double& tlbef, const double& delq,
                  const double& dt_veg, const double& t_veg, const double& t_grnd, const double& forc_pbot,
                  const double& qflx_tran_veg, const double& qflx_evap_veg, const double& eflx_sh_veg,
                  const double& forc_q, const double& forc_rho, const double& thm, const double& emv, const double& emg,
                  const double& forc_lwrad, const double& wtgq, const double& wtalq, const double& wtlq0,
                  const double& wtaq0, double& h2ocan, double& eflx_sh_grnd, double& eflx_sh_snow, double& eflx_sh_soil,
                  double& eflx_sh_h2osfc, double& qflx_evap_soi, double& qflx_ev_snow, double& qflx_ev_soil,
                  double& qflx_ev_h2osfc, double& dlrad, double& ulrad, double& cgrnds, double& cgrndl, double& cgrnd,
                  double& t_ref2m, double& q_ref2m, double& rh_ref2m)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_CF_COMPUTE_FLUX).in(dtime).in(snl).in(frac_veg_nosno).in(frac_sno).row(t_soisno, 20, false)
      .in(frac_h2osfc).in(t_h2osfc).in(sabv).in(qg_snow).in(qg_soil).in(qg_h2osfc).in(dqgdT).in(htvp).in(wtg).in(wtl0).in(wta0)
      .in(wtal).in(air).in(bir).in(cir).in(qsatl).in(qsatldT).in(dth).in(dqh).in(temp1).in(temp2).in(temp12m).in(temp22m)
      .in(tlbef).in(delq).in(dt_veg).in(t_veg).in(t_grnd).in(forc_pbot).in(qflx_tran_veg).in(qflx_evap_veg).in(eflx_sh_veg)
      .in(forc_q).in(forc_rho).in(thm).in(emv).in(emg).in(forc_lwrad).in(wtgq).in(wtalq).in(wtlq0).in(wtaq0)
      .io(h2ocan).io(eflx_sh_grnd).io(eflx_sh_snow).io(eflx_sh_soil).io(eflx_sh_h2osfc).io(qflx_evap_soi).io(qflx_ev_snow)
      .io(qflx_ev_soil).io(qflx_ev_h2osfc).io(dlrad).io(ulrad).io(cgrnds).io(cgrndl).io(cgrnd).io(t_ref2m).io(q_ref2m)
      .io(rh_ref2m).call();
}

} // namespace ELM::canopy_fluxes
