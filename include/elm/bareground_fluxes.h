// bareground_fluxes.h - ELM::bareground_fluxes::* of the reference (src/physics/bareground_fluxes.h,
// bareground_fluxes_impl.hh:7-170) on the B200 backend: identical names, namespace, argument order and meaning; each
// call runs the function's device code (elmkernels_b200/csrc/phys_bareground.h, namespace bgf, with the Monin-Obukhov
// functions of phys_friction.h) through elmk_fn_call.  The reference's test/test_BGFlux.cc compiles unchanged with
// -I<repo>/include/elm in place of -I<reference>/src/physics.
#pragma once
#include "elm_constants.h"   // the reference's data / constants headers (src/data)
#include "land_data.h"

#include "elm_b200_fn.hh"

namespace ELM::bareground_fluxes {

inline void initialize_flux(const LandType& Land, const int& frac_veg_nosno, const double& forc_u, const double& forc_v,
                            const double& forc_q, const double& forc_th, const double& forc_hgt_u_patch, const double& thm,
                            const double& thv, const double& t_grnd, const double& qg, const double& z0mg, double& dlrad,
                            double& ulrad, double& zldis, double& displa, double& dth, double& dqh, double& obu, double& ur,
                            double& um)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_BGF_INITIALIZE_FLUX).in(frac_veg_nosno).in(forc_u).in(forc_v).in(forc_q).in(forc_th)
      .in(forc_hgt_u_patch).in(thm).in(thv).in(t_grnd).in(qg).in(z0mg).io(dlrad).io(ulrad).io(zldis).io(displa).io(dth).io(dqh)
      .io(obu).io(ur).io(um).call();
}

inline void stability_iteration(const LandType& Land, const int& frac_veg_nosno, const double& forc_hgt_t_patch,
                                const double& forc_hgt_u_patch, const double& forc_hgt_q_patch, const double& z0mg,
                                const double& zldis, const double& displa, const double& dth, const double& dqh,
                                const double& ur, const double& forc_q, const double& forc_th, const double& thv,
                                double& z0hg, double& z0qg, double& obu, double& um, double& temp1, double& temp2,
                                double& temp12m, double& temp22m, double& ustar)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_BGF_STABILITY_ITERATION).in(frac_veg_nosno).in(forc_hgt_t_patch).in(forc_hgt_u_patch)
      .in(forc_hgt_q_patch).in(z0mg).in(zldis).in(displa).in(dth).in(dqh).in(ur).in(forc_q).in(forc_th).in(thv).io(z0hg).io(z0qg)
      .io(obu).io(um).io(temp1).io(temp2).io(temp12m).io(temp22m).io(ustar).call();
}

template <typename ArrayD1>
void compute_flux(const LandType& Land, const int& frac_veg_nosno, const int& snl, const double& forc_rho,
                  const double& soilbeta, const double& dqgdT, const double& htvp, const double& t_h2osfc,
                  const double& qg_snow, const double& qg_soil, const double& qg_h2osfc, const ArrayD1 t_soisno,
                  const double& forc_pbot, const double& dth, const double& dqh, const double& temp1, const double& temp2,
                  const double& temp12m, const double& temp22m, const double& ustar, const double& forc_q, const double& thm,
                  double& cgrnds, double& cgrndl, double& cgrnd, double& eflx_sh_grnd, double& eflx_sh_tot,
                  double& eflx_sh_snow, double& eflx_sh_soil, double& eflx_sh_h2osfc, double& qflx_evap_soi,
                  double& qflx_evap_tot, double& qflx_ev_snow, double& qflx_ev_soil, double& qflx_ev_h2osfc, double& t_ref2m,
                  double& q_ref2m, double& rh_ref2m)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_BGF_COMPUTE_FLUX).in(frac_veg_nosno).in(snl).in(forc_rho).in(soilbeta).in(dqgdT).in(htvp).in(t_h2osfc)
      .in(qg_snow).in(qg_soil).in(qg_h2osfc).row(t_soisno, 20, false).in(forc_pbot).in(dth).in(dqh).in(temp1).in(temp2)
      .in(temp12m).in(temp22m).in(ustar).in(forc_q).in(thm).io(cgrnds).io(cgrndl).io(cgrnd).io(eflx_sh_grnd).io(eflx_sh_tot)
      .io(eflx_sh_snow).io(eflx_sh_soil).io(eflx_sh_h2osfc).io(qflx_evap_soi).io(qflx_evap_tot).io(qflx_ev_snow)
      .io(qflx_ev_soil).io(qflx_ev_h2osfc).io(t_ref2m).io(q_ref2m).io(rh_ref2m).call();
}

} // namespace ELM::bareground_fluxes
