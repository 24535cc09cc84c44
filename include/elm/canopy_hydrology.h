// canopy_hydrology.h - ELM::canopy_hydrology::* of the reference (src/physics/canopy_hydrology.h:19-170,
// canopy_hydrology_impl.hh:8-357) on the B200 backend: identical names, namespace, argument order and meaning; each
// call runs the function's device code (elmkernels_b200/csrc/phys_hydrology.h, namespace hyd) through elmk_fn_call.
// A translation unit written against the reference's header - its test/test_CanHydro.cc - compiles unchanged with
// -I<repo>/include/elm in place of -I<reference>/src/physics.
#pragma once
#include "elm_constants.h"   // the reference's data / constants headers (src/data), as its own header includes them
#include "land_data.h"

#include "elm_b200_fn.hh"

namespace ELM::canopy_hydrology {

inline void interception(const LandType& Land, const int& frac_veg_nosno, const double& forc_rain, const double& forc_snow,
                         const double& dewmx, const double& elai, const double& esai, const double& dtime, double& h2ocan,
                         double& qflx_candrip, double& qflx_through_snow, double& qflx_through_rain, double& fracsnow,
                         double& fracrain)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_INTERCEPTION).in(frac_veg_nosno).in(forc_rain).in(forc_snow).in(dewmx).in(elai).in(esai).in(dtime)
      .io(h2ocan).io(qflx_candrip).io(qflx_through_snow).io(qflx_through_rain).io(fracsnow).io(fracrain).call();
}

// (host-side bookkeeping in the reference as well: canopy_hydrology_impl.hh:68-80)
inline void Irrigation(const LandType& Land, const double& irrig_rate, int& n_irrig_steps_left, double& qflx_irrig)
{
  if (!Land.lakpoi) {
    if (n_irrig_steps_left > 0) {
      qflx_irrig = irrig_rate;
      n_irrig_steps_left -= 1;
    } else {
      qflx_irrig = 0.0;
    }
  }
}

inline void ground_flux(const LandType& Land, const int& do_capsnow, const int& frac_veg_nosno, const double& forc_rain,
                        const double& forc_snow, const double& qflx_irrig, const double& qflx_candrip,
                        const double& qflx_through_snow, const double& qflx_through_rain, const double& fracsnow,
                        const double& fracrain, double& qflx_snwcp_liq, double& qflx_snwcp_ice, double& qflx_snow_grnd,
                        double& qflx_rain_grnd)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_GROUND_FLUX).in(do_capsnow).in(frac_veg_nosno).in(forc_rain).in(forc_snow).in(qflx_irrig)
      .in(qflx_candrip).in(qflx_through_snow).in(qflx_through_rain).in(fracsnow).in(fracrain)
      .io(qflx_snwcp_liq).io(qflx_snwcp_ice).io(qflx_snow_grnd).io(qflx_rain_grnd).call();
}

inline void fraction_wet(const LandType& Land, const int& frac_veg_nosno, const double& dewmx, const double& elai,
                         const double& esai, const double& h2ocan, double& fwet, double& fdry)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_FRACTION_WET).in(frac_veg_nosno).in(dewmx).in(elai).in(esai).in(h2ocan).io(fwet).io(fdry).call();
}

template <typename ArrayD1>
void snow_init(const LandType& Land, const double& dtime, const int& do_capsnow, const int& oldfflag, const double& forc_t,
               const double& t_grnd, const double& qflx_snow_grnd, const double& qflx_snow_melt, const double& n_melt,
               double& snow_depth, double& h2osno, double& int_snow, ArrayD1 swe_old, ArrayD1 h2osoi_liq,
               ArrayD1 h2osoi_ice, ArrayD1 t_soisno, ArrayD1 frac_iceold, int& snl, ArrayD1 dz, ArrayD1 z, ArrayD1 zi,
               ArrayD1 snw_rds, double& frac_sno_eff, double& frac_sno)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_SNOW_INIT).in(dtime).in(do_capsnow).in(oldfflag).in(forc_t).in(t_grnd).in(qflx_snow_grnd)
      .in(qflx_snow_melt).in(n_melt).io(snow_depth).io(h2osno).io(int_snow).row(swe_old, 5, true).row(h2osoi_liq, 20, true)
      .row(h2osoi_ice, 20, true).row(t_soisno, 20, true).row(frac_iceold, 5, true).io(snl).row(dz, 20, true).row(z, 20, true)
      .row(zi, 21, true).row(snw_rds, 5, true).io(frac_sno_eff).io(frac_sno).call();
}

template <typename ArrayD1>
void fraction_h2osfc(const LandType& Land, const double& micro_sigma, const double& h2osno, double& h2osfc,
                     ArrayD1 h2osoi_liq, double& frac_sno, double& frac_sno_eff, double& frac_h2osfc)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_FRACTION_H2OSFC).in(micro_sigma).in(h2osno).io(h2osfc).row(h2osoi_liq, 20, true).io(frac_sno)
      .io(frac_sno_eff).io(frac_h2osfc).call();
}

} // namespace ELM::canopy_hydrology
