// surface_radiation.h - ELM::surface_radiation::* of the reference (src/physics/surface_radiation.h,
// surface_radiation_impl.hh:9-240) on the B200 backend: identical names, namespace, argument order and meaning; each
// call runs the function's device code (elmkernels_b200/csrc/phys_radiation.h, namespace rad) through elmk_fn_call.
// The reference's test/test_SurfRad.cc and test/test_CanSunShade.cc compile unchanged with -I<repo>/include/elm in
// place of -I<reference>/src/physics.  Rows have the extents of the reference's state: numrad = 2, nlevsno + 1 = 6,
// nlevcan = 1.
#pragma once
#include <cassert>

#include "elm_constants.h"   // the reference's data / constants headers (src/data)
#include "land_data.h"

#include "elm_b200_fn.hh"

namespace ELM::surface_radiation {

template <typename ArrayD1>
void initialize_flux(const LandType& Land, double& sabg_soil, double& sabg_snow, double& sabg, double& sabv, double& fsa,
                     ArrayD1 sabg_lyr)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_RAD_INITIALIZE_FLUX).io(sabg_soil).io(sabg_snow).io(sabg).io(sabv).io(fsa).row(sabg_lyr, 6, true).call();
}

template <typename ArrayD1>
void total_absorbed_radiation(const LandType& Land, const int& snl, const ArrayD1 ftdd, const ArrayD1 ftid, const ArrayD1 ftii,
                              const ArrayD1 forc_solad, const ArrayD1 forc_solai, const ArrayD1 fabd, const ArrayD1 fabi,
                              const ArrayD1 albsod, const ArrayD1 albsoi, const ArrayD1 albsnd, const ArrayD1 albsni,
                              const ArrayD1 albgrd, const ArrayD1 albgri, double& sabv, double& fsa, double& sabg,
                              double& sabg_soil, double& sabg_snow, ArrayD1 trd, ArrayD1 tri)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_RAD_TOTAL_ABSORBED).in(snl).row(ftdd, 2, false).row(ftid, 2, false).row(ftii, 2, false)
      .row(forc_solad, 2, false).row(forc_solai, 2, false).row(fabd, 2, false).row(fabi, 2, false).row(albsod, 2, false)
      .row(albsoi, 2, false).row(albsnd, 2, false).row(albsni, 2, false).row(albgrd, 2, false).row(albgri, 2, false)
      .io(sabv).io(fsa).io(sabg).io(sabg_soil).io(sabg_snow).row(trd, 2, true).row(tri, 2, true).call();
}

template <typename ArrayD1>
void layer_absorbed_radiation(const LandType& Land, const int& snl, const double& sabg, const double& sabg_snow,
                              const double& snow_depth, const ArrayD1 flx_absdv, const ArrayD1 flx_absdn,
                              const ArrayD1 flx_absiv, const ArrayD1 flx_absin, const ArrayD1 trd, const ArrayD1 tri,
                              ArrayD1 sabg_lyr)
{
  b200::fn::require_soil(Land);
  double ok = 1.0;
  b200::fn::Args(ELMK_FN_RAD_LAYER_ABSORBED).in(snl).in(sabg).in(sabg_snow).in(snow_depth).row(flx_absdv, 6, false)
      .row(flx_absdn, 6, false).row(flx_absiv, 6, false).row(flx_absin, 6, false).row(trd, 2, false).row(tri, 2, false)
      .row(sabg_lyr, 6, true).io(ok).call();
  assert(ok != 0.0 && "absorbed solar radiation of the snow layers does not sum to sabg_snow");   // surface_radiation_impl.hh:173
  (void)ok;
}

template <typename ArrayD1>
void reflected_radiation(const LandType& Land, const ArrayD1 albd, const ArrayD1 albi, const ArrayD1 forc_solad,
                         const ArrayD1 forc_solai, double& fsr)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_RAD_REFLECTED).row(albd, 2, false).row(albi, 2, false).row(forc_solad, 2, false).row(forc_solai, 2, false)
      .io(fsr).call();
}

template <typename ArrayD1>
void canopy_sunshade_fractions(const LandType& Land, const int& nrad, const double& elai, const ArrayD1 tlai_z,
                               const ArrayD1 fsun_z, const ArrayD1 forc_solad, const ArrayD1 forc_solai,
                               const ArrayD1 fabd_sun_z, const ArrayD1 fabd_sha_z, const ArrayD1 fabi_sun_z,
                               const ArrayD1 fabi_sha_z, ArrayD1 parsun_z, ArrayD1 parsha_z, ArrayD1 laisun_z, ArrayD1 laisha_z,
                               double& laisun, double& laisha)
{
  b200::fn::require_soil(Land);
  b200::fn::Args(ELMK_FN_RAD_SUNSHADE).in(nrad).in(elai).row(tlai_z, 1, false).row(fsun_z, 1, false).row(forc_solad, 2, false)
      .row(forc_solai, 2, false).row(fabd_sun_z, 1, false).row(fabd_sha_z, 1, false).row(fabi_sun_z, 1, false)
      .row(fabi_sha_z, 1, false).row(parsun_z, 1, true).row(parsha_z, 1, true).row(laisun_z, 1, true).row(laisha_z, 1, true)
      .io(laisun).io(laisha).call();
}

} // namespace ELM::surface_radiation
