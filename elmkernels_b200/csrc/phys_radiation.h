// phys_radiation.h - surface radiation group (a4): sunlit/shaded canopy fractions and absorbed PAR,
// solar radiation absorbed by vegetation / ground / snow layers, reflected solar.
//
// Parity target: kokkos_surface_radiation, reference
// driver/kokkos/surface_radiation_kokkos.cc:7-97 ->
//   canopy_sunshade_fractions :202, initialize_flux :9, total_absorbed_radiation :30,
//   layer_absorbed_radiation :77, reflected_radiation :179   (src/physics/surface_radiation_impl.hh)
// Closed form, no transcendentals: bandwidth-bound.  trd/tri (the wrapper's two scratch Views) stay
// in registers.
#pragma once
#include "elmk_state.h"

namespace elmk {

ELMK_HD void column_surface_radiation(const Cols& S, const Tables&, const int c)
{
  const int snl = C1(snl);
  const int nrad = C1(nrad);
  double solad[NUMRAD], solai[NUMRAD];
#pragma unroll
  for (int ib = 0; ib < NUMRAD; ++ib) {
    solad[ib] = C2(forc_solad, ib);
    solai[ib] = C2(forc_solai, ib);
  }

  // -- canopy_sunshade_fractions (nlevcan == 1: at most one canopy layer) --
  double laisun = 0.0, laisha = 0.0;
  if (nrad > 0) {
    const double tlai = C2(tlai_z, 0), fsun = C2(fsun_z, 0);
    const double lsun_z = tlai * fsun;
    const double lsha_z = tlai * (1.0 - fsun);
    laisun += lsun_z;
    laisha += lsha_z;
    C2(laisun_z, 0) = lsun_z;
    C2(laisha_z, 0) = lsha_z;
    C2(parsun_z, 0) = solad[0] * C2(fabd_sun_z, 0) + solai[0] * C2(fabi_sun_z, 0);
    C2(parsha_z, 0) = solad[0] * C2(fabd_sha_z, 0) + solai[0] * C2(fabi_sha_z, 0);
  }
  C1(laisun) = laisun;
  C1(laisha) = laisha;

  // -- total_absorbed_radiation --
  double sabg_soil = 0.0, sabg_snow = 0.0, sabg = 0.0, sabv = 0.0, fsa = 0.0;
  double trd[NUMRAD], tri[NUMRAD];
#pragma unroll
  for (int ib = 0; ib < NUMRAD; ++ib) {
    const double cad = solad[ib] * C2(fabd, ib);
    const double cai = solai[ib] * C2(fabi, ib);
    sabv += cad + cai;
    fsa += cad + cai;
    trd[ib] = solad[ib] * C2(ftdd, ib);
    tri[ib] = solad[ib] * C2(ftid, ib) + solai[ib] * C2(ftii, ib);
    double absrad = trd[ib] * (1.0 - C2(albsod, ib)) + tri[ib] * (1.0 - C2(albsoi, ib));
    sabg_soil += absrad;
    absrad = trd[ib] * (1.0 - C2(albsnd, ib)) + tri[ib] * (1.0 - C2(albsni, ib));
    sabg_snow += absrad;
    absrad = trd[ib] * (1.0 - C2(albgrd, ib)) + tri[ib] * (1.0 - C2(albgri, ib));
    sabg += absrad;
    fsa += absrad;
    if (snl == 0) {
      sabg_snow = sabg;
      sabg_soil = sabg;
    }
  }

  // -- layer_absorbed_radiation --
  double lyr[NLEVSNO + 1];
#pragma unroll
  for (int i = 0; i <= NLEVSNO; ++i) lyr[i] = 0.0;
  if (snl == 0) {
    lyr[NLEVSNO] = sabg;
  } else {
    double snl_sum = 0.0;
#pragma unroll
    for (int i = 0; i <= NLEVSNO; ++i) {
      lyr[i] = C2(flx_absdv, i) * trd[0] + C2(flx_absdn, i) * trd[1] + C2(flx_absiv, i) * tri[0] + C2(flx_absin, i) * tri[1];
      if (i >= NLEVSNO - snl) snl_sum += lyr[i];
    }
    // the per-layer factors are stale when the number of snow layers changed since the albedo call:
    // redistribute over the top layers (reference :135-148)
    if (fabs(snl_sum - sabg_snow) > 0.00001) {
#pragma unroll
      for (int i = 0; i <= NLEVSNO; ++i) lyr[i] = 0.0;
      if (snl == 1) {
        lyr[NLEVSNO - 1] = sabg_snow * 0.6;
        lyr[NLEVSNO] = sabg_snow * 0.4;
      } else {
#pragma unroll
        for (int i = 0; i <= NLEVSNO; ++i) {
          if (i == NLEVSNO - snl) lyr[i] = sabg_snow * 0.75;
          if (i == NLEVSNO - snl + 1) lyr[i] = sabg_snow * 0.25;
        }
      }
    }
  }
  double err_sum = 0.0;
#pragma unroll
  for (int i = 0; i <= NLEVSNO; ++i) {
    err_sum += lyr[i];
    C2(sabg_lyr, i) = lyr[i];
  }
  if (fabs(err_sum - sabg_snow) > 0.00001) C1(errmask) |= (int)ERR_SABG_LAYERS;

  C1(sabg_soil) = sabg_soil;
  C1(sabg_snow) = sabg_snow;
  C1(sabg) = sabg;
  C1(sabv) = sabv;
  C1(fsa) = fsa;

  // -- reflected_radiation --
  const double rvis = C2(albd, 0) * solad[0] + C2(albi, 0) * solai[0];
  const double rnir = C2(albd, 1) * solad[1] + C2(albi, 1) * solai[1];
  C1(fsr) = rvis + rnir;
}

} // namespace elmk
