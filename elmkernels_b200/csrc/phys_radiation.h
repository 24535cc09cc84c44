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
#include "phys_hydrology.h"   // ColRow / ELMK_ROW

namespace elmk {

// The bodies at the reference's function granularity (rows through any accessor with operator[]), shared by the fused
// column kernel below and the library-level ELM::surface_radiation::* entry points (include/elm/surface_radiation.h).
namespace rad {
// canopy_sunshade_fractions (surface_radiation_impl.hh:202-240); nrad <= nlevcan == 1
ELMK_HD void canopy_sunshade_fractions(const int nrad, const double /*elai*/, const ColRow tlai_z, const ColRow fsun_z, const ColRow solad,
                                       const ColRow solai, const ColRow fabd_sun_z, const ColRow fabd_sha_z, const ColRow fabi_sun_z,
                                       const ColRow fabi_sha_z, const ColRow parsun_z, const ColRow parsha_z, const ColRow laisun_z,
                                       const ColRow laisha_z, double& laisun_out, double& laisha_out)
{
  double laisun = 0.0, laisha = 0.0;
  for (int iv = 0; iv < nrad; ++iv) {
    const double tlai = tlai_z[iv], fsun = fsun_z[iv];
    const double lsun_z = tlai * fsun;
    const double lsha_z = tlai * (1.0 - fsun);
    laisun += lsun_z;
    laisha += lsha_z;
    laisun_z[iv] = lsun_z;
    laisha_z[iv] = lsha_z;
    parsun_z[iv] = solad[0] * fabd_sun_z[iv] + solai[0] * fabi_sun_z[iv];
    parsha_z[iv] = solad[0] * fabd_sha_z[iv] + solai[0] * fabi_sha_z[iv];
  }
  laisun_out = laisun;
  laisha_out = laisha;
}

// initialize_flux (:9-27)
ELMK_HD void initialize_flux(double& sabg_soil, double& sabg_snow, double& sabg, double& sabv, double& fsa, const ColRow sabg_lyr)
{
  sabg_soil = 0.0; sabg_snow = 0.0; sabg = 0.0; sabv = 0.0; fsa = 0.0;
  for (int j = 0; j <= NLEVSNO; ++j) sabg_lyr[j] = 0.0;
}

// total_absorbed_radiation (:30-74), subgridflag == 1
ELMK_HD void total_absorbed_radiation(const int snl, const ColRow ftdd, const ColRow ftid, const ColRow ftii, const ColRow solad,
                                      const ColRow solai, const ColRow fabd, const ColRow fabi, const ColRow albsod, const ColRow albsoi,
                                      const ColRow albsnd, const ColRow albsni, const ColRow albgrd, const ColRow albgri, double& sabv_io,
                                      double& fsa_io, double& sabg_io, double& sabg_soil_io, double& sabg_snow_io, const ColRow trd,
                                      const ColRow tri)
{
  double sabv = sabv_io, fsa = fsa_io, sabg = sabg_io, sabg_soil = sabg_soil_io, sabg_snow = sabg_snow_io;
  for (int ib = 0; ib < NUMRAD; ++ib) {
    const double cad = solad[ib] * fabd[ib];
    const double cai = solai[ib] * fabi[ib];
    sabv += cad + cai;
    fsa += cad + cai;
    const double td = solad[ib] * ftdd[ib];
    const double ti = solad[ib] * ftid[ib] + solai[ib] * ftii[ib];
    trd[ib] = td;
    tri[ib] = ti;
    double absrad = td * (1.0 - albsod[ib]) + ti * (1.0 - albsoi[ib]);
    sabg_soil += absrad;
    absrad = td * (1.0 - albsnd[ib]) + ti * (1.0 - albsni[ib]);
    sabg_snow += absrad;
    absrad = td * (1.0 - albgrd[ib]) + ti * (1.0 - albgri[ib]);
    sabg += absrad;
    fsa += absrad;
    if (snl == 0) {
      sabg_snow = sabg;
      sabg_soil = sabg;
    }
  }
  sabv_io = sabv; fsa_io = fsa; sabg_io = sabg; sabg_soil_io = sabg_soil; sabg_snow_io = sabg_snow;
}

// layer_absorbed_radiation (:77-176), subgridflag == 1; returns false where the reference asserts (:173)
ELMK_HD bool layer_absorbed_radiation(const int snl, const double sabg, const double sabg_snow, const double /*snow_depth*/,
                                      const ColRow flx_absdv, const ColRow flx_absdn, const ColRow flx_absiv, const ColRow flx_absin,
                                      const ColRow trd, const ColRow tri, const ColRow sabg_lyr)
{
  double lyr[NLEVSNO + 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i <= NLEVSNO; ++i) lyr[i] = 0.0;
  if (snl == 0) {
    lyr[NLEVSNO] = sabg;
  } else {
    const double trd0 = trd[0], trd1 = trd[1], tri0 = tri[0], tri1 = tri[1];
    double snl_sum = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i <= NLEVSNO; ++i) {
      lyr[i] = flx_absdv[i] * trd0 + flx_absdn[i] * trd1 + flx_absiv[i] * tri0 + flx_absin[i] * tri1;
      if (i >= NLEVSNO - snl) snl_sum += lyr[i];
    }
    // the per-layer factors are stale when the number of snow layers changed since the albedo call:
    // redistribute over the top layers (reference :119-138)
    if (fabs(snl_sum - sabg_snow) > 0.00001) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int i = 0; i <= NLEVSNO; ++i) lyr[i] = 0.0;
      if (snl == 1) {
        lyr[NLEVSNO - 1] = sabg_snow * 0.6;
        lyr[NLEVSNO] = sabg_snow * 0.4;
      } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i <= NLEVSNO; ++i) {
          if (i == NLEVSNO - snl) lyr[i] = sabg_snow * 0.75;
          if (i == NLEVSNO - snl + 1) lyr[i] = sabg_snow * 0.25;
        }
      }
    }
  }
  double err_sum = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i <= NLEVSNO; ++i) {
    err_sum += lyr[i];
    sabg_lyr[i] = lyr[i];
  }
  return !(fabs(err_sum - sabg_snow) > 0.00001);
}

// reflected_radiation (:179-199)
ELMK_HD void reflected_radiation(const ColRow albd, const ColRow albi, const ColRow solad, const ColRow solai, double& fsr)
{
  const double rvis = albd[0] * solad[0] + albi[0] * solai[0];
  const double rnir = albd[1] * solad[1] + albi[1] * solai[1];
  fsr = rvis + rnir;
}
} // namespace rad

ELMK_HD void column_surface_radiation(const Cols& S, const Tables&, const int c)
{
  const int snl = C1(snl);
  const int nrad = C1(nrad);
  double solad_[NUMRAD], solai_[NUMRAD], trd_[NUMRAD], tri_[NUMRAD];
#pragma unroll
  for (int ib = 0; ib < NUMRAD; ++ib) {
    solad_[ib] = C2(forc_solad, ib);
    solai_[ib] = C2(forc_solai, ib);
  }
  const ColRow solad{solad_, 1}, solai{solai_, 1}, trd{trd_, 1}, tri{tri_, 1};   // local rows (the wrapper's scratch Views)

  double laisun, laisha;
  rad::canopy_sunshade_fractions(nrad > 0 ? 1 : 0, 0.0, ELMK_ROW(tlai_z), ELMK_ROW(fsun_z), solad, solai, ELMK_ROW(fabd_sun_z),
                                 ELMK_ROW(fabd_sha_z), ELMK_ROW(fabi_sun_z), ELMK_ROW(fabi_sha_z), ELMK_ROW(parsun_z),
                                 ELMK_ROW(parsha_z), ELMK_ROW(laisun_z), ELMK_ROW(laisha_z), laisun, laisha);
  C1(laisun) = laisun;
  C1(laisha) = laisha;

  double sabg_soil = 0.0, sabg_snow = 0.0, sabg = 0.0, sabv = 0.0, fsa = 0.0;   // initialize_flux
  rad::total_absorbed_radiation(snl, ELMK_ROW(ftdd), ELMK_ROW(ftid), ELMK_ROW(ftii), solad, solai, ELMK_ROW(fabd), ELMK_ROW(fabi),
                                ELMK_ROW(albsod), ELMK_ROW(albsoi), ELMK_ROW(albsnd), ELMK_ROW(albsni), ELMK_ROW(albgrd),
                                ELMK_ROW(albgri), sabv, fsa, sabg, sabg_soil, sabg_snow, trd, tri);
  if (!rad::layer_absorbed_radiation(snl, sabg, sabg_snow, 0.0, ELMK_ROW(flx_absdv), ELMK_ROW(flx_absdn), ELMK_ROW(flx_absiv),
                                     ELMK_ROW(flx_absin), trd, tri, ELMK_ROW(sabg_lyr)))
    C1(errmask) |= (int)ERR_SABG_LAYERS;

  C1(sabg_soil) = sabg_soil;
  C1(sabg_snow) = sabg_snow;
  C1(sabg) = sabg;
  C1(sabv) = sabv;
  C1(fsa) = fsa;

  double fsr;
  rad::reflected_radiation(ELMK_ROW(albd), ELMK_ROW(albi), solad, solai, fsr);
  C1(fsr) = fsr;
}

} // namespace elmk
