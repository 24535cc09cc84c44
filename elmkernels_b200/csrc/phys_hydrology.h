// phys_hydrology.h - canopy water: wetted fraction (group a1) and interception / snow
// initialisation / surface-water fraction (group a3).
//
// Parity targets:
//   kokkos_frac_wet          reference driver/kokkos/canopy_hydrology_kokkos.cc:98-112
//     -> canopy_hydrology::fraction_wet      src/physics/canopy_hydrology_impl.hh:123-142
//   kokkos_canopy_hydrology  canopy_hydrology_kokkos.cc:7-96
//     -> interception :8-66, ground_flux :83-119, snow_init :146-308, fraction_h2osfc :312-357
// Scope: soil/crop land units, no lake, no urban (elmk_set_tables rejects anything else), so the
// reference's LandType branches are resolved at compile time here.
#pragma once
#include "elmk_state.h"

namespace elmk {

// The group bodies are written at the reference's function granularity (namespace hyd: same argument lists minus the
// LandType, rows through any accessor with operator[]), so that the fused column kernels and the library-level
// ELM::canopy_hydrology::* entry points (include/elm/canopy_hydrology.h -> elmk_fn_call) run the same device code.


namespace hyd {
// ---- a1: fraction_wet (canopy_hydrology_impl.hh:123-142) ----
ELMK_HD void fraction_wet(const int veg, const double dewmx, const double lai, const double sai, const double canwat,
                          double& fwet, double& fdry)
{
  double wet = 0.0, dry = 0.0;
  if (veg == 1) {
    if (canwat > 0.0) {
      const double vegt = veg * (lai + sai);
      const double dewmxi = 1.0 / dewmx;
      // exponent literal as in the reference, not 2/3
      wet = dmin(m_pow((dewmxi / vegt) * canwat, 0.666666666666), 1.0);
    }
    dry = (1.0 - wet) * lai / (lai + sai);
  }
  fwet = wet;
  fdry = dry;
}

// ---- a3: interception (:8-66): throughfall and canopy drip ----
ELMK_HD void interception(const int veg, const double rain, const double snow, const double dewmx, const double elai,
                          const double esai, const double dtime, double& h2ocan, double& candrip, double& thru_snow,
                          double& thru_rain, double& fracsnow, double& fracrain)
{
  candrip = 0.0; thru_snow = 0.0; thru_rain = 0.0; fracsnow = 0.0; fracrain = 0.0;
  if (veg == 1 && (rain + snow) > 0.0) {
    const double lsai = elai + esai;
    double canwat = h2ocan;
    fracsnow = snow / (snow + rain);
    fracrain = rain / (snow + rain);
    const double canmax = dewmx * lsai;
    const double fpi = 0.25 * (1.0 - m_exp(-0.5 * lsai));
    thru_snow = snow * (1.0 - fpi);
    thru_rain = rain * (1.0 - fpi);
    const double intr = (snow + rain) * fpi;
    canwat = dmax(0.0, (canwat + dtime * intr));
    const double xrun = (canwat - canmax) / dtime;
    if (xrun > 0.0) {
      candrip = xrun;
      canwat = canmax;
    }
    h2ocan = canwat;
  }
}

// ---- ground_flux (:83-119): precipitation reaching the ground ----
ELMK_HD void ground_flux(const int capsnow, const int veg, const double rain, const double snow, const double irrig,
                         const double candrip, const double thru_snow, const double thru_rain, const double fracsnow,
                         const double fracrain, double& snwcp_liq, double& snwcp_ice, double& snow_grnd, double& rain_grnd)
{
  double grnd_snow, grnd_rain;
  if (veg == 0) {
    grnd_snow = snow;
    grnd_rain = rain;
  } else {
    grnd_snow = thru_snow + (candrip * fracsnow);
    grnd_rain = thru_rain + (candrip * fracrain);
  }
  grnd_rain = grnd_rain + irrig;
  if (capsnow) {
    snwcp_liq = grnd_rain;
    snwcp_ice = grnd_snow;
    snow_grnd = 0.0;
    rain_grnd = 0.0;
  } else {
    snwcp_liq = 0.0;
    snwcp_ice = 0.0;
    snow_grnd = grnd_snow;
    rain_grnd = grnd_rain;
  }
}
} // namespace hyd

// density of newly fallen snow [kg/m3] as a function of air temperature (Alta relationship)
ELMK_HD double fresh_snow_density(const double forc_t)
{
  if (forc_t > TFRZ + 2.0) return 50.0 + 1.7 * 0x1.185f05d1aebd7p+6;   // pow(17.0, 1.5) as glibc rounds it
  if (forc_t > TFRZ - 15.0) return 50.0 + 1.7 * m_pow((forc_t - TFRZ + 15.0), 1.5);
  return 50.0;
}

// Niu & Yang (2007) snow-cover fraction used when oldfflag == 1
ELMK_HD double fsca_niu_yang(const double snow_depth, const double swe)
{
  // m_pow(x, 1.0) == x exactly, kept out
  return m_tanh(snow_depth / (2.5 * ZLND * dmin(800.0, (swe / snow_depth / 100.0))));
}

namespace hyd {
// ---- snow_init (:146-308): snow depth, snow-cover fraction, birth of the first snow layer ----
template <class Row>
ELMK_HD void snow_init(const double dtime, const int capsnow, const int oldfflag, const double forc_t, const double /*t_grnd*/,
                       const double snow_grnd, const double qflx_snow_melt, const double nmelt, double& snow_depth,
                       double& h2osno, double& int_snow, const Row swe_old, const Row h2osoi_liq, const Row h2osoi_ice,
                       const Row t_soisno, const Row frac_iceold, int& snl_io, const Row dz, const Row z, const Row zi,
                       const Row snw_rds, double& frac_sno_eff, double& frac_sno)
{
  int snl = snl_io;
  double depth = snow_depth, swe = h2osno, intsnow = int_snow, fsno = frac_sno;
  const double depth_before = depth;
  constexpr double accum_factor = 0.1;
  for (int j = 0; j < NLEVSNO; ++j)
    swe_old[j] = (j < NLEVSNO - snl) ? 0.0 : h2osoi_liq[j] + h2osoi_ice[j];

  double dz_snowf, newsnow;
  if (capsnow) {
    dz_snowf = 0.0;
    newsnow = snow_grnd * dtime;
    fsno = 1.0;
    intsnow = 5.e2;
  } else {
    const double bifall = fresh_snow_density(forc_t);
    newsnow = snow_grnd * dtime;
    intsnow = dmax(intsnow, swe);
    const double snowmelt = qflx_snow_melt * dtime;
    if (swe > 0.0) {
      if (snowmelt > 0.0) {
        const double smr = dmin(1.0, (swe / intsnow));
        fsno = 1.0 - m_pow((m_acos(dmin(1.0, (2.0 * smr - 1.0))) / PI), nmelt);
      }
      if (newsnow > 0.0) {
        fsno = 1.0 - (1.0 - m_tanh(accum_factor * newsnow)) * (1.0 - fsno);
        const double t = (swe + newsnow) / (0.5 * (m_cos(PI * m_pow((1.0 - dmax(fsno, 1.e-6)), (1.0 / nmelt))) + 1.0));
        intsnow = dmin(1.e8, t);
      }
      if (fsno > 0.0) {
        depth = depth + newsnow / (bifall * fsno);
      } else {
        depth = 0.0;
      }
      if (oldfflag == 1) {
        if (depth > 0.0) fsno = fsca_niu_yang(depth, swe + newsnow);
        if (swe < 1.0) fsno = dmin(fsno, swe);
      }
    } else {
      if (newsnow > 0.0) {
        const double z_avg = newsnow / bifall;
        fsno = m_tanh(accum_factor * newsnow);
        const double t = (swe + newsnow) / (0.5 * (m_cos(PI * m_pow((1.0 - dmax(fsno, 1.e-6)), (1.0 / nmelt))) + 1.0));
        intsnow = dmin(1.e8, t);
        depth = z_avg / fsno;
        if (oldfflag == 1 && depth > 0.0) fsno = fsca_niu_yang(depth, swe + newsnow);
      } else {
        depth = 0.0;
        fsno = 0.0;
      }
    }
    swe = swe + newsnow;
    intsnow = intsnow + newsnow;
    dz_snowf = (depth - depth_before);
  }
  const double fsno_eff = fsno;   // soil/crop land unit with subgridflag == 1

  const int bot = NLEVSNO - 1;
  bool newnode = false;
  if (snl == 0 && snow_grnd > 0.0 && (fsno * depth) >= 0.01) {
    newnode = true;
    snl = 1;
    dz[bot] = depth;
    z[bot] = -0.5 * depth;
    zi[bot] = -depth;
    t_soisno[bot] = dmin(TFRZ, forc_t);
    h2osoi_ice[bot] = swe;
    h2osoi_liq[bot] = 0.0;
    frac_iceold[bot] = 1.0;
    snw_rds[bot] = SNW_RDS_MIN;
  }
  if (snl > 0 && !newnode) {
    const int top = NLEVSNO - snl;
    h2osoi_ice[top] = h2osoi_ice[top] + newsnow;
    dz[top] = dz[top] + dz_snowf;
  }
  snl_io = snl;
  snow_depth = depth;
  h2osno = swe;
  int_snow = intsnow;
  frac_sno = fsno;
  frac_sno_eff = fsno_eff;
}

// ---- fraction_h2osfc (:312-357): fraction of the column covered by standing surface water, 10 fixed Newton steps ----
template <class Row>
ELMK_HD void fraction_h2osfc(const double micro_sigma, const double h2osno, double& h2osfc, const Row h2osoi_liq,
                             double& frac_sno, double& frac_sno_eff, double& frac_h2osfc)
{
  double sfc = h2osfc, fsfc, fsno = frac_sno;
  constexpr double min_h2osfc = 1.e-8;
  if (sfc > min_h2osfc) {
    double d = 0.0;
    const double sigma = 1.0e3 * micro_sigma;
    for (int l = 0; l < 10; ++l) {
      const double fd = 0.5 * d * (1.0 + m_erf(d / (sigma * sqrt(2.0)))) +
                        sigma / sqrt(2.0 * PI) * m_exp(-sq(d) / (2.0 * sq(sigma))) - sfc;
      const double dfdd = 0.5 * (1.0 + m_erf(d / (sigma * sqrt(2.0))));
      d = d - fd / dfdd;
    }
    fsfc = 0.5 * (1.0 + m_erf(d / (sigma * sqrt(2.0))));
  } else {
    fsfc = 0.0;
    h2osoi_liq[NLEVSNO] = h2osoi_liq[NLEVSNO] + sfc;
    sfc = 0.0;
  }
  if (fsno > (1.0 - fsfc) && h2osno > 0.0) {
    if (fsfc > 0.01) {
      fsfc = dmax((1.0 - fsno), 0.01);
      fsno = 1.0 - fsfc;
    } else {
      fsno = 1.0 - fsfc;
    }
    frac_sno_eff = fsno;
  }
  frac_sno = fsno;
  h2osfc = sfc;
  frac_h2osfc = fsfc;
}
} // namespace hyd

// ---- a1 -------------------------------------------------------------------------------------
ELMK_HD void column_frac_wet(const Cols& S, const Tables& T, const int c)
{
  double wet, dry;
  hyd::fraction_wet(C1(frac_veg_nosno), T.dewmx, C1(elai), C1(esai), C1(h2ocan), wet, dry);
  C1(fwet) = wet;
  C1(fdry) = dry;
}

// ---- a3: kokkos_canopy_hydrology = interception, ground_flux (qflx_irrig hard-wired 0.0 in the wrapper), snow_init,
//      fraction_h2osfc (canopy_hydrology_kokkos.cc:7-96) ----------------------------------------
ELMK_HD void column_canopy_hydrology(const Cols& S, const Tables& T, const double dtime, const int c)
{
  const int veg = C1(frac_veg_nosno);
  const int capsnow = C1(do_capsnow);
  const double rain = C1(forc_rain), snow = C1(forc_snow);
  const double forc_t = C1(forc_tbot);

  double candrip, thru_snow, thru_rain, fracsnow, fracrain;
  double canwat = C1(h2ocan);
  hyd::interception(veg, rain, snow, T.dewmx, C1(elai), C1(esai), dtime, canwat, candrip, thru_snow, thru_rain, fracsnow, fracrain);
  if (veg == 1 && (rain + snow) > 0.0) C1(h2ocan) = canwat;

  double snwcp_liq, snwcp_ice, snow_grnd, rain_grnd;
  hyd::ground_flux(capsnow, veg, rain, snow, 0.0, candrip, thru_snow, thru_rain, fracsnow, fracrain, snwcp_liq, snwcp_ice,
                   snow_grnd, rain_grnd);
  C1(qflx_snwcp_liq) = snwcp_liq;
  C1(qflx_snwcp_ice) = snwcp_ice;
  C1(qflx_snow_grnd) = snow_grnd;
  C1(qflx_rain_grnd) = rain_grnd;

  int snl = C1(snl);
  double depth = C1(snow_depth), swe = C1(h2osno), intsnow = C1(int_snow), fsno = C1(frac_sno), fsno_eff;
  hyd::snow_init(dtime, capsnow, T.oldfflag, forc_t, 0.0, snow_grnd, C1(qflx_snow_melt), C1(n_melt), depth, swe, intsnow,
                 ELMK_ROW(swe_old), ELMK_ROW(h2osoi_liq), ELMK_ROW(h2osoi_ice), ELMK_ROW(t_soisno), ELMK_ROW(frac_iceold), snl,
                 ELMK_ROW(dz), ELMK_ROW(zsoi), ELMK_ROW(zisoi), ELMK_ROW(snw_rds), fsno_eff, fsno);

  double sfc = C1(h2osfc), fsfc;
  hyd::fraction_h2osfc(C1(micro_sigma), swe, sfc, ELMK_ROW(h2osoi_liq), fsno, fsno_eff, fsfc);

  C1(snl) = snl;
  C1(snow_depth) = depth;
  C1(h2osno) = swe;
  C1(int_snow) = intsnow;
  C1(frac_sno) = fsno;
  C1(frac_sno_eff) = fsno_eff;
  C1(h2osfc) = sfc;
  C1(frac_h2osfc) = fsfc;
}

} // namespace elmk
