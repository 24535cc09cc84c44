// phys_forcing.h - the per-step producers of the chain's inputs: atmospheric
// forcing from raw time series, satellite phenology from monthly values.
//
// Parity target:
//   ELM::get_forcing                      driver/kokkos/atm_forcing_kokkos.cc:48-63  - eight launches, one per variable
//   ProcessTBOT :40, ProcessPBOT :56, ProcessQBOT :76, ProcessFLDS :101, ProcessFSDS :127, ProcessPREC :160,
//   ProcessWIND :181, ProcessZBOT :203, interp_forcing :213, tdc :219, esatw :224, esati :238
//                                         src/physics/atm_physics_impl.hh
//   ComputePhenology::operator()          src/physics/phenology_physics_impl.hh:20-69
// The eight forcing functors are one pass over the column here, in the order of ELM::get_forcing (QBOT reads the
// TBOT/PBOT results, FLDS reads all three, PREC reads TBOT).  The raw series stay resident in HBM as
// series[t * stride + column] - the reference's own layout, AtmDataManager::data(ntimes, ncells) (atm_data.h) -
// so a step costs two reads per variable and no host traffic.
#pragma once
#include "elmk_state.h"

namespace elmk {

constexpr int ATM_TBOT = 0, ATM_PBOT = 1, ATM_QBOT = 2, ATM_FLDS = 3, ATM_FSDS = 4, ATM_PREC = 5, ATM_WIND = 6, ATM_NVARS = 7;
constexpr int PHEN_MLAI = 0, PHEN_MSAI = 1, PHEN_MHTOP = 2, PHEN_MHBOT = 3, PHEN_NVARS = 4;

// ---- solar geometry of kokkos_init_timestep (init_timestep_kokkos.cc:27-35) ----
// incident_shortwave::average_cosz (src/physics/incident_shortwave.cc:108-116) per column, and ELM::daylength /
// max_daylength (src/physics/day_length.cc:16-39).  The reference evaluates them once for its single site
// (S.lat_r, S.lon_r) on the host and assigns the cosine to every column; here every column may have its own
// coordinates.  The trigonometric functions of the latitude do not change with time: the host evaluates them once
// (elmk_set_coordinates), the per-step declination terms are kernel arguments, and the device is left with the
// half-day arc cosine and four sines per column, from the restatements of libm's routines (elmk_libm.h).
struct SolarStep {
  const double* sin_lat;   // [n] sin(lat), cos(lat), tan(ensure_tan_defined(lat)), longitude [rad]
  const double* cos_lat;
  const double* tan_lat;
  const double* lon;
  int per_column;          // 0: element 0 applies to every column (one site)
  double dtrad, frac2pi, sin_decl, cos_decl, tan_decl;
};
namespace solar {
constexpr double TWO_PI = PI * 2.0, PI_OVER_TWO = PI / 2.0;
ELMK_HD double ensure_tan_defined(const double v) { return (v == PI_OVER_TWO) ? v - 1.0e-05 : (v == -PI_OVER_TWO) ? v + 1.0e-05 : v; }
// declination_angle_sin (:17): host only (its sine is evaluated once per step)
inline double declination(const int doy) { return 23.45 * PI / 180.0 * sin(TWO_PI * (284.0 + doy) / 365.0); }
// ELM::daylength (day_length.cc:16-34).  QUIRK: the clamp `min(offset_pole, max(1.0 * offset_pole, lat))` always
// yields offset_pole, so the latitude does not enter (the lower bound was meant to be -offset_pole); reproduced.
inline double daylength(const double lat, const double decl)
{
  (void)lat;
  constexpr double secs_per_radian = 13750.9871;
  constexpr double lat_epsilon = 10.0 * 2.220446049250313e-16;
  constexpr double offset_pole = PI / 2.0 - lat_epsilon;
  const double my_lat = offset_pole;
  double temp = -(sin(my_lat) * sin(decl)) / (cos(my_lat) * cos(decl));
  temp = dmin(1.0, dmax(-1.0, temp));
  return 2.0 * secs_per_radian * acos(temp);
}
inline double max_daylength(const double lat) { return (lat < 0.0) ? daylength(lat, -0.409571) : daylength(lat, 0.409571); }
} // namespace solar

#if defined(__CUDA_ARCH__)
ELMK_HD_NOINLINE double m_sin(double x) { return lm::g_sin(x); }
#else
ELMK_HD_NOINLINE double m_sin(double x) { return sin(x); }
#endif

ELMK_HD void column_coszen(const Cols& S, const SolarStep& G, const int c)
{
  using namespace solar;
  const int k = G.per_column ? c : 0;
  const double dtrad = G.dtrad;
  // dt_start_rad (:39-44), dt_end_rad (:47-50)
  double t_start = G.frac2pi + G.lon[k] - PI;
  t_start = (t_start >= PI) ? t_start - TWO_PI : (t_start < -PI) ? t_start + TWO_PI : t_start;
  const double t_end = t_start + dtrad;
  // coshalfday (:54-58)
  const double ch = -G.tan_lat[k] * G.tan_decl;
  const double cos_h = (ch <= -1.0) ? PI : (ch >= 1.0) ? 0.0 : m_acos(ch);
  // avg_hourangle (:63-96)
  double ha0, ha1, ha2, ha3;
  if (t_end >= PI && t_start <= PI && PI - cos_h <= dtrad) {
    ha0 = dmin(dmax(t_start, -cos_h), cos_h);
    ha1 = cos_h;
    ha2 = TWO_PI - cos_h;
    ha3 = dmin(dmax(t_end, TWO_PI - cos_h), TWO_PI + cos_h);
  } else if (t_end >= -PI && t_start <= -PI && PI - cos_h <= dtrad) {
    ha0 = dmin(dmax(t_start, -TWO_PI - cos_h), -TWO_PI + cos_h);
    ha1 = -TWO_PI + cos_h;
    ha2 = -cos_h;
    ha3 = dmin(dmax(t_end, -cos_h), cos_h);
  } else {
    if (t_start > PI) ha0 = dmin(dmax(t_start - TWO_PI, -cos_h), cos_h);
    else if (t_start < -PI) ha0 = dmin(dmax(t_start + TWO_PI, -cos_h), cos_h);
    else ha0 = dmin(dmax(t_start, -cos_h), cos_h);
    if (t_end > PI) ha1 = dmin(dmax(t_end - TWO_PI, -cos_h), cos_h);
    else if (t_end < -PI) ha1 = dmin(dmax(t_end + TWO_PI, -cos_h), cos_h);
    else ha1 = dmin(dmax(t_end, -cos_h), cos_h);
    ha2 = 0.0;
    ha3 = 0.0;
  }
  // integrate_cosz (:100-113)
  const double aa = G.sin_lat[k] * G.sin_decl;
  const double bb = G.cos_lat[k] * G.cos_decl;
  double cosz = 0.0;
  if (ha1 > ha0 || ha3 > ha2)
    cosz = (aa * (ha1 - ha0) + bb * (m_sin(ha1) - m_sin(ha0))) / dtrad + (aa * (ha3 - ha2) + bb * (m_sin(ha3) - m_sin(ha2))) / dtrad;
  C1(coszen) = cosz;
}

struct AtmSeries {
  const double* v[ATM_NVARS];
  long long stride;   // elements between consecutive times
};
struct PhenSeries {
  const double* v[PHEN_NVARS];
  long long stride;   // elements between consecutive months
};

// Lowe (1977) saturation vapour pressure polynomials over water / ice [Pa], argument in deg C
ELMK_HD double atm_esatw(const double t)
{
  return 100.0 * (6.107799961 + t * (4.436518521e-01 + t * (1.428945805e-02 + t * (2.650648471e-04 +
                  t * (3.031240396e-06 + t * (2.034080948e-08 + t * 6.136820929e-11))))));
}
ELMK_HD double atm_esati(const double t)
{
  return 100.0 * (6.109177956 + t * (5.034698970e-01 + t * (1.886013408e-02 + t * (4.176223716e-04 +
                  t * (5.824720280e-06 + t * (4.838803174e-08 + t * 1.838826904e-10))))));
}

ELMK_HD void column_atm_forcing(const Cols& S, const AtmSeries& A, const int t, const double wt1, const double wt2,
                                const bool qbot_is_rh, const int c)
{
  const long long i0 = (long long)t * A.stride + c, i1 = i0 + A.stride;
  // temperature and potential temperature
  const double tbot = dmin(A.v[ATM_TBOT][i0] * wt1 + A.v[ATM_TBOT][i1] * wt2, 323.0);
  C1(forc_tbot) = tbot;
  C1(forc_thbot) = tbot;
  // pressure
  const double pbot = dmax(A.v[ATM_PBOT][i0] * wt1 + A.v[ATM_PBOT][i1] * wt2, 4.0e4);
  C1(forc_pbot) = pbot;
  // specific humidity, from relative humidity [%] when the series holds RH
  double qbot = dmax(A.v[ATM_QBOT][i0] * wt1 + A.v[ATM_QBOT][i1] * wt2, 1.0e-9);
  if (qbot_is_rh) {
    const double tc = dmin(50.0, dmax(-50.0, (tbot - TFRZ)));
    const double e = (tbot > TFRZ) ? atm_esatw(tc) : atm_esati(tc);
    const double qsat = 0.622 * e / (pbot - 0.378 * e);
    qbot *= qsat / 100.0;
  }
  C1(forc_qbot) = qbot;
  // downward longwave: the series value when plausible, else a clear-sky estimate
  const double flds = A.v[ATM_FLDS][i0] * wt1 + A.v[ATM_FLDS][i1] * wt2;
  if (flds <= 50.0 || flds >= 600.0) {
    const double e = pbot * qbot / (0.622 + 0.378 * qbot);
    const double ea = 0.70 + 5.95e-5 * 0.01 * e * m_exp(1500.0 / tbot);
    C1(forc_lwrad) = ea * STEBOL * pow4(tbot);
  } else {
    C1(forc_lwrad) = flds;
  }
  // shortwave: not interpolated; direct / diffuse split of the visible and near-infrared halves
  const double sw = dmax(A.v[ATM_FSDS][i0] * C1(coszen) * 0.5, 0.0);
  const double rvis = dmin(0.99, dmax(0.17639 + 0.00380 * sw - 9.0039e-06 * sq(sw) + 8.1351e-09 * cube(sw), 0.01));
  const double rnir = dmin(0.99, dmax(0.29548 + 0.00504 * sw - 1.4957e-05 * sq(sw) + 1.4881e-08 * cube(sw), 0.01));
  C2(forc_solad, 0) = rvis * sw;
  C2(forc_solad, 1) = rnir * sw;
  C2(forc_solai, 0) = (1.0 - rvis) * sw;
  C2(forc_solai, 1) = (1.0 - rnir) * sw;
  // precipitation: not interpolated; rain / snow ramp over 2 K above freezing
  const double frac1 = (tbot - TFRZ) * 0.5;
  const double frac2 = dmin(1.0, dmax(0.0, frac1));
  const double prec = dmax(A.v[ATM_PREC][i0], 0.0);
  C1(forc_rain) = frac2 * prec;
  C1(forc_snow) = (1.0 - frac2) * prec;
  // wind
  C1(forc_u) = A.v[ATM_WIND][i0] * wt1 + A.v[ATM_WIND][i1] * wt2;
  C1(forc_v) = 0.0;
  // forcing height: hard-wired by the reference (ProcessZBOT ignores its series)
  C1(forc_hgt) = 30.0;
  C1(forc_hgt_u_patch) = 30.0;
  C1(forc_hgt_t_patch) = 30.0;
  C1(forc_hgt_q_patch) = 30.0;
}

ELMK_HD void column_phenology(const Cols& S, const PhenSeries& P, const int m, const double wt1, const double wt2, const int c)
{
  const long long i0 = (long long)m * P.stride + c, i1 = i0 + P.stride;
  const int vtype = C1(vtype);
  double tlai, tsai, htop, hbot;
  if (vtype != 0) {
    tlai = wt1 * P.v[PHEN_MLAI][i0] + wt2 * P.v[PHEN_MLAI][i1];
    tsai = wt1 * P.v[PHEN_MSAI][i0] + wt2 * P.v[PHEN_MSAI][i1];
    htop = wt1 * P.v[PHEN_MHTOP][i0] + wt2 * P.v[PHEN_MHTOP][i1];
    hbot = wt1 * P.v[PHEN_MHBOT][i0] + wt2 * P.v[PHEN_MHBOT][i1];
  } else {
    tlai = 0.0; tsai = 0.0; htop = 0.0; hbot = 0.0;
  }
  C1(tlai) = tlai; C1(tsai) = tsai; C1(htop) = htop; C1(hbot) = hbot;
  // burial by snow: trees and shrubs (vtype 1..11) by the buried share of the crown, grasses and crops within 0.2 m
  const double snow_depth = C1(snow_depth), fsno = C1(frac_sno);
  double fb;
  if (vtype > 0 && vtype <= 11) {
    const double ol = dmin(dmax(snow_depth - hbot, 0.0), htop - hbot);
    fb = 1.0 - ol / dmax(1.e-06, htop - hbot);
  } else {
    fb = 1.0 - dmax(dmin(snow_depth, 0.2), 0.0) / 0.2;
  }
  double elai = dmax(tlai * (1.0 - fsno) + tlai * fb * fsno, 0.0);
  double esai = dmax(tsai * (1.0 - fsno) + tsai * fb * fsno, 0.0);
  if (elai < 0.05) elai = 0.0;
  if (esai < 0.05) esai = 0.0;
  C1(elai) = elai;
  C1(esai) = esai;
  C1(frac_veg_nosno_alb) = ((elai + esai) >= 0.05) ? 1 : 0;
}

} // namespace elmk
