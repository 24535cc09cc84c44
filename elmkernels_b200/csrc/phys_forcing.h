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
