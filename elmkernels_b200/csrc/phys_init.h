// phys_init.h - one-time cold-start initialisation of a column, soil / crop land
// units (the only ones the hot path accepts).
//
// Parity target: the per-column lambda of initialize_kokkos_elm (driver/kokkos/initialize_elm_kokkos.cc:374-431):
//   PFTData::get_pft_psn                       src/data/pft_data_impl.hh:67-98
//   init_topo_slope :7, init_melt_factor :14, init_micro_sigma :30         src/physics/init_topography_impl.hh
//   init_snow_layers :68, init_snow_state :12                               src/physics/init_snow_state_impl.hh
//   pedotransfer :6, soil_hydraulic_params :18, init_soil_hydraulics :92    src/physics/soil_texture_hydraulic_model_impl.hh
//   init_vegrootfr :190, init_soil_temp :11, init_soilh2o_state :62         src/physics/init_soil_state_impl.hh
// Inputs besides the state (vtype, raw topo_slope, topo_std, the soil part of dz / zsoi / zisoi): soil texture per
// layer and the initial snow depth.  init_snow_state is a cold start: it zeroes snow_depth / h2osno again after
// init_snow_layers has built the layers from the depth - reproduced.
#pragma once
#include "elmk_state.h"

namespace elmk {

struct InitInputs {
  const double* pct_sand;   // [NLEVGRND][stride]
  const double* pct_clay;
  const double* organic;
  const double* snow_depth; // [ncols]
  long long stride;
  double organic_max;
};

namespace init {
constexpr double SPVAL = 1.0e36;

// soil_hydraulic_params for one layer
ELMK_HD void soil_layer(const double pct_sand, const double pct_clay, const double zsoi, const double om_frac,
                        double& watsat, double& bsw, double& sucsat, double& watdry, double& watopt, double& watfc,
                        double& tkmg, double& tkdry, double& csol)
{
  constexpr double zsapric = 0.5, pcalpha = 0.5, pcbeta = 0.139, om_tkd = 0.05, om_tkm = 0.25, om_csol = 2.5;
  // Cosby et al. (1984), table 5
  watsat = 0.489 - 0.00126 * pct_sand;
  bsw = 2.91 + 0.159 * pct_clay;
  sucsat = 10.0 * m_pow(10.0, (1.88 - 0.0131 * pct_sand));
  const double xksat = 0.0070556 * m_pow(10.0, (-0.884 + 0.0153 * pct_sand));
  const double om_watsat = dmax(0.93 - 0.1 * (zsoi / zsapric), 0.83);
  const double om_b = dmin(2.7 + 9.3 * (zsoi / zsapric), 12.0);
  const double om_sucsat = dmin(10.3 - 0.2 * (zsoi / zsapric), 10.1);
  const double om_hksat = dmax(0.28 - 0.2799 * (zsoi / zsapric), 0.0001);
  const double bulk_den = (1.0 - watsat) * 2.7e3;
  const double tkm = (1.0 - om_frac) * (8.8 * pct_sand + 2.92 * pct_clay) / (pct_sand + pct_clay) + om_tkm * om_frac;
  watsat = (1.0 - om_frac) * watsat + om_watsat * om_frac;
  bsw = (1.0 - om_frac) * (2.91 + 0.159 * pct_clay) + om_frac * om_b;
  sucsat = (1.0 - om_frac) * sucsat + om_sucsat * om_frac;
  double perc_frac;
  if (om_frac > pcalpha) {
    const double perc_norm = m_pow((1.0 - pcalpha), -pcbeta);
    perc_frac = perc_norm * m_pow((om_frac - pcalpha), pcbeta);
  } else {
    perc_frac = 0.0;
  }
  const double uncon_frac = (1.0 - om_frac) + (1.0 - perc_frac) * om_frac;
  double uncon_hksat;
  if (om_frac < 1.0) {
    uncon_hksat = uncon_frac / ((1.0 - om_frac) / xksat + ((1.0 - perc_frac) * om_frac) / om_hksat);
  } else {
    uncon_hksat = 0.0;
  }
  const double hksat = uncon_frac * uncon_hksat + (perc_frac * om_frac) * om_hksat;
  tkmg = m_pow(tkm, (1.0 - watsat));
  tkdry = ((0.135 * bulk_den + 64.7) / (2.7e3 - 0.947 * bulk_den)) * (1.0 - om_frac) + om_tkd * om_frac;
  csol = ((1.0 - om_frac) * (2.128 * pct_sand + 2.385 * pct_clay) / (pct_sand + pct_clay) + om_csol * om_frac) * 1.0e6;
  watdry = watsat * m_pow((316230.0 / sucsat), (-1.0 / bsw));
  watopt = watsat * m_pow((158490.0 / sucsat), (-1.0 / bsw));
  watfc = watsat * m_pow((0.1 / (hksat * 86400.0)), (1.0 / (2.0 * bsw + 3.0)));
}
} // namespace init

ELMK_HD void column_init(const Cols& S, const Tables& T, const InitInputs& X, const int c)
{
  using namespace init;
  const int vtype = C1(vtype);
  const double snow_depth = X.snow_depth[c];

  // photosynthesis constants of the column's PFT
#pragma unroll
  for (int k = 0; k < 27; ++k) C2(psn_pft, k) = T.psn[k][vtype];

  // topography
  const double topo_slope = dmax(C1(topo_slope), 0.2);
  C1(topo_slope) = topo_slope;
  C1(n_melt) = 200.0 / dmax(10.0, C1(topo_std));
  {
    const double slopebeta = 3.0, slopemax = 0.4;
    const double slope0 = m_pow(slopemax, (-1.0 / slopebeta));
    C1(micro_sigma) = m_pow((topo_slope + slope0), -slopebeta);
  }

  // snow layers from the initial depth
  double dz[NLEVSNO], z[NLEVSNO], zi[NLEVSNO + 1];
  int snl = 0;
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i) { dz[i] = SPVAL; z[i] = SPVAL; zi[i] = SPVAL; }
  zi[NLEVSNO] = C2(zisoi, NLEVSNO);
  if (snow_depth < 0.01) {
    snl = 0;
#pragma unroll
    for (int i = 0; i < NLEVSNO; ++i) { dz[i] = 0.0; z[i] = 0.0; zi[i] = 0.0; }
    zi[NLEVSNO] = 0.0;
  } else {
    if ((snow_depth >= 0.01) && (snow_depth <= 0.03)) {
      snl = 1; dz[4] = snow_depth;
    } else if ((snow_depth > 0.03) && (snow_depth <= 0.04)) {
      snl = 2; dz[3] = snow_depth / 2.0; dz[4] = dz[3];
    } else if ((snow_depth > 0.04) && (snow_depth <= 0.07)) {
      snl = 2; dz[3] = 0.02; dz[4] = snow_depth - dz[3];
    } else if ((snow_depth > 0.07) && (snow_depth <= 0.12)) {
      snl = 3; dz[2] = 0.02; dz[3] = (snow_depth - 0.02) / 2.0; dz[4] = dz[3];
    } else if ((snow_depth > 0.12) && (snow_depth <= 0.18)) {
      snl = 3; dz[2] = 0.02; dz[3] = 0.05; dz[4] = snow_depth - dz[2] - dz[3];
    } else if ((snow_depth > 0.18) && (snow_depth <= 0.29)) {
      snl = 4; dz[1] = 0.02; dz[2] = 0.05; dz[3] = (snow_depth - dz[1] - dz[2]) / 2.0; dz[4] = dz[3];
    } else if ((snow_depth > 0.29) && (snow_depth <= 0.41)) {
      snl = 4; dz[1] = 0.02; dz[2] = 0.05; dz[3] = 0.11; dz[4] = snow_depth - dz[1] - dz[2] - dz[3];
    } else if ((snow_depth > 0.41) && (snow_depth <= 0.64)) {
      snl = 5; dz[0] = 0.02; dz[1] = 0.05; dz[2] = 0.11; dz[3] = (snow_depth - dz[0] - dz[1] - dz[2]) / 2.0; dz[4] = dz[3];
    } else if (snow_depth > 0.64) {
      snl = 5; dz[0] = 0.02; dz[1] = 0.05; dz[2] = 0.11; dz[3] = 0.23; dz[4] = snow_depth - dz[0] - dz[1] - dz[2] - dz[3];
    }
  }
#pragma unroll
  for (int j = NLEVSNO - 1; j >= 0; --j) {
    if (j >= NLEVSNO - snl) {
      z[j] = zi[j + 1] - 0.5 * dz[j];
      zi[j] = zi[j + 1] - dz[j];
    }
  }
  C1(snl) = snl;
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i) { C2(dz, i) = dz[i]; C2(zsoi, i) = z[i]; C2(zisoi, i) = zi[i]; }
  C2(zisoi, NLEVSNO) = zi[NLEVSNO];

  // soil hydraulic and thermal parameters; the five bedrock layers repeat the texture of the deepest soil layer
  double watsat_l[NLEVGRND];
#pragma unroll 1
  for (int i = 0; i < NLEVGRND; ++i) {
    const int src = (i < NLEVSOI) ? i : NLEVSOI - 1;
    const double sand = X.pct_sand[(long long)src * X.stride + c], clay = X.pct_clay[(long long)src * X.stride + c];
    const double om_frac = (i < NLEVSOI) ? sq(X.organic[(long long)i * X.stride + c] / X.organic_max) : 0.0;
    double watsat, bsw, sucsat, watdry, watopt, watfc, tkmg, tkdry, csol;
    soil_layer(sand, clay, C2(zsoi, i + NLEVSNO), om_frac, watsat, bsw, sucsat, watdry, watopt, watfc, tkmg, tkdry, csol);
    if (i >= NLEVSOI) csol = 2.0e6;
    C2(watsat, i) = watsat; C2(bsw, i) = bsw; C2(sucsat, i) = sucsat; C2(watdry, i) = watdry; C2(watopt, i) = watopt;
    C2(watfc, i) = watfc; C2(tkmg, i) = tkmg; C2(tkdry, i) = tkdry; C2(csol, i) = csol;
    watsat_l[i] = watsat;
  }

  // root fractions (Zeng et al. 1998)
  if (vtype != 0) {
    const double ra = T.roota[vtype], rb = T.rootb[vtype];
#pragma unroll 1
    for (int i = 0; i < NLEVSOI - 1; ++i) {
      const double zt = C2(zisoi, i + NLEVSNO), zb = C2(zisoi, i + 1 + NLEVSNO);
      C2(rootfr, i) = 0.5 * (m_exp(-ra * zt) + m_exp(-rb * zt) - m_exp(-ra * zb) - m_exp(-rb * zb));
    }
    const double zl = C2(zisoi, NLEVSOI - 1 + NLEVSNO);
    C2(rootfr, NLEVSOI - 1) = 0.5 * (m_exp(-ra * zl) + m_exp(-rb * zl));
  } else {
#pragma unroll
    for (int i = 0; i < NLEVSOI; ++i) C2(rootfr, i) = 0.0;
  }
#pragma unroll
  for (int i = NLEVSOI; i < NLEVGRND; ++i) C2(rootfr, i) = 0.0;

  // temperatures: 250 K in snow, 274 K in soil; t_grnd = top active layer
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i)
    if (i >= NLEVSNO - snl) C2(t_soisno, i) = 250.0;
#pragma unroll
  for (int i = NLEVSNO; i < NLEVTOT; ++i) C2(t_soisno, i) = 274.0;
  C1(t_grnd) = (snl > 0) ? 250.0 : 274.0;

  // cold-start snow state
  C1(h2osno) = 0.0; C1(int_snow) = 0.0; C1(snow_depth) = 0.0; C1(h2osfc) = 0.0; C1(h2ocan) = 0.0;
  C1(frac_h2osfc) = 0.0; C1(fwet) = 0.0; C1(fdry) = 0.0;
  C1(frac_sno) = 0.0;   // (snow_depth is zero by now, so the Niu-Yang branch never runs)
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i) C2(snw_rds, i) = (snl > 0 && i >= NLEVSNO - snl) ? SNW_RDS_MIN : 0.0;

  // soil water: 0.15 m3/m3 capped at saturation, frozen below TFRZ (all soil starts at 274 K: liquid); snow layers
  // hold ice at 250 kg/m3; slots above the snow pack keep the reference's 1e36 marker
#pragma unroll
  for (int i = 0; i < NLEVGRND; ++i) {
    const double vol = dmin(0.15, watsat_l[i]);
    C2(h2osoi_vol, i) = vol;
    C2(h2osoi_ice, i + NLEVSNO) = 0.0;
    C2(h2osoi_liq, i + NLEVSNO) = C2(dz, i + NLEVSNO) * DENH2O * vol;
  }
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i) {
    if (i >= NLEVSNO - snl) {
      C2(h2osoi_ice, i) = dz[i] * 250.0;
      C2(h2osoi_liq, i) = 0.0;
    } else {
      C2(h2osoi_ice, i) = SPVAL;
      C2(h2osoi_liq, i) = SPVAL;
    }
  }
}

} // namespace elmk
