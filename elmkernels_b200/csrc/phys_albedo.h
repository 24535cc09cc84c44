// phys_albedo.h - surface albedo group (a2): soil albedo, SNICAR snow radiative transfer
// (direct and diffuse), ground albedo, snow-layer absorption factors, two-stream canopy solution.
//
// Parity target: kokkos_albedo_snicar, reference
// driver/kokkos/albedo_kokkos.cc:10-377, which chains
//   surface_albedo::init_timestep :90, soil_albedo :690, ground_albedo :155,
//   flux_absorption_factor :171, canopy_layer_lai :215, two_stream_solver :323
//                                                        (src/physics/surface_albedo_impl.hh)
//   snow_snicar::init_timestep :9, snow_aerosol_mie_params :107,
//   snow_radiative_transfer_solver :313, snow_albedo_radiation_factor :673, each called twice
//                                                        (src/physics/snow_snicar_impl.hh)
// B200 shape: one thread owns one column; all per-layer / per-band intermediates (the 18 scratch
// Views the wrapper allocates per call, albedo_kokkos.cc:19-38) live in registers - layer loops are
// written over the fixed five snow slots with a predicate instead of [snl_top, snl_btm] bounds so
// that every array index is a compile-time constant after unrolling.  Night columns leave after
// writing the initial values.
#pragma once
#include "elmk_state.h"

// ELMK_SNICAR_ROLLED (experiment): the layer loops of one band solve stay loops
#if defined(__CUDA_ARCH__) && defined(ELMK_SNICAR_ROLLED)
#define ELMK_SNICAR_LOOP _Pragma("unroll 1")
#else
#define ELMK_SNICAR_LOOP _Pragma("unroll")
#endif
namespace elmk {

namespace alb {
constexpr double MPE = 1.e-06;
constexpr double EXTKN = 0.30;
constexpr double MIN_SNW = 1.0e-30;
constexpr int RDS_MIN_TBL = 30;
constexpr int RDS_MAX_TBL = 1500;
constexpr int MIE_N = 1471;
} // namespace alb

// SNICAR for one incident-flux type.  flg = 1 direct beam, 2 diffuse.
// cnc[i][j]: aerosol mass concentrations per snow slot.  Outputs: alb_out[2] (VIS, NIR) and
// flx_abs[6][2] (five snow slots + ground, VIS/NIR), both already zero on entry.
// Two Gauss points of the diffuse delta-Eddington integration (snow_snicar_impl.hh:459-478) in one called function
// with its divisions and exponentials IN LINE (the build leaves functions named *_inl alone, ptx_rewrite.py): the
// 2 x (three divisions + one exp) are independent of each other, so ptxas interleaves eight dependent chains
// where the loop body, written with called m_div / m_exp, ran them one after the other.  None of the numerators
// can be zero (optical depth > 0, 1 + g (1 - w) >= 1), so the in-line divisions stay on their fast path.
// Direct-beam delta-Eddington solution of one layer (snow_snicar_impl.hh:430-457), same arrangement: one called
// function, divisions / exponentials / square root in line, so that the independent ones overlap.
struct LayerDirect {
  double lm, rdif_a, tdif_a, trnlay, rdir, tdir;
};
ELMK_HD_NOINLINE LayerDirect snicar_layer_direct_inl(const double ts, const double ws, const double gs, const double mu_not,
                                                     const double exp_min)
{
  LayerDirect o;
  const double lm = sqrt(3.0 * (1.0 - ws) * (1.0 - ws * gs));
  const double ue = 1.5 * (1.0 - ws * gs) / lm;
  const double extins = dmax(exp_min, i_exp(-lm * ts));
  const double ne = ((ue + 1.0) * (ue + 1.0) / extins) - ((ue - 1.0) * (ue - 1.0) * extins);
  o.lm = lm;
  o.rdif_a = (sq(ue) - 1.0) * (1.0 / extins - extins) / ne;
  o.tdif_a = 4.0 * ue / ne;
  o.trnlay = dmax(exp_min, i_exp(-ts / mu_not));
  const double alp = 0.75 * ws * mu_not * ((1.0 + gs * (1.0 - ws)) / (1.0 - lm * lm * mu_not * mu_not));
  const double gam = 0.5 * ws * ((1.0 + 3.0 * gs * (1.0 - ws) * mu_not * mu_not) / (1.0 - lm * lm * mu_not * mu_not));
  const double apg = alp + gam;
  const double amg = alp - gam;
  o.rdir = apg * o.rdif_a + amg * (o.tdif_a * o.trnlay - 1.0);
  o.tdir = apg * o.tdif_a + (amg * o.rdif_a - apg + 1.0) * o.trnlay;
  return o;
}

struct GaussPair {
  double rdr0, tdr0, rdr1, tdr1;
};
ELMK_HD_NOINLINE GaussPair snicar_gauss_pair_inl(const double ts, const double ws, const double gs, const double lm,
                                                 const double R1, const double T1, const double mu0, const double mu1,
                                                 const double exp_min)
{
  GaussPair g;
  {
    const double mu = mu0;
    const double trn = dmax(exp_min, i_exp(-ts / mu));
    const double alp = 0.75 * ws * mu * ((1.0 + gs * (1.0 - ws)) / (1.0 - lm * lm * mu * mu));
    const double gam = 0.5 * ws * ((1.0 + 3.0 * gs * (1.0 - ws) * mu * mu) / (1.0 - lm * lm * mu * mu));
    const double apg = alp + gam;
    const double amg = alp - gam;
    g.rdr0 = apg * R1 + amg * T1 * trn - amg;
    g.tdr0 = apg * T1 + amg * R1 * trn - apg * trn + trn;
  }
  {
    const double mu = mu1;
    const double trn = dmax(exp_min, i_exp(-ts / mu));
    const double alp = 0.75 * ws * mu * ((1.0 + gs * (1.0 - ws)) / (1.0 - lm * lm * mu * mu));
    const double gam = 0.5 * ws * ((1.0 + 3.0 * gs * (1.0 - ws) * mu * mu) / (1.0 - lm * lm * mu * mu));
    const double apg = alp + gam;
    const double amg = alp - gam;
    g.rdr1 = apg * R1 + amg * T1 * trn - amg;
    g.tdr1 = apg * T1 + amg * R1 * trn - apg * trn + trn;
  }
  return g;
}

// The snow column as SNICAR sees it (snow_snicar::init_timestep :37-100): ice / liquid mass and the grain radius of
// the five snow slots, top = index of the top slot; a column without explicit layers is one layer of h2osno.
struct SnicarColumn {
  double ice[NLEVSNO], liq[NLEVSNO];
  int rds[NLEVSNO];
  int top;
};
ELMK_HD void snicar_column(const Cols& S, const int c, const double h2osno, const int snl, SnicarColumn& K, uint32_t& err)
{
  using namespace alb;
  int snl_lcl;
  if (snl == 0) {
    snl_lcl = 1;
#pragma unroll
    for (int i = 0; i < NLEVSNO; ++i) { K.ice[i] = 0.0; K.liq[i] = 0.0; K.rds[i] = 0; }
    K.ice[NLEVSNO - 1] = h2osno;
    K.liq[NLEVSNO - 1] = 0.0;
    K.rds[NLEVSNO - 1] = (int)round(SNW_RDS_MIN);
  } else {
    snl_lcl = snl;
#pragma unroll
    for (int i = 0; i < NLEVSNO; ++i) {
      K.liq[i] = C2(h2osoi_liq, i);
      K.ice[i] = C2(h2osoi_ice, i);
      K.rds[i] = (int)round(C2(snw_rds, i));
    }
  }
  K.top = NLEVSNO - snl_lcl;   // index of the top snow slot; bottom slot is NLEVSNO-1
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i)
    if (i >= K.top && ((K.rds[i] < RDS_MIN_TBL) || (K.rds[i] > RDS_MAX_TBL))) err |= ERR_SNICAR_RADIUS;
}

// aerosol mass concentrations of snow slot i as SNICAR orders them (OC ignored, :144-145)
ELMK_HD void snicar_cnc(const Cols& S, const int c, const int i, double (&cnc)[NAER])
{
  cnc[0] = C2(cnc_bcphi, i);
  cnc[1] = C2(cnc_bcpho, i);
  cnc[2] = 0.0;
  cnc[3] = 0.0;
  cnc[4] = C2(cnc_dst1, i);
  cnc[5] = C2(cnc_dst2, i);
  cnc[6] = C2(cnc_dst3, i);
  cnc[7] = C2(cnc_dst4, i);
}

// ONE spectral band b of one incident-flux type (flg = 1 direct beam, 2 diffuse): weighted Mie parameters, delta
// transform, Delta-Eddington adding-doubling (snow_radiative_transfer_solver :213-666 for one value of its band
// loop).  Outputs: the band's albedo and absorbed fluxes of the five snow slots + the ground.  The ten (flg, b)
// combinations of a column are independent of each other: the one-thread-per-column path runs them one after the
// other (snicar_solve below), the SNICAR kernel of the CUDA library gives each its own lane.
// CNC: callable (i, cnc[8]) that fills the aerosol concentrations of slot i (zeros for b >= 3, :146-152).
template <class CNC>
ELMK_HD void snicar_band(const Tables& T, const SnicarColumn& K, const int flg, const int b, const double mu_not,
                         const double (&albsoi)[NUMRAD], const CNC& cnc_of, double& albedo_out,
                         double (&flx_abs_b)[NLEVSNO + 1], uint32_t& err)
{
  using namespace alb;
  const int top = K.top;
  const double flx_slrd = (flg == 1) ? 1.0 / (mu_not * PI) : 0.0;
  const double flx_slri = (flg == 1) ? 0.0 : 1.0;
  const int d = (flg == 1) ? 0 : 1;    // drc / dfs optics
  // Gaussian quadrature for the diffuse re-integration (:349-352)
  const double gauspt[8] = {0.9894009, 0.9445750, 0.8656312, 0.7554044, 0.6178762, 0.4580168, 0.2816036, 0.0950125};
  const double gauswt[8] = {0.0271525, 0.0622535, 0.0951585, 0.1246290, 0.1495960, 0.1691565, 0.1826034, 0.1894506};
  constexpr double puny = 1.0e-11;
  constexpr double exp_min = 0x1.7cd79b5647c9bp-15;   // exp(-10), the reference's constant-folded value (:357)
  constexpr double trmin = 0.001;
  // BC optics indices: fixed 100 nm effective radii (:127-128,236-237) -> round(100/50) - 1 = 1
  constexpr int idx_nclrds = 1;
  const bool with_aer = (b < 3);   // aerosol concentrations are zeroed for bands 3 and 4 (:146-152)

  // ---- weighted Mie parameters and delta transform per layer (:213-305) ----
  double ts_[NLEVSNO], ws_[NLEVSNO], gs_[NLEVSNO];
ELMK_SNICAR_LOOP
  for (int i = 0; i < NLEVSNO; ++i) {
    ts_[i] = 0.0; ws_[i] = 0.0; gs_[i] = 0.0;
    if (i >= top) {
      const int ridx = K.rds[i] - RDS_MIN_TBL;
      const double ss_snw = T.snw[d][0][b * MIE_N + ridx];
      const double asm_snw = T.snw[d][1][b * MIE_N + ridx];
      const double ext_snw = T.snw[d][2][b * MIE_N + ridx];
      int idx_icerds;
      if (K.rds[i] < 125) idx_icerds = K.rds[i] / 50 - 1;
      else if (K.rds[i] < 175) idx_icerds = 1;
      else idx_icerds = (K.rds[i] / 250) + 2 - 1;
      idx_icerds = imin(imax(idx_icerds, 0), 7);
      const double enh = T.bcenh[idx_icerds][idx_nclrds][b];

      const double L_snw = K.ice[i] + K.liq[i];
      const double tau_snw = L_snw * ext_snw;
      double tau_sum = 0.0, omega_sum = 0.0, g_sum = 0.0;
      double cnc[NAER];
      if (with_aer) {
        cnc_of(i, cnc);
      } else {
ELMK_SNICAR_LOOP
        for (int j = 0; j < NAER; ++j) cnc[j] = 0.0;
      }
ELMK_SNICAR_LOOP
      for (int j = 0; j < NAER; ++j) {
        double ss, as, ex;
        if (j == 0) { ss = T.bc[0][0][idx_nclrds][b]; as = T.bc[0][1][idx_nclrds][b]; ex = T.bc[0][2][idx_nclrds][b] * enh; }
        else if (j == 1) { ss = T.bc[1][0][idx_nclrds][b]; as = T.bc[1][1][idx_nclrds][b]; ex = T.bc[1][2][idx_nclrds][b]; }
        else { ss = T.aer_band[j - 2][0][b]; as = T.aer_band[j - 2][1][b]; ex = T.aer_band[j - 2][2][b]; }
        const double L_aer = L_snw * cnc[j];
        const double tau_aer = L_aer * ex;
        tau_sum += tau_aer;
        omega_sum += (tau_aer * ss);
        g_sum += (tau_aer * ss * as);
      }
      const double tau = tau_sum + tau_snw;
      const double omega = (1.0 / tau) * (omega_sum + (ss_snw * tau_snw));
      const double g = (1.0 / (tau * omega)) * (g_sum + (asm_snw * ss_snw * tau_snw));
      gs_[i] = g / (1.0 + g);
      ws_[i] = ((1.0 - sq(g)) * omega) / (1.0 - (omega * sq(g)));
      ts_[i] = (1.0 - (omega * sq(g))) * tau;
    }
  }

  // ---- Delta-Eddington adding-doubling (:384-666) ----
  double trndir[NLEVSNO + 1], trntdr[NLEVSNO + 1], trndif[NLEVSNO + 1], rupdir[NLEVSNO + 1], rupdif[NLEVSNO + 1],
      rdndif[NLEVSNO + 1];
  double rdir[NLEVSNO], rdif_a[NLEVSNO], rdif_b[NLEVSNO], tdir[NLEVSNO], tdif_a[NLEVSNO], tdif_b[NLEVSNO],
      trnlay[NLEVSNO];
ELMK_SNICAR_LOOP
  for (int i = 0; i <= NLEVSNO; ++i) {
    trndir[i] = 0.0; trntdr[i] = 0.0; trndif[i] = 0.0; rupdir[i] = 0.0; rupdif[i] = 0.0; rdndif[i] = 0.0;
  }
ELMK_SNICAR_LOOP
  for (int i = 0; i <= NLEVSNO; ++i)
    if (i == top) { trndir[i] = 1.0; trntdr[i] = 1.0; trndif[i] = 1.0; rdndif[i] = 0.0; }

ELMK_SNICAR_LOOP
  for (int i = 0; i < NLEVSNO; ++i) {
    rdir[i] = 0.0; rdif_a[i] = 0.0; rdif_b[i] = 0.0; tdir[i] = 0.0; tdif_a[i] = 0.0; tdif_b[i] = 0.0; trnlay[i] = 0.0;
    if (i >= top) {
      if (trntdr[i] > trmin) {
        const double ts = ts_[i], ws = ws_[i], gs = gs_[i];
        const LayerDirect ld = snicar_layer_direct_inl(ts, ws, gs, mu_not, exp_min);
        const double lm = ld.lm;
        rdif_a[i] = ld.rdif_a;
        tdif_a[i] = ld.tdif_a;
        trnlay[i] = ld.trnlay;
        rdir[i] = ld.rdir;
        tdir[i] = ld.tdir;
        const double R1 = rdif_a[i];
        const double T1 = tdif_a[i];
        double swt = 0.0, smr = 0.0, smt = 0.0;
#pragma unroll 1
        for (int ng = 0; ng < 8; ng += 2) {
          const double mu0 = gauspt[ng], gwt0 = gauswt[ng], mu1 = gauspt[ng + 1], gwt1 = gauswt[ng + 1];
          const GaussPair g = snicar_gauss_pair_inl(ts, ws, gs, lm, R1, T1, mu0, mu1, exp_min);
          swt = swt + mu0 * gwt0;
          smr = smr + mu0 * g.rdr0 * gwt0;
          smt = smt + mu0 * g.tdr0 * gwt0;
          swt = swt + mu1 * gwt1;
          smr = smr + mu1 * g.rdr1 * gwt1;
          smt = smt + mu1 * g.tdr1 * gwt1;
        }
        rdif_a[i] = smr / swt;
        tdif_a[i] = smt / swt;
        rdif_b[i] = rdif_a[i];
        tdif_b[i] = tdif_a[i];
      }
      trndir[i + 1] = trndir[i] * trnlay[i];
      const double refkm1 = 1.0 / (1.0 - rdndif[i] * rdif_a[i]);
      const double tdrrdir = trndir[i] * rdir[i];
      const double tdndif = trntdr[i] - trndir[i];
      trntdr[i + 1] = trndir[i] * tdir[i] + (tdndif + tdrrdir * rdndif[i]) * refkm1 * tdif_a[i];
      rdndif[i + 1] = rdif_b[i] + (tdif_b[i] * rdndif[i] * refkm1 * tdif_a[i]);
      trndif[i + 1] = trndif[i] * refkm1 * tdif_a[i];
    }
  }

  // underlying ground: VIS albedo for band 0, NIR otherwise (:526-531)
  rupdir[NLEVSNO] = (b == 0) ? albsoi[0] : albsoi[1];
  rupdif[NLEVSNO] = rupdir[NLEVSNO];
ELMK_SNICAR_LOOP
  for (int i = NLEVSNO - 1; i >= 0; --i) {
    if (i >= top) {
      const double refkp1 = 1.0 / (1.0 - rdif_b[i] * rupdif[i + 1]);
      rupdir[i] = rdir[i] + (trnlay[i] * rupdir[i + 1] + (tdir[i] - trnlay[i]) * rupdif[i + 1]) * refkp1 * tdif_b[i];
      rupdif[i] = rdif_a[i] + tdif_a[i] * rupdif[i + 1] * refkp1 * tdif_b[i];
    }
  }

  // net (down - up) flux at every interface; absorbed flux per layer
  double dftmp[NLEVSNO + 1];
  double albedo = 0.0, F_sfc_pls = 0.0;
ELMK_SNICAR_LOOP
  for (int i = 0; i <= NLEVSNO; ++i) {
    dftmp[i] = 0.0;
    if (i >= top) {
      const double refk = 1.0 / (1.0 - rdndif[i] * rupdif[i]);
      double dfdir = trndir[i] + (trntdr[i] - trndir[i]) * (1.0 - rupdif[i]) * refk -
                     trndir[i] * rupdir[i] * (1.0 - rdndif[i]) * refk;
      if (dfdir < puny) dfdir = 0.0;
      double dfdif = trndif[i] * (1.0 - rupdif[i]) * refk;
      if (dfdif < puny) dfdif = 0.0;
      dftmp[i] = (flg == 1) ? dfdir : dfdif;
      if (i == top) {
        if (flg == 1) {
          albedo = rupdir[i];
          F_sfc_pls = (trndir[i] * rupdir[i] + (trntdr[i] - trndir[i]) * rupdif[i]) * refk;
        } else {
          albedo = rupdif[i];
          F_sfc_pls = trndif[i] * rupdif[i] * refk;
        }
      }
    }
  }
  double F_abs_sum = 0.0;
ELMK_SNICAR_LOOP
  for (int i = 0; i <= NLEVSNO; ++i) flx_abs_b[i] = 0.0;
ELMK_SNICAR_LOOP
  for (int i = 0; i < NLEVSNO; ++i) {
    if (i >= top) {
      const double F_abs = dftmp[i] - dftmp[i + 1];
      flx_abs_b[i] = F_abs;
      if (F_abs < -0.00001) err |= ERR_SNICAR_NEGABS;
      F_abs_sum = F_abs_sum + F_abs;
    }
  }
  const double F_btm_net = dftmp[NLEVSNO];
  flx_abs_b[NLEVSNO] = F_btm_net;
  // (flg_nosnl == 1 re-stores the same two values, :628-639)
ELMK_SNICAR_LOOP
  for (int i = 0; i <= NLEVSNO; ++i)
    if (i >= top && flx_abs_b[i] < 0.0) flx_abs_b[i] = 0.0;
  const double energy_sum = (mu_not * PI * flx_slrd) + flx_slri - (F_abs_sum + F_btm_net + F_sfc_pls);
  if (fabs(energy_sum) > 0.00001) err |= ERR_SNICAR_ENERGY;
  albedo_out = albedo;
  if (albedo > 1.0) err |= ERR_SNICAR_ALBEDO;
}

// Band weighting of the five band results to VIS / NIR (snow_albedo_radiation_factor :706-760).
// albout[b], flx_abs_lcl[i][b]: the band results; alb_out[2], flx_abs[6][2]: VIS / NIR (flx_abs zero on entry).
ELMK_HD void snicar_combine(const int flg, const double mu_not, const int top, const int rds_top,
                            const double (&albout_lcl)[NBND_SNW], const double (&flx_abs_lcl)[NLEVSNO + 1][NBND_SNW],
                            double (&alb_out)[NUMRAD], double (&flx_abs)[NLEVSNO + 1][NUMRAD])
{
  double wgt[NBND_SNW];
  wgt[0] = 1.0;
  if (flg == 1) {
    wgt[1] = 0.49352158521175; wgt[2] = 0.18099494230665; wgt[3] = 0.12094898498813; wgt[4] = 0.20453448749347;
  } else {
    wgt[1] = 0.58581507618433; wgt[2] = 0.20156903770812; wgt[3] = 0.10917889346386; wgt[4] = 0.10343699264369;
  }
  alb_out[0] = albout_lcl[0];
  double flx_sum = 0.0, wgt_sum = 0.0;
#pragma unroll
  for (int b = 1; b < NBND_SNW; ++b) {
    flx_sum += wgt[b] * albout_lcl[b];
    wgt_sum += wgt[b];
  }
  alb_out[1] = flx_sum / wgt_sum;
#pragma unroll
  for (int i = 0; i <= NLEVSNO; ++i) {
    flx_abs[i][0] = flx_abs_lcl[i][0];
    if (i >= top) {
      flx_sum = 0.0;
#pragma unroll
      for (int b = 1; b < NBND_SNW; ++b) flx_sum += wgt[b] * flx_abs_lcl[i][b];
      flx_abs[i][1] = flx_sum / wgt_sum;
    }
  }
  // NIR direct albedo adjustment for solar zenith angles beyond 75 degrees (:751-760)
  if ((mu_not < 0.2588) && (flg == 1)) {
    const double sza_c1 = 0.085730 + -0.630883 * mu_not + 1.303723 * sq(mu_not);
    const double sza_c0 = 1.467291 + -3.338043 * mu_not + 6.807489 * sq(mu_not);
    const double sza_factor = sza_c1 * (m_log10(rds_top * 1.0) - 6.0) + sza_c0;
    const double adjust = alb_out[1] * (sza_factor - 1.0) * wgt_sum;
    alb_out[1] *= sza_factor;
#pragma unroll
    for (int i = 0; i <= NLEVSNO; ++i) if (i == top) flx_abs[i][1] -= adjust;
  }
}

// SNICAR for one incident-flux type, all five bands by one thread.  Outputs: alb_out[2] (VIS, NIR) and flx_abs[6][2]
// (five snow slots + ground, VIS/NIR), both already zero on entry.
ELMK_HD_NOINLINE void snicar_solve(const Cols& S, const Tables& T, const int c, const int flg, const double coszen,
                          const double h2osno, const int snl, const double (&albsoi)[NUMRAD], double (&alb_out)[NUMRAD],
                          double (&flx_abs)[NLEVSNO + 1][NUMRAD], uint32_t& err)
{
  using namespace alb;
  if (!((coszen > 0.0) && (h2osno > MIN_SNW))) {
    // snow_albedo_radiation_factor :762-768
    if ((coszen > 0.0) && (h2osno < MIN_SNW) && (h2osno > 0.0)) {
      alb_out[0] = albsoi[0];
      alb_out[1] = albsoi[1];
    } else {
      alb_out[0] = 0.0;
      alb_out[1] = 0.0;
    }
    return;
  }
  SnicarColumn K;
  snicar_column(S, c, h2osno, snl, K, err);
  if (err & ERR_SNICAR_RADIUS) return;   // the reference throws here; a table gather would be out of range
  const double mu_not = dmax(coszen, 0.01);
  double albout_lcl[NBND_SNW];
  double flx_abs_lcl[NLEVSNO + 1][NBND_SNW];
  const auto cnc_of = [&](const int i, double (&cnc)[NAER]) { snicar_cnc(S, c, i, cnc); };
#pragma unroll 1
  for (int b = 0; b < NBND_SNW; ++b) {
    double fb[NLEVSNO + 1];
    snicar_band(T, K, flg, b, mu_not, albsoi, cnc_of, albout_lcl[b], fb, err);
#pragma unroll
    for (int i = 0; i <= NLEVSNO; ++i) flx_abs_lcl[i][b] = fb[i];
  }
  int rds_top = 0;
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i) if (i == K.top) rds_top = K.rds[i];
  snicar_combine(flg, mu_not, K.top, rds_top, albout_lcl, flx_abs_lcl, alb_out, flx_abs);
}

// two-stream canopy solution for one waveband; returns through references (two_stream_solver :391-498)
struct TwoStreamCommon {
  double cosz, chil, gdir, twostext, avmu, temp0, temp2, wl, ws;
};

// SNICAR_DONE: the SNICAR kernel of the CUDA library has already run for this step and left albsnd / albsni and the
// four flx_abs* fields of the sunlit snow columns in the state (elmk_lib.cu, k_snicar); they are read back instead of
// being recomputed.
template <bool SNICAR_DONE = false>
ELMK_HD void column_albedo(const Cols& S, const Tables& T, const int c)
{
  using namespace alb;
  const double coszen = C1(coszen);
  const double elai = C1(elai), esai = C1(esai);
  const int snl = C1(snl);
  uint32_t err = 0;

  // ---- surface_albedo::init_timestep :102-150 ----
  double albsod[NUMRAD] = {0.0, 0.0}, albsoi[NUMRAD] = {0.0, 0.0}, albgrd[NUMRAD] = {0.0, 0.0}, albgri[NUMRAD] = {0.0, 0.0};
  double albd[NUMRAD] = {1.0, 1.0}, albi[NUMRAD] = {1.0, 1.0};
  double fabd[NUMRAD] = {0.0, 0.0}, fabi[NUMRAD] = {0.0, 0.0}, fabi_sun[NUMRAD] = {0.0, 0.0}, fabi_sha[NUMRAD] = {0.0, 0.0};
  double fabd_sun[NUMRAD] = {0.0, 0.0}, fabd_sha[NUMRAD] = {0.0, 0.0};   // wrapper-local in the reference (quirk 9)
  double ftdd[NUMRAD] = {0.0, 0.0}, ftid[NUMRAD] = {0.0, 0.0}, ftii[NUMRAD] = {0.0, 0.0};
  double absdv[NLEVSNO + 1], absdn[NLEVSNO + 1], absiv[NLEVSNO + 1], absin[NLEVSNO + 1];
#pragma unroll
  for (int i = 0; i <= NLEVSNO; ++i) { absdv[i] = 0.0; absdn[i] = 0.0; absiv[i] = 0.0; absin[i] = 0.0; }
  double vcsun = 0.0;
  double vcsha = (1.0 - m_exp(-EXTKN * elai)) / EXTKN;
  if (elai > 0.0) vcsha /= elai; else vcsha = 0.0;

  double albsnd[NUMRAD] = {0.0, 0.0}, albsni[NUMRAD] = {0.0, 0.0};
  double fsun = 0.0, fabd_sun_z = 0.0, fabd_sha_z = 0.0, fabi_sun_z = 0.0, fabi_sha_z = 0.0;

  if (coszen > 0.0) {
    // ---- soil_albedo :702-709 ----
    const int col = C1(isoicol);
    const double inc = dmax(0.11 - 0.40 * C2(h2osoi_vol, 0), 0.0);
#pragma unroll
    for (int ib = 0; ib < NUMRAD; ++ib) {
      albsod[ib] = dmin(T.albsat[col][ib] + inc, T.albdry[col][ib]);
      albsoi[ib] = albsod[ib];
    }
  }

  // ---- SNICAR, direct then diffuse ----
  const double h2osno = C1(h2osno);
  double flx_absd_snw[NLEVSNO + 1][NUMRAD], flx_absi_snw[NLEVSNO + 1][NUMRAD];
#pragma unroll
  for (int i = 0; i <= NLEVSNO; ++i) {
    flx_absd_snw[i][0] = 0.0; flx_absd_snw[i][1] = 0.0; flx_absi_snw[i][0] = 0.0; flx_absi_snw[i][1] = 0.0;
  }
  const bool snicar_ran = SNICAR_DONE && (coszen > 0.0) && (h2osno > MIN_SNW);
  if (snicar_ran) {
    albsnd[0] = C2(albsnd, 0); albsnd[1] = C2(albsnd, 1); albsni[0] = C2(albsni, 0); albsni[1] = C2(albsni, 1);
  } else if (!SNICAR_DONE && (coszen > 0.0) && (h2osno > MIN_SNW)) {
    snicar_solve(S, T, c, 1, coszen, h2osno, snl, albsoi, albsnd, flx_absd_snw, err);
    snicar_solve(S, T, c, 2, coszen, h2osno, snl, albsoi, albsni, flx_absi_snw, err);
  } else if ((coszen > 0.0) && (h2osno < MIN_SNW) && (h2osno > 0.0)) {
    albsnd[0] = albsoi[0]; albsnd[1] = albsoi[1]; albsni[0] = albsoi[0]; albsni[1] = albsoi[1];
  }

  int nrad = 1;
  if (coszen > 0.0) {
    const double fsno = C1(frac_sno);
    // ---- ground_albedo :161-166, flux_absorption_factor (subgridflag == 1 branch) :199-207 ----
#pragma unroll
    for (int ib = 0; ib < NUMRAD; ++ib) {
      albgrd[ib] = albsod[ib] * (1.0 - fsno) + albsnd[ib] * fsno;
      albgri[ib] = albsoi[ib] * (1.0 - fsno) + albsni[ib] * fsno;
    }
#pragma unroll
    for (int i = 0; i <= NLEVSNO; ++i) {
      absdv[i] = flx_absd_snw[i][0] * (1.0 - albsnd[0]);
      absiv[i] = flx_absi_snw[i][0] * (1.0 - albsni[0]);
      absdn[i] = flx_absd_snw[i][1] * (1.0 - albsnd[1]);
      absin[i] = flx_absi_snw[i][1] * (1.0 - albsni[1]);
    }
  }

  // ---- canopy_layer_lai (nlevcan == 1) :224-229,262-317 ----
  // tlai_z = elai, tsai_z = esai; the cumulative check compares identical values and cannot fire
  const double tlai_z = elai;

  // ---- two_stream_solver :338-686 ----
  const bool sunlit = coszen > 0.0;
  if (sunlit && (elai + esai) > 0.0) {
    const int vt = C1(vtype);
    const double wl = elai / dmax(elai + esai, MPE);
    const double ws = esai / dmax(elai + esai, MPE);
    const double cosz = dmax(0.001, coszen);
    double chil = dmin(dmax(T.xl[vt], -0.4), 0.6);
    if (fabs(chil) <= 0.01) chil = 0.01;
    const double phi1 = 0.5 - 0.633 * chil - 0.330 * chil * chil;
    const double phi2 = 0.877 * (1.0 - 2.0 * phi1);
    const double gdir = phi1 + phi2 * cosz;
    const double twostext = gdir / cosz;
    const double avmu = (1.0 - phi1 / phi2 * m_log((phi1 + phi2) / phi1)) / phi2;
    const double temp0 = gdir + phi2 * cosz;
    const double temp1 = phi1 * cosz;
    const double temp2 = (1.0 - temp1 / temp0 * m_log((temp1 + temp0) / temp1));
    const double t_veg = C1(t_veg), fwet = C1(fwet);
    const double omegas[NUMRAD] = {0.8, 0.4};
    constexpr double betads = 0.5, betais = 0.5;

#pragma unroll
    for (int ib = 0; ib < NUMRAD; ++ib) {
      const double rho = dmax(T.rhol[vt][ib] * wl + T.rhos[vt][ib] * ws, MPE);
      const double tau = dmax(T.taul[vt][ib] * wl + T.taus[vt][ib] * ws, MPE);
      const double omegal = rho + tau;
      const double asu = 0.5 * omegal * gdir / temp0 * temp2;
      const double betadl = (1.0 + avmu * twostext) / (omegal * avmu * twostext) * asu;
      const double betail = 0.5 * ((rho + tau) + (rho - tau) * sq((1.0 + chil) / 2.0)) / omegal;
      double tmp0, tmp1, tmp2;
      if (t_veg > TFRZ) {
        tmp0 = omegal;
        tmp1 = betadl;
        tmp2 = betail;
      } else {
        tmp0 = (1.0 - fwet) * omegal + fwet * omegas[ib];
        tmp1 = ((1.0 - fwet) * omegal * betadl + fwet * omegas[ib] * betads) / tmp0;
        tmp2 = ((1.0 - fwet) * omegal * betail + fwet * omegas[ib] * betais) / tmp0;
      }
      const double omega = tmp0;
      const double betad = tmp1;
      const double betai = tmp2;
      const double bb = 1.0 - omega + omega * betai;
      const double c1 = omega * betai;
      tmp0 = avmu * twostext;
      const double dd = tmp0 * omega * betad;
      const double f = tmp0 * omega * (1.0 - betad);
      tmp1 = bb * bb - c1 * c1;
      const double h = sqrt(tmp1) / avmu;
      const double sigma = tmp0 * tmp0 - tmp1;
      const double p1 = bb + avmu * h;
      const double p2 = bb - avmu * h;
      const double p3 = bb + tmp0;
      const double p4 = bb - tmp0;
      double t1 = dmin(h * (elai + esai), 40.0);
      const double s1 = m_exp(-t1);
      t1 = dmin(twostext * (elai + esai), 40.0);
      const double s2 = m_exp(-t1);

      // direct beam
      double u1 = bb - c1 / albgrd[ib];
      double u2 = bb - c1 * albgrd[ib];
      const double u3 = f + c1 * albgrd[ib];
      tmp2 = u1 - avmu * h;
      double tmp3 = u1 + avmu * h;
      double d1 = p1 * tmp2 / s1 - p2 * tmp3 * s1;
      double tmp4 = u2 + avmu * h;
      double tmp5 = u2 - avmu * h;
      double d2 = tmp4 / s1 - tmp5 * s1;
      const double h1 = -dd * p4 - c1 * f;
      const double tmp6 = dd - h1 * p3 / sigma;
      const double tmp7 = (dd - c1 - h1 / sigma * (u1 + tmp0)) * s2;
      const double h2 = (tmp6 * tmp2 / s1 - p2 * tmp7) / d1;
      const double h3 = -(tmp6 * tmp3 * s1 - p1 * tmp7) / d1;
      const double h4 = -f * p3 - c1 * dd;
      const double tmp8 = h4 / sigma;
      const double tmp9 = (u3 - tmp8 * (u2 - tmp0)) * s2;
      const double h5 = -(tmp8 * tmp4 / s1 + tmp9) / d2;
      const double h6 = (tmp8 * tmp5 * s1 + tmp9) / d2;
      albd[ib] = h1 / sigma + h2 + h3;
      ftid[ib] = h4 * s2 / sigma + h5 * s1 + h6 / s1;
      ftdd[ib] = s2;
      fabd[ib] = 1.0 - albd[ib] - (1.0 - albgrd[ib]) * ftdd[ib] - (1.0 - albgri[ib]) * ftid[ib];
      double a1 = h1 / sigma * (1.0 - s2 * s2) / (2.0 * twostext) + h2 * (1.0 - s2 * s1) / (twostext + h) +
                  h3 * (1.0 - s2 / s1) / (twostext - h);
      double a2 = h4 / sigma * (1.0 - s2 * s2) / (2.0 * twostext) + h5 * (1.0 - s2 * s1) / (twostext + h) +
                  h6 * (1.0 - s2 / s1) / (twostext - h);
      fabd_sun[ib] = (1.0 - omega) * (1.0 - s2 + 1.0 / avmu * (a1 + a2));
      fabd_sha[ib] = fabd[ib] - fabd_sun[ib];

      // diffuse
      u1 = bb - c1 / albgri[ib];
      u2 = bb - c1 * albgri[ib];
      tmp2 = u1 - avmu * h;
      tmp3 = u1 + avmu * h;
      d1 = p1 * tmp2 / s1 - p2 * tmp3 * s1;
      tmp4 = u2 + avmu * h;
      tmp5 = u2 - avmu * h;
      d2 = tmp4 / s1 - tmp5 * s1;
      const double h7 = (c1 * tmp2) / (d1 * s1);
      const double h8 = (-c1 * tmp3 * s1) / d1;
      const double h9 = tmp4 / (d2 * s1);
      const double h10 = (-tmp5 * s1) / d2;
      albi[ib] = h7 + h8;
      ftii[ib] = h9 * s1 + h10 / s1;
      fabi[ib] = 1.0 - albi[ib] - (1.0 - albgri[ib]) * ftii[ib];
      a1 = h7 * (1.0 - s2 * s1) / (twostext + h) + h8 * (1.0 - s2 / s1) / (twostext - h);
      a2 = h9 * (1.0 - s2 * s1) / (twostext + h) + h10 * (1.0 - s2 / s1) / (twostext - h);
      fabi_sun[ib] = (1.0 - omega) / avmu * (a1 + a2);
      fabi_sha[ib] = fabi[ib] - fabi_sun[ib];

      if (ib == 0) {
        // sun/shade big-leaf quantities, visible band only (:511-532)
        fsun = (1.0 - s2) / t1;
        const double laisum = elai + esai;
        fabd_sun_z = fabd_sun[ib] / (fsun * laisum);
        fabi_sun_z = fabi_sun[ib] / (fsun * laisum);
        fabd_sha_z = fabd_sha[ib] / ((1.0 - fsun) * laisum);
        fabi_sha_z = fabi_sha[ib] / ((1.0 - fsun) * laisum);
        const double extkb = twostext;
        vcsun = (1.0 - m_exp(-(EXTKN + extkb) * elai)) / (EXTKN + extkb);
        vcsha = (1.0 - m_exp(-EXTKN * elai)) / EXTKN - vcsun;
        if (elai > 0.0) {
          vcsun = vcsun / (fsun * elai);
          vcsha = vcsha / ((1.0 - fsun) * elai);
        } else {
          vcsun = 0.0;
          vcsha = 0.0;
        }
      }
    }
  } else if (sunlit) {
    // no vegetation: everything reaches the ground (:672-685)
#pragma unroll
    for (int ib = 0; ib < NUMRAD; ++ib) {
      fabd[ib] = 0.0; fabd_sun[ib] = 0.0; fabd_sha[ib] = 0.0; fabi[ib] = 0.0; fabi_sun[ib] = 0.0; fabi_sha[ib] = 0.0;
      ftdd[ib] = 1.0; ftid[ib] = 0.0; ftii[ib] = 1.0;
      albd[ib] = albgrd[ib];
      albi[ib] = albgri[ib];
    }
  }

  // ---- write back ----
#pragma unroll
  for (int ib = 0; ib < NUMRAD; ++ib) {
    C2(albsod, ib) = albsod[ib]; C2(albsoi, ib) = albsoi[ib]; C2(albgrd, ib) = albgrd[ib]; C2(albgri, ib) = albgri[ib];
    C2(albd, ib) = albd[ib]; C2(albi, ib) = albi[ib]; C2(fabd, ib) = fabd[ib]; C2(fabi, ib) = fabi[ib];
    C2(fabi_sun, ib) = fabi_sun[ib]; C2(fabi_sha, ib) = fabi_sha[ib];
    C2(ftdd, ib) = ftdd[ib]; C2(ftid, ib) = ftid[ib]; C2(ftii, ib) = ftii[ib];
    if (!snicar_ran) { C2(albsnd, ib) = albsnd[ib]; C2(albsni, ib) = albsni[ib]; }
  }
  if (!snicar_ran) {   // (else the SNICAR kernel has stored them)
#pragma unroll
    for (int i = 0; i <= NLEVSNO; ++i) {
      C2(flx_absdv, i) = absdv[i]; C2(flx_absdn, i) = absdn[i]; C2(flx_absiv, i) = absiv[i]; C2(flx_absin, i) = absin[i];
    }
  }
  C1(vcmaxcintsun) = vcsun;
  C1(vcmaxcintsha) = vcsha;
  C1(nrad) = nrad;
  C2(tlai_z, 0) = tlai_z;
  C2(fsun_z, 0) = fsun;
  C2(fabd_sun_z, 0) = fabd_sun_z; C2(fabd_sha_z, 0) = fabd_sha_z;
  C2(fabi_sun_z, 0) = fabi_sun_z; C2(fabi_sha_z, 0) = fabi_sha_z;
  if (err) C1(errmask) |= (int)err;
}

} // namespace elmk
