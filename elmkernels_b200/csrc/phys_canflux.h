// phys_canflux.h - canopy fluxes group (a7): root-zone moisture stress, the leaf-temperature Newton
// iteration coupled to Monin-Obukhov stability and to sunlit/shaded Farquhar-Collatz photosynthesis
// with Ball-Berry stomatal conductance (secant -> Brent root find on ci), canopy and ground fluxes.
//
// Parity target: kokkos_canopy_fluxes, reference
// driver/kokkos/canopy_fluxes_kokkos.cc:6-265 ->
//   canopy_fluxes::initialize_flux :95, stability_iteration :187, compute_flux :456
//                                                        (src/physics/canopy_fluxes_impl.hh)
//   photosynthesis::photosynthesis :9, hybrid :517, brent :396, ci_func :308, quadratic :286,
//   ft :623, fth :628, fth25 :633                        (src/physics/photosynthesis_impl.hh)
//   soil_moist_stress::calc_effective_soilporosity :62, calc_volumetric_h2oliq :77,
//   calc_root_moist_stress :89                           (src/physics/soil_moist_stress_impl.hh)
// The branch structure of the solvers is kept identical: the order of comparisons decides which
// iterate is returned, hence the bits of every downstream flux.  The 30 scratch Views of the wrapper
// (:11-40) are registers.
#pragma once
#include "elmk_state.h"
#include "phys_bareground.h"
#include "phys_cantemp.h"
#include "phys_friction.h"

namespace elmk {

// per-PFT photosynthesis constants of one column (the reference's struct PFTDataPSN, pft_data.h:20-24)
struct PsnPft {
  double fnr, act25, kcha, koha, cpha, vcmaxha, jmaxha, tpuha, lmrha;
  double vcmaxhd, jmaxhd, tpuhd, lmrhd, lmrse, qe, theta_cj, bbbopt, mbbopt;
  double c3psn, slatop, leafcn, flnr, fnitr, dleaf, smpso, smpsc, tc_stress;
};

ELMK_HD PsnPft load_psn_pft(const Cols& S, const int c)
{
  PsnPft p;
  p.fnr = C2(psn_pft, 0); p.act25 = C2(psn_pft, 1); p.kcha = C2(psn_pft, 2); p.koha = C2(psn_pft, 3);
  p.cpha = C2(psn_pft, 4); p.vcmaxha = C2(psn_pft, 5); p.jmaxha = C2(psn_pft, 6); p.tpuha = C2(psn_pft, 7);
  p.lmrha = C2(psn_pft, 8); p.vcmaxhd = C2(psn_pft, 9); p.jmaxhd = C2(psn_pft, 10); p.tpuhd = C2(psn_pft, 11);
  p.lmrhd = C2(psn_pft, 12); p.lmrse = C2(psn_pft, 13); p.qe = C2(psn_pft, 14); p.theta_cj = C2(psn_pft, 15);
  p.bbbopt = C2(psn_pft, 16); p.mbbopt = C2(psn_pft, 17); p.c3psn = C2(psn_pft, 18); p.slatop = C2(psn_pft, 19);
  p.leafcn = C2(psn_pft, 20); p.flnr = C2(psn_pft, 21); p.fnitr = C2(psn_pft, 22); p.dleaf = C2(psn_pft, 23);
  p.smpso = C2(psn_pft, 24); p.smpsc = C2(psn_pft, 25); p.tc_stress = C2(psn_pft, 26);
  return p;
}

// ---- photosynthesis -------------------------------------------------------------------------

// temperature response functions
ELMK_HD double psn_ft(const double tl, const double ha)
{
  return m_exp(ha / (RGAS * 1.0e-3 * (TFRZ + 25.0)) * (1.0 - (TFRZ + 25.0) / tl));
}
ELMK_HD double psn_fth(const double tl, const double hd, const double se, const double scale)
{
  return scale / (1.0 + m_exp((-hd + se * tl) / (RGAS * 1.0e-3 * tl)));
}
ELMK_HD double psn_fth25(const double hd, const double se)
{
  return 1.0 + m_exp((-hd + se * (TFRZ + 25.0)) / (RGAS * 1.0e-3 * (TFRZ + 25.0)));
}

// ELMK_INLINE_DIV_BEGIN  (the quadratics and ci_func: eleven divisions per evaluation, pairwise independent - see
//                          elmkernels_b200/ptx_rewrite.py, marked_ranges)
// roots of a x^2 + b x + c, numerically stable form; a == 0 is an error in the reference
ELMK_HD void psn_quadratic(const double a, const double b, const double cc, double& r1, double& r2, uint32_t& err)
{
  if (a == 0.0) err |= ERR_QUADRATIC;
  double q;
  if (b >= 0.0) {
    q = -0.5 * (b + sqrt(b * b - 4.0 * a * cc));
  } else {
    q = -0.5 * (b - sqrt(b * b - 4.0 * a * cc));
  }
  r1 = q / a;
  if (q != 0.0) {
    r2 = cc / q;
  } else {
    r2 = 1.0e36;
  }
}

// everything ci_func needs besides ci; gs_mol and the assimilation rates are its side outputs
struct LeafPsn {
  // inputs
  double gb_mol, je, cair, oair, lmr, par, rh_can, vcmax, pbot, cp, kc, ko, qe, tpu, kp, theta_cj, bbb, mbb;
  bool c3;
  // sub-expressions of ci_func that do not depend on ci (same operations, same order - evaluated once per root
  // find instead of once per function evaluation; two of them are divisions)
  double kco, cp8, r14gb, gb16, gbmb;   // kc (1 + oair/ko), 8 cp, 1.4 / gb, 1.6 gb, gb - bbb
  // outputs of the last evaluation
  double gs_mol, ac, aj, ap, ag, an;
};

// f(ci) = ci - (ca - (1.4/gb + 1.6/gs) p an).  Inlined at its single call site (psn_hybrid below).
ELMK_HD double psn_ci_func(const double ci, LeafPsn& L, uint32_t& err)
{
  constexpr double theta_ip = 0.95;
  if (L.c3) {
    L.ac = L.vcmax * dmax(ci - L.cp, 0.0) / (ci + L.kco);
    L.aj = L.je * dmax(ci - L.cp, 0.0) / (4.0 * ci + L.cp8);
    L.ap = 3.0 * L.tpu;
  } else {
    L.ac = L.vcmax;
    L.aj = L.qe * L.par * 4.6;
    L.ap = L.kp * dmax(ci, 0.0) / L.pbot;
  }
  double r1, r2;
  psn_quadratic(L.theta_cj, -(L.ac + L.aj), L.ac * L.aj, r1, r2, err);
  const double ai = dmin(r1, r2);
  psn_quadratic(theta_ip, -(ai + L.ap), ai * L.ap, r1, r2, err);
  L.ag = dmin(r1, r2);
  L.an = L.ag - L.lmr;
  if (L.an < 0.0) return 0.0;
  double cs = L.cair - L.r14gb * L.an * L.pbot;
  cs = dmax(cs, 1.e-6);
  const double aquad = cs;
  const double bquad = cs * L.gbmb - L.mbb * L.an * L.pbot;
  const double cquad = -L.gb_mol * (cs * L.bbb + L.mbb * L.an * L.pbot * L.rh_can);
  psn_quadratic(aquad, bquad, cquad, r1, r2, err);
  L.gs_mol = dmax(r1, r2);
  return ci - L.cair + L.an * L.pbot * (1.4 * L.gs_mol + L.gb16) / (L.gb_mol * L.gs_mol);
}
// ELMK_INLINE_DIV_END

// Root of ci_func: secant search for a sign change, Brent's method (Numerical Recipes form) once a root is
// bracketed; leaves the side outputs of the last ci_func evaluation in L.  Reference: hybrid
// (photosynthesis_impl.hh:517-600) calling brent (:396-515), both calling ci_func at several places.
//
// Written here as ONE loop around ONE ci_func evaluation, with the position inside the reference's control flow
// kept in `state`: the sequence of evaluation points and every arithmetic operation are the reference's, but
//   * the lanes of a warp meet at the same ci_func code whatever phase (first two points, secant, Brent, fall-back)
//     each of them is in, instead of executing the phases one after the other;
//   * with a single call site ci_func is inlined and LeafPsn lives in registers - as a called function with the
//     state behind a reference it cost ~40 local-memory accesses per evaluation (round-1 ncu: 18 M local loads
//     and 18 M local stores per 512k columns in the iteration kernel).
ELMK_HD void psn_hybrid(double x0, LeafPsn& L, uint32_t& err)
{
  constexpr double eps = 1.0e-2;
  constexpr double eps1 = 1.0e-4;
  constexpr int itmax = 40;
  constexpr int BRENT_ITMAX = 20;
  constexpr double BRENT_EPS = 1.0e-2;
  enum { AT_X0, AT_X1, IN_SECANT, IN_BRENT, AT_MINX };
  int state = AT_X0;
  double x = x0;                            // where ci_func is evaluated next
  double f0 = 0.0, x1 = 0.0, f1 = 0.0, minx = x0, minf = 0.0, tol = 0.0;
  int iter = 0;
  double a = 0.0, b = 0.0, cc = 0.0, fa = 0.0, fb = 0.0, fc = 0.0, d = 0.0, e = 0.0;   // Brent
  int biter = 0;
#pragma unroll 1
  while (true) {
    const double f = psn_ci_func(x, L, err);
    bool brent_step = false;
    if (state == AT_X0) {
      f0 = f;
      if (f0 == 0.0) break;
      minx = x0;
      minf = f0;
      x1 = x0 * 0.99;
      x = x1;
      state = AT_X1;
      continue;
    } else if (state == AT_X1) {
      f1 = f;
      if (f1 == 0.0) break;
      if (f1 < minf) {
        minx = x1;
        minf = f1;
      }
      iter = 0;
    } else if (state == IN_SECANT) {
      f1 = f;
      if (f1 < minf) {
        minx = x1;
        minf = f1;
      }
      if (fabs(f1) <= eps1) break;
      if (f1 * f0 < 0.0) {
        // bracketed: Brent on [x0, x1] with the tolerance of the last secant step
        a = x0; b = x1; fa = f0; fb = f1;
        if ((fa > 0.0 && fb > 0.0) || (fa < 0.0 && fb < 0.0)) err |= ERR_BRENT_BRACKET;
        cc = b; fc = fb;
        d = 0.0; e = 0.0;
        biter = 0;
        brent_step = true;
      } else if (iter > itmax) {
        // not converged: fall back to the evaluation with the smallest residual
        x = minx;
        state = AT_MINX;
        continue;
      }
    } else if (state == IN_BRENT) {
      fb = f;
      if (fb == 0.0) break;
      brent_step = true;
    } else {
      break;   // AT_MINX: that evaluation was the last one
    }

    if (!brent_step) {
      // one secant step
      iter += 1;
      const double dx = -f1 * (x1 - x0) / (f1 - f0);
      const double xn = x1 + dx;
      tol = fabs(xn) * eps;
      if (fabs(dx) < tol) break;
      x0 = x1;
      f0 = f1;
      x1 = xn;
      x = x1;
      state = IN_SECANT;
      continue;
    }

    // one Brent step: the body of the reference's loop up to its ci_func call
    if (biter == BRENT_ITMAX) break;
    biter += 1;
    if ((fb > 0.0 && fc > 0.0) || (fb < 0.0 && fc < 0.0)) {
      cc = a;
      fc = fa;
      d = b - a;
      e = d;
    }
    if (fabs(fc) < fabs(fb)) {
      a = b;
      b = cc;
      cc = a;
      fa = fb;
      fb = fc;
      fc = fa;
    }
    const double tol1 = 2.0 * BRENT_EPS * fabs(b) + 0.5 * tol;
    const double xm = 0.5 * (cc - b);
    if (fabs(xm) <= tol1 || fb == 0.0) break;
    if (fabs(e) >= tol1 && fabs(fa) > fabs(fb)) {
      const double s = fb / fa;
      double p, q;
      if (a == cc) {
        p = 2.0 * xm * s;
        q = 1.0 - s;
      } else {
        q = fa / fc;
        const double r = fb / fc;
        p = s * (2.0 * xm * q * (q - r) - (b - a) * (r - 1.0));
        q = (q - 1.0) * (r - 1.0) * (s - 1.0);
      }
      if (p > 0.0) q *= -1.0;
      p = fabs(p);
      if (2.0 * p < dmin(3.0 * xm * q - fabs(tol1 * q), fabs(e * q))) {
        e = d;
        d = p / q;
      } else {
        d = xm;
        e = d;
      }
    } else {
      d = xm;
      e = d;
    }
    a = b;
    fa = fb;
    if (fabs(d) > tol1) {
      b = b + d;
    } else {
      b = b + copysign(tol1, xm);
    }
    x = b;
    state = IN_BRENT;
  }
}

// Loop-invariant part of the photosynthesis model for one column: everything in photosynthesis() (:9-283) that
// depends only on the PFT constants, the 10-day temperature, pressure and day length.  The reference
// recomputes it in both calls (sunlit, shaded) of every stability pass; the values are identical, so it is
// computed once per column here.
struct PsnColumn {
  bool c3;
  double vcmax25top, jmax25top, tpu25top, kp25top, lmr25top, vcmaxse, jmaxse, tpuse;
  double lmrc, vcmaxc, jmaxc, tpuc;   // high-temperature inhibition scaled to 1 at 25 C (fth25)
  double cf, kc25, ko25, cp25;
};

ELMK_HD PsnColumn psn_column(const PsnPft& P, const double t10, const double pbot, const double thm, const double oair,
                             const double dayl_factor)
{
  constexpr double sco = 0.5 * 0.209 / (42.75 / 1.e06);
  PsnColumn C;
  C.c3 = (round(P.c3psn) == 1);   // anything else is treated as C4, as in the reference (:22-27)
  const double lnc = 1.0 / (P.slatop * P.leafcn);
  const double act25 = P.act25 * 1000.0 / 60.0;
  double vcmax25top = lnc * P.flnr * P.fnr * act25 * dayl_factor;
  vcmax25top *= P.fnitr;
  const double t10c = dmin(dmax((t10 - TFRZ), 11.0), 35.0);
  C.vcmax25top = vcmax25top;
  C.jmax25top = (2.59 - 0.035 * t10c) * vcmax25top;
  C.tpu25top = 0.167 * vcmax25top;
  C.kp25top = 20000.0 * vcmax25top;
  C.lmr25top = C.c3 ? vcmax25top * 0.015 : vcmax25top * 0.025;
  C.vcmaxse = 668.39 - 1.07 * t10c;
  C.jmaxse = 659.70 - 0.75 * t10c;
  C.tpuse = C.vcmaxse;
  C.lmrc = psn_fth25(P.lmrhd, P.lmrse);
  C.vcmaxc = psn_fth25(P.vcmaxhd, C.vcmaxse);
  C.jmaxc = psn_fth25(P.jmaxhd, C.jmaxse);
  C.tpuc = psn_fth25(P.tpuhd, C.tpuse);
  C.cf = pbot / (RGAS * 1.0e-3 * thm) * 1.e06;
  C.kc25 = (404.9 / 1.e06) * pbot;
  C.ko25 = (278.4 / 1.e03) * pbot;
  C.cp25 = 0.5 * oair / sco;
  return C;
}

// Leaf-temperature response factors of one stability pass: the sunlit and the shaded call see the same leaf
// temperature, so the (up to 11) exponentials are evaluated once per pass instead of twice.
struct PsnPass {
  double lmr_a, lmr_b;                       // C3: ft, fth;  C4: 2^((T-25)/10), 1 + i_exp(1.3 (T-55))
  double vc_a, vc_b, jm_a, jm_b, tp_a, tp_b; // ft, fth of vcmax, jmax, tpu (daytime only)
  double p2, c4d1, c4d2;                     // 2^((T-25)/10) and the two C4 vcmax inhibition denominators
  double kc, ko, cp;                         // Michaelis-Menten constants and CO2 compensation point
};

// ELMK_INLINE_DIV_BEGIN
// the same two with the exponential in line: psn_pass evaluates up to eleven of them at one leaf temperature, all
// independent - called, they would run one after the other (elmk_common.h, m_div2)
ELMK_HD double psn_ft_i(const double tl, const double ha)
{
  return i_exp(ha / (RGAS * 1.0e-3 * (TFRZ + 25.0)) * (1.0 - (TFRZ + 25.0) / tl));
}
ELMK_HD double psn_fth_i(const double tl, const double hd, const double se, const double scale)
{
  return scale / (1.0 + i_exp((-hd + se * tl) / (RGAS * 1.0e-3 * tl)));
}
ELMK_HD PsnPass psn_pass(const PsnPft& P, const PsnColumn& C, const double t_veg, const bool day)
{
  PsnPass T;
  T.p2 = 0.0; T.c4d1 = 1.0; T.c4d2 = 1.0;
  T.vc_a = 0.0; T.vc_b = 0.0; T.jm_a = 0.0; T.jm_b = 0.0; T.tp_a = 0.0; T.tp_b = 0.0;
  if (!C.c3 || day) T.p2 = pow_cbase(2.0, ELMK_LN_2, ((t_veg - (TFRZ + 25.0)) / 10.0));
  if (C.c3) {
    T.lmr_a = psn_ft_i(t_veg, P.lmrha);
    T.lmr_b = psn_fth_i(t_veg, P.lmrhd, P.lmrse, C.lmrc);
  } else {
    T.lmr_a = T.p2;
    T.lmr_b = (1.0 + i_exp(1.3 * (t_veg - (TFRZ + 55.0))));
  }
  if (day) {
    T.vc_a = psn_ft_i(t_veg, P.vcmaxha);
    T.vc_b = psn_fth_i(t_veg, P.vcmaxhd, C.vcmaxse, C.vcmaxc);
    T.jm_a = psn_ft_i(t_veg, P.jmaxha);
    T.jm_b = psn_fth_i(t_veg, P.jmaxhd, C.jmaxse, C.jmaxc);
    T.tp_a = psn_ft_i(t_veg, P.tpuha);
    T.tp_b = psn_fth_i(t_veg, P.tpuhd, C.tpuse, C.tpuc);
    if (!C.c3) {
      T.c4d1 = (1.0 + i_exp(0.2 * ((TFRZ + 15.0) - t_veg)));
      T.c4d2 = (1.0 + i_exp(0.3 * (t_veg - (TFRZ + 40.0))));
    }
  }
  T.kc = C.kc25 * psn_ft_i(t_veg, P.kcha);
  T.ko = C.ko25 * psn_ft_i(t_veg, P.koha);
  T.cp = C.cp25 * psn_ft_i(t_veg, P.cpha);
  return T;
}
// ELMK_INLINE_DIV_END

// stomatal resistance of the sunlit or the shaded canopy fraction (nlevcan == 1, nrad == 1)
ELMK_HD double psn_stomatal_resistance(const PsnPft& P, const PsnColumn& C, const PsnPass& T, const int nrad,
                                                const double pbot, const double esat_tv, const double eair,
                                                const double oair, const double cair, const double rb,
                                                const double btran, const double vcmaxcint, const double par,
                                                const double lai, uint32_t& err)
{
  constexpr double fnps = 0.15;
  constexpr double theta_psii = 0.7;
  const bool c3 = C.c3;
  if (nrad < 1) return 0.0;
  const double nscaler = vcmaxcint;

  // leaf maintenance respiration (always) and the carboxylation capacities (daytime only)
  double lmr_z;
  const double lmr25 = C.lmr25top * nscaler;
  if (c3) {
    lmr_z = lmr25 * T.lmr_a * T.lmr_b;
  } else {
    lmr_z = lmr25 * T.lmr_a;
    lmr_z /= T.lmr_b;
  }
  double vcmax_z, jmax_z, tpu_z, kp_z;
  if (par <= 0.0) {
    vcmax_z = 0.0; jmax_z = 0.0; tpu_z = 0.0; kp_z = 0.0;
  } else {
    const double vcmax25 = C.vcmax25top * nscaler;
    const double jmax25 = C.jmax25top * nscaler;
    const double tpu25 = C.tpu25top * nscaler;
    const double kp25 = C.kp25top * nscaler;
    vcmax_z = vcmax25 * T.vc_a * T.vc_b;
    jmax_z = jmax25 * T.jm_a * T.jm_b;
    tpu_z = tpu25 * T.tp_a * T.tp_b;
    if (!c3) {
      vcmax_z = vcmax25 * T.p2;
      vcmax_z /= T.c4d1;
      vcmax_z /= T.c4d2;
    }
    kp_z = kp25 * T.p2;
  }
  vcmax_z *= btran;
  lmr_z *= btran;

  const double cf = C.cf;
  const double gb = 1.0 / rb;
  const double gb_mol = gb * cf;
  const double bbb = dmax(P.bbbopt * btran, 1.0);
  constexpr double rsmax0 = 2.0e4;
  const double kc = T.kc, ko = T.ko, cp = T.cp;

  double rs_z;
  if (par <= 0.0) {
    rs_z = dmin(rsmax0, 1.0 / bbb * cf);
  } else {
    const double ceair = dmin(eair, esat_tv);
    const double rh_can = ceair / esat_tv;
    const double qabs = 0.5 * (1.0 - fnps) * par * 4.6;
    double r1, r2;
    psn_quadratic(theta_psii, -(qabs + jmax_z), qabs * jmax_z, r1, r2, err);
    LeafPsn L;
    L.gb_mol = gb_mol; L.je = dmin(r1, r2); L.cair = cair; L.oair = oair; L.lmr = lmr_z; L.par = par;
    L.rh_can = rh_can; L.vcmax = vcmax_z; L.pbot = pbot; L.cp = cp; L.kc = kc; L.ko = ko; L.qe = P.qe;
    L.tpu = tpu_z; L.kp = kp_z; L.theta_cj = P.theta_cj; L.bbb = bbb; L.mbb = P.mbbopt; L.c3 = c3;
    L.kco = kc * (1.0 + oair / ko); L.cp8 = 8.0 * cp; L.r14gb = 1.4 / gb_mol; L.gb16 = 1.6 * gb_mol; L.gbmb = gb_mol - bbb;
    L.gs_mol = 0.0; L.ac = 0.0; L.aj = 0.0; L.ap = 0.0; L.ag = 0.0; L.an = 0.0;
    // every call restarts from the atmospheric CO2 guess: iterations do not inherit the previous root
    psn_hybrid(c3 ? 0.7 * cair : 0.4 * cair, L, err);
    if (L.an < 0.0) L.gs_mol = bbb;
    const double gs = L.gs_mol / cf;
    rs_z = dmin(1.0 / gs, rsmax0);
    if (L.gs_mol < 0.0) err |= ERR_NEG_STOMATAL;
  }
  // canopy integration over the single layer
  const double gscan = lai / (rb + rs_z);
  const double laican = lai;
  return (laican > 0.0) ? laican / gscan - rb : 0.0;
}

// ---- canopy fluxes --------------------------------------------------------------------------
//
// The group is written as three pieces around one explicit state object so that the CUDA library can
// run the data-dependent stability iteration with warp-level re-packing (elmk_lib.cu, k_canflux_iterate):
//   canflux_begin    initialize_flux (:95-181): moisture stress, canopy aerodynamics, first guess
//   canflux_iterate  ONE pass of the stability_iteration loop body (:215-451); returns true when done
//   canflux_end      compute_flux (:456-539) and the write-back
// column_canopy_fluxes = begin; while (!iterate) ; end  - the reference's control flow, used by the host
// port and by the one-launch-per-group plan.

// doubles of the iteration state.  STATE: plain copies of per-column state fields that CanopyFluxes does not change
// before its last phase (the re-packed kernels re-read them from the state instead of carrying them through
// scratch memory); CONST: values derived once in canflux_begin; CARRIED: the loop-carried values.
// (STATE_A: read once or twice per pass - the members the iteration kernel keeps in shared memory, elmk_lib.cu IterView)
#define ELMK_CANFLUX_STATE_A(X)                                                                                   \
  X(hgt_u, C1(forc_hgt_u_patch)) X(hgt_t, C1(forc_hgt_t_patch)) X(hgt_q, C1(forc_hgt_q_patch)) X(displa, C1(displa)) \
  X(z0mv, C1(z0mv)) X(z0mg, C1(z0mg)) X(fwet, C1(fwet)) X(fdry, C1(fdry)) X(laisun, C1(laisun)) X(laisha, C1(laisha)) \
  X(snow_depth, C1(snow_depth)) X(soilbeta, C1(soilbeta)) X(sabv, C1(sabv)) X(htop, C1(htop)) X(h2ocan0, C1(h2ocan)) \
  X(vcsha, C1(vcmaxcintsha)) X(vcsun, C1(vcmaxcintsun)) X(t10, C1(t10))
#define ELMK_CANFLUX_STATE_B(X)                                                                                   \
  X(pbot, C1(forc_pbot)) X(forc_q, C1(forc_qbot)) X(forc_th, C1(forc_thbot)) X(forc_lwrad, C1(forc_lwrad))        \
  X(thm, C1(thm)) X(thv, C1(thv)) X(tg, C1(t_grnd)) X(qg, C1(qg)) X(elai, C1(elai)) X(esai, C1(esai))             \
  X(emv, C1(emv)) X(emg, C1(emg)) X(fsno, C1(frac_sno)) X(fsfc, C1(frac_h2osfc))                                  \
  X(parsha, C2(parsha_z, 0)) X(parsun, C2(parsun_z, 0)) X(laisha_z, C2(laisha_z, 0)) X(laisun_z, C2(laisun_z, 0))  \
  X(t_soil1, C2(t_soisno, NLEVSNO)) X(t_sfc, C1(t_h2osfc))
#define ELMK_CANFLUX_STATE(X) ELMK_CANFLUX_STATE_A(X) ELMK_CANFLUX_STATE_B(X)
#define ELMK_CANFLUX_CONST(X)                                                                                     \
  X(forc_po2) X(forc_pco2) X(forc_rho) X(dayl_factor) X(air) X(bir) X(cir) X(ur) X(zldis) X(lw_grnd) X(t_snotop)  \
  X(dtime)
// (CARRIED_SET: given a value by canflux_begin; CARRIED_ZERO: start the iteration at zero - the set-up launch does not
//  store them and a lane of the iteration kernel that takes a new column does not load them)
#define ELMK_CANFLUX_CARRIED_SET(X)                                                                               \
  X(btran) X(t_veg) X(el) X(qsatl) X(qsatldT) X(taf) X(qaf) X(dth) X(dqh) X(delq) X(um) X(obu) X(qflx_tran_veg)   \
  X(qflx_evap_veg) X(eflx_sh_veg)
#define ELMK_CANFLUX_CARRIED_ZERO(X)                                                                              \
  X(obuold) X(del) X(efeb) X(wtg) X(wtl0) X(wta0) X(wtal) X(wtgq) X(wtalq) X(wtlq0) X(wtaq0) X(tlbef) X(dt_veg)  \
  X(p_ustar) X(p_temp1) X(p_temp2) X(p_obu)
#define ELMK_CANFLUX_CARRIED(X) ELMK_CANFLUX_CARRIED_SET(X) ELMK_CANFLUX_CARRIED_ZERO(X)
#define ELMK_CANFLUX_INT(X) X(nrad) X(veg) X(soybean) X(itlef) X(nmozsgn) X(err)

struct CanopyIter {
#define X(n, e) double n;
  ELMK_CANFLUX_STATE(X)
#undef X
#define X(n) double n;
  ELMK_CANFLUX_CONST(X)
  ELMK_CANFLUX_CARRIED(X)
#undef X
#define X(n) int n;
  ELMK_CANFLUX_INT(X)
#undef X
};

// Non-vegetated columns: initialize_flux (:121-131) and the unconditional zeroing of compute_flux.
// Returns false when the column has no exposed vegetation (nothing else to do).
ELMK_HD bool canflux_begin(const Cols& S, const int vtype, const StepArgs& A, const PsnPft& P, const int c, CanopyIter& I)
{
  // canopy_fluxes::compute_flux zeroes these for every column (:475-480) - including the bare
  // columns whose values kokkos_bareground_fluxes has just computed
  C1(cgrnd) = 0.0;
  C1(cgrnds) = 0.0;
  C1(cgrndl) = 0.0;

  const int veg = C1(frac_veg_nosno);
  const double forc_t = C1(forc_tbot);
  if (veg == 0) {
    C1(btran) = 0.0;
    C1(t_veg) = forc_t;
#pragma unroll
    for (int i = 0; i < NLEVGRND; ++i) C2(rootr, i) = 0.0;
    return false;
  }

  I.err = 0;
  I.veg = veg;
  I.dtime = A.dtime;
  const int snl = C1(snl);
  I.pbot = C1(forc_pbot); I.forc_q = C1(forc_qbot); I.forc_th = C1(forc_thbot);
  I.forc_lwrad = C1(forc_lwrad);
  I.thm = C1(thm); I.thv = C1(thv); I.tg = C1(t_grnd); I.qg = C1(qg);
  I.elai = C1(elai); I.esai = C1(esai); I.emv = C1(emv); I.emg = C1(emg);
  I.z0mg = C1(z0mg);
  I.hgt_u = C1(forc_hgt_u_patch); I.hgt_t = C1(forc_hgt_t_patch); I.hgt_q = C1(forc_hgt_q_patch);
  // derive_forc_po2 / derive_forc_pco2 (canopy_fluxes_kokkos.cc:49-51) unless the caller supplies the partial pressures
  I.forc_po2 = S.po2_in ? S.po2_in[c] : O2_MOLAR_CONST * I.pbot;
  I.forc_pco2 = S.pco2_in ? S.pco2_in[c] : CO2_PPMV * 1.0e-6 * I.pbot;
  I.forc_rho = air_density(I.pbot, I.forc_q, forc_t);

  // ---- initialize_flux (:133-181) ----
  I.dayl_factor = dmin(1.0, dmax(0.01, (A.dayl * A.dayl) / (A.max_dayl * A.max_dayl)));
  // root-zone moisture stress: effective porosity, liquid volume, per-layer resistance
  double btran = 0.0;
  double rootr[NLEVGRND];
  // The inputs of layer i + 1 are requested before the power function of layer i is called: the loop is rolled and the
  // call is a barrier for the compiler, so without this a warp has one layer's loads in flight at a time (set-up launch
  // 1.62 -> 1.29 ms).
  double n_watsat = C2(watsat, 0), n_dz = C2(dz, NLEVSNO), n_ice = C2(h2osoi_ice, NLEVSNO), n_liq = C2(h2osoi_liq, NLEVSNO),
         n_t = C2(t_soisno, NLEVSNO), n_sucsat = C2(sucsat, 0), n_bsw = C2(bsw, 0), n_rootfr = C2(rootfr, 0);
#pragma unroll 1
  for (int i = 0; i < NLEVGRND; ++i) {
    const double watsat = n_watsat, dzk = n_dz, ice_k = n_ice, liq_k = n_liq, t_k = n_t, sucsat_i = n_sucsat, bsw_i = n_bsw,
                 rootfr_i = n_rootfr;
    if (i + 1 < NLEVGRND) {
      const int k1 = NLEVSNO + i + 1;
      n_watsat = C2(watsat, i + 1); n_dz = C2(dz, k1); n_ice = C2(h2osoi_ice, k1); n_liq = C2(h2osoi_liq, k1);
      n_t = C2(t_soisno, k1); n_sucsat = C2(sucsat, i + 1); n_bsw = C2(bsw, i + 1); n_rootfr = C2(rootfr, i + 1);
    }
    const double vol_ice = dmin(watsat, (ice_k / (DENICE * dzk)));
    const double eff_por = watsat - vol_ice;
    C2(eff_porosity, i) = eff_por;
    const double liqvol = dmin(eff_por, (liq_k / (dzk * DENH2O)));
    if (liqvol <= 0.0 || t_k <= TFRZ + P.tc_stress) {
      rootr[i] = 0.0;
    } else {
      const double s_node = dmax(liqvol / eff_por, 0.01);
      double smp_node = -sucsat_i * m_pow(s_node, (-bsw_i));
      smp_node = dmax(P.smpsc, smp_node);
      const double rresis = dmin((eff_por / watsat) * (smp_node - P.smpsc) / (P.smpso - P.smpsc), 1.0);
      rootr[i] = rootfr_i * rresis;
      btran += dmax(rootr[i], 0.0);
    }
  }
#pragma unroll 1
  for (int i = 0; i < NLEVGRND; ++i) {
    if (btran > 0.0) rootr[i] /= btran; else rootr[i] = 0.0;
    C2(rootr, i) = rootr[i];
  }
  I.btran = btran;

  // sparse/dense canopy aerodynamic parameters
  double displa = C1(displa), z0mv = C1(z0mv);
  const double lt = dmin(I.elai + I.esai, 2.0);
  const double egvf = (1.0 - m_exp(-lt)) / (1.0 - m_exp(-2.0));
  displa *= egvf;
  z0mv = m_exp(egvf * m_log(z0mv) + (1.0 - egvf) * m_log(I.z0mg));
  C1(displa) = displa;
  C1(z0mv) = z0mv;
  C1(z0hv) = z0mv;
  C1(z0qv) = z0mv;
  I.displa = displa;
  I.z0mv = z0mv;

  // net absorbed longwave coefficients
  I.air = I.emv * (1.0 + (1.0 - I.emv) * (1.0 - I.emg)) * I.forc_lwrad;
  I.bir = -(2.0 - I.emv * (1.0 - I.emg)) * I.emv * STEBOL;
  I.cir = I.emv * I.emg * STEBOL;

  I.t_veg = C1(t_veg);
  double deldT;
  qsat(I.t_veg, I.pbot, I.el, deldT, I.qsatl, I.qsatldT);
  I.taf = (I.tg + I.thm) / 2.0;
  I.qaf = (I.forc_q + I.qg) / 2.0;
  const double fu = C1(forc_u), fv = C1(forc_v);
  I.ur = dmax(1.0, sqrt(fu * fu + fv * fv));
  I.dth = I.thm - I.taf;
  I.dqh = I.forc_q - I.qaf;
  I.delq = I.qg - I.qaf;
  const double dthv = I.dth * (1.0 + 0.61 * I.forc_q) + 0.61 * I.forc_th * I.dqh;
  I.zldis = I.hgt_u - displa;
  if (!(I.zldis >= 0.0)) I.err |= ERR_FORC_HEIGHT;
  mo_initial_length(I.ur, I.thv, dthv, I.zldis, z0mv, I.um, I.obu);

  // ---- loop-invariant inputs of stability_iteration (:215-451) ----
  I.fwet = C1(fwet); I.fdry = C1(fdry); I.laisun = C1(laisun); I.laisha = C1(laisha);
  I.snow_depth = C1(snow_depth); I.soilbeta = C1(soilbeta);
  I.fsno = C1(frac_sno); I.fsfc = C1(frac_h2osfc); I.t_sfc = C1(t_h2osfc);
  I.sabv = C1(sabv); I.htop = C1(htop); I.t10 = C1(t10);
  I.h2ocan0 = C1(h2ocan);
  I.nrad = C1(nrad);
  I.vcsha = C1(vcmaxcintsha); I.vcsun = C1(vcmaxcintsun);
  I.parsha = C2(parsha_z, 0); I.parsun = C2(parsun_z, 0);
  I.laisha_z = C2(laisha_z, 0); I.laisun_z = C2(laisun_z, 0);
  I.t_snotop = C2(t_soisno, NLEVSNO - snl); I.t_soil1 = C2(t_soisno, NLEVSNO);
  I.soybean = (vtype == PFT_SOYBEAN || vtype == PFT_SOYBEAN_IRRIG) ? 1 : 0;
  // ground-emitted longwave does not change during the iteration
  I.lw_grnd = (I.fsno * pow4(I.t_snotop) + (1.0 - I.fsno - I.fsfc) * pow4(I.t_soil1) + I.fsfc * pow4(I.t_sfc));

  I.itlef = 0; I.nmozsgn = 0;
  I.del = 0.0; I.efeb = 0.0; I.obuold = 0.0;
  I.qflx_tran_veg = C1(qflx_tran_veg); I.qflx_evap_veg = C1(qflx_evap_veg); I.eflx_sh_veg = C1(eflx_sh_veg);
  I.wtg = 0.0; I.wtl0 = 0.0; I.wta0 = 0.0; I.wtal = 0.0; I.wtgq = 0.0; I.wtalq = 0.0; I.wtlq0 = 0.0; I.wtaq0 = 0.0;
  I.tlbef = 0.0; I.dt_veg = 0.0;
  I.p_ustar = 0.0; I.p_temp1 = 0.0; I.p_temp2 = 0.0; I.p_obu = 0.0;
  return true;
}

// One pass of the stability iteration.  Returns true when the loop of the reference would end
// (converged, or 41 passes done).
template <class IT = CanopyIter>
ELMK_HD bool canflux_iterate(const PsnPft& P, const PsnColumn& PC, IT& I)
{
  constexpr double ria = 0.5, dlemin = 0.1, dtmin = 0.01;
  constexpr int itmax = 40, itmin = 2;
  uint32_t err = (uint32_t)I.err;
  const double elai = I.elai, esai = I.esai, forc_q = I.forc_q, pbot = I.pbot, thm = I.thm, tg = I.tg, qg = I.qg;
  const double forc_rho = I.forc_rho, dtime = I.dtime, h2ocan0 = I.h2ocan0;
  const int veg = I.veg;

  // friction velocity and the profile relations at the forcing heights (z0h = z0q = z0m for vegetation).  The
  // reference also evaluates the 2 m relations in every pass (canopy_fluxes_impl.hh:232-240) but only reads those of the last one
  // (t_ref2m, q_ref2m): they are evaluated once, in canflux_end, from the Obukhov length this pass started with.
  MoProfiles p;
  {
    const MoPair mp = mo_pair_inl(I.hgt_u - I.displa, I.hgt_t - I.displa, I.um, I.obu, I.z0mv, I.z0mv);
    p.ustar = mp.ustar;
    p.temp1 = mp.temp;
  }
  p.temp2 = (I.hgt_q == I.hgt_t) ? p.temp1 : mo_scalar_profile(I.hgt_q - I.displa, I.obu, I.z0mv);
  I.p_ustar = p.ustar; I.p_temp1 = p.temp1; I.p_temp2 = p.temp2; I.p_obu = I.obu;
  double t_veg = I.t_veg;
  const double tlbef = t_veg;
  I.tlbef = tlbef;
  const double del2 = I.del;
  const double ram = 1.0 / (p.ustar * p.ustar / I.um);
  const double rah0 = 1.0 / (p.temp1 * p.ustar);
  const double raw0 = 1.0 / (p.temp2 * p.ustar);
  const double uaf = I.um * sqrt(1.0 / (ram * I.um));
  const double cf = 0.01 / (sqrt(uaf) * sqrt(P.dleaf));
  const double rb = 1.0 / (cf * uaf);
  const double w = m_exp(-(elai + esai));
  const double csoilb = (VKC / (0.13 * m_pow((I.z0mg * uaf / 1.5e-5), 0.45)));
  const double ri = (GRAV * I.htop * (I.taf - tg)) / (I.taf * sq(uaf));
  double csoilcn;
  if ((I.taf - tg) > 0.0) {
    const double ricsoilc = CSOILC / (1.0 + ria * dmin(ri, 10.0));
    csoilcn = csoilb * w + ricsoilc * (1.0 - w);
  } else {
    csoilcn = csoilb * w + CSOILC * (1.0 - w);
  }
  const double rah1 = 1.0 / (csoilcn * uaf);
  const double raw1 = rah1;
  const double svpts = I.el;
  const double eah = pbot * I.qaf / 0.622;

  double btran = I.btran;
  const PsnPass PT = psn_pass(P, PC, t_veg, (I.parsun > 0.0) || (I.parsha > 0.0));
  // sunlit, then shaded leaves: one copy of the (inlined) photosynthesis code, run twice
  double rssun = 0.0, rssha = 0.0;
#pragma unroll 1
  for (int leaf = 0; leaf < 2; ++leaf) {
    if (I.soybean) btran = dmin(1.0, btran * 1.25);
    const double rs = psn_stomatal_resistance(P, PC, PT, I.nrad, pbot, svpts, eah, I.forc_po2, I.forc_pco2, rb, btran,
                                              leaf ? I.vcsha : I.vcsun, leaf ? I.parsha : I.parsun,
                                              leaf ? I.laisha_z : I.laisun_z, err);
    if (leaf) rssha = rs; else rssun = rs;
  }
  I.btran = btran;

  // sensible-heat conductances: air, leaf, ground
  const double wta = 1.0 / rah0;
  const double wtl = (elai + esai) / rb;
  const double wtg = 1.0 / rah1;
  const double wtshi = 1.0 / (wta + wtl + wtg);
  const double wtl0 = wtl * wtshi;
  const double wtg0 = wtg * wtshi;
  const double wta0 = wta * wtshi;
  const double wtga = wta0 + wtg0;
  I.wtg = wtg; I.wtl0 = wtl0; I.wta0 = wta0;
  I.wtal = wta0 + wtl0;

  // fraction of potential evaporation from the leaf
  double rppdry;
  if (I.fdry > 0.0) {
    rppdry = I.fdry * rb * (I.laisun / (rb + rssun) + I.laisha / (rb + rssha)) / elai;
  } else {
    rppdry = 0.0;
  }
  double efpot = forc_rho * wtl * (I.qsatl - I.qaf);
  double rpp;
  double qflx_tran_veg;
  if (efpot > 0.0) {
    if (btran > 0.0) {
      qflx_tran_veg = efpot * rppdry;
      rpp = rppdry + I.fwet;
    } else {
      rpp = I.fwet;
      qflx_tran_veg = 0.0;
    }
    rpp = dmin(rpp, (qflx_tran_veg + h2ocan0 / dtime) / efpot);
  } else {
    rpp = 1.0;
    qflx_tran_veg = 0.0;
  }

  // latent-heat conductances, with the dry-litter layer resistance
  const double wtaq = veg / raw0;
  const double wtlq = veg * (elai + esai) / rb * rpp;
  const double fsno_dl = I.snow_depth / 0.05;
  const double elai_dl = 0.5 * (1.0 - dmin(fsno_dl, 1.0));
  const double rdl = (1.0 - m_exp(-elai_dl)) / (0.004 * uaf);
  double wtgq;
  if (I.delq < 0.0) {
    wtgq = veg / (raw1 + rdl);
  } else {
    wtgq = I.soilbeta * veg / (raw1 + rdl);
  }
  const double wtsqi = 1.0 / (wtaq + wtlq + wtgq);
  const double wtgq0 = wtgq * wtsqi;
  const double wtlq0 = wtlq * wtsqi;
  const double wtaq0 = wtaq * wtsqi;
  const double wtgaq = wtaq0 + wtgq0;
  I.wtgq = wtgq; I.wtlq0 = wtlq0; I.wtaq0 = wtaq0;
  I.wtalq = wtaq0 + wtlq0;
  const double dc1 = forc_rho * CPAIR * wtl;
  const double dc2 = HVAP * forc_rho * wtlq;
  const double efsh = dc1 * (wtga * t_veg - wtg0 * tg - wta0 * thm);
  double efe = dc2 * (wtgaq * I.qsatl - wtgq0 * qg - wtaq0 * forc_q);
  double erre = 0.0;
  if ((efe * I.efeb) < 0.0) {
    const double efeold = efe;
    efe = 0.1 * efeold;
    erre = efe - efeold;
  }

  // leaf energy balance: Newton step on t_veg, limited to 1 K per iteration
  double dt_veg = (I.sabv + I.air + I.bir * pow4(t_veg) + I.cir * I.lw_grnd - efsh - efe) /
                  (-4.0 * I.bir * cube(t_veg) + dc1 * wtga + dc2 * wtgaq * I.qsatldT);
  t_veg = tlbef + dt_veg;
  const double dels = dt_veg;
  const double del = fabs(dels);
  I.del = del;
  double errb = 0.0;
  if (del > 1.0) {
    dt_veg = dels / del;
    t_veg = tlbef + dt_veg;
    errb = I.sabv + I.air + I.bir * cube(tlbef) * (tlbef + 4.0 * dt_veg) + I.cir * I.lw_grnd - (efsh + dc1 * wtga * dt_veg) -
           (efe + dc2 * wtgaq * I.qsatldT * dt_veg);
  }
  I.dt_veg = dt_veg;

  // fluxes from leaves to canopy air
  efpot = forc_rho * wtl * (wtgaq * (I.qsatl + I.qsatldT * dt_veg) - wtgq0 * qg - wtaq0 * forc_q);
  double qflx_evap_veg = rpp * efpot;
  if (efpot > 0.0 && btran > 0.0) {
    qflx_tran_veg = efpot * rppdry;
  } else {
    qflx_tran_veg = 0.0;
  }
  const double ecidif = dmax(0.0, qflx_evap_veg - qflx_tran_veg - h2ocan0 / dtime);
  qflx_evap_veg = dmin(qflx_evap_veg, qflx_tran_veg + h2ocan0 / dtime);
  I.qflx_tran_veg = qflx_tran_veg;
  I.qflx_evap_veg = qflx_evap_veg;
  I.eflx_sh_veg = efsh + dc1 * wtga * dt_veg + errb + erre + HVAP * ecidif;
  double deldT;
  qsat(t_veg, pbot, I.el, deldT, I.qsatl, I.qsatldT);
  I.t_veg = t_veg;

  // canopy air state and Monin-Obukhov length for the next pass
  I.taf = wtg0 * tg + wta0 * thm + wtl0 * t_veg;
  I.qaf = wtlq0 * I.qsatl + wtgq0 * qg + forc_q * wtaq0;
  I.dth = thm - I.taf;
  I.dqh = forc_q - I.qaf;
  I.delq = I.wtalq * qg - wtlq0 * I.qsatl - wtaq0 * forc_q;
  const double tstar = p.temp1 * I.dth;
  const double qstar = p.temp2 * I.dqh;
  const double thvstar = tstar * (1.0 + 0.61 * forc_q) + 0.61 * I.forc_th * qstar;
  double zeta = I.zldis * VKC * GRAV * thvstar / (sq(p.ustar) * I.thv);
  if (zeta >= 0.0) {
    zeta = dmin(2.0, dmax(zeta, 0.01));
    I.um = dmax(I.ur, 0.1);
  } else {
    zeta = dmax(-100.0, dmin(zeta, -0.01));
    const double wc = 1.0 * m_pow((-GRAV * p.ustar * thvstar * 1000.0 / I.thv), 0.333);
    I.um = sqrt(I.ur * I.ur + wc * wc);
  }
  double obu = I.zldis / zeta;
  if (I.obuold * obu < 0.0) I.nmozsgn += 1;
  if (I.nmozsgn >= 4) obu = I.zldis / (-0.01);
  I.obu = obu;
  I.obuold = obu;
  I.err = (int)err;

  // convergence: at least three passes, leaf temperature within 0.01 K twice in a row and the
  // latent heat flux within 0.1 W/m2
  bool stop = false;
  I.itlef += 1;
  if (I.itlef > itmin) {
    const double dele = fabs(efe - I.efeb);
    I.efeb = efe;
    const double det = dmax(del, del2);
    if ((det < dtmin) && (dele < dlemin)) stop = true;
  }
  return stop || !(I.itlef <= itmax);
}

// compute_flux (:482-539) and the write-back of the iteration results; p_temp12m / p_temp22m: the 2 m profile relations
// of the last pass
ELMK_HD void canflux_end_with(const Cols& S, const int c, const CanopyIter& I, const double p_temp12m, const double p_temp22m)
{
  const double t_veg = I.t_veg, thm = I.thm, tg = I.tg, forc_q = I.forc_q, forc_rho = I.forc_rho;
  const double wtg = I.wtg, wtl0 = I.wtl0, wta0 = I.wta0, wtal = I.wtal, wtgq = I.wtgq, wtalq = I.wtalq,
               wtlq0 = I.wtlq0, wtaq0 = I.wtaq0, qsatl = I.qsatl;
  const double emv = I.emv, emg = I.emg, forc_lwrad = I.forc_lwrad, tlbef = I.tlbef, dt_veg = I.dt_veg;
#ifdef ELMK_DEBUG_ITER
  C1(altmax_lastyear_indx) = I.itlef;   // development aid: outer iteration count into a field the chain never reads
#endif
  C1(btran) = I.btran;
  C1(t_veg) = t_veg;
  C1(qflx_tran_veg) = I.qflx_tran_veg;
  C1(qflx_evap_veg) = I.qflx_evap_veg;
  C1(eflx_sh_veg) = I.eflx_sh_veg;

  const double htvp = C1(htvp);
  const double delt = wtal * tg - wtl0 * t_veg - wta0 * thm;
  C1(eflx_sh_grnd) = CPAIR * forc_rho * wtg * delt;
  const double delt_snow = wtal * I.t_snotop - wtl0 * t_veg - wta0 * thm;
  C1(eflx_sh_snow) = CPAIR * forc_rho * wtg * delt_snow;
  const double delt_soil = wtal * I.t_soil1 - wtl0 * t_veg - wta0 * thm;
  C1(eflx_sh_soil) = CPAIR * forc_rho * wtg * delt_soil;
  const double delt_h2osfc = wtal * I.t_sfc - wtl0 * t_veg - wta0 * thm;
  C1(eflx_sh_h2osfc) = CPAIR * forc_rho * wtg * delt_h2osfc;
  C1(qflx_evap_soi) = forc_rho * wtgq * I.delq;
  const double delq_snow = wtalq * C1(qg_snow) - wtlq0 * qsatl - wtaq0 * forc_q;
  C1(qflx_ev_snow) = forc_rho * wtgq * delq_snow;
  const double delq_soil = wtalq * C1(qg_soil) - wtlq0 * qsatl - wtaq0 * forc_q;
  C1(qflx_ev_soil) = forc_rho * wtgq * delq_soil;
  const double delq_h2osfc = wtalq * C1(qg_h2osfc) - wtlq0 * qsatl - wtaq0 * forc_q;
  C1(qflx_ev_h2osfc) = forc_rho * wtgq * delq_h2osfc;
  const double t_ref2m = thm + I.p_temp1 * I.dth * (1.0 / p_temp12m - 1.0 / I.p_temp1);
  const double q_ref2m = forc_q + I.p_temp2 * I.dqh * (1.0 / p_temp22m - 1.0 / I.p_temp2);
  double e2m, de2m, qsat2m, dqsat2m;
  qsat(t_ref2m, I.pbot, e2m, de2m, qsat2m, dqsat2m);
  C1(t_ref2m) = t_ref2m;
  C1(q_ref2m) = q_ref2m;
  C1(rh_ref2m) = dmin(100.0, (q_ref2m / qsat2m) * 100.0);
  // (products kept in the reference's left-to-right association)
  C1(dlrad) = (1.0 - emv) * emg * forc_lwrad + emv * emg * STEBOL * cube(tlbef) * (tlbef + 4.0 * dt_veg);
  C1(ulrad) = ((1.0 - emg) * (1.0 - emv) * (1.0 - emv) * forc_lwrad +
               emv * (1.0 + (1.0 - emg) * (1.0 - emv)) * STEBOL * cube(tlbef) * (tlbef + 4.0 * dt_veg) +
               emg * (1.0 - emv) * STEBOL * I.lw_grnd);
  double cgrnds = 0.0, cgrndl = 0.0;
  cgrnds += CPAIR * forc_rho * wtg * wtal;
  cgrndl += forc_rho * wtgq * wtalq * C1(dqgdT);
  C1(cgrnds) = cgrnds;
  C1(cgrndl) = cgrndl;
  C1(cgrnd) = cgrnds + cgrndl * htvp;
  C1(h2ocan) = dmax(0.0, I.h2ocan0 + (I.qflx_tran_veg - I.qflx_evap_veg) * I.dtime);
  if (I.err) C1(errmask) |= I.err;
}

// the 2 m relations the reference evaluates in every pass (friction_velocity_temp2m / _humidity2m with z0h = z0q = z0m),
// from the Obukhov length the last pass started with
ELMK_HD double canflux_temp12m(const CanopyIter& I) { return mo_scalar_profile(2.0 + I.z0mv, I.p_obu, I.z0mv, true); }
ELMK_HD void canflux_end(const Cols& S, const int c, const CanopyIter& I)
{
  const double p_temp12m = canflux_temp12m(I);
  canflux_end_with(S, c, I, p_temp12m, p_temp12m);   // same roughness length for heat and moisture
}

ELMK_HD void column_canopy_fluxes(const Cols& S, const Tables& T, const StepArgs& A, const int c)
{
  const PsnPft P = load_psn_pft(S, c);
  CanopyIter I;
  if (!canflux_begin(S, T.vtype, A, P, c, I)) return;
  const PsnColumn PC = psn_column(P, I.t10, I.pbot, I.thm, I.forc_po2, I.dayl_factor);
#pragma unroll 1
  while (!canflux_iterate(P, PC, I)) {
  }
  canflux_end(S, c, I);
}

} // namespace elmk
