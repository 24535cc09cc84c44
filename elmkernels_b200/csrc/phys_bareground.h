// phys_bareground.h - bare-ground surface fluxes (group a6): three fixed Monin-Obukhov stability
// iterations, then sensible/latent heat fluxes and the 2 m diagnostics, for columns without
// exposed vegetation (frac_veg_nosno == 0).
//
// Parity target: kokkos_bareground_fluxes, reference
// driver/kokkos/bareground_fluxes_kokkos.cc:7-123 -> initialize_flux :7, stability_iteration :30,
// compute_flux :82 (src/physics/bareground_fluxes_impl.hh), derive_forc_rho
// (src/physics/atm_physics_impl.hh:249-263).  The wrapper's 12 scratch Views are registers here.
#pragma once
#include "elmk_state.h"
#include "phys_cantemp.h"
#include "phys_friction.h"

namespace elmk {

// air density from pressure, specific humidity and temperature
ELMK_HD double air_density(const double pbot, const double qbot, const double tbot)
{
  const double vp = qbot * pbot / (0.622 + 0.378 * qbot);
  return (pbot - 0.378 * vp) / (RAIR * tbot);
}

ELMK_HD void column_bareground_fluxes(const Cols& S, const Tables&, const int c)
{
  // compute_flux zeroes these for every column, vegetated or not (:104-108)
  C1(cgrnd) = 0.0;
  C1(cgrnds) = 0.0;
  C1(cgrndl) = 0.0;
  if (C1(frac_veg_nosno) != 0) return;

  const double forc_q = C1(forc_qbot), forc_th = C1(forc_thbot), pbot = C1(forc_pbot);
  const double thm = C1(thm), thv = C1(thv), tg = C1(t_grnd);
  const double z0mg = C1(z0mg);
  const double hgt_u = C1(forc_hgt_u_patch), hgt_t = C1(forc_hgt_t_patch), hgt_q = C1(forc_hgt_q_patch);
  const double forc_rho = air_density(pbot, forc_q, C1(forc_tbot));

  // -- initialize_flux --
  const double fu = C1(forc_u), fv = C1(forc_v);
  const double ur = dmax(1.0, sqrt(fu * fu + fv * fv));
  const double dth = thm - tg;
  const double dqh = forc_q - C1(qg);
  const double zldis = hgt_u;
  const double dthv = dth * (1.0 + 0.61 * forc_q) + 0.61 * forc_th * dqh;
  constexpr double displa = 0.0;
  C1(dlrad) = 0.0;
  C1(ulrad) = 0.0;
  double um, obu;
  mo_initial_length(ur, thv, dthv, zldis, z0mg, um, obu);

  // -- stability_iteration: exactly three passes, no convergence test --
  double z0hg = C1(z0hg), z0qg = C1(z0qg);
  // (friction velocity and the temperature relation are evaluated together; the 2 m relations, which the reference
  //  evaluates in every pass but reads after the last one only, once after the loop from that pass's inputs)
  MoProfiles p;
  double obu_p = obu, z0h_p = z0hg, z0q_p = z0qg;
#pragma unroll 1
  for (int it = 0; it < 3; ++it) {
    obu_p = obu; z0h_p = z0hg; z0q_p = z0qg;
    {
      const MoPair mp = mo_pair_inl(hgt_u - displa, hgt_t - displa, um, obu, z0mg, z0hg);
      p.ustar = mp.ustar;
      p.temp1 = mp.temp;
    }
    p.temp2 = (hgt_q == hgt_t && z0qg == z0hg) ? p.temp1 : mo_scalar_profile(hgt_q - displa, obu, z0qg);
    const double tstar = p.temp1 * dth;
    const double qstar = p.temp2 * dqh;
    const double thvstar = tstar * (1.0 + 0.61 * forc_q) + 0.61 * forc_th * qstar;
    z0hg = z0mg / m_exp(0.13 * m_pow((p.ustar * z0mg / 1.5e-5), 0.45));
    z0qg = z0hg;
    double zeta = zldis * VKC * GRAV * thvstar / (sq(p.ustar) * thv);
    if (zeta >= 0.0) {
      zeta = dmin(2.0, dmax(zeta, 0.01));
      um = dmax(ur, 0.1);
    } else {
      zeta = dmax(-100.0, dmin(zeta, -0.01));
      const double wc = 1.0 * m_pow((-GRAV * p.ustar * thvstar * 1000.0 / thv), 0.333);
      um = sqrt(ur * ur + wc * wc);
    }
    obu = zldis / zeta;
  }
  C1(z0hg) = z0hg;
  C1(z0qg) = z0qg;
  p.temp12m = mo_scalar_profile(2.0 + z0h_p, obu_p, z0h_p, true);
  p.temp22m = (z0q_p == z0h_p) ? p.temp12m : mo_scalar_profile(2.0 + z0q_p, obu_p, z0q_p);

  // -- compute_flux --
  const double rah = 1.0 / (p.temp1 * p.ustar);
  const double raw = 1.0 / (p.temp2 * p.ustar);
  const double raih = forc_rho * CPAIR / rah;
  const double raiw = (dqh > 0.0) ? forc_rho / raw : C1(soilbeta) * forc_rho / raw;
  const double htvp = C1(htvp);
  const double cgrnds = raih;
  const double cgrndl = raiw * C1(dqgdT);
  C1(cgrnds) = cgrnds;
  C1(cgrndl) = cgrndl;
  C1(cgrnd) = cgrnds + htvp * cgrndl;
  const double sh_grnd = -raih * dth;
  C1(eflx_sh_grnd) = sh_grnd;
  C1(eflx_sh_tot) = sh_grnd;
  const int snl = C1(snl);
  C1(eflx_sh_snow) = -raih * (thm - C2(t_soisno, NLEVSNO - snl));
  C1(eflx_sh_soil) = -raih * (thm - C2(t_soisno, NLEVSNO));
  C1(eflx_sh_h2osfc) = -raih * (thm - C1(t_h2osfc));
  const double evap_soi = -raiw * dqh;
  C1(qflx_evap_soi) = evap_soi;
  C1(qflx_evap_tot) = evap_soi;
  C1(qflx_ev_snow) = -raiw * (forc_q - C1(qg_snow));
  C1(qflx_ev_soil) = -raiw * (forc_q - C1(qg_soil));
  C1(qflx_ev_h2osfc) = -raiw * (forc_q - C1(qg_h2osfc));
  const double t_ref2m = thm + p.temp1 * dth * (1.0 / p.temp12m - 1.0 / p.temp1);
  const double q_ref2m = forc_q + p.temp2 * dqh * (1.0 / p.temp22m - 1.0 / p.temp2);
  double e2m, de2m, qsat2m, dqsat2m;
  qsat(t_ref2m, pbot, e2m, de2m, qsat2m, dqsat2m);
  C1(t_ref2m) = t_ref2m;
  C1(q_ref2m) = q_ref2m;
  C1(rh_ref2m) = dmin(100.0, (q_ref2m / qsat2m * 100.0));
}

} // namespace elmk
