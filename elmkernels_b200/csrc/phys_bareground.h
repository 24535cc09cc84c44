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

// The three functions of the group (reference bareground_fluxes_impl.hh) for a column with frac_veg_nosno == 0; the
// library-level API of include/elm/bareground_fluxes.h calls them one by one through elmk_fn_call, the column body
// below composes them.
namespace bgf {

// initialize_flux :7-27
ELMK_HD void initialize_flux(const double forc_u, const double forc_v, const double forc_q, const double forc_th,
                             const double hgt_u, const double thm, const double thv, const double t_grnd, const double qg,
                             const double z0mg, double& dlrad, double& ulrad, double& zldis, double& displa, double& dth,
                             double& dqh, double& obu, double& ur, double& um)
{
  ur = dmax(1.0, sqrt(forc_u * forc_u + forc_v * forc_v));
  dth = thm - t_grnd;
  dqh = forc_q - qg;
  zldis = hgt_u;
  const double dthv = dth * (1.0 + 0.61 * forc_q) + 0.61 * forc_th * dqh;
  displa = 0.0;
  dlrad = 0.0;
  ulrad = 0.0;
  mo_initial_length(ur, thv, dthv, zldis, z0mg, um, obu);
}

// stability_iteration :30-79: exactly three passes, no convergence test.
// Friction velocity and the temperature relation of a pass are evaluated together (mo_pair_inl); the 2 m relations,
// which the reference evaluates in every pass but reads after the last one only, once after the loop from the last
// pass's inputs.
ELMK_HD void stability_iteration(const double hgt_t, const double hgt_u, const double hgt_q, const double z0mg,
                                 const double zldis, const double displa, const double dth, const double dqh, const double ur,
                                 const double forc_q, const double forc_th, const double thv, double& z0hg, double& z0qg,
                                 double& obu, double& um, double& temp1, double& temp2, double& temp12m, double& temp22m,
                                 double& ustar)
{
  double obu_p = obu, z0h_p = z0hg, z0q_p = z0qg;
#pragma unroll 1
  for (int it = 0; it < 3; ++it) {
    obu_p = obu; z0h_p = z0hg; z0q_p = z0qg;
    {
      const MoPair mp = mo_pair_inl(hgt_u - displa, hgt_t - displa, um, obu, z0mg, z0hg);
      ustar = mp.ustar;
      temp1 = mp.temp;
    }
    temp2 = (hgt_q == hgt_t && z0qg == z0hg) ? temp1 : mo_scalar_profile(hgt_q - displa, obu, z0qg);
    const double tstar = temp1 * dth;
    const double qstar = temp2 * dqh;
    const double thvstar = tstar * (1.0 + 0.61 * forc_q) + 0.61 * forc_th * qstar;
    z0hg = z0mg / m_exp(0.13 * m_pow((ustar * z0mg / 1.5e-5), 0.45));
    z0qg = z0hg;
    double zeta = zldis * VKC * GRAV * thvstar / (sq(ustar) * thv);
    if (zeta >= 0.0) {
      zeta = dmin(2.0, dmax(zeta, 0.01));
      um = dmax(ur, 0.1);
    } else {
      zeta = dmax(-100.0, dmin(zeta, -0.01));
      const double wc = 1.0 * m_pow((-GRAV * ustar * thvstar * 1000.0 / thv), 0.333);
      um = sqrt(ur * ur + wc * wc);
    }
    obu = zldis / zeta;
  }
  temp12m = mo_scalar_profile(2.0 + z0h_p, obu_p, z0h_p, true);
  temp22m = (z0q_p == z0h_p) ? temp12m : mo_scalar_profile(2.0 + z0q_p, obu_p, z0q_p);
}

// compute_flux :82-170 for an unvegetated column; t_top = t_soisno(nlevsno - snl), t_soil1 = t_soisno(nlevsno)
struct Fluxes {
  double cgrnds, cgrndl, cgrnd, eflx_sh_grnd, eflx_sh_snow, eflx_sh_soil, eflx_sh_h2osfc, qflx_evap_soi, qflx_ev_snow,
      qflx_ev_soil, qflx_ev_h2osfc, t_ref2m, q_ref2m, rh_ref2m;   // (eflx_sh_tot = eflx_sh_grnd, qflx_evap_tot = qflx_evap_soi)
};
ELMK_HD Fluxes compute_flux(const double forc_rho, const double soilbeta, const double dqgdT, const double htvp,
                            const double t_h2osfc, const double qg_snow, const double qg_soil, const double qg_h2osfc,
                            const double t_top, const double t_soil1, const double pbot, const double dth, const double dqh,
                            const double temp1, const double temp2, const double temp12m, const double temp22m,
                            const double ustar, const double forc_q, const double thm)
{
  Fluxes f;
  const double rah = 1.0 / (temp1 * ustar);
  const double raw = 1.0 / (temp2 * ustar);
  const double raih = forc_rho * CPAIR / rah;
  const double raiw = (dqh > 0.0) ? forc_rho / raw : soilbeta * forc_rho / raw;
  f.cgrnds = raih;
  f.cgrndl = raiw * dqgdT;
  f.cgrnd = f.cgrnds + htvp * f.cgrndl;
  f.eflx_sh_grnd = -raih * dth;
  f.eflx_sh_snow = -raih * (thm - t_top);
  f.eflx_sh_soil = -raih * (thm - t_soil1);
  f.eflx_sh_h2osfc = -raih * (thm - t_h2osfc);
  f.qflx_evap_soi = -raiw * dqh;
  f.qflx_ev_snow = -raiw * (forc_q - qg_snow);
  f.qflx_ev_soil = -raiw * (forc_q - qg_soil);
  f.qflx_ev_h2osfc = -raiw * (forc_q - qg_h2osfc);
  f.t_ref2m = thm + temp1 * dth * (1.0 / temp12m - 1.0 / temp1);
  f.q_ref2m = forc_q + temp2 * dqh * (1.0 / temp22m - 1.0 / temp2);
  double e2m, de2m, qsat2m, dqsat2m;
  qsat(f.t_ref2m, pbot, e2m, de2m, qsat2m, dqsat2m);
  f.rh_ref2m = dmin(100.0, (f.q_ref2m / qsat2m * 100.0));
  return f;
}

} // namespace bgf

ELMK_HD void column_bareground_fluxes(const Cols& S, const Tables&, const int c)
{
  // compute_flux zeroes these for every column, vegetated or not (:104-108)
  C1(cgrnd) = 0.0;
  C1(cgrnds) = 0.0;
  C1(cgrndl) = 0.0;
  if (C1(frac_veg_nosno) != 0) return;

  const double forc_q = C1(forc_qbot), forc_th = C1(forc_thbot), pbot = C1(forc_pbot);
  const double thm = C1(thm), thv = C1(thv);
  const double z0mg = C1(z0mg);
  const double hgt_u = C1(forc_hgt_u_patch), hgt_t = C1(forc_hgt_t_patch), hgt_q = C1(forc_hgt_q_patch);
  const double forc_rho = air_density(pbot, forc_q, C1(forc_tbot));

  double dlrad, ulrad, zldis, displa, dth, dqh, obu, ur, um;
  bgf::initialize_flux(C1(forc_u), C1(forc_v), forc_q, forc_th, hgt_u, thm, thv, C1(t_grnd), C1(qg), z0mg, dlrad, ulrad, zldis,
                       displa, dth, dqh, obu, ur, um);
  C1(dlrad) = dlrad;
  C1(ulrad) = ulrad;

  double z0hg = C1(z0hg), z0qg = C1(z0qg), temp1, temp2, temp12m, temp22m, ustar;
  bgf::stability_iteration(hgt_t, hgt_u, hgt_q, z0mg, zldis, displa, dth, dqh, ur, forc_q, forc_th, thv, z0hg, z0qg, obu, um,
                           temp1, temp2, temp12m, temp22m, ustar);
  C1(z0hg) = z0hg;
  C1(z0qg) = z0qg;

  const int snl = C1(snl);
  const bgf::Fluxes f = bgf::compute_flux(forc_rho, C1(soilbeta), C1(dqgdT), C1(htvp), C1(t_h2osfc), C1(qg_snow), C1(qg_soil),
                                          C1(qg_h2osfc), C2(t_soisno, NLEVSNO - snl), C2(t_soisno, NLEVSNO), pbot, dth, dqh,
                                          temp1, temp2, temp12m, temp22m, ustar, forc_q, thm);
  C1(cgrnds) = f.cgrnds;
  C1(cgrndl) = f.cgrndl;
  C1(cgrnd) = f.cgrnd;
  C1(eflx_sh_grnd) = f.eflx_sh_grnd;
  C1(eflx_sh_tot) = f.eflx_sh_grnd;
  C1(eflx_sh_snow) = f.eflx_sh_snow;
  C1(eflx_sh_soil) = f.eflx_sh_soil;
  C1(eflx_sh_h2osfc) = f.eflx_sh_h2osfc;
  C1(qflx_evap_soi) = f.qflx_evap_soi;
  C1(qflx_evap_tot) = f.qflx_evap_soi;
  C1(qflx_ev_snow) = f.qflx_ev_snow;
  C1(qflx_ev_soil) = f.qflx_ev_soil;
  C1(qflx_ev_h2osfc) = f.qflx_ev_h2osfc;
  C1(t_ref2m) = f.t_ref2m;
  C1(q_ref2m) = f.q_ref2m;
  C1(rh_ref2m) = f.rh_ref2m;
}

} // namespace elmk
