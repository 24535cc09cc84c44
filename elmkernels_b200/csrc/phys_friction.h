// phys_friction.h - Monin-Obukhov surface-layer similarity (Zeng et al. 1998): friction velocity and
// the temperature / humidity profile relations, shared by the bare-ground (a6) and canopy (a7) fluxes.
//
// Parity target: reference src/physics/friction_velocity_impl.hh:17-173
//   StabilityFunc1 :17, StabilityFunc2 :26, monin_obukhov_length :35, friction_velocity_wind :62,
//   friction_velocity_temp :84, friction_velocity_humidity :105, friction_velocity_temp2m :133,
//   friction_velocity_humidity2m :152.
// The literal exponent 0.333 (not 1/3) is the reference's.
#pragma once
#include "elmk_common.h"

namespace elmk {

ELMK_HD double mo_psi_m(const double zeta)   // wind-profile stability correction, unstable branch
{
  const double chik2 = sqrt(1.0 - 16.0 * zeta);
  const double chik = sqrt(chik2);
  return 2.0 * m_log((1.0 + chik) * 0.5) + m_log((1.0 + chik2) * 0.5) - 2.0 * m_atan(chik) + PI * 0.5;
}

ELMK_HD double mo_psi_h(const double zeta)   // scalar-profile stability correction, unstable branch
{
  const double chik2 = sqrt(1.0 - 16.0 * zeta);
  return 2.0 * m_log((1.0 + chik2) * 0.5);
}

// initial Monin-Obukhov length and wind speed from the bulk Richardson number
ELMK_HD void mo_initial_length(const double ur, const double thv, const double dthv, const double zldis,
                               const double z0m, double& um, double& obu)
{
  constexpr double wc = 0.5;
  if (dthv >= 0.0) {
    um = dmax(ur, 0.1);
  } else {
    um = sqrt(ur * ur + wc * wc);
  }
  const double rib = GRAV * zldis * dthv / (thv * um * um);
  double zeta;
  if (rib >= 0.0) {
    zeta = rib * m_log(zldis / z0m) / (1.0 - 5.0 * dmin(rib, 0.19));
    zeta = dmin(2.0, dmax(zeta, 0.01));
  } else {
    zeta = rib * m_log(zldis / z0m);
    zeta = dmax(-100.0, dmin(zeta, -0.01));
  }
  obu = zldis / zeta;
}

// The four stability regimes of the two profile functions below share their leading logarithm, and the two
// unstable (resp. stable) regimes share everything but one term.  A warp holds columns of every regime, so the
// functions are written regime-convergent: pick the regime's operands first, then evaluate each transcendental
// once for all lanes that need it, instead of once per regime branch (ncu, round 1: 10 active lanes per
// instruction in this file with the reference's if / else-if ladder).  Every lane still performs exactly the
// reference's operations in the reference's order.
//
// psi(-zetam), psi(-zetat) and the two constant powers are values of the reference's expressions at constant
// arguments (evaluated with the same libm, tools/mo_constants.py); the host checker build keeps the calls.
#ifdef ELMK_EXACT_POW
#define ELMK_MO_PSI_M_ZETAM mo_psi_m(-1.574)
#define ELMK_MO_PSI_H_ZETAT mo_psi_h(-0.465)
#define ELMK_MO_POW_ZETAM m_pow(1.574, 0.333)
#define ELMK_MO_POW_ZETAT m_pow(0.465, -0.333)
#else
#define ELMK_MO_PSI_M_ZETAM 0x1.5ba94811aa58ep+0
#define ELMK_MO_PSI_H_ZETAT 0x1.569b4c1ef37bbp+0
#define ELMK_MO_POW_ZETAM 0x1.29be614584642p+0
#define ELMK_MO_POW_ZETAT 0x1.4a5a581293329p+0
#endif

ELMK_HD_NOINLINE double mo_ustar(const double forc_hgt_u, const double displa, const double um, const double obu,
                        const double z0m)
{
  constexpr double zetam = 1.574;
  const double zldis = forc_hgt_u - displa;
  const double zeta = zldis / obu;
  const bool unstable = (zeta < 0.0);
  const bool far = unstable ? (zeta < (-zetam)) : !(zeta <= 1.0);   // very unstable / very stable
  const double num = far ? (unstable ? -zetam * obu : obu) : zldis;
  const double lead = m_log(num / z0m);
  double den;
  if (unstable) {
    const double p1 = far ? ELMK_MO_PSI_M_ZETAM : mo_psi_m(zeta);
    den = lead - p1 + mo_psi_m(z0m / obu);
    if (far) den = den + 1.14 * (m_pow((-zeta), 0.333) - ELMK_MO_POW_ZETAM);
  } else {
    const double t = 5.0 * z0m / obu;
    if (far) den = lead + 5.0 - t + (5.0 * m_log(zeta) + zeta - 1.0);
    else den = lead + 5.0 * zeta - t;
  }
  return VKC * um / den;
}

// scalar (temperature or humidity) profile relation for a reference height `zldis` above the
// displacement height and roughness length z0
// `grouped`: friction_velocity_temp2m writes the very-stable branch as 5 (z0/L) instead of (5 z0)/L
// (:147), which can differ in the last bit; the other four callers use the ungrouped product.
ELMK_HD_NOINLINE double mo_scalar_profile(const double zldis, const double obu, const double z0, const bool grouped = false)
{
  constexpr double zetat = 0.465;
  const double zeta = zldis / obu;
  const bool unstable = (zeta < 0.0);
  const bool far = unstable ? (zeta < (-zetat)) : !(zeta <= 1.0);
  const double num = far ? (unstable ? -zetat * obu : obu) : zldis;
  const double lead = m_log(num / z0);
  double den;
  if (unstable) {
    const double p1 = far ? ELMK_MO_PSI_H_ZETAT : mo_psi_h(zeta);
    den = lead - p1 + mo_psi_h(z0 / obu);
    if (far) den = den + 0.8 * (ELMK_MO_POW_ZETAT - m_pow((-zeta), -0.333));
  } else if (far) {
    const double stable = grouped ? 5.0 * (z0 / obu) : 5.0 * z0 / obu;
    den = lead + 5.0 - stable + (5.0 * m_log(zeta) + zeta - 1.0);
  } else {
    den = lead + 5.0 * zeta - 5.0 * z0 / obu;
  }
  return VKC / den;
}

// Friction velocity and one scalar profile relation at the same Obukhov length in ONE called function whose
// logarithms, arctangents, square roots and divisions are in line (`_inl`: the build leaves its divisions alone).
// Called one after the other, mo_ustar and mo_scalar_profile are two strings of up to six serial transcendental
// calls; here the two leading logarithms, the four stability corrections of the unstable case and the quotients
// are independent chains that ptxas interleaves.  Same operations per value as the two functions above; for a
// very unstable lane psi is evaluated at the constant argument instead of being a stored constant.
struct MoPair {
  double ustar, temp;
};
ELMK_HD double mo_psi_m_i(const double zeta)
{
  const double chik2 = sqrt(1.0 - 16.0 * zeta);
  const double chik = sqrt(chik2);
  return 2.0 * i_log((1.0 + chik) * 0.5) + i_log((1.0 + chik2) * 0.5) - 2.0 * i_atan(chik) + PI * 0.5;
}
ELMK_HD double mo_psi_h_i(const double zeta)
{
  const double chik2 = sqrt(1.0 - 16.0 * zeta);
  return 2.0 * i_log((1.0 + chik2) * 0.5);
}
ELMK_HD_NOINLINE MoPair mo_pair_inl(const double zldis_u, const double zldis_s, const double um, const double obu,
                                    const double z0m, const double z0s)
{
  constexpr double zetam = 1.574, zetat = 0.465;
  const double zeta_u = zldis_u / obu, zeta_s = zldis_s / obu;
  const bool un_u = (zeta_u < 0.0), un_s = (zeta_s < 0.0);
  if (un_u != un_s) {   // (cannot happen for heights above the displacement height; keeps the semantics anyway)
    MoPair r;
    r.ustar = mo_ustar(zldis_u, 0.0, um, obu, z0m);
    r.temp = mo_scalar_profile(zldis_s, obu, z0s);
    return r;
  }
  const bool far_u = un_u ? (zeta_u < (-zetam)) : !(zeta_u <= 1.0);
  const bool far_s = un_s ? (zeta_s < (-zetat)) : !(zeta_s <= 1.0);
  const double num_u = far_u ? (un_u ? -zetam * obu : obu) : zldis_u;
  const double num_s = far_s ? (un_s ? -zetat * obu : obu) : zldis_s;
  const double lead_u = i_log(num_u / z0m);
  const double lead_s = i_log(num_s / z0s);
  double den_u, den_s;
  if (un_u) {
    const double p1u = mo_psi_m_i(far_u ? -zetam : zeta_u);
    const double p2u = mo_psi_m_i(z0m / obu);
    const double p1s = mo_psi_h_i(far_s ? -zetat : zeta_s);
    const double p2s = mo_psi_h_i(z0s / obu);
    den_u = lead_u - p1u + p2u;
    den_s = lead_s - p1s + p2s;
    if (far_u) den_u = den_u + 1.14 * (m_pow((-zeta_u), 0.333) - ELMK_MO_POW_ZETAM);
    if (far_s) den_s = den_s + 0.8 * (ELMK_MO_POW_ZETAT - m_pow((-zeta_s), -0.333));
  } else {
    const double tu = 5.0 * z0m / obu;
    const double ts = 5.0 * z0s / obu;
    if (far_u) den_u = lead_u + 5.0 - tu + (5.0 * i_log(zeta_u) + zeta_u - 1.0);
    else den_u = lead_u + 5.0 * zeta_u - tu;
    if (far_s) den_s = lead_s + 5.0 - ts + (5.0 * i_log(zeta_s) + zeta_s - 1.0);
    else den_s = lead_s + 5.0 * zeta_s - ts;
  }
  MoPair r;
  r.ustar = VKC * um / den_u;
  r.temp = VKC / den_s;
  return r;
}

// the five profile quantities of one stability iteration
struct MoProfiles {
  double ustar, temp1, temp2, temp12m, temp22m;
};

ELMK_HD MoProfiles mo_profiles(const double hgt_u, const double hgt_t, const double hgt_q, const double displa,
                               const double um, const double obu, const double z0m, const double z0h,
                               const double z0q)
{
  MoProfiles p;
  p.ustar = mo_ustar(hgt_u, displa, um, obu, z0m);
  p.temp1 = mo_scalar_profile(hgt_t - displa, obu, z0h);
  p.temp2 = (hgt_q == hgt_t && z0q == z0h) ? p.temp1 : mo_scalar_profile(hgt_q - displa, obu, z0q);
  p.temp12m = mo_scalar_profile(2.0 + z0h, obu, z0h, true);
  p.temp22m = (z0q == z0h) ? p.temp12m : mo_scalar_profile(2.0 + z0q, obu, z0q);
  return p;
}

} // namespace elmk
