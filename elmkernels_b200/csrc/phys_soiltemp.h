// phys_soiltemp.h - soil/snow temperature group (a8): thermal properties, surface heat flux and its
// derivative, Crank-Nicolson heat diffusion through snow + standing surface water + soil as a
// 21-row pentadiagonal system, melt/freeze with supercooled soil water, new ground temperature.
//
// Parity target: kokkos_soil_temperature, reference
// driver/kokkos/soil_temperature_kokkos.cc:6-278, i.e. its nine launches fused into one pass:
//   soil_thermal::calc_soil_tk :22, calc_snow_tk :93, calc_face_tk :129, calc_soil_heat_capacity :160,
//   calc_snow_heat_capacity :202, calc_h2osfc_tk :234, calc_h2osfc_heat_capacity :249,
//   calc_h2osfc_height :263                        (src/physics/soil_thermal_properties_impl.hh)
//   soil_temp::calc_surface_heat_flux :16, calc_dhsdT :31, check_absorbed_solar :37,
//   calc_diffusive_heat_flux :49, calc_heat_flux_matrix_factor :94, update_temperature :154,
//   update_t_grnd :180                             (src/physics/soil_temperature_impl.hh)
//   set_RHS :31 (+ :79,114,138,181)                (src/physics/soil_temp_rhs_impl.hh)
//   set_LHS :105 (+ the block builders :160-481)   (src/physics/soil_temp_lhs_impl.hh)
//   solver::PDMA :16-76                            (src/physics/pentadiagonal_solver_impl.hh)
//   phase_change_h2osfc :13, phase_change_soisno :186   (src/physics/phase_change_impl.hh)
// One thread per column: the recurrences are 21 sequential steps, a warp-cooperative solve would idle
// most lanes.  The ~420 scratch doubles per column that the reference moves through 29 Views between
// its launches never leave the thread.
//
// The system is pentadiagonal only through two entries: the bottom snow layer (row 4) and the top soil
// layer (row 6) are coupled across the surface-water row 5.  Band k of row i multiplies unknown
// i + 2 - k (band 2 = diagonal), as in soil_temp_lhs_impl.hh:11-15.
#pragma once
#include "elmk_state.h"

namespace elmk {

namespace st {
constexpr double TKICE = 2.290, TKWAT = 0.57, TKAIR = 0.023, THIN_SFCLAYER = 1.0e-6;
constexpr double CNFAC = 0.5, CAPR = 0.34;
} // namespace st

ELMK_HD void column_soil_temperature(const Cols& S, const Tables&, const double dtime, const int c)
{
  using namespace st;
  const int snl = C1(snl);
  const int top = NLEVSNO - snl;
  const double fsno = C1(frac_sno), fse = C1(frac_sno_eff), fsfc = C1(frac_h2osfc);
  double h2osfc = C1(h2osfc), h2osno = C1(h2osno);
  double t_sfc = C1(t_h2osfc);

  double t[NLEVTOT], liq[NLEVTOT], ice[NLEVTOT], dz[NLEVTOT], z[NLEVTOT], zi[NLEVTOT + 1];
#pragma unroll
  for (int i = 0; i < NLEVTOT; ++i) {
    t[i] = C2(t_soisno, i); liq[i] = C2(h2osoi_liq, i); ice[i] = C2(h2osoi_ice, i);
    dz[i] = C2(dz, i); z[i] = C2(zsoi, i); zi[i] = C2(zisoi, i);
  }
  zi[NLEVTOT] = C2(zisoi, NLEVTOT);
  double watsat[NLEVGRND];
#pragma unroll
  for (int i = 0; i < NLEVGRND; ++i) watsat[i] = C2(watsat, i);

  // ---- thermal conductivity of the layers (Johansen) and at the interfaces, heat capacities ----
  double thk[NLEVTOT], tk[NLEVTOT], cv[NLEVTOT];
#pragma unroll
  for (int i = NLEVSNO; i < NLEVTOT; ++i) {
    const int k = i - NLEVSNO;
    double satw = (liq[i] / DENH2O + ice[i] / DENICE) / (dz[i] * watsat[k]);
    satw = dmin(1.0, satw);
    const double tkdry = C2(tkdry, k);
    if (satw > 1.0e-6) {
      const double dke = (t[i] >= TFRZ) ? dmax(0.0, m_log10(satw) + 1.0) : satw;
      const double fl = (liq[i] / (DENH2O * dz[i])) / (liq[i] / (DENH2O * dz[i]) + ice[i] / (DENICE * dz[i]));
      const double dksat = C2(tkmg, k) * pow_cbase(TKWAT, ELMK_LN_TKWAT, fl * watsat[k]) * pow_cbase(TKICE, ELMK_LN_TKICE, (1.0 - fl) * watsat[k]);
      thk[i] = dke * dksat + (1.0 - dke) * tkdry;
    } else {
      thk[i] = tkdry;
    }
    // (no layer lies below nlevbed == nlevgrnd, so the bedrock override never applies)
    cv[i] = C2(csol, i) * (1.0 - watsat[k]) * dz[i] + (ice[i] * CPICE + liq[i] * CPWAT);
    if (i == NLEVSNO && snl == 0 && h2osno > 0.0) cv[i] += CPICE * h2osno;
  }
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i) {
    if (i < top) {
      thk[i] = 0.0;
      cv[i] = 0.0;
    } else {
      const double bw = (ice[i] + liq[i]) / (fsno * dz[i]);
      thk[i] = TKAIR + (7.75e-5 * bw + 1.105e-6 * bw * bw) * (TKICE - TKAIR);
      cv[i] = (fsno > 0.0) ? dmax(THIN_SFCLAYER, (CPWAT * liq[i] + CPICE * ice[i]) / fsno) : THIN_SFCLAYER;
    }
  }
#pragma unroll
  for (int i = 0; i < NLEVTOT - 1; ++i) {
    tk[i] = (i < top) ? 0.0
                      : thk[i] * thk[i + 1] * (z[i + 1] - z[i]) /
                            (thk[i] * (z[i + 1] - zi[i + 1]) + thk[i + 1] * (zi[i + 1] - z[i]));
  }
  tk[NLEVTOT - 1] = 0.0;
  const double zh2osfc = 1.0e-3 * (0.5 * h2osfc);
  const double tk_sfc = TKWAT * thk[NLEVSNO] * (z[NLEVSNO] + zh2osfc) / (TKWAT * z[NLEVSNO] + thk[NLEVSNO] * zh2osfc);
  const bool ponded = (h2osfc > THIN_SFCLAYER) && (fsfc > THIN_SFCLAYER);
  const double c_sfc = ponded ? dmax(THIN_SFCLAYER, CPWAT * h2osfc / fsfc) : THIN_SFCLAYER;
  const double dz_sfc = ponded ? dmax(THIN_SFCLAYER, 1.0e-3 * h2osfc / fsfc) : THIN_SFCLAYER;

  // ---- surface heat fluxes and their temperature derivative ----
  const int veg = C1(frac_veg_nosno);
  const double dlrad = C1(dlrad), emg = C1(emg), lwrad = C1(forc_lwrad), htvp = C1(htvp);
  const double sabg_soil = C1(sabg_soil), sabg_snow = C1(sabg_snow);
  double sabg_lyr[NLEVSNO + 1];
#pragma unroll
  for (int i = 0; i <= NLEVSNO; ++i) sabg_lyr[i] = C2(sabg_lyr, i);
  C1(sabg_chk) = fse * sabg_snow + (1.0 - fse) * sabg_soil;
  const double lw_in = (1.0 - veg) * emg * lwrad;
  const double hs_soil = sabg_soil + dlrad + lw_in - emg * STEBOL * pow4(t[NLEVSNO]) -
                         (C1(eflx_sh_soil) + C1(qflx_ev_soil) * htvp);
  const double hs_sfc = sabg_soil + dlrad + lw_in - emg * STEBOL * pow4(t_sfc) -
                        (C1(eflx_sh_h2osfc) + C1(qflx_ev_h2osfc) * htvp);
  double t_top = t[NLEVSNO], sabg_top = sabg_lyr[NLEVSNO];
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i)
    if (i == top) { t_top = t[i]; sabg_top = sabg_lyr[i]; }
  const double hs_top_snow = sabg_top + dlrad + lw_in - emg * STEBOL * pow4(t_top) -
                             (C1(eflx_sh_snow) + C1(qflx_ev_snow) * htvp);
  const double dhsdT = -C1(cgrnd) - 4.0 * emg * STEBOL * cube(C1(t_grnd));

  // ---- diffusive heat flux at the interfaces and the time-step factor of each layer ----
  double fn[NLEVTOT], fact[NLEVTOT];
#pragma unroll
  for (int i = 0; i < NLEVTOT - 1; ++i) fn[i] = (i < top) ? 0.0 : tk[i] * (t[i + 1] - t[i]) / (z[i + 1] - z[i]);
  fn[NLEVTOT - 1] = 0.0;
#pragma unroll
  for (int i = 0; i < NLEVTOT; ++i) {
    if (i < top) fact[i] = 0.0;
    else if (i == top) fact[i] = dtime / cv[i] * dz[i] / (0.5 * (z[i] - zi[i] + CAPR * (z[i + 1] - zi[i])));
    else fact[i] = dtime / cv[i];
    C2(fact, i) = fact[i];
  }

  // ---- right-hand side: rows 0-4 snow, row 5 surface water, rows 6-20 soil ----
  double rhs[NROWS];
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i) {
    if (i < top) rhs[i] = 0.0;
    else if (i == top) rhs[i] = t[i] + fact[i] * (hs_top_snow - dhsdT * t[i] + CNFAC * fn[i]);
    else rhs[i] = t[i] + CNFAC * fact[i] * (fn[i] - fn[i - 1]) + fact[i] * sabg_lyr[i];
  }
  const double fn_sfc = tk_sfc * (t[NLEVSNO] - t_sfc) / (0.5 * dz_sfc + z[NLEVSNO]);
  rhs[NLEVSNO] = t_sfc + (dtime / c_sfc) * (hs_sfc - dhsdT * t_sfc + CNFAC * fn_sfc);
  {
    const int s = NLEVSNO;
    if (snl == 0) {
      rhs[s + 1] = t[s] + fact[s] * (hs_top_snow - dhsdT * t[s] + CNFAC * fn[s]);
    } else {
      double r = t[s] + fact[s] * ((1.0 - fse) * (hs_soil - dhsdT * t[s]) + CNFAC * (fn[s] - fse * fn[s - 1]));
      r += fse * fact[s] * sabg_lyr[s];
      rhs[s + 1] = r;
    }
  }
#pragma unroll
  for (int j = NLEVSNO + 1; j < NLEVTOT - 1; ++j) rhs[j + 1] = t[j] + CNFAC * fact[j] * (fn[j] - fn[j - 1]);
  {
    const int b = NLEVTOT - 1;
    rhs[b + 1] = t[b] - CNFAC * fact[b] * fn[b - 1] + fact[b] * fn[b];
  }

  // ---- left-hand side: five bands per row ----
  double b0[NROWS], b1[NROWS], b2[NROWS], b3[NROWS], b4[NROWS];
#pragma unroll
  for (int i = 0; i < NROWS; ++i) { b0[i] = 0.0; b1[i] = 0.0; b2[i] = 0.0; b3[i] = 0.0; b4[i] = 0.0; }
  constexpr double OMC = 1.0 - CNFAC;
  if (snl > 0) {
    // snow rows
#pragma unroll
    for (int i = 0; i < NLEVSNO; ++i) {
      if (i == top) {
        const double dzp = z[i + 1] - z[i];
        b2[i] = 1.0 + OMC * fact[i] * tk[i] / dzp - fact[i] * dhsdT;
        if (snl > 1) b1[i] = -OMC * fact[i] * tk[i] / dzp;
      } else if (i > top) {
        const double dzm = z[i] - z[i - 1];
        const double dzp = z[i + 1] - z[i];
        b3[i] = -OMC * fact[i] * tk[i - 1] / dzm;
        b2[i] = 1.0 + OMC * fact[i] * (tk[i] / dzp + tk[i - 1] / dzm);
        if (i != NLEVSNO - 1) b1[i] = -OMC * fact[i] * tk[i] / dzp;
      }
    }
    // bottom snow layer -> top soil layer, across the surface-water row
    b0[NLEVSNO - 1] = -OMC * fact[NLEVSNO - 1] * tk[NLEVSNO - 1] / (z[NLEVSNO] - z[NLEVSNO - 1]);
  }
  // surface-water row
  b2[NLEVSNO] = 1.0 + OMC * (dtime / c_sfc) * tk_sfc / (0.5 * dz_sfc + z[NLEVSNO]) - (dtime / c_sfc) * dhsdT;
  b1[NLEVSNO] = -OMC * (dtime / c_sfc) * tk_sfc / (0.5 * dz_sfc + z[NLEVSNO]);
  // top soil row
  {
    const int s = NLEVSNO, r = NLEVSNO + 1;
    const double dzp = z[s + 1] - z[s];
    if (snl == 0) {
      b2[r] = 1.0 + OMC * fact[s] * tk[s] / dzp - fact[s] * dhsdT;
      b1[r] = -OMC * fact[s] * tk[s] / dzp;
    } else {
      const double dzm = z[s] - z[s - 1];
      b2[r] = 1.0 + OMC * fact[s] * (tk[s] / dzp + fse * tk[s - 1] / dzm) - (1.0 - fse) * fact[s] * dhsdT;
      b1[r] = -OMC * fact[s] * tk[s] / dzp;
      b4[r] = -fse * OMC * fact[s] * tk[s - 1] / dzm;
    }
    if (fsfc != 0.0) {
      const double dzm = 0.5 * dz_sfc + z[s];
      b2[r] += fsfc * (OMC * fact[s] * tk_sfc / dzm + fact[s] * dhsdT);
      b3[r] = -fsfc * OMC * fact[s] * tk_sfc / (0.5 * dz_sfc + z[s]);
    }
  }
  // interior soil rows and the bottom row
#pragma unroll
  for (int j = NLEVSNO + 1; j < NLEVTOT - 1; ++j) {
    const double dzm = z[j] - z[j - 1];
    const double dzp = z[j + 1] - z[j];
    b3[j + 1] = -OMC * fact[j] * tk[j - 1] / dzm;
    b2[j + 1] = 1.0 + OMC * fact[j] * (tk[j] / dzp + tk[j - 1] / dzm);
    b1[j + 1] = -OMC * fact[j] * tk[j] / dzp;
  }
  {
    const int b = NLEVTOT - 1;
    const double dzm = z[b] - z[b - 1];
    b3[b + 1] = -OMC * fact[b] * tk[b - 1] / dzm;
    b2[b + 1] = 1.0 + OMC * fact[b] * tk[b - 1] / dzm;
  }

  // ---- pentadiagonal solve (PDMA); A, B, Z are zero above the first active row ----
  {
    constexpr int N = NROWS;
    double A[N], B[N], Z[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { A[i] = 0.0; B[i] = 0.0; Z[i] = 0.0; }
    double U1 = 1.0 / b2[top];
    A[top] = b1[top] * U1;
    B[top] = b0[top] * U1;
    Z[top] = rhs[top] * U1;
    double Y1 = b3[top + 1];
    U1 = 1.0 / (b2[top + 1] - A[top] * Y1);
    A[top + 1] = (b1[top + 1] - B[top] * Y1) * U1;
    B[top + 1] = b0[top + 1] * U1;
    Z[top + 1] = (rhs[top + 1] - Z[top] * Y1) * U1;
    for (int i = top + 2; i < N - 2; ++i) {
      Y1 = b3[i] - A[i - 2] * b4[i];
      U1 = 1.0 / (b2[i] - B[i - 2] * b4[i] - A[i - 1] * Y1);
      A[i] = (b1[i] - B[i - 1] * Y1) * U1;
      B[i] = b0[i] * U1;
      Z[i] = (rhs[i] - Z[i - 2] * b4[i] - Z[i - 1] * Y1) * U1;
    }
    Y1 = b3[N - 2] - A[N - 4] * b4[N - 2];
    U1 = 1.0 / (b2[N - 2] - B[N - 4] * b4[N - 2] - A[N - 3] * Y1);
    A[N - 2] = (b1[N - 2] - B[N - 3] * Y1) * U1;
    const double Y2 = b3[N - 1] - A[N - 3] * b4[N - 1];
    const double U2 = 1.0 / (b2[N - 1] - B[N - 3] * b4[N - 1] - A[N - 2] * Y2);
    // the reference uses Z(N-3) and Z(N-2) twice here; band 4 is zero in these rows
    Z[N - 2] = (rhs[N - 2] - Z[N - 3] * b4[N - 2] - Z[N - 3] * Y1) * U1;
    Z[N - 1] = (rhs[N - 1] - Z[N - 2] * b4[N - 1] - Z[N - 2] * Y2) * U2;
    rhs[N - 1] = Z[N - 1];
    rhs[N - 2] = Z[N - 2] - A[N - 2] * rhs[N - 1];
    for (int i = N - 3; i >= 0; --i) rhs[i] = Z[i] - A[i] * rhs[i + 1] - B[i] * rhs[i + 2];
  }

  // ---- new temperatures ----
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i)
    if (i >= top) t[i] = rhs[i];
#pragma unroll
  for (int i = NLEVSNO; i < NLEVTOT; ++i) t[i] = rhs[i + 1];
  t_sfc = (fsfc != 0.0) ? rhs[NLEVSNO] : t[NLEVSNO];

  // ---- phase change of standing surface water ----
  double int_snow = C1(int_snow), snow_depth = C1(snow_depth);
  double xmf_sfc = 0.0, q_sfc_ice = 0.0, e_sfc_snow = 0.0;
  {
    const int b = NLEVSNO - 1;   // bottom snow slot
    if (fsfc > 0.0 && t_sfc <= TFRZ) {
      const double tinc = TFRZ - t_sfc;
      t_sfc = TFRZ;
      const double hm = fsfc * (dhsdT * tinc - tinc * c_sfc / dtime);
      const double xm = hm * dtime / HFUS;
      const double temp1 = h2osfc + xm;
      const double z_avg = fsno * snow_depth;
      double rho_avg = (z_avg > 0.0) ? dmin(800.0, h2osno / z_avg) : 200.0;
      if (temp1 >= 0.0) {
        // part of the pond freezes onto the snow pack
        h2osno -= xm;
        int_snow -= xm;
        if (snl > 0) ice[b] -= xm;
        h2osfc += xm;
        xmf_sfc = hm;
        q_sfc_ice = -xm / dtime;
        snow_depth = (fsno > 0 && snl > 0) ? h2osno / (rho_avg * fsno) : h2osno / DENICE;
        if (snl == 0) {
          t[b] = t_sfc;
          e_sfc_snow = 0.0;
        } else {
          const double c1 = (snl == 1) ? fsno * (dtime / fact[b] - dhsdT * dtime) : fsno / fact[b] * dtime;
          const double c2 = (fsfc != 0.0) ? (-CPWAT * xm - fsfc * dhsdT * dtime) : 0.0;
          t[b] = (c1 * t[b] + c2 * t_sfc) / (c1 + c2);
          e_sfc_snow = (t_sfc - t[b]) * c2 / dtime;
        }
      } else {
        // the whole pond freezes
        rho_avg = (h2osno * rho_avg + h2osfc * DENICE) / (h2osno + h2osfc);
        h2osno += h2osfc;
        int_snow += h2osfc;
        q_sfc_ice = h2osfc / dtime;
        if (snl > 0) ice[b] = ice[b] + h2osfc;
        t_sfc = t_sfc - temp1 * HFUS / (dtime * dhsdT - c_sfc);
        xmf_sfc = hm - fsfc * temp1 * HFUS / dtime;
        if (snl == 0) {
          t[b] = t_sfc;
        } else {
          const double c1 = (snl == 1) ? fsno * (dtime / fact[b] - dhsdT * dtime) : fsno / fact[b] * dtime;
          const double c2 = (fsfc != 0.0) ? fsfc * (c_sfc - dtime * dhsdT) : 0.0;
          t[b] = (c1 * t[b] + c2 * t_sfc) / (c1 + c2);
          t_sfc = t[b];
        }
        h2osfc = 0.0;
        snow_depth = (fsno > 0.0 && snl > 0) ? h2osno / (rho_avg * fsno) : h2osno / DENICE;
      }
    }
  }
  C1(xmf_h2osfc) = xmf_sfc;
  C1(qflx_h2osfc_ice) = q_sfc_ice;
  C1(eflx_h2osfc_snow) = e_sfc_snow;

  // ---- phase change in snow and soil layers ----
  double xmf = 0.0, q_snomelt = 0.0, q_snow_melt = 0.0;
  double snofrz_lyr[NLEVSNO];
  int imelt[NLEVTOT];
  double tinc[NLEVTOT], supercool[NLEVGRND];
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i) snofrz_lyr[i] = 0.0;
#pragma unroll
  for (int i = 0; i < NLEVTOT; ++i) {
    imelt[i] = (i >= top) ? 0 : C2(imelt, i);   // rows above the snow pack keep their stale flags (quirk 13)
    tinc[i] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i) {
    if (i >= top) {
      if (ice[i] > 0.0 && t[i] > TFRZ) { imelt[i] = 1; tinc[i] = TFRZ - t[i]; t[i] = TFRZ; }
      if (liq[i] > 0.0 && t[i] < TFRZ) { imelt[i] = 2; tinc[i] = TFRZ - t[i]; t[i] = TFRZ; }
    }
  }
#pragma unroll
  for (int i = NLEVSNO; i < NLEVTOT; ++i) {
    const int k = i - NLEVSNO;
    if (ice[i] > 0.0 && t[i] > TFRZ) { imelt[i] = 1; tinc[i] = TFRZ - t[i]; t[i] = TFRZ; }
    supercool[k] = 0.0;
    if (t[i] < TFRZ) {
      const double smp = HFUS * (TFRZ - t[i]) / (GRAV * t[i]) * 1000.0;
      supercool[k] = watsat[k] * m_pow(smp / C2(sucsat, k), -1.0 / C2(bsw, k));
      supercool[k] *= dz[i] * 1000.0;
    }
    if (liq[i] > supercool[k] && t[i] < TFRZ) { imelt[i] = 2; tinc[i] = TFRZ - t[i]; t[i] = TFRZ; }
    if (snl == 0 && h2osno > 0.0 && i == NLEVSNO) {
      if (t[i] > TFRZ) { imelt[i] = 1; tinc[i] = TFRZ - t[i]; t[i] = TFRZ; }
    }
  }
#pragma unroll
  for (int i = 0; i < NLEVTOT; ++i) {
    if (i < top) continue;
    double hm = 0.0;
    if (imelt[i] > 0) {
      if (i == top) {
        if (i < NLEVSNO) {
          hm = fse * (dhsdT * tinc[i] - tinc[i] / fact[i]);
        } else {
          const double temp_hm = dhsdT * tinc[i] - tinc[i] / fact[i];
          hm = (fsfc != 0.0) ? temp_hm - fsfc * (dhsdT * tinc[i]) : temp_hm;
        }
      } else if (i == NLEVSNO) {
        hm = (1.0 - fse - fsfc) * dhsdT * tinc[i] - tinc[i] / fact[i];
      } else {
        hm = (i < NLEVSNO) ? -fse * (tinc[i] / fact[i]) : -tinc[i] / fact[i];
      }
    }
    if (imelt[i] == 1 && hm < 0.0) { hm = 0.0; imelt[i] = 0; }
    if (imelt[i] == 2 && hm > 0.0) { hm = 0.0; imelt[i] = 0; }
    if (imelt[i] > 0 && fabs(hm) > 0.0) {
      double xm = hm * dtime / HFUS;
      if (i == NLEVSNO) {
        if (snl == 0 && h2osno > 0.0 && xm > 0.0) {
          // thin snow without a layer melts before the soil ice does
          const double temp1 = h2osno;
          h2osno = dmax(0.0, temp1 - xm);
          const double propor = h2osno / temp1;
          snow_depth *= propor;
          const double heatr = hm - HFUS * (temp1 - h2osno) / dtime;
          if (heatr > 0.0) {
            xm = heatr * dtime / HFUS;
            hm = heatr;
          } else {
            xm = 0.0;
            hm = 0.0;
          }
          q_snomelt = dmax(0.0, temp1 - h2osno) / dtime;
          xmf = HFUS * q_snomelt;
          q_snow_melt = q_snomelt;
        }
      }
      double heatr = 0.0;
      const double wmass0 = ice[i] + liq[i];
      const double wice0 = ice[i];
      if (xm > 0.0) {
        ice[i] = dmax(0.0, wice0 - xm);
        heatr = hm - HFUS * (wice0 - ice[i]) / dtime;
      } else if (xm < 0.0) {
        if (i < NLEVSNO) {
          ice[i] = dmin(wmass0, wice0 - xm);
        } else {
          const double sc = supercool[(i >= NLEVSNO) ? i - NLEVSNO : 0];
          ice[i] = (wmass0 < sc) ? 0.0 : dmin(wmass0 - sc, wice0 - xm);
        }
        heatr = hm - HFUS * (wice0 - ice[i]) / dtime;
      }
      liq[i] = dmax(0.0, wmass0 - ice[i]);
      if (fabs(heatr) > 0.0) {
        if (i == top) {
          if (snl == 0) t[i] += fact[i] * heatr / (1.0 - (1.0 - fsfc) * fact[i] * dhsdT);
          else t[i] += (fact[i] / fse) * heatr / (1.0 - fact[i] * dhsdT);
        } else if (i == NLEVSNO) {
          t[i] += fact[i] * heatr / (1.0 - (1.0 - fse - fsfc) * fact[i] * dhsdT);
        } else {
          if (i >= NLEVSNO) t[i] += fact[i] * heatr;
          else if (fse > 0.0) t[i] += (fact[i] / fse) * heatr;
        }
        if (i < NLEVSNO) {
          if (liq[i] * ice[i] > 0.0) t[i] = TFRZ;
        }
      }
      xmf += HFUS * (wice0 - ice[i]) / dtime;
      if (imelt[i] == 1 && i < NLEVSNO) q_snomelt += dmax(0.0, (wice0 - ice[i])) / dtime;
      if (imelt[i] == 2 && i < NLEVSNO) snofrz_lyr[i] = dmax(0.0, (ice[i] - wice0)) / dtime;
    }
  }
  double q_snofrz = 0.0;
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i) {
    if (imelt[i] == 2) q_snofrz += snofrz_lyr[i];
    C2(qflx_snofrz_lyr, i) = snofrz_lyr[i];
  }
  C1(xmf) = xmf;
  C1(qflx_snofrz) = q_snofrz;
  C1(qflx_snow_melt) = q_snow_melt;
  C1(qflx_snomelt) = q_snomelt;
  C1(eflx_snomelt) = q_snomelt * HFUS;

  // ---- new ground temperature ----
  double tg;
  t_top = t[NLEVSNO];
#pragma unroll
  for (int i = 0; i < NLEVSNO; ++i)
    if (i == top) t_top = t[i];
  if (snl > 0) {
    tg = (fsfc != 0.0) ? fse * t_top + (1.0 - fse - fsfc) * t[NLEVSNO] + fsfc * t_sfc
                       : fse * t_top + (1.0 - fse) * t[NLEVSNO];
  } else {
    tg = (fsfc != 0.0) ? (1.0 - fsfc) * t[NLEVSNO] + fsfc * t_sfc : t[NLEVSNO];
  }
  C1(t_grnd) = tg;

  // ---- write back the prognostic column ----
  // Rows above the snow pack are untouched by the reference, except slot 4 which the surface-water
  // phase change may initialise when there is no snow layer (phase_change_impl.hh:79-83,123-125).
#pragma unroll
  for (int i = 0; i < NLEVTOT; ++i) {
    if (i >= top || i == NLEVSNO - 1) {
      C2(t_soisno, i) = t[i];
      C2(h2osoi_ice, i) = ice[i];
    }
    if (i >= top) {
      C2(h2osoi_liq, i) = liq[i];
      C2(imelt, i) = imelt[i];
    }
  }
  C1(t_h2osfc) = t_sfc;
  C1(h2osfc) = h2osfc;
  C1(h2osno) = h2osno;
  C1(int_snow) = int_snow;
  C1(snow_depth) = snow_depth;
}

} // namespace elmk
