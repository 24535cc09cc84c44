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

// On the device the layer / row loops stay loops (dynamic layer index, per-thread arrays in local memory) instead of
// being unrolled 20x into ~350 KB of straight-line SASS that no instruction cache holds: 63 KB, 4.4 -> 3.85 ms per 2M
// columns on B200 (-DELMK_SOIL_UNROLLED restores the unrolled form for A/B runs).
#if defined(__CUDA_ARCH__) && !defined(ELMK_SOIL_UNROLLED)
#define ELMK_SOIL_LOOP _Pragma("unroll 1")
#else
#define ELMK_SOIL_LOOP _Pragma("unroll")
#endif
template <bool REALIGN = false>
ELMK_HD void column_soil_temperature(const Cols& S, const Tables&, const double dtime, const int c)
{
  using namespace st;
  const int snl = C1(snl);
  const int top = NLEVSNO - snl;
  const double fsno = C1(frac_sno), fse = C1(frac_sno_eff), fsfc = C1(frac_h2osfc);
  double h2osfc = C1(h2osfc), h2osno = C1(h2osno);
  double t_sfc = C1(t_h2osfc);

  // The function streams over the layers three times - properties, matrix rows + forward elimination,
  // phase change - and re-reads the few per-layer inputs each pass needs instead of holding the ~340 doubles
  // of the column (state, conductivities, five bands, right-hand side, solver work arrays) at once: the first
  // version of this kernel ran at 255 registers with 2.9 KB of spills per thread.
  // (the old temperatures, the node depths, the absorbed radiation of the snow layers and - once computed - fact are
  //  read from the state where they are needed: per-thread copies of them were half of a 1.5 KB frame, and at 28 resident
  //  warps per SM the frames of the launch exceed L2; the rows are L1 / L2 hits for the block that has just read them)
#define t_old(i) C2(t_soisno, i)
#define z_node(i) C2(zsoi, i)

  // ---- pass 1: thermal conductivity (Johansen) and heat capacity of every layer -> fact; the conductivity at the
  //      interface above and the diffusive heat flux through it as soon as both neighbours are known ----
  double tk[NLEVTOT], fn[NLEVTOT];
  double thk_above = 0.0, thk_soil1 = 0.0;
ELMK_SOIL_LOOP
  for (int i = 0; i < NLEVTOT; ++i) {
    if (i % 5 == 0 && i > 0) ELMK_REALIGN(REALIGN);
    const double liq = C2(h2osoi_liq, i), ice = C2(h2osoi_ice, i), dz = C2(dz, i);
    double cv, thk_i;
    if (i >= NLEVSNO) {
      const int k = i - NLEVSNO;
      const double watsat = C2(watsat, k);
      const Div2 hold = m_div2(liq, DENH2O * dz, ice, DENICE * dz);   // (explicit uses keep m_div / m_div2 in the module)
      double satw = m_div(liq / DENH2O + ice / DENICE, dz * watsat);
      satw = dmin(1.0, satw);
      const double tkdry = C2(tkdry, k);
      if (satw > 1.0e-6) {
        const double dke = (t_old(i) >= TFRZ) ? dmax(0.0, m_log10(satw) + 1.0) : satw;
        const double fl = hold.q0 / (hold.q0 + hold.q1);
        const double dksat = C2(tkmg, k) * pow_cbase(TKWAT, ELMK_LN_TKWAT, fl * watsat) * pow_cbase(TKICE, ELMK_LN_TKICE, (1.0 - fl) * watsat);
        thk_i = dke * dksat + (1.0 - dke) * tkdry;
      } else {
        thk_i = tkdry;
      }
      // (no layer lies below nlevbed == nlevgrnd, so the bedrock override never applies)
      cv = C2(csol, i) * (1.0 - watsat) * dz + (ice * CPICE + liq * CPWAT);
      if (i == NLEVSNO && snl == 0 && h2osno > 0.0) cv += CPICE * h2osno;
    } else if (i < top) {
      thk_i = 0.0;
      cv = 0.0;
    } else {
      const double bw = (ice + liq) / (fsno * dz);
      thk_i = TKAIR + (7.75e-5 * bw + 1.105e-6 * bw * bw) * (TKICE - TKAIR);
      cv = (fsno > 0.0) ? dmax(THIN_SFCLAYER, (CPWAT * liq + CPICE * ice) / fsno) : THIN_SFCLAYER;
    }
    const double zi = z_node(i);
    double fact_i;
    if (i < top) fact_i = 0.0;
    else if (i == top) fact_i = dtime / cv * dz / (0.5 * (zi - C2(zisoi, i) + CAPR * (z_node((i + 1 < NLEVTOT) ? i + 1 : i) - C2(zisoi, i))));
    else fact_i = dtime / cv;
    C2(fact, i) = fact_i;
    if (i == NLEVSNO) thk_soil1 = thk_i;
    if (i > 0) {
      // interface between layers i - 1 and i
      const int j = i - 1;
      if (j < top) {
        tk[j] = 0.0;
        fn[j] = 0.0;
      } else {
        const double zj = z_node(j), zi1 = C2(zisoi, i);
        tk[j] = thk_above * thk_i * (zi - zj) / (thk_above * (zi - zi1) + thk_i * (zi1 - zj));
        fn[j] = tk[j] * (t_old(i) - t_old(j)) / (zi - zj);
      }
    }
    thk_above = thk_i;
  }
  ELMK_REALIGN(REALIGN);
  tk[NLEVTOT - 1] = 0.0;
  fn[NLEVTOT - 1] = 0.0;
  const double zh2osfc = 1.0e-3 * (0.5 * h2osfc);
  const double z_soil1 = z_node(NLEVSNO);
  const double tk_sfc = TKWAT * thk_soil1 * (z_soil1 + zh2osfc) / (TKWAT * z_soil1 + thk_soil1 * zh2osfc);
  const bool ponded = (h2osfc > THIN_SFCLAYER) && (fsfc > THIN_SFCLAYER);
  const double c_sfc = ponded ? dmax(THIN_SFCLAYER, CPWAT * h2osfc / fsfc) : THIN_SFCLAYER;
  const double dz_sfc = ponded ? dmax(THIN_SFCLAYER, 1.0e-3 * h2osfc / fsfc) : THIN_SFCLAYER;

  // ---- surface heat fluxes and their temperature derivative ----
  const int veg = C1(frac_veg_nosno);
  const double dlrad = C1(dlrad), emg = C1(emg), lwrad = C1(forc_lwrad), htvp = C1(htvp);
  const double sabg_soil = C1(sabg_soil), sabg_snow = C1(sabg_snow);
  C1(sabg_chk) = fse * sabg_snow + (1.0 - fse) * sabg_soil;
  const double lw_in = (1.0 - veg) * emg * lwrad;
  const double hs_soil = sabg_soil + dlrad + lw_in - emg * STEBOL * pow4(t_old(NLEVSNO)) -
                         (C1(eflx_sh_soil) + C1(qflx_ev_soil) * htvp);
  const double hs_sfc = sabg_soil + dlrad + lw_in - emg * STEBOL * pow4(t_sfc) -
                        (C1(eflx_sh_h2osfc) + C1(qflx_ev_h2osfc) * htvp);
  const double t_top = t_old(top), sabg_top = C2(sabg_lyr, top);   // (top == NLEVSNO without snow layers)
  const double hs_top_snow = sabg_top + dlrad + lw_in - emg * STEBOL * pow4(t_top) -
                             (C1(eflx_sh_snow) + C1(qflx_ev_snow) * htvp);
  const double dhsdT = -C1(cgrnd) - 4.0 * emg * STEBOL * cube(C1(t_grnd));

  ELMK_REALIGN(REALIGN);
  // ---- pass 2: one row of the band system at a time (rows 0-4 snow, row 5 surface water, rows 6-20 soil;
  //      band k of row i multiplies unknown i + 2 - k), eliminated as soon as it is built (PDMA forward
  //      sweep).  Rows above the first active one leave A = B = Z = 0; with those zeros the general
  //      elimination step reproduces the reference's special first and second steps bit for bit. ----
  constexpr double OMC = 1.0 - CNFAC;
  constexpr int N = NROWS;
  double A[N], B[N], Z[N];
  const double fn_sfc = tk_sfc * (t_old(NLEVSNO) - t_sfc) / (0.5 * dz_sfc + z_node(NLEVSNO));
ELMK_SOIL_LOOP
  for (int r = 0; r < N; ++r) {
    if (r % 5 == 0 && r > 0) ELMK_REALIGN(REALIGN);
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0, b4 = 0.0, rhs = 0.0;
    bool active = true;
    if (r < NLEVSNO) {
      const int i = r;
      active = (i >= top);
      if (i == top) {
        const double dzp = z_node(i + 1) - z_node(i);
        b2 = 1.0 + OMC * C2(fact, i) * tk[i] / dzp - C2(fact, i) * dhsdT;
        if (snl > 1) b1 = -OMC * C2(fact, i) * tk[i] / dzp;
        rhs = t_old(i) + C2(fact, i) * (hs_top_snow - dhsdT * t_old(i) + CNFAC * fn[i]);
      } else if (i > top) {
        const int im = (i > 0) ? i - 1 : 0;
        const double dzm = z_node(i) - z_node(im);
        const double dzp = z_node(i + 1) - z_node(i);
        b3 = -OMC * C2(fact, i) * tk[im] / dzm;
        b2 = 1.0 + OMC * C2(fact, i) * (tk[i] / dzp + tk[im] / dzm);
        if (i != NLEVSNO - 1) b1 = -OMC * C2(fact, i) * tk[i] / dzp;
        rhs = t_old(i) + CNFAC * C2(fact, i) * (fn[i] - fn[im]) + C2(fact, i) * C2(sabg_lyr, i);
      }
      // bottom snow layer -> top soil layer, across the surface-water row
      if (i == NLEVSNO - 1 && snl > 0) b0 = -OMC * C2(fact, i) * tk[i] / (z_node(NLEVSNO) - z_node(i));
    } else if (r == NLEVSNO) {
      b2 = 1.0 + OMC * (dtime / c_sfc) * tk_sfc / (0.5 * dz_sfc + z_node(NLEVSNO)) - (dtime / c_sfc) * dhsdT;
      b1 = -OMC * (dtime / c_sfc) * tk_sfc / (0.5 * dz_sfc + z_node(NLEVSNO));
      rhs = t_sfc + (dtime / c_sfc) * (hs_sfc - dhsdT * t_sfc + CNFAC * fn_sfc);
    } else if (r == NLEVSNO + 1) {
      const int s = NLEVSNO;
      const double dzp = z_node(s + 1) - z_node(s);
      if (snl == 0) {
        b2 = 1.0 + OMC * C2(fact, s) * tk[s] / dzp - C2(fact, s) * dhsdT;
        b1 = -OMC * C2(fact, s) * tk[s] / dzp;
        rhs = t_old(s) + C2(fact, s) * (hs_top_snow - dhsdT * t_old(s) + CNFAC * fn[s]);
      } else {
        const double dzm = z_node(s) - z_node(s - 1);
        b2 = 1.0 + OMC * C2(fact, s) * (tk[s] / dzp + fse * tk[s - 1] / dzm) - (1.0 - fse) * C2(fact, s) * dhsdT;
        b1 = -OMC * C2(fact, s) * tk[s] / dzp;
        b4 = -fse * OMC * C2(fact, s) * tk[s - 1] / dzm;
        double rr = t_old(s) + C2(fact, s) * ((1.0 - fse) * (hs_soil - dhsdT * t_old(s)) + CNFAC * (fn[s] - fse * fn[s - 1]));
        rr += fse * C2(fact, s) * C2(sabg_lyr, s);
        rhs = rr;
      }
      if (fsfc != 0.0) {
        const double dzm = 0.5 * dz_sfc + z_node(s);
        b2 += fsfc * (OMC * C2(fact, s) * tk_sfc / dzm + C2(fact, s) * dhsdT);
        b3 = -fsfc * OMC * C2(fact, s) * tk_sfc / (0.5 * dz_sfc + z_node(s));
      }
    } else if (r < N - 1) {
      const int j = r - 1;
      const double dzm = z_node(j) - z_node(j - 1);
      const double dzp = z_node(j + 1) - z_node(j);
      b3 = -OMC * C2(fact, j) * tk[j - 1] / dzm;
      b2 = 1.0 + OMC * C2(fact, j) * (tk[j] / dzp + tk[j - 1] / dzm);
      b1 = -OMC * C2(fact, j) * tk[j] / dzp;
      rhs = t_old(j) + CNFAC * C2(fact, j) * (fn[j] - fn[j - 1]);
    } else {
      const int b = NLEVTOT - 1;
      const double dzm = z_node(b) - z_node(b - 1);
      b3 = -OMC * C2(fact, b) * tk[b - 1] / dzm;
      b2 = 1.0 + OMC * C2(fact, b) * tk[b - 1] / dzm;
      rhs = t_old(b) - CNFAC * C2(fact, b) * fn[b - 1] + C2(fact, b) * fn[b];
    }

    // forward elimination of row r
    const double Am2 = (r >= 2) ? A[(r >= 2) ? r - 2 : 0] : 0.0, Bm2 = (r >= 2) ? B[(r >= 2) ? r - 2 : 0] : 0.0,
                 Zm2 = (r >= 2) ? Z[(r >= 2) ? r - 2 : 0] : 0.0;
    const double Am1 = (r >= 1) ? A[(r >= 1) ? r - 1 : 0] : 0.0, Bm1 = (r >= 1) ? B[(r >= 1) ? r - 1 : 0] : 0.0,
                 Zm1 = (r >= 1) ? Z[(r >= 1) ? r - 1 : 0] : 0.0;
    if (!active) {
      A[r] = 0.0; B[r] = 0.0; Z[r] = 0.0;
    } else if (r < N - 2) {
      const double Y1 = b3 - Am2 * b4;
      const double U1 = 1.0 / (b2 - Bm2 * b4 - Am1 * Y1);
      A[r] = (b1 - Bm1 * Y1) * U1;
      B[r] = b0 * U1;
      Z[r] = (rhs - Zm2 * b4 - Zm1 * Y1) * U1;
    } else if (r == N - 2) {
      const double Y1 = b3 - Am2 * b4;
      const double U1 = 1.0 / (b2 - Bm2 * b4 - Am1 * Y1);
      A[r] = (b1 - Bm1 * Y1) * U1;
      B[r] = 0.0;
      // the reference uses Z(N-3) twice here where Z(N-4), Z(N-3) are meant; band 4 is zero in this row
      Z[r] = (rhs - Zm1 * b4 - Zm1 * Y1) * U1;
    } else {
      // last row: Y2 / U2 of the reference, again with its doubled Z(N-2)
      const double Y2 = b3 - Am2 * b4;
      const double U2 = 1.0 / (b2 - Bm2 * b4 - Am1 * Y2);
      A[r] = 0.0;
      B[r] = 0.0;
      Z[r] = (rhs - Zm1 * b4 - Zm1 * Y2) * U2;
    }
  }
  ELMK_REALIGN(REALIGN);
  // back substitution; the solution overwrites Z
  double* const sol = Z;
  sol[N - 2] = Z[N - 2] - A[N - 2] * sol[N - 1];
ELMK_SOIL_LOOP
  for (int i = N - 3; i >= 0; --i) sol[i] = Z[i] - A[i] * sol[i + 1] - B[i] * sol[i + 2];

  // surface-water temperature: the solution of its row when there is surface water, else the top soil value
  t_sfc = (fsfc != 0.0) ? sol[NLEVSNO] : sol[NLEVSNO + 1];

  // ---- phase change of standing surface water (touches the bottom snow slot only) ----
  constexpr int SB = NLEVSNO - 1;   // bottom snow slot
  const double fact_sb = C2(fact, SB);
  double t_sb = (SB >= top) ? sol[SB] : C2(t_soisno, SB);
  double ice_sb = C2(h2osoi_ice, SB);
  double int_snow = C1(int_snow), snow_depth = C1(snow_depth);
  double xmf_sfc = 0.0, q_sfc_ice = 0.0, e_sfc_snow = 0.0;
  {
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
        if (snl > 0) ice_sb -= xm;
        h2osfc += xm;
        xmf_sfc = hm;
        q_sfc_ice = -xm / dtime;
        snow_depth = (fsno > 0 && snl > 0) ? h2osno / (rho_avg * fsno) : h2osno / DENICE;
        if (snl == 0) {
          t_sb = t_sfc;
          e_sfc_snow = 0.0;
        } else {
          const double c1 = (snl == 1) ? fsno * (dtime / fact_sb - dhsdT * dtime) : fsno / fact_sb * dtime;
          const double c2 = (fsfc != 0.0) ? (-CPWAT * xm - fsfc * dhsdT * dtime) : 0.0;
          t_sb = (c1 * t_sb + c2 * t_sfc) / (c1 + c2);
          e_sfc_snow = (t_sfc - t_sb) * c2 / dtime;
        }
      } else {
        // the whole pond freezes
        rho_avg = (h2osno * rho_avg + h2osfc * DENICE) / (h2osno + h2osfc);
        h2osno += h2osfc;
        int_snow += h2osfc;
        q_sfc_ice = h2osfc / dtime;
        if (snl > 0) ice_sb = ice_sb + h2osfc;
        t_sfc = t_sfc - temp1 * HFUS / (dtime * dhsdT - c_sfc);
        xmf_sfc = hm - fsfc * temp1 * HFUS / dtime;
        if (snl == 0) {
          t_sb = t_sfc;
        } else {
          const double c1 = (snl == 1) ? fsno * (dtime / fact_sb - dhsdT * dtime) : fsno / fact_sb * dtime;
          const double c2 = (fsfc != 0.0) ? fsfc * (c_sfc - dtime * dhsdT) : 0.0;
          t_sb = (c1 * t_sb + c2 * t_sfc) / (c1 + c2);
          t_sfc = t_sb;
        }
        h2osfc = 0.0;
        snow_depth = (fsno > 0.0 && snl > 0) ? h2osno / (rho_avg * fsno) : h2osno / DENICE;
      }
    }
  }
  C1(xmf_h2osfc) = xmf_sfc;
  C1(qflx_h2osfc_ice) = q_sfc_ice;
  C1(eflx_h2osfc_snow) = e_sfc_snow;

  ELMK_REALIGN(REALIGN);
  // ---- pass 3: phase change in snow and soil layers, one layer at a time; the water state of a layer is
  //      read only now (the solve does not need it) and written back at once ----
  double xmf = 0.0, q_snomelt = 0.0, q_snow_melt = 0.0, q_snofrz = 0.0;
  double t_new_top = 0.0, t_new_soil1 = 0.0;
ELMK_SOIL_LOOP
  for (int i = 0; i < NLEVTOT; ++i) {
    if (i % 5 == 0 && i > 0) ELMK_REALIGN(REALIGN);
    if (i < top) {
      // rows above the snow pack are untouched (their melt flags stay stale, quirk 13), except the bottom
      // snow slot, which the surface-water phase change may have initialised (phase_change_impl.hh:79-83,123-125)
      if (i < NLEVSNO) C2(qflx_snofrz_lyr, i) = 0.0;
      if (i == SB) {
        C2(t_soisno, i) = t_sb;
        C2(h2osoi_ice, i) = ice_sb;
      }
      continue;
    }
    double ti = (i == SB) ? t_sb : ((i < NLEVSNO) ? sol[i] : sol[(i + 1 < N) ? i + 1 : i]);
    double liq = C2(h2osoi_liq, i);
    double ice = (i == SB) ? ice_sb : C2(h2osoi_ice, i);
    const double fi = C2(fact, i);
    int imelt = 0;
    double tinc = 0.0, supercool = 0.0;
    if (i < NLEVSNO) {
      if (ice > 0.0 && ti > TFRZ) { imelt = 1; tinc = TFRZ - ti; ti = TFRZ; }
      if (liq > 0.0 && ti < TFRZ) { imelt = 2; tinc = TFRZ - ti; ti = TFRZ; }
    } else {
      const int k = i - NLEVSNO;
      if (ice > 0.0 && ti > TFRZ) { imelt = 1; tinc = TFRZ - ti; ti = TFRZ; }
      if (ti < TFRZ) {
        const double smp = HFUS * (TFRZ - ti) / (GRAV * ti) * 1000.0;
        supercool = C2(watsat, k) * m_pow(smp / C2(sucsat, k), -1.0 / C2(bsw, k));
        supercool *= C2(dz, i) * 1000.0;
      }
      if (liq > supercool && ti < TFRZ) { imelt = 2; tinc = TFRZ - ti; ti = TFRZ; }
      if (snl == 0 && h2osno > 0.0 && i == NLEVSNO) {
        if (ti > TFRZ) { imelt = 1; tinc = TFRZ - ti; ti = TFRZ; }
      }
    }
    double hm = 0.0;
    if (imelt > 0) {
      if (i == top) {
        if (i < NLEVSNO) {
          hm = fse * (dhsdT * tinc - tinc / fi);
        } else {
          const double temp_hm = dhsdT * tinc - tinc / fi;
          hm = (fsfc != 0.0) ? temp_hm - fsfc * (dhsdT * tinc) : temp_hm;
        }
      } else if (i == NLEVSNO) {
        hm = (1.0 - fse - fsfc) * dhsdT * tinc - tinc / fi;
      } else {
        hm = (i < NLEVSNO) ? -fse * (tinc / fi) : -tinc / fi;
      }
    }
    if (imelt == 1 && hm < 0.0) { hm = 0.0; imelt = 0; }
    if (imelt == 2 && hm > 0.0) { hm = 0.0; imelt = 0; }
    double snofrz = 0.0;
    if (imelt > 0 && fabs(hm) > 0.0) {
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
      const double wmass0 = ice + liq;
      const double wice0 = ice;
      if (xm > 0.0) {
        ice = dmax(0.0, wice0 - xm);
        heatr = hm - HFUS * (wice0 - ice) / dtime;
      } else if (xm < 0.0) {
        if (i < NLEVSNO) {
          ice = dmin(wmass0, wice0 - xm);
        } else {
          ice = (wmass0 < supercool) ? 0.0 : dmin(wmass0 - supercool, wice0 - xm);
        }
        heatr = hm - HFUS * (wice0 - ice) / dtime;
      }
      liq = dmax(0.0, wmass0 - ice);
      if (fabs(heatr) > 0.0) {
        if (i == top) {
          if (snl == 0) ti += fi * heatr / (1.0 - (1.0 - fsfc) * fi * dhsdT);
          else ti += (fi / fse) * heatr / (1.0 - fi * dhsdT);
        } else if (i == NLEVSNO) {
          ti += fi * heatr / (1.0 - (1.0 - fse - fsfc) * fi * dhsdT);
        } else {
          if (i >= NLEVSNO) ti += fi * heatr;
          else if (fse > 0.0) ti += (fi / fse) * heatr;
        }
        if (i < NLEVSNO) {
          if (liq * ice > 0.0) ti = TFRZ;
        }
      }
      xmf += HFUS * (wice0 - ice) / dtime;
      if (imelt == 1 && i < NLEVSNO) q_snomelt += dmax(0.0, (wice0 - ice)) / dtime;
      if (imelt == 2 && i < NLEVSNO) snofrz = dmax(0.0, (ice - wice0)) / dtime;
    }
    if (i < NLEVSNO) {
      if (imelt == 2) q_snofrz += snofrz;
      C2(qflx_snofrz_lyr, i) = snofrz;
    }
    if (i == top) t_new_top = ti;
    if (i == NLEVSNO) t_new_soil1 = ti;
    C2(t_soisno, i) = ti;
    C2(h2osoi_ice, i) = ice;
    C2(h2osoi_liq, i) = liq;
    C2(imelt, i) = imelt;
  }
  C1(xmf) = xmf;
  C1(qflx_snofrz) = q_snofrz;
  C1(qflx_snow_melt) = q_snow_melt;
  C1(qflx_snomelt) = q_snomelt;
  C1(eflx_snomelt) = q_snomelt * HFUS;

  // ---- new ground temperature ----
  double tg;
  if (snl > 0) {
    tg = (fsfc != 0.0) ? fse * t_new_top + (1.0 - fse - fsfc) * t_new_soil1 + fsfc * t_sfc
                       : fse * t_new_top + (1.0 - fse) * t_new_soil1;
  } else {
    tg = (fsfc != 0.0) ? (1.0 - fsfc) * t_new_soil1 + fsfc * t_sfc : t_new_soil1;
  }
  C1(t_grnd) = tg;

  C1(t_h2osfc) = t_sfc;
  C1(h2osfc) = h2osfc;
  C1(h2osno) = h2osno;
  C1(int_snow) = int_snow;
  C1(snow_depth) = snow_depth;
}

#undef t_old
#undef z_node

} // namespace elmk
