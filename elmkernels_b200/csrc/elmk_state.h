// elmk_state.h - the column state as the kernels see it.
//
// Layout in HBM: structure of arrays, COLUMN INNERMOST.  A field with nlev elements per column
// is stored as base[lev * np + col], np = padded column count.  Thirty-two consecutive threads of a
// warp (32 consecutive columns) therefore read one level of a field as one contiguous 256-byte
// segment, whatever the number of levels - the transpose of the reference's host layout
// d_[lev + col*nlev] (reference src/utils/array.hh:176-183), which would make every warp load a
// stride-nlev gather.
#pragma once
#include "elmk_common.h"

namespace elmk {

typedef double elmk_F64;
typedef int elmk_I32;
typedef unsigned char elmk_U8;

// one pointer per field of include/elmk_fields.def
struct Cols {
  long long np;   // distance (in elements) between consecutive levels of a field
  int ncols;      // number of valid columns
  int npi;        // np again, as a 32-bit value: element offsets lev * np + col fit 31 bits for every field (checked at create)
  // optional per-column CO2 / O2 partial pressures [Pa] (elmk_set_gas_pressures); null: the wrapper's constants
  const double* pco2_in;
  const double* po2_in;
#define ELMK_FIELD(name, type, nlev, cls) elmk_##type* name;
#include "../../include/elmk_fields.def"
#undef ELMK_FIELD
};

// global tables (device copies of struct elmk_tables, include/elmk_b200.h)
struct Tables {
  int ltype, ctype, vtype, urbpoi, lakpoi, oldfflag;
  double dewmx;
  // PFT constants used on the path (pft_data.h:38-77)
  double z0mr[17], displar[17];
  double xl[17], rhol[17][2], rhos[17][2], taul[17][2], taus[17][2];
  double albsat[20][2], albdry[20][2];
  // photosynthesis constants per PFT in the member order of PFTDataPSN (row 26 = tc_stress, one value for all) and the
  // root-profile parameters: only the one-time column initialisation reads them (phys_init.h)
  double psn[27][17], roota[17], rootb[17];
  // SNICAR optics (snicar_data.h:40-70)
  double aer_band[6][3][NBND_SNW];       // [oc1,oc2,dst1..dst4][ss_alb,asm_prm,ext_cff_mss][band]
  double bc[2][3][10][NBND_SNW];         // [bc1,bc2][ss_alb,asm_prm,ext_cff_mss][nclrds][band]
  double bcenh[8][10][NBND_SNW];
  const double* snw[2][3];               // [drc,dfs][ss_alb,asm_prm,ext_cff_mss] -> [band][1471] in HBM
  const double* snowage[3];              // tau, kappa, drdt0 -> [11][31][8] in HBM
};

// per-step scalars
struct StepArgs {
  double dtime, dayl, max_dayl;
#ifdef ELMK_BULK_PREFETCH   // (experiment: row base pointers for the bulk L2 prefetch of csrc/elmk_lib.cu)
  const char* const* pf_rows = nullptr;
  int pf_nrows = 0;
#endif
};

// field accessors used by the physics bodies: S is a `const Cols&`, c the column index
#define C1(field) S.field[c]
#ifdef ELMK_INDEX64
#define C2(field, lev) S.field[(long long)(lev) * S.np + c]
#else
// 32-bit element offset: one IMAD + one IMAD.WIDE per access instead of a 64-bit multiply-add chain
#define C2(field, lev) S.field[(int)(lev) * S.npi + (int)c]
#endif

// row of a multi-level field of one column: in the column-innermost layout of the state (ELMK_ROW), or contiguous
// (stride 1: per-thread rows and the flat argument arrays of elmk_fn_call)
struct ColRow {
  double* p;
  int stride;   // (32-bit element offsets: i * stride < 2^31, as for C2)
  ELMK_HD double& operator[](const int i) const { return p[i * stride]; }
  ELMK_HD ColRow from(const int first) const { return ColRow{p + first * stride, stride}; }   // the row from element `first` on
};
#define ELMK_ROW(field) ColRow{S.field + c, S.npi}

} // namespace elmk
