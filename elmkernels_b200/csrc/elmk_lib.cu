// elmk_lib.cu - libelmk_b200.so: the C ABI of include/elmk_b200.h over hand-written FP64 CUDA kernels
// for sm_100a.  One handle = one device, one stream, one contiguous range of land columns resident in
// HBM as structure-of-arrays with the column index innermost (elmk_state.h).
//
// Kernels: one thread per column.  A launch runs a compile-time set of kernel groups (a bit mask of
// ELMK_G_*) back to back for its column, in the chain order of the reference's
// ELMInterface::advance (driver/kokkos/elm_kokkos_interface.cc:289-318); elmk_step covers the requested
// mask with the launches of the active plan (see kPlans).  The reference issues 23 parallel_for
// dispatches and 107 scratch allocations per step for the same work.
//
// There is no host fallback: every entry point that computes needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/elmk_b200.h"
#include "elmk_state.h"
#include "phys_albedo.h"
#include "phys_bareground.h"
#include "phys_canflux.h"
#include "phys_cantemp.h"
#include "phys_forcing.h"
#include "phys_hydrology.h"
#include "phys_init.h"
#include "phys_radiation.h"
#include "phys_snow.h"
#include "phys_soiltemp.h"
#include "phys_surfflux.h"

namespace {
using namespace elmk;

// ------------------------------------------------------------------------------------------------
// field table
// ------------------------------------------------------------------------------------------------
struct Spec { const char* name; int dtype; int nlev; };
const Spec kSpecs[] = {
#define F64 ELMK_F64
#define I32 ELMK_I32
#define U8 ELMK_U8
#define ELMK_FIELD(name, type, nlev, cls) {#name, type, nlev},
#include "../../include/elmk_fields.def"
#undef ELMK_FIELD
#undef F64
#undef I32
#undef U8
};
constexpr int kNumFields = sizeof(kSpecs) / sizeof(kSpecs[0]);
inline size_t esize(int dt) { return dt == ELMK_F64 ? 8 : dt == ELMK_I32 ? 4 : 1; }

constexpr int kBlock = 128;              // threads per block of the column kernels
constexpr int kColAlign = 128;           // columns are padded to a multiple of this
constexpr size_t kStageBytes = 64u << 20; // device staging buffer for layout conversion (per slot)

// ------------------------------------------------------------------------------------------------
// column kernels
// ------------------------------------------------------------------------------------------------
// launches whose blocks re-align their warps (instruction-cache sharing, see ELMK_REALIGN in elmk_common.h): soil
// temperature alone (barriers inside its body) and the fused unsorted launches (barriers between their groups)
template <uint32_t MASK> constexpr bool kWholeBlocks = (MASK == ELMK_G_SOIL_TEMPERATURE) || ((MASK & (MASK - 1u)) != 0 && !(MASK & ELMK_G_SOIL_TEMPERATURE));

// `live`: the thread owns a real column.  With REALIGN the threads of a block meet again between the groups of a fused
// launch (and, for soil temperature, inside the group's body), so threads without a column stay until the end.
// internal launch-mask bit (never part of the public group mask): the albedo group for columns whose SNICAR results the
// SNICAR kernel has already stored (k_snicar below)
constexpr uint32_t G_ALBEDO_REST = 1u << 20;

template <uint32_t MASK, bool REALIGN = false>
__device__ __forceinline__ void run_groups(const Cols& S, const Tables& T, const StepArgs& A, const int c, const bool live = true)
{
#define ELMK_GROUP(BIT, CALL)                                        \
  if (MASK & (BIT)) {                                                \
    if (live) { CALL; }                                              \
    if (REALIGN && (MASK & ~(((BIT) << 1) - 1u))) __syncthreads();   \
  }
  ELMK_GROUP(ELMK_G_FRAC_WET, column_frac_wet(S, T, c))
  ELMK_GROUP(ELMK_G_ALBEDO, column_albedo<false>(S, T, c))
  ELMK_GROUP(G_ALBEDO_REST, column_albedo<true>(S, T, c))
  ELMK_GROUP(ELMK_G_CANOPY_HYDROLOGY, column_canopy_hydrology(S, T, A.dtime, c))
  ELMK_GROUP(ELMK_G_SURFACE_RADIATION, column_surface_radiation(S, T, c))
  ELMK_GROUP(ELMK_G_CANOPY_TEMPERATURE, column_canopy_temperature(S, T, c))
  ELMK_GROUP(ELMK_G_BAREGROUND_FLUXES, column_bareground_fluxes(S, T, c))
  ELMK_GROUP(ELMK_G_CANOPY_FLUXES, column_canopy_fluxes(S, T, A, c))
  if (MASK & ELMK_G_SOIL_TEMPERATURE) {
    // barriers inside the body: every thread runs it, the padding columns (ncols..np, zero-filled, never downloaded) included
    if (REALIGN || live) column_soil_temperature<REALIGN>(S, T, A.dtime, c);
  }
  ELMK_GROUP(ELMK_G_SNOW_HYDROLOGY, column_snow_hydrology(S, T, A.dtime, c))
  ELMK_GROUP(ELMK_G_SURFACE_FLUXES, column_surface_fluxes(S, T, A.dtime, c))
  ELMK_GROUP(ELMK_G_CONSERVATION, column_conservation(S, T, A.dtime, c))
#undef ELMK_GROUP
}

// ---- (experiment, -DELMK_BULK_PREFETCH) bulk prefetch of a block's input rows into L2 through the TMA engine:
// cp.async.bulk.prefetch.L2, SASS UBLKPF.  With the column-innermost layout every (field, level) row of a block's 128
// columns is one contiguous, 16-byte aligned segment of 1 KB.  The host passes the list of row base pointers the
// launch is going to read (StepArgs::pf_rows, built once per handle); thread t of a block asks for rows t, t + 128, ...
// before the column work starts.  Measured and not adopted: profiles/r2_experiments.md.
#ifdef ELMK_BULK_PREFETCH
__device__ __forceinline__ void prefetch_rows(const StepArgs& A)
{
  const long long off = (long long)blockIdx.x * kBlock * (long long)sizeof(double);
  for (int r = threadIdx.x; r < A.pf_nrows; r += kBlock)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.pf_rows[r] + off), "r"((unsigned)(kBlock * sizeof(double))) : "memory");
}
#endif

template <uint32_t MASK>
__global__ void __launch_bounds__(kBlock) k_groups(const Cols S, const Tables* __restrict__ Tp, const StepArgs A)
{
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (!kWholeBlocks<MASK> && c >= S.ncols) return;
  run_groups<MASK, kWholeBlocks<MASK>>(S, *Tp, A, c, c < S.ncols);
}
// same, with a floor on the resident blocks per SM (caps the registers per thread)
template <uint32_t MASK, int MINBLOCKS>
__global__ void __launch_bounds__(kBlock, MINBLOCKS) k_groups_occ(const Cols S, const Tables* __restrict__ Tp, const StepArgs A)
{
#ifdef ELMK_BULK_PREFETCH
  if (A.pf_nrows) prefetch_rows(A);
#endif
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (!kWholeBlocks<MASK> && c >= S.ncols) return;
  run_groups<MASK, kWholeBlocks<MASK>>(S, *Tp, A, c, c < S.ncols);
}

// ---- (experiment, development builds) one thread per column, the columns of a block re-ordered by work class ----------
// The snow state machine only does work for columns that carry snow layers, and how much depends on the layer count; in
// column order a warp holds a mixture (ncu: 15.9 of 32 lanes active in the snow launch).  The block's 128 columns are
// counting-sorted (stable) by class in shared memory and thread t takes the t-th column of that order: warps become
// homogeneous, while every (field, level) access of the block still falls into the same 1 KB row segment it reads in
// column order - the same sectors reach the SM, only their assignment to warps changes.  KEY: 1 = snl (0..5),
// 2 = snl > 0, 3 = exposed vegetation.  Padding columns sort last.  Measured and not adopted (profiles/r2_experiments.md:
// snow launch 3.0 -> 3.9 / 3.6 ms with keys 1 / 2): the 8-byte accesses of a class-ordered warp touch up to 32 sectors
// instead of 8, and these launches are bound by that load / store path, not by idle lanes.
#ifdef ELMK_DEV_VARIANTS
template <int KEY>
__device__ __forceinline__ int work_class(const Cols& S, const int c)
{
  if (KEY == 1) return S.snl[c];
  if (KEY == 2) return S.snl[c] > 0 ? 0 : 1;
  return S.frac_veg_nosno[c] != 0 ? 0 : 1;
}
template <int KEY>
__device__ __forceinline__ int block_order_by_class(const Cols& S)
{
  constexpr int NCLS = 7, NW = kBlock / 32;
  __shared__ unsigned char perm[kBlock];
  __shared__ unsigned char cnt[NW][NCLS];
  const int t = threadIdx.x, w = t >> 5;
  const unsigned lane = t & 31u, below = (1u << lane) - 1u;
  const int c = blockIdx.x * kBlock + t;
  const int cls = (c < S.ncols) ? work_class<KEY>(S, c) : NCLS - 1;
  unsigned mine = 0;
#pragma unroll
  for (int k = 0; k < NCLS; ++k) {
    const unsigned m = __ballot_sync(0xffffffffu, cls == k);
    if (lane == 0) cnt[w][k] = (unsigned char)__popc(m);
    if (cls == k) mine = m;
  }
  __syncthreads();
  int pos = __popc(mine & below);
#pragma unroll
  for (int k = 0; k < NCLS; ++k)
#pragma unroll
    for (int v = 0; v < NW; ++v)
      if (k < cls || (k == cls && v < w)) pos += cnt[v][k];
  perm[pos] = (unsigned char)t;
  __syncthreads();
  return blockIdx.x * kBlock + perm[t];
}
template <uint32_t MASK, int MINBLOCKS, int KEY>
__global__ void __launch_bounds__(kBlock, MINBLOCKS) k_groups_classed(const Cols S, const Tables* __restrict__ Tp, const StepArgs A)
{
  const int c = block_order_by_class<KEY>(S);
  if (!kWholeBlocks<MASK> && c >= S.ncols) return;
  run_groups<MASK, kWholeBlocks<MASK>>(S, *Tp, A, c, c < S.ncols);
}
#endif

// ---- SNICAR with one warp-task per (32 columns, incident-flux type, spectral band) -------------------------------
// kokkos_albedo_snicar spends ~85 % of its instructions in the snow radiative transfer (round-1 ncu source view), and
// that part is ten independent solves per column: direct / diffuse x five bands (snow_snicar_impl.hh:313-670), each a
// loop over up to five layers with eight Gauss points per layer.  One thread per column ran them one after the other
// with the union of their arrays live (246 registers, 2.7 KB of local memory, 7 warps per SM).  Here a block owns a
// window of columns, lists its sunlit snow columns in shared memory ordered by their number of layers, and its warps
// pull work items (chunk of 32 listed columns, flg, band) from a shared counter: lane = column, the whole warp in
// the same band.  (Layers below the depth where a band's transmission falls under 0.001 skip their solve; that cut-off
// depends mostly on the band, so lanes of one band stay together - with one lane per band of the same column only 17
// of 30 lanes were active in the layer solve.)  The seven results of a solve go to a scratch slice of the block
// (L2-resident); after a block barrier one thread per (column, flg) adds the five bands in the reference's order
// (snow_albedo_radiation_factor :706-760) and stores albsnd / albsni and the absorption factors
// (flux_absorption_factor :199-207).  Every value is computed by the same operations in the same order as in
// column_albedo<false>: bit-identical (tests: split plan = one thread per column, fused plan = this kernel).
constexpr int kSnicarWindow = 1024;
constexpr int kSnicarBlock = 128;
constexpr int kSnicarTasks = 2 * NBND_SNW;                    // (flg, band)
constexpr int kSnicarValues = NLEVSNO + 2;                    // albedo + absorbed flux of five slots and the ground
constexpr size_t kSnicarSlice = (size_t)kSnicarWindow * kSnicarTasks * kSnicarValues;   // doubles of scratch per block
template <int MINBLOCKS>
__global__ void __launch_bounds__(kSnicarBlock, MINBLOCKS) k_snicar(const Cols S, const Tables* __restrict__ Tp,
                                                                     double* __restrict__ scratch_all, const int nwindows)
{
  __shared__ unsigned short order[kSnicarWindow];
  __shared__ int count[8], start[8], next_item, nactive;
  const Tables& T = *Tp;
  const int lane = threadIdx.x & 31;
  double* const scratch = scratch_all + (size_t)blockIdx.x * kSnicarSlice;   // [task][value][listed column]
  for (int w = blockIdx.x; w < nwindows; w += gridDim.x) {
    const int base = w * kSnicarWindow;
    const int nvalid = (S.ncols - base < kSnicarWindow) ? (S.ncols - base) : kSnicarWindow;
    __syncthreads();   // the previous window's list is no longer read
    if (threadIdx.x < 8) count[threadIdx.x] = 0;
    if (threadIdx.x == 0) next_item = 0;
    __syncthreads();
    constexpr int R = kSnicarWindow / kSnicarBlock;
    int key[R], rank[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = r * kSnicarBlock + threadIdx.x;
      key[r] = -1;
      if (i < nvalid) {
        const int c = base + i;
        if (S.coszen[c] > 0.0 && S.h2osno[c] > alb::MIN_SNW) {
          const int snl = S.snl[c];
          key[r] = snl > 0 ? snl : 1;          // layers the solve walks through
          rank[r] = atomicAdd(&count[key[r]], 1);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0;
      for (int k = 7; k >= 0; --k) { start[k] = acc; acc += count[k]; }   // deepest snow packs first
      nactive = acc;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (key[r] >= 0) order[start[key[r]] + rank[r]] = (unsigned short)(r * kSnicarBlock + threadIdx.x);
    __syncthreads();
    const int nact = nactive;
    const int nitems = ((nact + 31) / 32) * kSnicarTasks;
    // ---- the band solves ----
    while (true) {
      int item = 0;
      if (lane == 0) item = atomicAdd(&next_item, 1);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= nitems) break;
      const int chunk = item / kSnicarTasks, task = item - chunk * kSnicarTasks;   // the tasks of a chunk run side by side
      const int flg = task / NBND_SNW + 1, b = task - (flg - 1) * NBND_SNW;
      const int idx = chunk * 32 + lane;
      if (idx < nact) {
        const int c = base + order[idx];
        const double coszen = C1(coszen), h2osno = C1(h2osno);
        const int snl = C1(snl);
        // soil_albedo :702-709 (the ground under the snow pack)
        const int col = C1(isoicol);
        const double inc = dmax(0.11 - 0.40 * C2(h2osoi_vol, 0), 0.0);
        double albsoi[NUMRAD];
#pragma unroll
        for (int ib = 0; ib < NUMRAD; ++ib) albsoi[ib] = dmin(T.albsat[col][ib] + inc, T.albdry[col][ib]);
        uint32_t err = 0;
        SnicarColumn K;
        snicar_column(S, c, h2osno, snl, K, err);
        double albedo = 0.0, fb[NLEVSNO + 1];
#pragma unroll
        for (int i = 0; i <= NLEVSNO; ++i) fb[i] = 0.0;
        if (!(err & ERR_SNICAR_RADIUS)) {   // (the reference throws; results stay zero)
          const auto cnc_of = [&](const int i, double (&cnc)[NAER]) { snicar_cnc(S, c, i, cnc); };
          snicar_band(T, K, flg, b, dmax(coszen, 0.01), albsoi, cnc_of, albedo, fb, err);
        }
        if (err) atomicOr(&S.errmask[c], (int)err);
        double* out = scratch + (size_t)task * kSnicarValues * kSnicarWindow + idx;
        out[0] = albedo;
#pragma unroll
        for (int i = 0; i <= NLEVSNO; ++i) out[(size_t)(i + 1) * kSnicarWindow] = fb[i];
      }
    }
    __syncthreads();   // (block-scope visibility of the scratch slice)
    // ---- band weighting to VIS / NIR, one thread per (listed column, flg) ----
    for (int j = threadIdx.x; j < 2 * nact; j += kSnicarBlock) {
      const int idx = j >> 1, flg = (j & 1) + 1;
      const int c = base + order[idx];
      const double coszen = C1(coszen), h2osno = C1(h2osno);
      const int snl = C1(snl);
      uint32_t err = 0;
      SnicarColumn K;
      snicar_column(S, c, h2osno, snl, K, err);
      int rds_top = 0;
#pragma unroll
      for (int i = 0; i < NLEVSNO; ++i) if (i == K.top) rds_top = K.rds[i];
      double albout_lcl[NBND_SNW], flx_abs_lcl[NLEVSNO + 1][NBND_SNW];
#pragma unroll
      for (int bb = 0; bb < NBND_SNW; ++bb) {
        const double* in = scratch + (size_t)((flg - 1) * NBND_SNW + bb) * kSnicarValues * kSnicarWindow + idx;
        albout_lcl[bb] = in[0];
#pragma unroll
        for (int i = 0; i <= NLEVSNO; ++i) flx_abs_lcl[i][bb] = in[(size_t)(i + 1) * kSnicarWindow];
      }
      double alb_out[NUMRAD] = {0.0, 0.0}, flx_abs[NLEVSNO + 1][NUMRAD];
#pragma unroll
      for (int i = 0; i <= NLEVSNO; ++i) { flx_abs[i][0] = 0.0; flx_abs[i][1] = 0.0; }
      if (!(err & ERR_SNICAR_RADIUS))
        snicar_combine(flg, dmax(coszen, 0.01), K.top, rds_top, albout_lcl, flx_abs_lcl, alb_out, flx_abs);
      if (flg == 1) {
        C2(albsnd, 0) = alb_out[0]; C2(albsnd, 1) = alb_out[1];
#pragma unroll
        for (int i = 0; i <= NLEVSNO; ++i) {
          C2(flx_absdv, i) = flx_abs[i][0] * (1.0 - alb_out[0]);
          C2(flx_absdn, i) = flx_abs[i][1] * (1.0 - alb_out[1]);
        }
      } else {
        C2(albsni, 0) = alb_out[0]; C2(albsni, 1) = alb_out[1];
#pragma unroll
        for (int i = 0; i <= NLEVSNO; ++i) {
          C2(flx_absiv, i) = flx_abs[i][0] * (1.0 - alb_out[0]);
          C2(flx_absin, i) = flx_abs[i][1] * (1.0 - alb_out[1]);
        }
      }
    }
  }
}

// ---- CanopyFluxes with warp-level re-packing of the stability iteration ---------------------------
// The number of passes of the leaf-temperature / stability loop is data dependent: 3..41, mean ~6, and a
// warp of 32 consecutive (or even class-sorted) columns runs as long as its slowest lane - measured mean
// of the per-warp maximum ~15, i.e. ~40 % lane utilisation.  The loop is therefore taken out of the
// column kernel:
//   k_canflux_begin    one thread per column: initialize_flux (moisture stress = 15 pow per column, first
//                      guess), iteration state -> scratch (column-innermost, coalesced), column appended to
//                      the day list (stomatal root-find every pass) or the night list
//   k_canflux_iterate  persistent warps; every lane owns one column of the queue and runs ONE pass per
//                      round; a lane whose column has converged stores its state and takes the next
//                      queue entry (warp-aggregated atomic), so lanes stay busy whatever the pass counts
//   k_canflux_end      one thread per column: compute_flux from the stored state
// The arithmetic is canflux_begin / canflux_iterate / canflux_end of phys_canflux.h, the same functions
// the one-launch column_canopy_fluxes chains, so the result is bit-identical to the plain kernel.
constexpr int kCanfluxDoubles = 0
#define X(n) +1
    ELMK_CANFLUX_CONST(X) ELMK_CANFLUX_CARRIED(X) ELMK_CANFLUX_INT(X)
#undef X
    ;

// 384-thread lock-step blocks: fastest of 128...768 threads with and without lock-step on B200
constexpr int kIterBlock = 384;
// which read-only values of a column in flight the iteration kernel keeps in shared memory (IterView below)
constexpr int kIterSet = 3;
// row of the scratch that holds the pass count of a column (itlef) after the launch
constexpr int kCanfluxItlefRow = 0
#define X(n) +1
    ELMK_CANFLUX_CONST(X) ELMK_CANFLUX_CARRIED(X)
#undef X
    + 3;   // nrad, veg, soybean, itlef
struct CanfluxQueue {
  double* scratch;   // [kCanfluxDoubles][np]
  int* list;         // [np]: day columns from the front, night columns from the back
  int* counters;     // [0] day count, [1] night count, [2] queue head
  long long np;
};

template <class IT>
__device__ __forceinline__ void canflux_store(const CanfluxQueue& Q, const int c, const IT& I, const bool all)
{
  double* p = Q.scratch + c;
  long long k = 0;
  // `all` (set-up launch): constants + the carried values that do not start at zero; else (a converged column of
  // the iteration kernel): every carried value
#define X(n) if (all) p[k * Q.np] = I.n; ++k;
  ELMK_CANFLUX_CONST(X)
#undef X
#define X(n) p[k * Q.np] = I.n; ++k;
  ELMK_CANFLUX_CARRIED_SET(X)
#undef X
#define X(n) if (!all) p[k * Q.np] = I.n; ++k;
  ELMK_CANFLUX_CARRIED_ZERO(X)
#undef X
#define X(n) p[k * Q.np] = (double)I.n; ++k;
  ELMK_CANFLUX_INT(X)
#undef X
}
// FRESH: the column comes from the set-up launch (the values that start at zero are not in the scratch)
template <bool FRESH = false, class IT = CanopyIter>
__device__ __forceinline__ void canflux_load(const Cols& S, const CanfluxQueue& Q, const int c, IT& I)
{
#define X(n, e) I.n = e;
  ELMK_CANFLUX_STATE(X)
#undef X
  const double* p = Q.scratch + c;
  long long k = 0;
#define X(n) I.n = p[k * Q.np]; ++k;
  ELMK_CANFLUX_CONST(X)
  ELMK_CANFLUX_CARRIED_SET(X)
#undef X
#define X(n) I.n = FRESH ? 0.0 : p[k * Q.np]; ++k;
  ELMK_CANFLUX_CARRIED_ZERO(X)
#undef X
#define X(n) I.n = (int)p[k * Q.np]; ++k;
  ELMK_CANFLUX_INT(X)
#undef X
}

__global__ void __launch_bounds__(kBlock, 8) k_canflux_begin(const Cols S, const Tables* __restrict__ Tp, const StepArgs A,
                                                          const CanfluxQueue Q)
{
  const int c = blockIdx.x * kBlock + threadIdx.x;
  int cls = 0;   // 0: nothing to iterate, 1: night, 2: day
  if (c < S.ncols) {
    const PsnPft P = load_psn_pft(S, c);
    // Unvegetated columns store zeros: every 32-byte sector of a scratch row is then written whole.  With only the
    // vegetated lanes storing, the partially written sectors cost a read-modify-write in HBM (ECC granule): 1.21 -> 1.16 ms.
    CanopyIter I = {};
    const bool veg = canflux_begin(S, Tp->vtype, A, P, c, I);
    if (!veg) I = CanopyIter{};
    canflux_store(Q, c, I, true);
    if (veg) cls = (I.parsun > 0.0 || I.parsha > 0.0) ? 2 : 1;
  }
  // warp-aggregated append: one atomic per warp and list
  const unsigned lane = threadIdx.x & 31u;
  const unsigned day = __ballot_sync(0xffffffffu, cls == 2), night = __ballot_sync(0xffffffffu, cls == 1);
  int base_day = 0, base_night = 0;
  if (lane == 0) {
    if (day) base_day = atomicAdd(&Q.counters[0], __popc(day));
    if (night) base_night = atomicAdd(&Q.counters[1], __popc(night));
  }
  base_day = __shfl_sync(0xffffffffu, base_day, 0);
  base_night = __shfl_sync(0xffffffffu, base_night, 0);
  const unsigned below = (1u << lane) - 1u;
  if (cls == 2) Q.list[base_day + __popc(day & below)] = c;
  if (cls == 1) Q.list[Q.np - 1 - (base_night + __popc(night & below))] = c;
}

// "any thread of my lock-step group still owns a column": a named barrier over GROUP threads (warps of one group are
// consecutive), GROUP == 0: the whole block
template <int GROUP>
__device__ __forceinline__ bool lockstep_any(const bool have)
{
  if (GROUP == 0) return __syncthreads_or(have);
  unsigned r;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
               : "=r"(r)
               : "r"((unsigned)have), "r"(1u + threadIdx.x / (GROUP ? GROUP : 1)), "r"((unsigned)GROUP)
               : "memory");
  return r != 0;
}

// Read-only values of a column in flight live in shared memory instead of registers / the spilled frame - one record per
// thread, an ODD number of doubles apart, so that the 64-bit accesses of a half-warp fall into 16 different bank pairs:
//   SET 1: the per-column constants of the photosynthesis model (PsnPft: 27 PFT values, PsnColumn: 16 derived ones)
//   SET 2: + the 12 CONST members of the iteration state
//   SET 3: + the 18 STATE_A members (read once or twice per pass)
// IterView is the iteration state with those members as references into the record; canflux_iterate / load / store
// are templates on the state type.
struct IterConst {
  PsnPft P;
  PsnColumn PC;
};
constexpr int kIterConstDoubles = (int)(sizeof(IterConst) / sizeof(double));
enum {
#define X(n) kIterSlot_##n,
  ELMK_CANFLUX_CONST(X)
#undef X
#define X(n, e) kIterSlot_##n,
  ELMK_CANFLUX_STATE_A(X)
#undef X
  kIterSlotCount
};
constexpr int kIterConstSlots = 0
#define X(n) +1
    ELMK_CANFLUX_CONST(X)
#undef X
    ;
constexpr int iter_record_doubles(const int set)
{
  return (kIterConstDoubles + (set >= 3 ? kIterSlotCount : set == 2 ? kIterConstSlots : 0)) | 1;
}
static_assert(iter_record_doubles(3) % 2 == 1 && iter_record_doubles(2) % 2 == 1 && iter_record_doubles(1) % 2 == 1,
              "iteration records: odd stride in doubles (conflict-free 64-bit shared-memory accesses)");
static_assert((size_t)kIterBlock * iter_record_doubles(kIterSet) * sizeof(double) <= 227u * 1024u,
              "iteration records of one block exceed the 227 KB of shared memory a block can opt into on sm_100a");
template <bool REF> struct IterSlot { typedef double type; };
template <> struct IterSlot<true> { typedef double& type; };
template <int SET>
struct IterView {
#define X(n, e) typename IterSlot<(SET >= 3)>::type n;
  ELMK_CANFLUX_STATE_A(X)
#undef X
#define X(n, e) double n;
  ELMK_CANFLUX_STATE_B(X)
#undef X
#define X(n) typename IterSlot<(SET >= 2)>::type n;
  ELMK_CANFLUX_CONST(X)
#undef X
#define X(n) double n;
  ELMK_CANFLUX_CARRIED(X)
#undef X
#define X(n) int n;
  ELMK_CANFLUX_INT(X)
#undef X
  // ro: the thread's slots in shared memory; members that are plain doubles start at zero (`zero` is a local 0.0)
  __device__ __forceinline__ IterView(double* ro, double& zero)
      :
#define X(n, e) n(SET >= 3 ? ro[kIterSlot_##n] : zero),
        ELMK_CANFLUX_STATE_A(X)
#undef X
#define X(n, e) n(0.0),
        ELMK_CANFLUX_STATE_B(X)
#undef X
#define X(n) n(SET >= 2 ? ro[kIterSlot_##n] : zero),
        ELMK_CANFLUX_CONST(X)
#undef X
#define X(n) n(0.0),
        ELMK_CANFLUX_CARRIED(X)
#undef X
        nrad(0), veg(0), soybean(0), itlef(0), nmozsgn(0), err(0)
  {
  }
};
template <int BLOCK, bool LOCKSTEP, int GROUP = 0, int SET = 1>
__global__ void __launch_bounds__(BLOCK) k_canflux_iterate(const Cols S, const CanfluxQueue Q)
{
  extern __shared__ double iter_smem[];
  const int nday = Q.counters[0], total = nday + Q.counters[1];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned below = (1u << lane) - 1u;
  bool have = false;
  int c = 0;
  PsnPft P_reg = {};
  PsnColumn PC_reg = {};
  double* const record = iter_smem + (SET ? threadIdx.x * iter_record_doubles(SET) : 0);
  IterConst* const mine = reinterpret_cast<IterConst*>(record);
  PsnPft& P = SET ? mine->P : P_reg;
  PsnColumn& PC = SET ? mine->PC : PC_reg;
  double zero = 0.0;
  IterView<SET> I(record + kIterConstDoubles, zero);
  while (true) {
    // ---- refill idle lanes from the queue ----
    const unsigned need = __ballot_sync(0xffffffffu, !have);
    if (need) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&Q.counters[2], __popc(need));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (!have) {
        const int q = base + __popc(need & below);
        if (q < total) {
          c = (q < nday) ? Q.list[q] : Q.list[Q.np - 1 - (q - nday)];
          canflux_load<true>(S, Q, c, I);
          P = load_psn_pft(S, c);
          PC = psn_column(P, I.t10, I.pbot, I.thm, I.forc_po2, I.dayl_factor);
          have = true;
        }
      }
    }
    if (LOCKSTEP) {
      // The pass body is several times the 32 KB L1.5 instruction cache (kernel: 126 KB of SASS).  Starting every pass
      // together keeps the warps of the block inside the same stretch of code, so that one warp's
      // instruction fetch serves the others (ncu: "no_instruction" was the top stall reason; measured
      // 7.4 ms -> 4.9 ms per 512k columns with 384-thread lock-step blocks; more barriers inside the pass
      // cost more than they saved).
      if (!lockstep_any<GROUP>(have)) break;
    } else {
      if (!__any_sync(0xffffffffu, have)) break;
    }
    // ---- one pass for every lane that owns a column ----
    if (have) {
      if (canflux_iterate(P, PC, I)) {
        canflux_store(Q, c, I, false);
        have = false;
      }
    }
  }
}

__global__ void __launch_bounds__(kBlock) k_canflux_end(const Cols S, const CanfluxQueue Q)
{
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= S.ncols || S.frac_veg_nosno[c] == 0) return;
  CanopyIter I;
  canflux_load(S, Q, c, I);
  canflux_end(S, c, I);
}

__global__ void __launch_bounds__(kBlock) k_init_timestep(const Cols S, const Tables* __restrict__ Tp, const int reset)
{
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= S.ncols) return;
  column_init_timestep(S, *Tp, reset, c);
}

__global__ void __launch_bounds__(kBlock) k_init_columns(const Cols S, const Tables* __restrict__ Tp, const InitInputs X)
{
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= S.ncols) return;
  column_init(S, *Tp, X, c);
}
__global__ void __launch_bounds__(kBlock) k_atm_forcing(const Cols S, const AtmSeries A, const int t, const double wt1,
                                                        const double wt2, const int qbot_is_rh)
{
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= S.ncols) return;
  column_atm_forcing(S, A, t, wt1, wt2, qbot_is_rh != 0, c);
}
// histogram of the stability-iteration pass counts of the vegetated columns (elmk_canflux_pass_histogram)
__global__ void __launch_bounds__(256) k_pass_histogram(const Cols S, const double* __restrict__ itlef, unsigned long long* hist)
{
  __shared__ unsigned int h[42];
  if (threadIdx.x < 42) h[threadIdx.x] = 0;
  __syncthreads();
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < S.ncols; c += gridDim.x * blockDim.x)
    if (S.frac_veg_nosno[c] != 0) {
      int it = (int)itlef[c];
      it = it < 0 ? 0 : it > 41 ? 41 : it;
      atomicAdd(&h[it], 1u);
    }
  __syncthreads();
  if (threadIdx.x < 42 && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}
__global__ void __launch_bounds__(kBlock) k_coszen(const Cols S, const SolarStep G)
{
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= S.ncols) return;
  column_coszen(S, G, c);
}
__global__ void __launch_bounds__(kBlock) k_phenology(const Cols S, const PhenSeries P, const int m, const double wt1,
                                                      const double wt2)
{
  const int c = blockIdx.x * kBlock + threadIdx.x;
  if (c >= S.ncols) return;
  column_phenology(S, P, m, wt1, wt2, c);
}

// ---- bare-ground fluxes on the compacted list of unvegetated columns -----------------------------------------------
// Only columns without exposed vegetation do work in this group (three Monin-Obukhov passes; ncu: 10 of 32 lanes active
// with one thread per column).  A block owns a window of 1024 consecutive columns, zeroes what compute_flux zeroes
// for every column, compacts the indices of the bare ones in shared memory (warp-aggregated append) and its threads
// then take the list entries: full warps, and the 8-byte accesses of a warp stay within the window's 8 KB per row.
#ifndef ELMK_BARE_WINDOW
#define ELMK_BARE_WINDOW 1024
#endif
#ifndef ELMK_BARE_OCC
#define ELMK_BARE_OCC 8
#endif
constexpr int kBareWindow = ELMK_BARE_WINDOW;
__global__ void __launch_bounds__(kBlock, ELMK_BARE_OCC) k_bareground_compact(const Cols S, const Tables* __restrict__ Tp, const StepArgs)
{
  __shared__ unsigned short list[kBareWindow];
  __shared__ int count;
  const long long w0 = (long long)blockIdx.x * kBareWindow;
  const unsigned lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
#pragma unroll 1
  for (int i = threadIdx.x; i < kBareWindow; i += kBlock) {
    const long long c = w0 + i;
    const bool bare = (c < S.ncols) && (S.frac_veg_nosno[c] == 0);
    if (c < S.ncols && !bare) { S.cgrnd[c] = 0.0; S.cgrnds[c] = 0.0; S.cgrndl[c] = 0.0; }
    const unsigned m = __ballot_sync(0xffffffffu, bare);
    int at = 0;
    if (lane == 0 && m) at = atomicAdd(&count, __popc(m));
    at = __shfl_sync(0xffffffffu, at, 0);
    if (bare) list[at + __popc(m & below)] = (unsigned short)i;
  }
  __syncthreads();
  const int n = count;
#pragma unroll 1
  for (int i = threadIdx.x; i < n; i += kBlock) column_bareground_fluxes(S, *Tp, (int)(w0 + list[i]));
}

// elmk_fn_call: one library-level physics function on the flat argument array of ONE column (include/elm/*.h)
#define FlatRow(ptr) ColRow{(ptr), 1}
constexpr int kFnSlots[ELMK_FN_COUNT] = {13, 14, 7, 150, 26, 11, 36, 39, 9, 18, 42, 25, 160, 93, 35, 61, 10, 6, 20, 22, 56, 217, 127, 83};
__global__ void k_fn_call(const int fn, double* __restrict__ a)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  switch (fn) {
    case ELMK_FN_INTERCEPTION:
      hyd::interception((int)a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12]);
      break;
    case ELMK_FN_GROUND_FLUX:
      hyd::ground_flux((int)a[0], (int)a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13]);
      break;
    case ELMK_FN_FRACTION_WET:
      hyd::fraction_wet((int)a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
      break;
    case ELMK_FN_SNOW_INIT: {
      // dtime capsnow oldfflag forc_t t_grnd snow_grnd snow_melt n_melt | snow_depth h2osno int_snow | swe_old[5] liq[20]
      // ice[20] t[20] frac_iceold[5] | snl | dz[20] z[20] zi[21] snw_rds[5] | frac_sno_eff frac_sno
      int snl = (int)a[81];
      hyd::snow_init(a[0], (int)a[1], (int)a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], FlatRow(a + 11), FlatRow(a + 16),
                     FlatRow(a + 36), FlatRow(a + 56), FlatRow(a + 76), snl, FlatRow(a + 82), FlatRow(a + 102), FlatRow(a + 122),
                     FlatRow(a + 143), a[148], a[149]);
      a[81] = (double)snl;
    } break;
    case ELMK_FN_FRACTION_H2OSFC:
      hyd::fraction_h2osfc(a[0], a[1], a[2], FlatRow(a + 3), a[23], a[24], a[25]);
      break;
    case ELMK_FN_RAD_INITIALIZE_FLUX:
      rad::initialize_flux(a[0], a[1], a[2], a[3], a[4], FlatRow(a + 5));
      break;
    case ELMK_FN_RAD_TOTAL_ABSORBED:
      // snl | ftdd ftid ftii solad solai fabd fabi albsod albsoi albsnd albsni albgrd albgri [2 each] | sabv fsa sabg sabg_soil
      // sabg_snow | trd[2] tri[2]
      rad::total_absorbed_radiation((int)a[0], FlatRow(a + 1), FlatRow(a + 3), FlatRow(a + 5), FlatRow(a + 7), FlatRow(a + 9),
                                    FlatRow(a + 11), FlatRow(a + 13), FlatRow(a + 15), FlatRow(a + 17), FlatRow(a + 19),
                                    FlatRow(a + 21), FlatRow(a + 23), FlatRow(a + 25), a[27], a[28], a[29], a[30], a[31],
                                    FlatRow(a + 32), FlatRow(a + 34));
      break;
    case ELMK_FN_RAD_LAYER_ABSORBED:
      // snl sabg sabg_snow snow_depth | flx_absdv flx_absdn flx_absiv flx_absin [6 each] | trd[2] tri[2] | sabg_lyr[6] | ok
      a[38] = rad::layer_absorbed_radiation((int)a[0], a[1], a[2], a[3], FlatRow(a + 4), FlatRow(a + 10), FlatRow(a + 16),
                                            FlatRow(a + 22), FlatRow(a + 28), FlatRow(a + 30), FlatRow(a + 32)) ? 1.0 : 0.0;
      break;
    case ELMK_FN_RAD_REFLECTED:
      rad::reflected_radiation(FlatRow(a + 0), FlatRow(a + 2), FlatRow(a + 4), FlatRow(a + 6), a[8]);
      break;
    case ELMK_FN_RAD_SUNSHADE:
      // nrad elai | tlai_z fsun_z | solad[2] solai[2] | fabd_sun_z fabd_sha_z fabi_sun_z fabi_sha_z | parsun_z parsha_z laisun_z
      // laisha_z | laisun laisha
      rad::canopy_sunshade_fractions((int)a[0], a[1], FlatRow(a + 2), FlatRow(a + 3), FlatRow(a + 4), FlatRow(a + 6), FlatRow(a + 8),
                                     FlatRow(a + 9), FlatRow(a + 10), FlatRow(a + 11), FlatRow(a + 12), FlatRow(a + 13),
                                     FlatRow(a + 14), FlatRow(a + 15), a[16], a[17]);
      break;
    // ---- canopy_temperature (phys_cantemp.h, namespace tmp); layer rows are [nlevsno + nlevgrnd] = 20, soil
    //      property rows [nlevgrnd] = 15 ----
    case ELMK_FN_TMP_OLD_GROUND_TEMP:   // t_h2osfc | t_soisno[20] | t_h2osfc_bef | tssbef[20]
      tmp::old_ground_temp(a[0], FlatRow(a + 1), a[21], FlatRow(a + 22));
      break;
    case ELMK_FN_TMP_GROUND_TEMP: {   // snl frac_sno_eff frac_h2osfc t_h2osfc | t_soisno[20] | t_grnd
      const int snl = (int)a[0];
      a[24] = tmp::ground_temp(snl, a[1], a[2], a[3], a[4 + NLEVSNO - snl], a[4 + NLEVSNO]);
    } break;
    case ELMK_FN_TMP_CALC_SOILALPHA:
      // frac_sno frac_h2osfc | h2osoi_liq[20] h2osoi_ice[20] dz[20] t_soisno[20] | watsat[15] sucsat[15] bsw[15] watdry[15]
      // watopt[15] | qred hr soilalpha
      tmp::calc_soilalpha(a[0], a[1], a[2 + NLEVSNO], a[22 + NLEVSNO], a[42 + NLEVSNO], a[62 + NLEVSNO], a[82], a[97], a[112],
                          a[157], a[158], a[159]);
      break;
    case ELMK_FN_TMP_CALC_SOILBETA:   // frac_sno frac_h2osfc | watsat[15] watfc[15] | h2osoi_liq[20] h2osoi_ice[20] dz[20] | soilbeta
      a[92] = tmp::calc_soilbeta(a[0], a[1], a[17], a[32 + NLEVSNO], a[52 + NLEVSNO], a[72 + NLEVSNO]);
      break;
    case ELMK_FN_TMP_HUMIDITIES: {
      // snl forc_q forc_pbot t_h2osfc t_grnd frac_sno frac_sno_eff frac_h2osfc qred hr | t_soisno[20] | qg_snow qg_soil qg
      // qg_h2osfc dqgdT
      const int snl = (int)a[0];
      tmp::humidities(snl, a[1], a[2], a[3], a[5], a[6], a[7], a[9], a[10 + NLEVSNO - snl], a[10 + NLEVSNO], a[30], a[31], a[32],
                      a[33], a[34]);
    } break;
    case ELMK_FN_TMP_GROUND_PROPERTIES: {
      // snl frac_sno forc_th forc_q elai esai htop | displar(vtype) z0mr(vtype) | h2osoi_liq[20] h2osoi_ice[20] | emg emv htvp
      // z0mg z0hg z0qg z0mv z0hv z0qv thv z0m displa
      const int top = NLEVSNO - (int)a[0];
      tmp::ground_properties(a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9 + top], a[29 + top], a[49], a[50], a[51], a[52],
                             a[53], a[54], a[55], a[56], a[57], a[58], a[59], a[60]);
    } break;
    case ELMK_FN_TMP_FORCING_HEIGHT:   // veg_active frac_veg_nosno z0m z0mg forc_t displa | hgt_u hgt_t hgt_q thm
      tmp::forcing_height(a[0] != 0.0, (int)a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]);
      break;
    case ELMK_FN_TMP_INIT_ENERGY_FLUXES:
      tmp::init_energy_fluxes(a[0], a[1], a[2], a[3], a[4], a[5]);
      break;
    // ---- bareground_fluxes (phys_bareground.h, namespace bgf): nothing happens on a column with exposed vegetation,
    //      except that compute_flux zeroes cgrnd / cgrnds / cgrndl (bareground_fluxes_impl.hh:104-108) ----
    case ELMK_FN_BGF_INITIALIZE_FLUX:
      // frac_veg_nosno forc_u forc_v forc_q forc_th forc_hgt_u_patch thm thv t_grnd qg z0mg | dlrad ulrad zldis displa dth dqh
      // obu ur um
      if ((int)a[0] == 0)
        bgf::initialize_flux(a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15],
                             a[16], a[17], a[18], a[19]);
      break;
    case ELMK_FN_BGF_STABILITY_ITERATION:
      // frac_veg_nosno hgt_t hgt_u hgt_q z0mg zldis displa dth dqh ur forc_q forc_th thv | z0hg z0qg obu um temp1 temp2 temp12m
      // temp22m ustar
      if ((int)a[0] == 0)
        bgf::stability_iteration(a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15],
                                 a[16], a[17], a[18], a[19], a[20], a[21]);
      break;
    case ELMK_FN_BGF_COMPUTE_FLUX: {
      // frac_veg_nosno snl forc_rho soilbeta dqgdT htvp t_h2osfc qg_snow qg_soil qg_h2osfc | t_soisno[20] | forc_pbot dth dqh
      // temp1 temp2 temp12m temp22m ustar forc_q thm | cgrnds cgrndl cgrnd eflx_sh_grnd eflx_sh_tot eflx_sh_snow eflx_sh_soil
      // eflx_sh_h2osfc qflx_evap_soi qflx_evap_tot qflx_ev_snow qflx_ev_soil qflx_ev_h2osfc t_ref2m q_ref2m rh_ref2m
      a[40] = 0.0; a[41] = 0.0; a[42] = 0.0;
      if ((int)a[0] == 0) {
        const int snl = (int)a[1];
        const bgf::Fluxes f = bgf::compute_flux(a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10 + NLEVSNO - snl],
                                                a[10 + NLEVSNO], a[30], a[31], a[32], a[33], a[34], a[35], a[36], a[37], a[38], a[39]);
        a[40] = f.cgrnds; a[41] = f.cgrndl; a[42] = f.cgrnd; a[43] = f.eflx_sh_grnd; a[44] = f.eflx_sh_grnd;
        a[45] = f.eflx_sh_snow; a[46] = f.eflx_sh_soil; a[47] = f.eflx_sh_h2osfc; a[48] = f.qflx_evap_soi; a[49] = f.qflx_evap_soi;
        a[50] = f.qflx_ev_snow; a[51] = f.qflx_ev_soil; a[52] = f.qflx_ev_h2osfc; a[53] = f.t_ref2m; a[54] = f.q_ref2m;
        a[55] = f.rh_ref2m;
      }
    } break;
    // ---- canopy_fluxes (phys_canflux.h): the three functions are canflux_begin, the loop over canflux_iterate and
    //      canflux_end of the re-packed kernels, run here on a one-column view of the argument array: the fields of
    //      the state the device code reads or writes point at the argument slots (np == 1), fields that are not
    //      arguments of the function at zeros / a sink ----
    case ELMK_FN_CF_INITIALIZE_FLUX: {
      // snl veg frac_sno hgt_u thm thv max_dayl dayl altmax altmax_last | t_soisno[20] ice[20] liq[20] dz[20] rootfr[15]
      // | tc_stress | sucsat[15] watsat[15] bsw[15] | smpso smpsc elai esai emv emg qg t_grnd forc_t pbot lwrad u v q th z0mg
      // || btran displa z0mv z0hv z0qv rootr[15] eff_porosity[15] dayl_factor air bir cir el qsatl qsatldT taf qaf um ur
      // obu zldis delq t_veg
      double zero[NLEVTOT + 1] = {}, sink[4] = {};
      int iv[3] = {(int)a[0], (int)a[1], 0};
      Cols S = {};
      S.np = 1; S.npi = 1; S.ncols = 1;
      S.snl = &iv[0]; S.frac_veg_nosno = &iv[1]; S.nrad = &iv[2];
      S.frac_sno = a + 2; S.forc_hgt_u_patch = a + 3; S.forc_hgt_t_patch = a + 3; S.forc_hgt_q_patch = a + 3;
      S.thm = a + 4; S.thv = a + 5; S.t_soisno = a + 10; S.h2osoi_ice = a + 30; S.h2osoi_liq = a + 50; S.dz = a + 70;
      S.rootfr = a + 90; S.sucsat = a + 106; S.watsat = a + 121; S.bsw = a + 136; S.elai = a + 153; S.esai = a + 154;
      S.emv = a + 155; S.emg = a + 156; S.qg = a + 157; S.t_grnd = a + 158; S.forc_tbot = a + 159; S.forc_pbot = a + 160;
      S.forc_lwrad = a + 161; S.forc_u = a + 162; S.forc_v = a + 163; S.forc_qbot = a + 164; S.forc_thbot = a + 165;
      S.z0mg = a + 166; S.btran = a + 167; S.displa = a + 168; S.z0mv = a + 169; S.z0hv = a + 170; S.z0qv = a + 171;
      S.rootr = a + 172; S.eff_porosity = a + 187; S.t_veg = a + 216;
      S.cgrnd = sink; S.cgrnds = sink + 1; S.cgrndl = sink + 2;
      S.fwet = zero; S.fdry = zero; S.laisun = zero; S.laisha = zero; S.snow_depth = zero; S.soilbeta = zero;
      S.frac_h2osfc = zero; S.t_h2osfc = zero; S.sabv = zero; S.htop = zero; S.t10 = zero; S.h2ocan = zero;
      S.vcmaxcintsha = zero; S.vcmaxcintsun = zero; S.parsha_z = zero; S.parsun_z = zero; S.laisha_z = zero;
      S.laisun_z = zero; S.qflx_tran_veg = zero; S.qflx_evap_veg = zero; S.eflx_sh_veg = zero;
      PsnPft P = {};
      P.tc_stress = a[105]; P.smpso = a[151]; P.smpsc = a[152];
      StepArgs A{0.0, a[7], a[6]};
      CanopyIter I = {};
      if (canflux_begin(S, 0, A, P, 0, I)) {
        a[167] = I.btran; a[202] = I.dayl_factor; a[203] = I.air; a[204] = I.bir; a[205] = I.cir; a[206] = I.el;
        a[207] = I.qsatl; a[208] = I.qsatldT; a[209] = I.taf; a[210] = I.qaf; a[211] = I.um; a[212] = I.ur;
        a[213] = I.obu; a[214] = I.zldis; a[215] = I.delq;
      }
    } break;
    case ELMK_FN_CF_STABILITY_ITERATION: {
      // vtype | dtime snl veg frac_sno hgt_u hgt_t hgt_q fwet fdry laisun laisha forc_rho snow_depth soilbeta frac_h2osfc
      // t_h2osfc sabv h2ocan htop | t_soisno[20] | air bir cir ur zldis displa elai esai t_grnd pbot forc_q forc_th z0mg
      // z0mv z0hv z0qv thm thv qg | psn_pft[27] | nrad t10 tlai_z vcmaxcintsha vcmaxcintsun parsha_z parsun_z laisha_z
      // laisun_z forc_pco2 forc_po2 dayl_factor || btran qflx_tran_veg qflx_evap_veg eflx_sh_veg wtg wtl0 wta0 wtal el
      // qsatl qsatldT taf qaf um dth dqh obu temp1 temp2 temp12m temp22m tlbef delq dt_veg t_veg wtgq wtalq wtlq0 wtaq0
      if ((int)a[3] == 0) break;   // (no exposed vegetation: the reference's function does nothing)
      const int vtype = (int)a[0], snl = (int)a[2];
      PsnPft P;
      {
        double* q = reinterpret_cast<double*>(&P);
        for (int k = 0; k < 27; ++k) q[k] = a[59 + k];
      }
      CanopyIter I = {};
      I.veg = (int)a[3]; I.dtime = a[1]; I.fsno = a[4]; I.hgt_u = a[5]; I.hgt_t = a[6]; I.hgt_q = a[7]; I.fwet = a[8];
      I.fdry = a[9]; I.laisun = a[10]; I.laisha = a[11]; I.forc_rho = a[12]; I.snow_depth = a[13]; I.soilbeta = a[14];
      I.fsfc = a[15]; I.t_sfc = a[16]; I.sabv = a[17]; I.h2ocan0 = a[18]; I.htop = a[19];
      I.t_snotop = a[20 + NLEVSNO - snl]; I.t_soil1 = a[20 + NLEVSNO];
      I.air = a[40]; I.bir = a[41]; I.cir = a[42]; I.ur = a[43]; I.zldis = a[44]; I.displa = a[45]; I.elai = a[46];
      I.esai = a[47]; I.tg = a[48]; I.pbot = a[49]; I.forc_q = a[50]; I.forc_th = a[51]; I.z0mg = a[52]; I.z0mv = a[53];
      I.thm = a[56]; I.thv = a[57]; I.qg = a[58];
      I.nrad = (int)a[86]; I.t10 = a[87]; I.vcsha = a[89]; I.vcsun = a[90]; I.parsha = a[91]; I.parsun = a[92];
      I.laisha_z = a[93]; I.laisun_z = a[94]; I.forc_pco2 = a[95]; I.forc_po2 = a[96]; I.dayl_factor = a[97];
      I.btran = a[98]; I.qflx_tran_veg = a[99]; I.qflx_evap_veg = a[100]; I.eflx_sh_veg = a[101];
      I.el = a[106]; I.qsatl = a[107]; I.qsatldT = a[108]; I.taf = a[109]; I.qaf = a[110]; I.um = a[111]; I.obu = a[114];
      I.delq = a[120]; I.t_veg = a[122];
      I.dth = I.thm - I.taf; I.dqh = I.forc_q - I.qaf;   // (both are assigned inside the loop before their first use)
      I.soybean = (vtype == PFT_SOYBEAN || vtype == PFT_SOYBEAN_IRRIG) ? 1 : 0;
      I.lw_grnd = (I.fsno * pow4(I.t_snotop) + (1.0 - I.fsno - I.fsfc) * pow4(I.t_soil1) + I.fsfc * pow4(I.t_sfc));
      const PsnColumn PC = psn_column(P, I.t10, I.pbot, I.thm, I.forc_po2, I.dayl_factor);
#pragma unroll 1
      while (!canflux_iterate(P, PC, I)) {
      }
      const double t12 = canflux_temp12m(I);
      a[98] = I.btran; a[99] = I.qflx_tran_veg; a[100] = I.qflx_evap_veg; a[101] = I.eflx_sh_veg; a[102] = I.wtg;
      a[103] = I.wtl0; a[104] = I.wta0; a[105] = I.wtal; a[106] = I.el; a[107] = I.qsatl; a[108] = I.qsatldT;
      a[109] = I.taf; a[110] = I.qaf; a[111] = I.um; a[112] = I.dth; a[113] = I.dqh; a[114] = I.obu; a[115] = I.p_temp1;
      a[116] = I.p_temp2; a[117] = t12; a[118] = t12; a[119] = I.tlbef; a[120] = I.delq; a[121] = I.dt_veg;
      a[122] = I.t_veg; a[123] = I.wtgq; a[124] = I.wtalq; a[125] = I.wtlq0; a[126] = I.wtaq0;
    } break;
    case ELMK_FN_CF_COMPUTE_FLUX: {
      // dtime snl veg frac_sno | t_soisno[20] | frac_h2osfc t_h2osfc sabv qg_snow qg_soil qg_h2osfc dqgdT htvp wtg wtl0 wta0
      // wtal air bir cir qsatl qsatldT dth dqh temp1 temp2 temp12m temp22m tlbef delq dt_veg t_veg t_grnd pbot qflx_tran_veg
      // qflx_evap_veg eflx_sh_veg forc_q forc_rho thm emv emg forc_lwrad wtgq wtalq wtlq0 wtaq0 || h2ocan eflx_sh_grnd
      // eflx_sh_snow eflx_sh_soil eflx_sh_h2osfc qflx_evap_soi qflx_ev_snow qflx_ev_soil qflx_ev_h2osfc dlrad ulrad cgrnds
      // cgrndl cgrnd t_ref2m q_ref2m rh_ref2m
      a[77] = 0.0; a[78] = 0.0; a[79] = 0.0;
      if ((int)a[2] == 0) break;
      const int snl = (int)a[1];
      double sink[8] = {};
      int isink[1] = {0};
      Cols S = {};
      S.np = 1; S.npi = 1; S.ncols = 1;
      S.btran = sink; S.t_veg = sink + 1; S.qflx_tran_veg = sink + 2; S.qflx_evap_veg = sink + 3; S.eflx_sh_veg = sink + 4;
      S.errmask = isink;
      S.htvp = a + 31; S.qg_snow = a + 27; S.qg_soil = a + 28; S.qg_h2osfc = a + 29; S.dqgdT = a + 30;
      S.h2ocan = a + 66; S.eflx_sh_grnd = a + 67; S.eflx_sh_snow = a + 68; S.eflx_sh_soil = a + 69; S.eflx_sh_h2osfc = a + 70;
      S.qflx_evap_soi = a + 71; S.qflx_ev_snow = a + 72; S.qflx_ev_soil = a + 73; S.qflx_ev_h2osfc = a + 74;
      S.dlrad = a + 75; S.ulrad = a + 76; S.cgrnds = a + 77; S.cgrndl = a + 78; S.cgrnd = a + 79; S.t_ref2m = a + 80;
      S.q_ref2m = a + 81; S.rh_ref2m = a + 82;
      CanopyIter I = {};
      I.dtime = a[0]; I.fsno = a[3]; I.t_snotop = a[4 + NLEVSNO - snl]; I.t_soil1 = a[4 + NLEVSNO]; I.fsfc = a[24];
      I.t_sfc = a[25]; I.sabv = a[26]; I.wtg = a[32]; I.wtl0 = a[33]; I.wta0 = a[34]; I.wtal = a[35]; I.air = a[36];
      I.bir = a[37]; I.cir = a[38]; I.qsatl = a[39]; I.qsatldT = a[40]; I.dth = a[41]; I.dqh = a[42]; I.p_temp1 = a[43];
      I.p_temp2 = a[44]; I.tlbef = a[47]; I.delq = a[48]; I.dt_veg = a[49]; I.t_veg = a[50]; I.tg = a[51]; I.pbot = a[52];
      I.qflx_tran_veg = a[53]; I.qflx_evap_veg = a[54]; I.eflx_sh_veg = a[55]; I.forc_q = a[56]; I.forc_rho = a[57];
      I.thm = a[58]; I.emv = a[59]; I.emg = a[60]; I.forc_lwrad = a[61]; I.wtgq = a[62]; I.wtalq = a[63]; I.wtlq0 = a[64];
      I.wtaq0 = a[65]; I.h2ocan0 = a[66];
      I.lw_grnd = (I.fsno * pow4(I.t_snotop) + (1.0 - I.fsno - I.fsfc) * pow4(I.t_soil1) + I.fsfc * pow4(I.t_sfc));
      canflux_end_with(S, 0, I, a[45], a[46]);
    } break;
    default: break;
  }
}

// elmk_math_eval: the library's transcendentals at caller-given arguments (parity diagnostic)
__global__ void __launch_bounds__(256) k_math_eval(const int fn, const long long n, const double* __restrict__ x,
                                                   const double* __restrict__ y, double* __restrict__ out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = x[i];
  double r;
  switch (fn) {
    case ELMK_MATH_EXP: r = m_exp(a); break;
    case ELMK_MATH_LOG: r = m_log(a); break;
    case ELMK_MATH_LOG10: r = m_log10(a); break;
    case ELMK_MATH_POW: r = m_pow(a, y[i]); break;
    case ELMK_MATH_ATAN: r = m_atan(a); break;
    case ELMK_MATH_COS: r = m_cos(a); break;
    case ELMK_MATH_TANH: r = m_tanh(a); break;
    case ELMK_MATH_ERF: r = m_erf(a); break;
    case ELMK_MATH_ACOS: r = m_acos(a); break;
    default: r = a / y[i]; break;
  }
  out[i] = r;
}

typedef void (*GroupKernel)(const Cols, const Tables*, const StepArgs);
// kind: how elmk_step issues the launch when all of its groups are requested
enum LaunchKind { kPlain = 0, kSnicarFirst = 1, kCanfluxRepacked = 2 };
struct Launch { uint32_t mask; GroupKernel fn; const char* name; int cols_per_block; int block; int kind; };
#define ELMK_LAUNCH(M, NAME) {(M), k_groups<(M)>, NAME, kBlock, kBlock, kPlain}
#define ELMK_LAUNCH_OCC(M, NAME, MINBLOCKS) {(M), k_groups_occ<(M), MINBLOCKS>, NAME, kBlock, kBlock, kPlain}

// plan "split": one launch per kernel group, one thread per column (the reference's wrapper granularity); the plain
// form of every group, against which the fused plan is tested for bit-identity
const Launch kSplit[] = {
    ELMK_LAUNCH(ELMK_G_FRAC_WET, "frac_wet"),
    ELMK_LAUNCH(ELMK_G_ALBEDO, "albedo_snicar"),
    ELMK_LAUNCH(ELMK_G_CANOPY_HYDROLOGY, "canopy_hydrology"),
    ELMK_LAUNCH(ELMK_G_SURFACE_RADIATION, "surface_radiation"),
    ELMK_LAUNCH(ELMK_G_CANOPY_TEMPERATURE, "canopy_temperature"),
    ELMK_LAUNCH(ELMK_G_BAREGROUND_FLUXES, "bareground_fluxes"),
    ELMK_LAUNCH(ELMK_G_CANOPY_FLUXES, "canopy_fluxes"),
    ELMK_LAUNCH(ELMK_G_SOIL_TEMPERATURE, "soil_temperature"),
    ELMK_LAUNCH(ELMK_G_SNOW_HYDROLOGY, "snow_hydrology"),
    ELMK_LAUNCH(ELMK_G_SURFACE_FLUXES, "surface_fluxes"),
    ELMK_LAUNCH(ELMK_G_CONSERVATION, "conservation"),
};
// plan "fused" (the default): the chain cut where register pressure changes character -
//   radiative transfer (SNICAR kernel, then soil / ground albedo + two-stream) | closed-form hydrology | radiation |
//   temperature | bare-ground fluxes | canopy-flux iteration (re-packed) | banded solve | snow state machine + flux
//   update + diagnostics
constexpr uint32_t M_RAD = ELMK_G_FRAC_WET | ELMK_G_ALBEDO;
constexpr uint32_t M_RAD_REST = ELMK_G_FRAC_WET | G_ALBEDO_REST;
constexpr uint32_t M_SFC = ELMK_G_CANOPY_HYDROLOGY | ELMK_G_SURFACE_RADIATION | ELMK_G_CANOPY_TEMPERATURE |
                           ELMK_G_BAREGROUND_FLUXES;
constexpr uint32_t M_END = ELMK_G_SNOW_HYDROLOGY | ELMK_G_SURFACE_FLUXES | ELMK_G_CONSERVATION;
// (register caps = resident blocks per SM, chosen by A/B runs on B200 at 2M columns: profiles/r2_experiments.md)
const Launch kFused[] = {
    {M_RAD, k_groups_occ<M_RAD_REST, 6>, "fracwet+albedo", kBlock, kBlock, kSnicarFirst},
    // (the closed-form surface groups run faster as four launches than fused - 1.54 vs 1.82 ms per 2M columns: each is
    //  near its own HBM or latency bound at its own register count; profiles/r2_experiments.md)
    ELMK_LAUNCH(ELMK_G_CANOPY_HYDROLOGY, "canopy_hydrology"),
    ELMK_LAUNCH(ELMK_G_SURFACE_RADIATION, "surface_radiation"),
    ELMK_LAUNCH(ELMK_G_CANOPY_TEMPERATURE, "canopy_temperature"),
    {ELMK_G_BAREGROUND_FLUXES, k_bareground_compact, "bareground_fluxes", kBareWindow, kBlock, kPlain},
    {ELMK_G_CANOPY_FLUXES, k_groups<ELMK_G_CANOPY_FLUXES>, "canopy_fluxes", kBlock, kBlock, kCanfluxRepacked},
    ELMK_LAUNCH_OCC(ELMK_G_SOIL_TEMPERATURE, "soil_temperature", 10),
    ELMK_LAUNCH_OCC(M_END, "snow+surface_fluxes+conservation", 6),
};
#ifdef ELMK_DEV_VARIANTS
// development builds only (elmkernels_b200/build.py --dev): register caps of the unsorted launches selected by
// environment variables (ELMK_OCC_RAD / _SFC / _SOIL / _END = blocks per SM) for A/B runs
template <uint32_t M> GroupKernel occ_variant(int minblocks) {
  switch (minblocks) {
    case 2: return k_groups_occ<M, 2>;
    case 3: return k_groups_occ<M, 3>;
    case 4: return k_groups_occ<M, 4>;
    case 5: return k_groups_occ<M, 5>;
    case 6: return k_groups_occ<M, 6>;
    case 8: return k_groups_occ<M, 8>;
    case 10: return k_groups_occ<M, 10>;
    default: return nullptr;
  }
}
#endif

// ------------------------------------------------------------------------------------------------
// layout conversion between the reference's host layout (column outer) and the device layout
// ------------------------------------------------------------------------------------------------
constexpr int kTileCols = 64;
// src: staged host rows [n][nlev]; dst: field base [nlev][np], columns col0..col0+n
template <typename T>
__global__ void __launch_bounds__(256) k_outer_to_inner(const T* __restrict__ src, T* __restrict__ dst, const long long n,
                                                         const int nlev, const long long np, const long long col0)
{
  extern __shared__ unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  const long long c0 = (long long)blockIdx.x * kTileCols;
  const int cols = (int)((n - c0 < kTileCols) ? (n - c0) : kTileCols);
  const int count = cols * nlev;
  const int pitch = nlev | 1;   // odd pitch: conflict-free shared-memory transposition
  for (int e = threadIdx.x; e < count; e += blockDim.x) {
    const int cl = e / nlev, lev = e - cl * nlev;
    tile[cl * pitch + lev] = src[c0 * nlev + e];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kTileCols * nlev; e += blockDim.x) {
    const int lev = e / kTileCols, cl = e - lev * kTileCols;
    if (cl < cols) dst[(long long)lev * np + col0 + c0 + cl] = tile[cl * pitch + lev];
  }
}
template <typename T>
__global__ void __launch_bounds__(256) k_inner_to_outer(const T* __restrict__ src, T* __restrict__ dst, const long long n,
                                                         const int nlev, const long long np, const long long col0)
{
  extern __shared__ unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  const long long c0 = (long long)blockIdx.x * kTileCols;
  const int cols = (int)((n - c0 < kTileCols) ? (n - c0) : kTileCols);
  const int pitch = nlev | 1;
  for (int e = threadIdx.x; e < kTileCols * nlev; e += blockDim.x) {
    const int lev = e / kTileCols, cl = e - lev * kTileCols;
    if (cl < cols) tile[cl * pitch + lev] = src[(long long)lev * np + col0 + c0 + cl];
  }
  __syncthreads();
  const int count = cols * nlev;
  for (int e = threadIdx.x; e < count; e += blockDim.x) {
    const int cl = e / nlev, lev = e - cl * nlev;
    dst[c0 * nlev + e] = tile[cl * pitch + lev];
  }
}

// The padding columns [ncols, np) of the last block: the soil-temperature launch keeps their threads alive (its block
// barriers need whole blocks), so they run the group body on whatever the padding holds.  A physically plausible
// constant column there keeps every value they compute and store finite (all-zero state divides by zero layer
// thicknesses); nothing reads the padding back, elmk_errors scans the valid columns only, and
// tests/test_gpu_padding.py checks that it stays finite and free of error bits.
__global__ void k_init_padding(const Cols S)
{
  const int c = S.ncols + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= S.np) return;
  for (int i = 0; i < NLEVTOT; ++i) {
    C2(dz, i) = (i < NLEVSNO) ? 0.0 : 0.5;
    C2(zsoi, i) = (i < NLEVSNO) ? 0.0 : 0.5 * (i - NLEVSNO) + 0.25;
    C2(t_soisno, i) = (i < NLEVSNO) ? 0.0 : 280.0;
    C2(h2osoi_liq, i) = (i < NLEVSNO) ? 0.0 : 50.0;
    C2(csol, i) = 2.0e6;
  }
  for (int i = 0; i <= NLEVTOT; ++i) C2(zisoi, i) = (i < NLEVSNO) ? 0.0 : 0.5 * (i - NLEVSNO);
  for (int i = 0; i < NLEVGRND; ++i) {
    C2(watsat, i) = 0.4; C2(tkdry, i) = 0.2; C2(tkmg, i) = 1.5; C2(bsw, i) = 5.0; C2(sucsat, i) = 100.0;
  }
  C1(t_grnd) = 280.0; C1(t_h2osfc) = 280.0; C1(emg) = 0.96; C1(htvp) = HVAP; C1(forc_lwrad) = 300.0;
}

template <typename T> __global__ void k_fill(T* __restrict__ p, const long long count, const T v)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// reductions: error words and the balance diagnostics
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_errors(const int* __restrict__ errmask, const int n, unsigned int* any,
                                                long long* first)
{
  unsigned int acc = 0;
  long long lo = 0x7fffffffffffffffLL;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned int w = (unsigned int)errmask[i];
    if (w) {
      acc |= w;
      if (i < lo) lo = i;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc |= __shfl_xor_sync(0xffffffffu, acc, o);
    const long long other = __shfl_xor_sync(0xffffffffu, lo, o);
    lo = (other < lo) ? other : lo;
  }
  if ((threadIdx.x & 31) == 0 && acc) {
    atomicOr(any, acc);
    atomicMin(first, lo);
  }
}

// out[k] = sum, out[8+k] = min, out[16+k] = max of diagnostic k; one block per (diagnostic, slice),
// per-block partials combined by the host in a fixed order (deterministic result)
constexpr int kDiagSlices = 64;
__global__ void __launch_bounds__(256) k_diag(const Cols S, double* __restrict__ partial)
{
  const double* d;
  switch (blockIdx.y) {
    case 0: d = S.dtend_column_h2o; break;
    case 1: d = S.errh2o; break;
    case 2: d = S.errh2osno; break;
    case 3: d = S.dwb; break;
    case 4: d = S.errsol; break;
    case 5: d = S.errlon; break;
    case 6: d = S.errseb; break;
    default: d = S.netrad; break;
  }
  const int n = S.ncols;
  const int per = (n + kDiagSlices - 1) / kDiagSlices;
  const int lo_i = blockIdx.x * per;
  const int hi_i = (lo_i + per < n) ? lo_i + per : n;
  double s = 0.0, lo = INFINITY, hi = -INFINITY;
  for (int i = lo_i + threadIdx.x; i < hi_i; i += blockDim.x) {
    const double v = d[i];
    s += v;
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
  __shared__ double sh[3][256];
  sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = lo; sh[2][threadIdx.x] = hi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] = fmin(sh[1][threadIdx.x], sh[1][threadIdx.x + o]);
      sh[2][threadIdx.x] = fmax(sh[2][threadIdx.x], sh[2][threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double* p = partial + ((size_t)blockIdx.y * kDiagSlices + blockIdx.x) * 3;
    p[0] = sh[0][0]; p[1] = sh[1][0]; p[2] = sh[2][0];
  }
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct Ctx {
  int device = 0;
  int64_t ncols = 0, np = 0;
  cudaStream_t stream = nullptr;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<void*> base;
  Cols cols;
  Tables* d_tables = nullptr;
  double* d_table_data = nullptr;
  char* stage[2] = {nullptr, nullptr};
  int stage_next = 0;
  double* atm[ATM_NVARS] = {};     // raw forcing series [ntimes][np], resident (elmk_atm_series)
  int atm_ntimes[ATM_NVARS] = {};
  double* phen[PHEN_NVARS] = {};   // monthly phenology values [nmonths][np] (elmk_phen_series)
  int phen_nmonths[PHEN_NVARS] = {};
  cudaStream_t s_series = nullptr;  // copy stream of elmk_atm_series_row
  cudaEvent_t ev_series = nullptr, ev_forcing = nullptr;
  bool series_pending = false;
  double* coords = nullptr;        // sin(lat), cos(lat), tan(lat), lon: [4][ncoords] (elmk_set_coordinates)
  double* gas = nullptr;           // [2][np]: CO2, O2 partial pressures (elmk_set_gas_pressures)
  int64_t ncoords = 0;
  double lat0 = 0.0;
  CanfluxQueue cq = {nullptr, nullptr, nullptr, 0};   // CanopyFluxes re-packing scratch (allocated on first use)
  int iterate_blocks = 0;
  bool repack = true;
  void (*snicar_fn)(const Cols, const Tables*, double*, int) = k_snicar<4>;
  double* snicar_scratch = nullptr;   // kSnicarSlice doubles per resident block (allocated on first use)
#ifdef ELMK_BULK_PREFETCH
  const char** pf_rows = nullptr;
  int pf_nrows = 0;
#endif
  int snicar_blocks = 0;
  void (*iterate_fn)(const Cols, const CanfluxQueue) = k_canflux_iterate<kIterBlock, true, 0, kIterSet>;
  int iterate_block = kIterBlock;
  int iterate_smem = kIterBlock * iter_record_doubles(kIterSet) * (int)sizeof(double);   // dynamic shared memory of the iteration kernel: one record per thread
  unsigned int* d_err = nullptr;   // [0] any, then long long first at +8
  double* d_diag = nullptr;
  void* h_pinned = nullptr;        // small pinned buffer for scalar read-backs
  bool tables_set = false;
  int64_t launches = 0;
  // optional per-launch timing
  bool timing = false, timing_detail = false;
  struct Timed { const char* name; uint32_t mask; cudaEvent_t t0, t1; };
  std::vector<Timed> timed;          // recorded, not yet accumulated
  std::vector<cudaEvent_t> ev_pool;  // recycled events
  struct Acc { const char* name; uint32_t mask; double ms; int64_t n; };
  std::vector<Acc> acc;
  const Launch* plan = kFused;
  int plan_len = sizeof(kFused) / sizeof(kFused[0]);
  std::vector<Launch> plan_own;   // a modified copy of the fused plan (development builds)
  std::string last_error;
};
Ctx* ctx(elmk_handle h) { return reinterpret_cast<Ctx*>(h); }

int fail(Ctx* c, cudaError_t e, const char* what) {
  if (c) c->last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return ELMK_ECUDA;
}
#define CU(call)                                        \
  do {                                                  \
    cudaError_t e_ = (call);                            \
    if (e_ != cudaSuccess) return fail(c, e_, #call);   \
  } while (0)

cudaEvent_t take_event(Ctx* c) {
  if (!c->ev_pool.empty()) {
    cudaEvent_t e = c->ev_pool.back();
    c->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
// brackets one kernel launch with events when timing is on
struct TimedScope {
  Ctx* c;
  Ctx::Timed t;
  TimedScope(Ctx* c_, const char* name, uint32_t mask) : c(c_) {
    if (!c->timing) { c = nullptr; return; }
    t = {name, mask, take_event(c), take_event(c)};
    cudaEventRecord(t.t0, c->stream);
  }
  ~TimedScope() {
    if (!c) return;   // timing off, or the scope was closed / discarded by hand
    cudaEventRecord(t.t1, c->stream);
    c->timed.push_back(t);
  }
};
int drain_timing(Ctx* c) {
  if (c->timed.empty()) return ELMK_OK;
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { c->last_error = cudaGetErrorString(e); return ELMK_ECUDA; }
  for (auto& t : c->timed) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t.t0, t.t1);
    bool found = false;
    for (auto& a : c->acc)
      if (a.name == t.name) { a.ms += ms; a.n += 1; found = true; break; }
    if (!found) c->acc.push_back({t.name, t.mask, (double)ms, 1});
    c->ev_pool.push_back(t.t0);
    c->ev_pool.push_back(t.t1);
  }
  c->timed.clear();
  return ELMK_OK;
}

int bind(Ctx* c) {
  cudaError_t e = cudaSetDevice(c->device);
  return e == cudaSuccess ? ELMK_OK : fail(c, e, "cudaSetDevice");
}

template <typename T>
int convert(Ctx* c, bool up, const T* staged_or_field, T* dst, int64_t n, int nlev, int64_t col0) {
  const unsigned grid = (unsigned)((n + kTileCols - 1) / kTileCols);
  const size_t sh = (size_t)kTileCols * (nlev | 1) * sizeof(T);
  if (up) k_outer_to_inner<T><<<grid, 256, sh, c->stream>>>(staged_or_field, dst, n, nlev, c->np, col0);
  else k_inner_to_outer<T><<<grid, 256, sh, c->stream>>>(staged_or_field, dst, n, nlev, c->np, col0);
  c->launches += 1;
  CU(cudaGetLastError());
  return ELMK_OK;
}

// host <-> device movement of one field; asynchronous on the stream when the host buffer is pinned
int move_raw(Ctx* c, char* dev, int dt, int nlev, void* host, int64_t col0, int64_t n, int layout, bool up);
int move_field(Ctx* c, int field, void* host, int64_t col0, int64_t n, int layout, bool up) {
  if (field < 0 || field >= kNumFields || col0 < 0 || n < 0 || col0 + n > c->ncols || !host) {
    c->last_error = "bad field / column range";
    return ELMK_EINVAL;
  }
  if (n == 0) return ELMK_OK;
  return move_raw(c, static_cast<char*>(c->base[field]), kSpecs[field].dtype, kSpecs[field].nlev, host, col0, n, layout, up);
}
// the same for any device array laid out like a field ([nlev][np])
int move_raw(Ctx* c, char* dev, int dt, int nlev, void* host, int64_t col0, int64_t n, int layout, bool up) {
  const size_t es = esize(dt);
  char* hb = static_cast<char*>(host);
  const cudaMemcpyKind kind = up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  if (nlev == 1) {
    if (up) CU(cudaMemcpyAsync(dev + col0 * es, hb, n * es, kind, c->stream));
    else CU(cudaMemcpyAsync(hb, dev + col0 * es, n * es, kind, c->stream));
    return ELMK_OK;
  }
  if (layout == ELMK_COL_INNER) {
    // host[(lev)*n + col]: one strided 2-D copy
    if (up) CU(cudaMemcpy2DAsync(dev + col0 * es, c->np * es, hb, n * es, n * es, nlev, kind, c->stream));
    else CU(cudaMemcpy2DAsync(hb, n * es, dev + col0 * es, c->np * es, n * es, nlev, kind, c->stream));
    return ELMK_OK;
  }
  // reference layout host[(col)*nlev + lev]: stage the rows on the device, transpose there
  const int64_t chunk_cols = std::max<int64_t>(kTileCols, (int64_t)(kStageBytes / (es * nlev)) / kTileCols * kTileCols);
  for (int64_t done = 0; done < n; done += chunk_cols) {
    const int64_t m = std::min(chunk_cols, n - done);
    const int slot = c->stage_next;
    c->stage_next ^= 1;
    char* st = c->stage[slot];
    char* hchunk = hb + (size_t)done * nlev * es;
    if (up) {
      CU(cudaMemcpyAsync(st, hchunk, (size_t)m * nlev * es, kind, c->stream));
      int rc;
      if (dt == ELMK_F64) rc = convert<double>(c, true, (const double*)st, (double*)dev, m, nlev, col0 + done);
      else if (dt == ELMK_I32) rc = convert<int>(c, true, (const int*)st, (int*)dev, m, nlev, col0 + done);
      else rc = convert<unsigned char>(c, true, (const unsigned char*)st, (unsigned char*)dev, m, nlev, col0 + done);
      if (rc) return rc;
    } else {
      int rc;
      if (dt == ELMK_F64) rc = convert<double>(c, false, (const double*)dev, (double*)st, m, nlev, col0 + done);
      else if (dt == ELMK_I32) rc = convert<int>(c, false, (const int*)dev, (int*)st, m, nlev, col0 + done);
      else rc = convert<unsigned char>(c, false, (const unsigned char*)dev, (unsigned char*)st, m, nlev, col0 + done);
      if (rc) return rc;
      CU(cudaMemcpyAsync(hchunk, st, (size_t)m * nlev * es, kind, c->stream));
    }
  }
  return ELMK_OK;
}

// CanopyFluxes as begin / re-packed iterate / end (three launches)
int launch_canflux_repacked(Ctx* c, const StepArgs& A) {
  if (!c->cq.scratch) {
    CU(cudaMalloc(&c->cq.scratch, sizeof(double) * (size_t)kCanfluxDoubles * c->np));
    CU(cudaMalloc(&c->cq.list, sizeof(int) * (size_t)c->np));
    CU(cudaMalloc(&c->cq.counters, sizeof(int) * 4));
    c->cq.np = c->np;
    int per_sm = 0, sms = 0;
    if (c->iterate_smem) CU(cudaFuncSetAttribute(c->iterate_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, c->iterate_smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c->iterate_fn, c->iterate_block, c->iterate_smem));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    c->iterate_blocks = std::max(1, per_sm) * std::max(1, sms);
  }
  const unsigned grid = (unsigned)((c->ncols + kBlock - 1) / kBlock);
  CU(cudaMemsetAsync(c->cq.counters, 0, sizeof(int) * 4, c->stream));
  const bool split = c->timing && c->timing_detail;   // time the three launches apart
  auto mark = [&](Ctx::Timed& t, const char* name) {
    t = {name, 0u, take_event(c), take_event(c)};
    cudaEventRecord(t.t0, c->stream);
  };
  auto done = [&](Ctx::Timed& t) {
    cudaEventRecord(t.t1, c->stream);
    c->timed.push_back(t);
  };
  Ctx::Timed tb{}, ti{}, te{};
  if (split) mark(tb, "canopy_fluxes:begin");
  k_canflux_begin<<<grid, kBlock, 0, c->stream>>>(c->cols, c->d_tables, A, c->cq);
  if (split) done(tb), mark(ti, "canopy_fluxes:iterate");
  const unsigned persistent = (unsigned)std::min<int64_t>(c->iterate_blocks, (c->ncols + kBlock - 1) / kBlock);
  c->iterate_fn<<<persistent, c->iterate_block, c->iterate_smem, c->stream>>>(c->cols, c->cq);
  if (split) done(ti), mark(te, "canopy_fluxes:end");
  k_canflux_end<<<grid, kBlock, 0, c->stream>>>(c->cols, c->cq);
  if (split) done(te);
  c->launches += 3;
  CU(cudaGetLastError());
  return ELMK_OK;
}

// ---- overlapped per-step exchange (see include/elmk_b200.h) ----
struct Exchange {
  Ctx* c = nullptr;
  std::vector<int> in_fields, out_fields;
  std::vector<size_t> in_off, out_off;   // byte offsets of the fields inside one staging slot
  size_t in_bytes = 0, out_bytes = 0;
  char* in_stage[2] = {nullptr, nullptr};
  char* out_stage[2] = {nullptr, nullptr};
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t in_ready[2] = {nullptr, nullptr};     // H2D of the slot finished            (recorded on s_in)
  cudaEvent_t in_consumed[2] = {nullptr, nullptr};  // commit has read the slot            (recorded on the step stream)
  cudaEvent_t out_ready[2] = {nullptr, nullptr};    // snapshot written into the slot      (recorded on the step stream)
  cudaEvent_t out_done[2] = {nullptr, nullptr};     // D2H of the slot finished            (recorded on s_out)
  int64_t posts = 0, commits = 0, fetches = 0, waits = 0, post_waits = 0;
};
Exchange* xch(elmk_exchange x) { return reinterpret_cast<Exchange*>(x); }

size_t field_bytes(const Ctx* c, int field) { return (size_t)c->ncols * kSpecs[field].nlev * esize(kSpecs[field].dtype); }

// staging (reference layout rows) <-> field (column-innermost), on the step stream
int exchange_convert(Ctx* c, int field, char* staged, bool up) {
  const int nlev = kSpecs[field].nlev, dt = kSpecs[field].dtype;
  char* dev = static_cast<char*>(c->base[field]);
  if (nlev == 1) {
    if (up) CU(cudaMemcpyAsync(dev, staged, field_bytes(c, field), cudaMemcpyDeviceToDevice, c->stream));
    else CU(cudaMemcpyAsync(staged, dev, field_bytes(c, field), cudaMemcpyDeviceToDevice, c->stream));
    return ELMK_OK;
  }
  if (dt == ELMK_F64) return up ? convert<double>(c, true, (const double*)staged, (double*)dev, c->ncols, nlev, 0)
                                : convert<double>(c, false, (const double*)dev, (double*)staged, c->ncols, nlev, 0);
  if (dt == ELMK_I32) return up ? convert<int>(c, true, (const int*)staged, (int*)dev, c->ncols, nlev, 0)
                                : convert<int>(c, false, (const int*)dev, (int*)staged, c->ncols, nlev, 0);
  return up ? convert<unsigned char>(c, true, (const unsigned char*)staged, (unsigned char*)dev, c->ncols, nlev, 0)
            : convert<unsigned char>(c, false, (const unsigned char*)dev, (unsigned char*)staged, c->ncols, nlev, 0);
}

} // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int elmk_abi_version(void) { return ELMK_ABI_VERSION; }
const char* elmk_backend(void) { return "cuda-sm100a"; }
int elmk_field_count(void) { return kNumFields; }
int elmk_field_id(const char* name) {
  if (!name) return -1;
  for (int i = 0; i < kNumFields; ++i)
    if (std::strcmp(kSpecs[i].name, name) == 0) return i;
  return -1;
}
int elmk_field_info(int field, const char** name, int* dtype, int* nlev) {
  if (field < 0 || field >= kNumFields) return ELMK_EINVAL;
  if (name) *name = kSpecs[field].name;
  if (dtype) *dtype = kSpecs[field].dtype;
  if (nlev) *nlev = kSpecs[field].nlev;
  return ELMK_OK;
}

int elmk_create(elmk_handle* out, int device, int64_t ncols) {
  // (element offsets lev * np + col are 32-bit in the kernels: 21 levels at most)
  if (!out || ncols <= 0 || ncols > (INT32_MAX - kColAlign) / 21) return ELMK_EINVAL;
  *out = nullptr;
  Ctx* c = new Ctx;
  c->device = device;
  c->ncols = ncols;
  c->np = (ncols + kColAlign - 1) / kColAlign * kColAlign;
  auto bail = [&](int rc) {
    std::fprintf(stderr, "elmk_create: %s\n", c->last_error.c_str());
    elmk_destroy(reinterpret_cast<elmk_handle>(c));
    return rc;
  };
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(fail(c, e, "cudaSetDevice"));
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess)
    return bail(fail(c, e, "cudaStreamCreate"));
  // one arena for all per-column fields; every field starts on a 256-byte boundary and, because np is
  // a multiple of 128 columns, so does every level row of every field
  std::vector<size_t> offset(kNumFields);
  size_t total = 0;
  for (int f = 0; f < kNumFields; ++f) {
    offset[f] = total;
    total += ((size_t)c->np * kSpecs[f].nlev * esize(kSpecs[f].dtype) + 255) / 256 * 256;
  }
  c->arena_bytes = total;
  if ((e = cudaMalloc(&c->arena, total)) != cudaSuccess) return bail(fail(c, e, "cudaMalloc(column state)"));
  if ((e = cudaMemsetAsync(c->arena, 0, total, c->stream)) != cudaSuccess) return bail(fail(c, e, "cudaMemset"));
  c->base.resize(kNumFields);
  for (int f = 0; f < kNumFields; ++f) c->base[f] = c->arena + offset[f];
  c->cols.np = c->np;
  c->cols.ncols = (int)ncols;
  c->cols.npi = (int)c->np;
  c->cols.pco2_in = nullptr;
  c->cols.po2_in = nullptr;
  {
    int f = 0;
#define ELMK_FIELD(name, type, nlev, cls) c->cols.name = static_cast<elmk_##type*>(c->base[f++]);
#include "../../include/elmk_fields.def"
#undef ELMK_FIELD
  }
  for (int s = 0; s < 2; ++s) {
    if ((e = cudaMalloc(&c->stage[s], kStageBytes)) != cudaSuccess) return bail(fail(c, e, "cudaMalloc(stage)"));
  }
  if ((e = cudaMalloc(&c->d_tables, sizeof(Tables))) != cudaSuccess) return bail(fail(c, e, "cudaMalloc(tables)"));
  const size_t table_doubles = (size_t)6 * NBND_SNW * ELMK_MIE_SNW + (size_t)3 * 11 * 31 * 8;
  if ((e = cudaMalloc(&c->d_table_data, table_doubles * sizeof(double))) != cudaSuccess)
    return bail(fail(c, e, "cudaMalloc(table data)"));
  if ((e = cudaMalloc(&c->d_err, 16)) != cudaSuccess) return bail(fail(c, e, "cudaMalloc(err)"));
  if ((e = cudaMalloc(&c->d_diag, sizeof(double) * 8 * kDiagSlices * 3)) != cudaSuccess)
    return bail(fail(c, e, "cudaMalloc(diag)"));
  if ((e = cudaMallocHost(&c->h_pinned, sizeof(double) * 8 * kDiagSlices * 3)) != cudaSuccess)
    return bail(fail(c, e, "cudaMallocHost"));
  if (c->np > c->ncols) {
    k_init_padding<<<1, kColAlign, 0, c->stream>>>(c->cols);
    if ((e = cudaGetLastError()) != cudaSuccess) return bail(fail(c, e, "k_init_padding"));
  }
  if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return bail(fail(c, e, "cudaStreamSynchronize"));
#ifdef ELMK_DEV_VARIANTS
  {
    const char* rp = std::getenv("ELMK_CANFLUX_REPACK");
    if (rp && rp[0] == '0') c->repack = false;
    const char* sn = std::getenv("ELMK_SNICAR_OCC");
    if (sn) { const int o = std::atoi(sn); c->snicar_fn = o == 2 ? k_snicar<2> : o == 3 ? k_snicar<3> : o == 5 ? k_snicar<5> : o == 6 ? k_snicar<6> : o == 8 ? k_snicar<8> : k_snicar<4>; }
    const char* ib = std::getenv("ELMK_ITER");   // e.g. "256l" = 256-thread lock-step blocks, "128" = 128 threads free-running
    if (ib) {
      const int nb = std::atoi(ib);
      const bool ls = std::strchr(ib, 'l') != nullptr;
      c->iterate_block = nb;
      c->iterate_smem = 0;
      const char* mm = std::strchr(ib, 'm');
      const int set = mm ? (mm[1] >= '1' && mm[1] <= '3' ? mm[1] - '0' : 1) : 0;
      if (nb == 128) c->iterate_fn = ls ? k_canflux_iterate<128, true, 0, 0> : k_canflux_iterate<128, false, 0, 0>;
      else if (nb == 256) c->iterate_fn = ls ? k_canflux_iterate<256, true, 0, 0> : k_canflux_iterate<256, false, 0, 0>;
      else if (nb == 512) c->iterate_fn = k_canflux_iterate<512, true, 0, 0>;
      else if (nb == 384 && set) {
        c->iterate_fn = set == 3 ? k_canflux_iterate<384, true, 0, 3> : set == 2 ? k_canflux_iterate<384, true, 0, 2> : k_canflux_iterate<384, true, 0, 1>;
        c->iterate_smem = 384 * iter_record_doubles(set) * (int)sizeof(double);
      }
      else if (nb == 384 && std::strstr(ib, "g64")) c->iterate_fn = k_canflux_iterate<384, true, 64, 0>;
      else if (nb == 384 && std::strstr(ib, "g128")) c->iterate_fn = k_canflux_iterate<384, true, 128, 0>;
      else if (nb == 384 && std::strstr(ib, "g192")) c->iterate_fn = k_canflux_iterate<384, true, 192, 0>;
      else { c->iterate_block = kIterBlock; c->iterate_fn = ls ? k_canflux_iterate<384, true, 0, 0> : k_canflux_iterate<384, false, 0, 0>; }
    }
    // launches of the fused plan found by the set of groups they cover
    auto slot = [&](uint32_t mask) -> Launch* {
      for (Launch& L : c->plan_own) if (L.mask == mask) return &L;
      return nullptr;
    };
    const char* e0 = std::getenv("ELMK_OCC_RAD");
    const char* e2 = std::getenv("ELMK_OCC_SOIL");
    const char* e3 = std::getenv("ELMK_OCC_END");
    const char* e4 = std::getenv("ELMK_OCC_BG");
    const char* sm = std::getenv("ELMK_SFC_MODE");   // how the closed-form surface groups (a3..a6) are cut into launches
    const char* em = std::getenv("ELMK_END_MODE");   // same for snow hydrology + surface fluxes + conservation
    if (e0 || e2 || e3 || e4 || sm || em || std::getenv("ELMK_END_CLASSED") || std::getenv("ELMK_SOIL_CLASSED") || std::getenv("ELMK_HYD_CLASSED")) {
      c->plan_own.assign(kFused, kFused + c->plan_len);
      GroupKernel k;
      if (e0 && (k = occ_variant<M_RAD_REST>(std::atoi(e0)))) slot(M_RAD)->fn = k;
      if (e2 && (k = occ_variant<ELMK_G_SOIL_TEMPERATURE>(std::atoi(e2)))) slot(ELMK_G_SOIL_TEMPERATURE)->fn = k;
      if (e3 && (k = occ_variant<M_END>(std::atoi(e3)))) slot(M_END)->fn = k;
      if (const char* pe = std::getenv("ELMK_END_CLASSED")) {   // snow launch on class-ordered blocks: key, blocks/SM as "16", "26", "18" ...
        const int v = std::atoi(pe);
        slot(M_END)->fn = v == 16 ? k_groups_classed<M_END, 6, 1> : v == 26 ? k_groups_classed<M_END, 6, 2> : v == 18 ? k_groups_classed<M_END, 8, 1> : k_groups_classed<M_END, 5, 1>;
      }
      if (const char* pe = std::getenv("ELMK_SOIL_CLASSED")) {
        const int v = std::atoi(pe);
        slot(ELMK_G_SOIL_TEMPERATURE)->fn = v == 2 ? k_groups_classed<ELMK_G_SOIL_TEMPERATURE, 10, 2> : k_groups_classed<ELMK_G_SOIL_TEMPERATURE, 10, 1>;
      }
      if (const char* pe = std::getenv("ELMK_HYD_CLASSED")) {
        (void)pe;
        slot(ELMK_G_CANOPY_HYDROLOGY)->fn = k_groups_classed<ELMK_G_CANOPY_HYDROLOGY, 4, 3>;
      }
      if (e4 && (k = occ_variant<ELMK_G_BAREGROUND_FLUXES>(std::atoi(e4)))) { slot(ELMK_G_BAREGROUND_FLUXES)->fn = k; slot(ELMK_G_BAREGROUND_FLUXES)->cols_per_block = kBlock; }
      if (sm) {
        constexpr uint32_t M_HRT = ELMK_G_CANOPY_HYDROLOGY | ELMK_G_SURFACE_RADIATION | ELMK_G_CANOPY_TEMPERATURE;
        const int mode = std::atoi(sm);
        std::vector<Launch> cut;
        if (mode == 0) cut = {ELMK_LAUNCH_OCC(M_SFC, "hydrology+radiation+temperature+bareground", 8)};
        else if (mode == 2) cut = {ELMK_LAUNCH_OCC(M_HRT, "hydrology+radiation+temperature", 8), ELMK_LAUNCH(ELMK_G_BAREGROUND_FLUXES, "bareground_fluxes")};
        if (!cut.empty()) {
          const auto first = c->plan_own.begin() + (slot(ELMK_G_CANOPY_HYDROLOGY) - c->plan_own.data());
          c->plan_own.erase(first, first + 4);
          c->plan_own.insert(c->plan_own.begin() + 1, cut.begin(), cut.end());
        }
      }
      if (em) {
        constexpr uint32_t M_FC = ELMK_G_SURFACE_FLUXES | ELMK_G_CONSERVATION;
        const int emode = std::atoi(em);
        std::vector<Launch> end;
        if (emode == 1) end = {ELMK_LAUNCH(ELMK_G_SNOW_HYDROLOGY, "snow_hydrology"), ELMK_LAUNCH(M_FC, "surface_fluxes+conservation")};
        else if (emode == 3) end = {ELMK_LAUNCH_OCC(ELMK_G_SNOW_HYDROLOGY, "snow_hydrology", 6), ELMK_LAUNCH_OCC(M_FC, "surface_fluxes+conservation", 8)};
        if (!end.empty()) {
          c->plan_own.pop_back();
          c->plan_own.insert(c->plan_own.end(), end.begin(), end.end());
        }
      }
      c->plan_len = (int)c->plan_own.size();
    }
  }
#endif
  *out = reinterpret_cast<elmk_handle>(c);
  return ELMK_OK;
}

int elmk_destroy(elmk_handle h) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->arena);
  cudaFree(c->stage[0]);
  cudaFree(c->stage[1]);
  cudaFree(c->d_tables);
  cudaFree(c->d_table_data);
  cudaFree(c->d_err);
  cudaFree(c->d_diag);
  for (double* p : c->atm) cudaFree(p);
  for (double* p : c->phen) cudaFree(p);
  if (c->s_series) { cudaStreamSynchronize(c->s_series); cudaStreamDestroy(c->s_series); }
  if (c->ev_series) cudaEventDestroy(c->ev_series);
  if (c->ev_forcing) cudaEventDestroy(c->ev_forcing);
  cudaFree(c->coords);
  cudaFree(c->snicar_scratch);
  cudaFree(c->cq.scratch);
  cudaFree(c->cq.list);
  cudaFree(c->cq.counters);
  cudaFree(c->gas);
  for (auto& t : c->timed) { cudaEventDestroy(t.t0); cudaEventDestroy(t.t1); }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return ELMK_OK;
}

const char* elmk_last_error(elmk_handle h) { return h ? ctx(h)->last_error.c_str() : "null handle"; }
int64_t elmk_ncols(elmk_handle h) { return h ? ctx(h)->ncols : 0; }

int elmk_set_tables(elmk_handle h, const elmk_tables* t) {
  Ctx* c = ctx(h);
  if (!c || !t) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  if (!(t->ltype == ISTSOIL || t->ltype == ISTCROP) || t->urbpoi || t->lakpoi) {
    c->last_error = "only soil/crop land units without lake or urban points are on the hot path";
    return ELMK_EUNSUPPORTED;
  }
  if (t->vtype < 0 || t->vtype >= ELMK_NUMPFT) return ELMK_EINVAL;
  Tables T;
  std::memset(&T, 0, sizeof(T));
  T.ltype = t->ltype; T.ctype = t->ctype; T.vtype = t->vtype; T.urbpoi = t->urbpoi; T.lakpoi = t->lakpoi;
  T.oldfflag = t->oldfflag; T.dewmx = t->dewmx;
  for (int v = 0; v < ELMK_NUMPFT; ++v) {
    T.z0mr[v] = t->pft[27][v]; T.displar[v] = t->pft[28][v]; T.xl[v] = t->pft[29][v];
    T.rhol[v][0] = t->pft[32][v]; T.rhol[v][1] = t->pft[33][v];
    T.rhos[v][0] = t->pft[34][v]; T.rhos[v][1] = t->pft[35][v];
    T.taul[v][0] = t->pft[36][v]; T.taul[v][1] = t->pft[37][v];
    T.taus[v][0] = t->pft[38][v]; T.taus[v][1] = t->pft[39][v];
    for (int k = 0; k < 26; ++k) T.psn[k][v] = t->pft[k][v];
    T.psn[26][v] = t->pft[26][0];
    T.roota[v] = t->pft[30][v]; T.rootb[v] = t->pft[31][v];
  }
  std::memcpy(T.albsat, t->albsat, sizeof(T.albsat));
  std::memcpy(T.albdry, t->albdry, sizeof(T.albdry));
  for (int s = 0; s < 6; ++s)
    for (int k = 0; k < 3; ++k) std::memcpy(T.aer_band[s][k], t->snicar_band[s * 3 + k], sizeof(double) * NBND_SNW);
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < 3; ++k) std::memcpy(T.bc[s][k], t->snicar_bc[s * 3 + k], sizeof(double) * 10 * NBND_SNW);
  std::memcpy(T.bcenh, t->bcenh, sizeof(T.bcenh));
  const size_t snw_n = (size_t)NBND_SNW * ELMK_MIE_SNW, age_n = (size_t)11 * 31 * 8;
  double* p = c->d_table_data;
  for (int d = 0; d < 2; ++d)
    for (int k = 0; k < 3; ++k) {
      CU(cudaMemcpyAsync(p, t->snicar_snow[d * 3 + k], snw_n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      T.snw[d][k] = p;
      p += snw_n;
    }
  for (int k = 0; k < 3; ++k) {
    CU(cudaMemcpyAsync(p, t->snowage[k], age_n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    T.snowage[k] = p;
    p += age_n;
  }
  CU(cudaMemcpyAsync(c->d_tables, &T, sizeof(T), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));   // T and the caller's arrays may go out of scope
  c->tables_set = true;
  return ELMK_OK;
}

int elmk_upload(elmk_handle h, int field, const void* host, int64_t col0, int64_t n, int layout) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  return move_field(c, field, const_cast<void*>(host), col0, n, layout, true);
}
int elmk_download(elmk_handle h, int field, void* host, int64_t col0, int64_t n, int layout) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  if (int rc = move_field(c, field, host, col0, n, layout, false)) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return ELMK_OK;
}
int elmk_upload_many(elmk_handle h, int nf, const int* fields, const void* const* hosts, int64_t col0, int64_t n,
                     int layout) {
  Ctx* c = ctx(h);
  if (!c || nf < 0 || (nf && (!fields || !hosts))) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  for (int i = 0; i < nf; ++i)
    if (int rc = move_field(c, fields[i], const_cast<void*>(hosts[i]), col0, n, layout, true)) return rc;
  return ELMK_OK;
}
int elmk_download_many(elmk_handle h, int nf, const int* fields, void* const* hosts, int64_t col0, int64_t n,
                       int layout) {
  Ctx* c = ctx(h);
  if (!c || nf < 0 || (nf && (!fields || !hosts))) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  for (int i = 0; i < nf; ++i)
    if (int rc = move_field(c, fields[i], hosts[i], col0, n, layout, false)) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return ELMK_OK;
}

int elmk_exchange_create(elmk_handle h, int n_in, const int* in_fields, int n_out, const int* out_fields,
                         elmk_exchange* out) {
  Ctx* c = ctx(h);
  if (!c || !out || n_in < 0 || n_out < 0 || (n_in && !in_fields) || (n_out && !out_fields)) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  auto* x = new Exchange();
  x->c = c;
  auto layout = [&](int n, const int* f, std::vector<int>& ids, std::vector<size_t>& off, size_t& total) {
    for (int i = 0; i < n; ++i) {
      if (f[i] < 0 || f[i] >= kNumFields) return false;
      ids.push_back(f[i]);
      off.push_back(total);
      total += (field_bytes(c, f[i]) + 255) / 256 * 256;
    }
    return true;
  };
  if (!layout(n_in, in_fields, x->in_fields, x->in_off, x->in_bytes) ||
      !layout(n_out, out_fields, x->out_fields, x->out_off, x->out_bytes)) {
    delete x;
    c->last_error = "elmk_exchange_create: bad field id";
    return ELMK_EINVAL;
  }
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  ok(cudaStreamCreateWithFlags(&x->s_in, cudaStreamNonBlocking));
  ok(cudaStreamCreateWithFlags(&x->s_out, cudaStreamNonBlocking));
  for (int s = 0; s < 2; ++s) {
    if (x->in_bytes) ok(cudaMalloc(&x->in_stage[s], x->in_bytes));
    if (x->out_bytes) ok(cudaMalloc(&x->out_stage[s], x->out_bytes));
    ok(cudaEventCreateWithFlags(&x->in_ready[s], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&x->in_consumed[s], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&x->out_ready[s], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&x->out_done[s], cudaEventDisableTiming));
  }
  if (e != cudaSuccess) {
    elmk_exchange_destroy(reinterpret_cast<elmk_exchange>(x));
    return fail(c, e, "elmk_exchange_create");
  }
  *out = reinterpret_cast<elmk_exchange>(x);
  return ELMK_OK;
}

int elmk_exchange_destroy(elmk_exchange xh) {
  Exchange* x = xch(xh);
  if (!x) return ELMK_EINVAL;
  cudaSetDevice(x->c->device);
  if (x->s_in) cudaStreamSynchronize(x->s_in);
  if (x->s_out) cudaStreamSynchronize(x->s_out);
  cudaStreamSynchronize(x->c->stream);
  for (int s = 0; s < 2; ++s) {
    cudaFree(x->in_stage[s]);
    cudaFree(x->out_stage[s]);
    if (x->in_ready[s]) cudaEventDestroy(x->in_ready[s]);
    if (x->in_consumed[s]) cudaEventDestroy(x->in_consumed[s]);
    if (x->out_ready[s]) cudaEventDestroy(x->out_ready[s]);
    if (x->out_done[s]) cudaEventDestroy(x->out_done[s]);
  }
  if (x->s_in) cudaStreamDestroy(x->s_in);
  if (x->s_out) cudaStreamDestroy(x->s_out);
  delete x;
  return ELMK_OK;
}

int elmk_exchange_post(elmk_exchange xh, const void* const* in_hosts) {
  Exchange* x = xch(xh);
  if (!x || (!in_hosts && !x->in_fields.empty())) return ELMK_EINVAL;
  Ctx* c = x->c;
  if (int rc = bind(c)) return rc;
  if (x->posts - x->commits >= 2) {
    c->last_error = "elmk_exchange_post: two posts already await elmk_exchange_commit";
    return ELMK_EINVAL;
  }
  const int slot = (int)(x->posts & 1);
  if (x->posts >= 2) CU(cudaStreamWaitEvent(x->s_in, x->in_consumed[slot], 0));   // the slot's previous contents were committed
  for (size_t i = 0; i < x->in_fields.size(); ++i)
    CU(cudaMemcpyAsync(x->in_stage[slot] + x->in_off[i], in_hosts[i], field_bytes(c, x->in_fields[i]),
                       cudaMemcpyHostToDevice, x->s_in));
  CU(cudaEventRecord(x->in_ready[slot], x->s_in));
  x->posts += 1;
  return ELMK_OK;
}

int elmk_exchange_commit(elmk_exchange xh) {
  Exchange* x = xch(xh);
  if (!x) return ELMK_EINVAL;
  Ctx* c = x->c;
  if (int rc = bind(c)) return rc;
  if (x->commits >= x->posts) {
    c->last_error = "elmk_exchange_commit without a matching elmk_exchange_post";
    return ELMK_EINVAL;
  }
  const int slot = (int)(x->commits & 1);
  CU(cudaStreamWaitEvent(c->stream, x->in_ready[slot], 0));
  for (size_t i = 0; i < x->in_fields.size(); ++i)
    if (int rc = exchange_convert(c, x->in_fields[i], x->in_stage[slot] + x->in_off[i], true)) return rc;
  CU(cudaEventRecord(x->in_consumed[slot], c->stream));
  x->commits += 1;
  return ELMK_OK;
}

int elmk_exchange_fetch(elmk_exchange xh, void* const* out_hosts) {
  Exchange* x = xch(xh);
  if (!x || (!out_hosts && !x->out_fields.empty())) return ELMK_EINVAL;
  Ctx* c = x->c;
  if (int rc = bind(c)) return rc;
  if (x->fetches - x->waits >= 2) {
    c->last_error = "elmk_exchange_fetch: two fetches already await elmk_exchange_wait";
    return ELMK_EINVAL;
  }
  const int slot = (int)(x->fetches & 1);
  if (x->fetches >= 2) CU(cudaStreamWaitEvent(c->stream, x->out_done[slot], 0));   // the slot's previous snapshot left the device
  for (size_t i = 0; i < x->out_fields.size(); ++i)
    if (int rc = exchange_convert(c, x->out_fields[i], x->out_stage[slot] + x->out_off[i], false)) return rc;
  CU(cudaEventRecord(x->out_ready[slot], c->stream));
  CU(cudaStreamWaitEvent(x->s_out, x->out_ready[slot], 0));
  for (size_t i = 0; i < x->out_fields.size(); ++i)
    CU(cudaMemcpyAsync(out_hosts[i], x->out_stage[slot] + x->out_off[i], field_bytes(c, x->out_fields[i]),
                       cudaMemcpyDeviceToHost, x->s_out));
  CU(cudaEventRecord(x->out_done[slot], x->s_out));
  x->fetches += 1;
  return ELMK_OK;
}

int elmk_exchange_wait(elmk_exchange xh) {
  Exchange* x = xch(xh);
  if (!x) return ELMK_EINVAL;
  Ctx* c = x->c;
  if (int rc = bind(c)) return rc;
  if (x->waits >= x->fetches) return ELMK_OK;   // nothing outstanding
  const int slot = (int)(x->waits & 1);
  CU(cudaEventSynchronize(x->out_done[slot]));
  x->waits += 1;
  return ELMK_OK;
}

int elmk_exchange_post_wait(elmk_exchange xh) {
  Exchange* x = xch(xh);
  if (!x) return ELMK_EINVAL;
  Ctx* c = x->c;
  if (int rc = bind(c)) return rc;
  if (x->post_waits < x->posts - 2) x->post_waits = x->posts - 2;   // older slots were reused: their copies are done
  if (x->post_waits >= x->posts) return ELMK_OK;                     // nothing outstanding
  const int slot = (int)(x->post_waits & 1);
  CU(cudaEventSynchronize(x->in_ready[slot]));
  x->post_waits += 1;
  return ELMK_OK;
}

int elmk_fn_call(int device, int fn, double* args, int64_t nargs) {
  if (fn < 0 || fn >= ELMK_FN_COUNT || !args || nargs != kFnSlots[fn]) return ELMK_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return ELMK_ECUDA;
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(double) * nargs) != cudaSuccess) return ELMK_ECUDA;
  cudaError_t e = cudaMemcpy(d, args, sizeof(double) * nargs, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_fn_call<<<1, 32>>>(fn, d);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(args, d, sizeof(double) * nargs, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? ELMK_OK : ELMK_ECUDA;
}

int elmk_math_eval(elmk_handle h, int fn, int64_t n, const double* x, const double* y, double* out) {
  Ctx* c = ctx(h);
  if (!c || !x || !out || fn < 0 || fn >= ELMK_MATH_COUNT || n < 0) return ELMK_EINVAL;
  if ((fn == ELMK_MATH_POW || fn == ELMK_MATH_DIV) && !y) return ELMK_EINVAL;
  if (n == 0) return ELMK_OK;
  if (int rc = bind(c)) return rc;
  double* d = nullptr;
  CU(cudaMalloc(&d, sizeof(double) * 3 * (size_t)n));
  cudaError_t e = cudaMemcpyAsync(d, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && y) e = cudaMemcpyAsync(d + n, y, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    k_math_eval<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(fn, n, d, d + n, d + 2 * n);
    c->launches += 1;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  return e == cudaSuccess ? ELMK_OK : fail(c, e, "elmk_math_eval");
}

// ---- per-step producers of the forcing and phenology inputs ----
namespace {
int set_series(Ctx* c, double** slot, int* count, const double* host, int n) {
  if (!host || n < 2) {
    c->last_error = "series need a host pointer and at least two time levels";
    return ELMK_EINVAL;
  }
  if (*slot && *count != n) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(*slot));
    *slot = nullptr;
  }
  if (!*slot) {
    CU(cudaMalloc(slot, sizeof(double) * (size_t)n * c->np));
    CU(cudaMemsetAsync(*slot, 0, sizeof(double) * (size_t)n * c->np, c->stream));
  }
  *count = n;
  // host rows of ncols values -> device rows of np values
  CU(cudaMemcpy2DAsync(*slot, sizeof(double) * c->np, host, sizeof(double) * c->ncols, sizeof(double) * c->ncols, n,
                       cudaMemcpyHostToDevice, c->stream));
  return ELMK_OK;
}
} // namespace

// ---- one-time cold-start initialisation of every column ----
int elmk_init_columns(elmk_handle h, const double* pct_sand, const double* pct_clay, const double* organic,
                      double organic_max, const double* snow_depth) {
  Ctx* c = ctx(h);
  if (!c || !pct_sand || !pct_clay || !organic || !snow_depth) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  if (!c->tables_set) return ELMK_ENOTABLES;
  double* buf = nullptr;
  const size_t per = (size_t)NLEVGRND * c->np;
  CU(cudaMalloc(&buf, sizeof(double) * (3 * per + c->np)));
  int rc = ELMK_OK;
  const double* hosts[3] = {pct_sand, pct_clay, organic};
  for (int k = 0; k < 3 && rc == ELMK_OK; ++k)
    rc = move_raw(c, reinterpret_cast<char*>(buf + k * per), ELMK_F64, NLEVGRND, const_cast<double*>(hosts[k]), 0, c->ncols,
                  ELMK_COL_OUTER, true);
  if (rc == ELMK_OK) {
    cudaError_t e = cudaMemcpyAsync(buf + 3 * per, snow_depth, sizeof(double) * c->ncols, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) rc = fail(c, e, "cudaMemcpyAsync(snow_depth)");
  }
  if (rc == ELMK_OK) {
    const InitInputs X{buf, buf + per, buf + 2 * per, buf + 3 * per, c->np, organic_max};
    const unsigned grid = (unsigned)((c->ncols + kBlock - 1) / kBlock);
    TimedScope ts(c, "init_columns", 0u);
    k_init_columns<<<grid, kBlock, 0, c->stream>>>(c->cols, c->d_tables, X);
    c->launches += 1;
  }
  cudaStreamSynchronize(c->stream);
  cudaFree(buf);
  if (rc == ELMK_OK) CU(cudaGetLastError());
  return rc;
}

int elmk_set_gas_pressures(elmk_handle h, const double* forc_pco2, const double* forc_po2) {
  Ctx* c = ctx(h);
  if (!c || ((forc_pco2 == nullptr) != (forc_po2 == nullptr))) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  if (!forc_pco2) {   // back to the constants of the reference's wrapper
    c->cols.pco2_in = nullptr;
    c->cols.po2_in = nullptr;
    return ELMK_OK;
  }
  if (!c->gas) {
    CU(cudaMalloc(&c->gas, sizeof(double) * 2 * (size_t)c->np));
    CU(cudaMemsetAsync(c->gas, 0, sizeof(double) * 2 * (size_t)c->np, c->stream));
  }
  CU(cudaMemcpyAsync(c->gas, forc_pco2, sizeof(double) * (size_t)c->ncols, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->gas + c->np, forc_po2, sizeof(double) * (size_t)c->ncols, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->cols.pco2_in = c->gas;
  c->cols.po2_in = c->gas + c->np;
  return ELMK_OK;
}

int elmk_set_coordinates(elmk_handle h, const double* lat_r, const double* lon_r, int64_t n) {
  Ctx* c = ctx(h);
  if (!c || !lat_r || !lon_r || !(n == 1 || n == c->ncols)) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  std::vector<double> host((size_t)4 * n);
  for (int64_t i = 0; i < n; ++i) {
    host[i] = std::sin(lat_r[i]);
    host[n + i] = std::cos(lat_r[i]);
    host[2 * n + i] = std::tan(solar::ensure_tan_defined(lat_r[i]));
    host[3 * n + i] = lon_r[i];
  }
  if (c->coords) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(c->coords));
    c->coords = nullptr;
  }
  CU(cudaMalloc(&c->coords, sizeof(double) * host.size()));
  CU(cudaMemcpyAsync(c->coords, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->ncoords = n;
  c->lat0 = lat_r[0];
  return ELMK_OK;
}

int elmk_solar_step(elmk_handle h, double dtime, double decday, int doy1, double* dayl, double* max_dayl) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  if (!c->coords) {
    c->last_error = "elmk_solar_step before elmk_set_coordinates";
    return ELMK_EINVAL;
  }
  // the per-step scalars of average_cosz (incident_shortwave.cc:108-116), on the host like the reference's
  const double declin = solar::declination((int)decday);
  SolarStep G;
  G.sin_lat = c->coords; G.cos_lat = c->coords + c->ncoords; G.tan_lat = c->coords + 2 * c->ncoords;
  G.lon = c->coords + 3 * c->ncoords;
  G.per_column = c->ncoords > 1;
  G.dtrad = dtime * solar::TWO_PI / 86400.0;
  G.frac2pi = (decday - std::floor(decday)) * solar::TWO_PI;
  G.sin_decl = std::sin(declin);
  G.cos_decl = std::cos(declin);
  G.tan_decl = std::tan(solar::ensure_tan_defined(declin));
  const unsigned grid = (unsigned)((c->ncols + kBlock - 1) / kBlock);
  {
    TimedScope ts(c, "coszen", 0u);
    k_coszen<<<grid, kBlock, 0, c->stream>>>(c->cols, G);
  }
  c->launches += 1;
  CU(cudaGetLastError());
  if (dayl) *dayl = solar::daylength(c->lat0, solar::declination(doy1));
  if (max_dayl) *max_dayl = solar::max_daylength(c->lat0);
  return ELMK_OK;
}

int elmk_atm_series(elmk_handle h, int var, const double* host, int ntimes) {
  Ctx* c = ctx(h);
  if (!c || var < 0 || var >= ATM_NVARS) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  return set_series(c, &c->atm[var], &c->atm_ntimes[var], host, ntimes);
}

int elmk_atm_series_row(elmk_handle h, int var, int t, const double* host) {
  Ctx* c = ctx(h);
  if (!c || var < 0 || var >= ATM_NVARS || !host) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  if (!c->atm[var] || t < 0 || t >= c->atm_ntimes[var]) {
    c->last_error = "elmk_atm_series_row: no such series / time level";
    return ELMK_EINVAL;
  }
  if (!c->s_series) {
    CU(cudaStreamCreateWithFlags(&c->s_series, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_series, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_forcing, cudaEventDisableTiming));
    CU(cudaEventRecord(c->ev_forcing, c->stream));
  }
  // the row may still be read by the forcing kernel issued last: the copy waits for it, then runs beside the step
  CU(cudaStreamWaitEvent(c->s_series, c->ev_forcing, 0));
  CU(cudaMemcpyAsync(c->atm[var] + (size_t)t * c->np, host, sizeof(double) * c->ncols, cudaMemcpyHostToDevice, c->s_series));
  CU(cudaEventRecord(c->ev_series, c->s_series));
  c->series_pending = true;
  return ELMK_OK;
}

int elmk_canflux_pass_histogram(elmk_handle h, int64_t hist[42]) {
  Ctx* c = ctx(h);
  if (!c || !hist) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  if (!c->cq.scratch) {
    c->last_error = "elmk_canflux_pass_histogram: the re-packed CanopyFluxes launch has not run on this handle";
    return ELMK_EINVAL;
  }
  unsigned long long* d = nullptr;
  CU(cudaMalloc(&d, sizeof(unsigned long long) * 42));
  cudaError_t e = cudaMemsetAsync(d, 0, sizeof(unsigned long long) * 42, c->stream);
  if (e == cudaSuccess) {
    k_pass_histogram<<<148 * 8, 256, 0, c->stream>>>(c->cols, c->cq.scratch + (size_t)kCanfluxItlefRow * c->np, d);
    c->launches += 1;
    e = cudaGetLastError();
  }
  unsigned long long host[42];
  if (e == cudaSuccess) e = cudaMemcpyAsync(host, d, sizeof(host), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(c, e, "elmk_canflux_pass_histogram");
  for (int i = 0; i < 42; ++i) hist[i] = (int64_t)host[i];
  return ELMK_OK;
}

int elmk_atm_forcing(elmk_handle h, int t_idx, double wt1, double wt2, int qbot_is_rh) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  AtmSeries A;
  A.stride = c->np;
  for (int v = 0; v < ATM_NVARS; ++v) {
    if (!c->atm[v] || t_idx < 0 || t_idx + 1 >= c->atm_ntimes[v]) {
      c->last_error = "elmk_atm_forcing: series missing or t_idx + 1 outside it";
      return ELMK_EINVAL;
    }
    A.v[v] = c->atm[v];
  }
  const unsigned grid = (unsigned)((c->ncols + kBlock - 1) / kBlock);
  {
    if (c->series_pending) {   // rows posted by elmk_atm_series_row have to be on the device first
      CU(cudaStreamWaitEvent(c->stream, c->ev_series, 0));
      c->series_pending = false;
    }
    TimedScope ts(c, "atm_forcing", 0u);
    k_atm_forcing<<<grid, kBlock, 0, c->stream>>>(c->cols, A, t_idx, wt1, wt2, qbot_is_rh);
  }
  if (c->s_series) CU(cudaEventRecord(c->ev_forcing, c->stream));
  c->launches += 1;
  CU(cudaGetLastError());
  return ELMK_OK;
}

int elmk_phen_series(elmk_handle h, int var, const double* host, int nmonths) {
  Ctx* c = ctx(h);
  if (!c || var < 0 || var >= PHEN_NVARS) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  return set_series(c, &c->phen[var], &c->phen_nmonths[var], host, nmonths);
}

int elmk_phenology(elmk_handle h, int start_idx, double wt1, double wt2) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  PhenSeries P;
  P.stride = c->np;
  for (int v = 0; v < PHEN_NVARS; ++v) {
    if (!c->phen[v] || start_idx < 0 || start_idx + 1 >= c->phen_nmonths[v]) {
      c->last_error = "elmk_phenology: series missing or start_idx + 1 outside it";
      return ELMK_EINVAL;
    }
    P.v[v] = c->phen[v];
  }
  const unsigned grid = (unsigned)((c->ncols + kBlock - 1) / kBlock);
  {
    TimedScope ts(c, "phenology", 0u);
    k_phenology<<<grid, kBlock, 0, c->stream>>>(c->cols, P, start_idx, wt1, wt2);
  }
  c->launches += 1;
  CU(cudaGetLastError());
  return ELMK_OK;
}

int elmk_fill(elmk_handle h, int field, double value) {
  Ctx* c = ctx(h);
  if (!c || field < 0 || field >= kNumFields) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  const long long count = (long long)c->np * kSpecs[field].nlev;
  const unsigned grid = (unsigned)std::min<long long>((count + 255) / 256, 148 * 16);
  if (kSpecs[field].dtype == ELMK_F64) k_fill<double><<<grid, 256, 0, c->stream>>>((double*)c->base[field], count, value);
  else if (kSpecs[field].dtype == ELMK_I32) k_fill<int><<<grid, 256, 0, c->stream>>>((int*)c->base[field], count, (int)value);
  else k_fill<unsigned char><<<grid, 256, 0, c->stream>>>((unsigned char*)c->base[field], count, (unsigned char)(value != 0.0));
  c->launches += 1;
  CU(cudaGetLastError());
  return ELMK_OK;
}

int elmk_init_timestep(elmk_handle h, int reset_forc_hgt) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  if (!c->tables_set) return ELMK_ENOTABLES;
  const unsigned grid = (unsigned)((c->ncols + kBlock - 1) / kBlock);
  {
    TimedScope ts(c, "init_timestep", 0u);
    k_init_timestep<<<grid, kBlock, 0, c->stream>>>(c->cols, c->d_tables, reset_forc_hgt);
  }
  c->launches += 1;
  CU(cudaGetLastError());
  return ELMK_OK;
}

int elmk_step(elmk_handle h, double dtime, double dayl, double max_dayl, uint32_t mask) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  if (!c->tables_set) {
    c->last_error = "elmk_step before elmk_set_tables";
    return ELMK_ENOTABLES;
  }
  mask &= ELMK_G_ALL;
  const StepArgs A{dtime, dayl, max_dayl};
  auto grid_for = [&](const Launch& L) { return (unsigned)((c->ncols + L.cols_per_block - 1) / L.cols_per_block); };
  // cover the requested groups, in chain order, with the launches of the plan; a launch whose group
  // set is only partly requested falls back to one launch per requested group
  for (int i = 0; i < c->plan_len; ++i) {
    const Launch& L = (c->plan == kFused && !c->plan_own.empty()) ? c->plan_own[i] : c->plan[i];
    const uint32_t want = L.mask & mask;
    if (!want) continue;
    if (want == L.mask && L.kind == kCanfluxRepacked && c->repack) {
      TimedScope ts(c, L.name, L.mask);
      if (int rc = launch_canflux_repacked(c, A)) return rc;
    } else if (want == L.mask && L.kind == kSnicarFirst) {
      // the ten band solves of every sunlit snow column, then the rest of the albedo group for every column
      TimedScope ts(c, L.name, L.mask);
      const bool split = c->timing && c->timing_detail;
      Ctx::Timed t1{}, t2{};
      if (split) { t1 = {"albedo:snicar", 0u, take_event(c), take_event(c)}; cudaEventRecord(t1.t0, c->stream); }
      const int nwindows = (int)((c->ncols + kSnicarWindow - 1) / kSnicarWindow);
      if (!c->snicar_scratch) {
        int per_sm = 0, sms = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c->snicar_fn, kSnicarBlock, 0));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        c->snicar_blocks = std::min(nwindows, std::max(1, per_sm) * std::max(1, sms));
        CU(cudaMalloc(&c->snicar_scratch, sizeof(double) * kSnicarSlice * (size_t)c->snicar_blocks));
      }
      c->snicar_fn<<<(unsigned)c->snicar_blocks, kSnicarBlock, 0, c->stream>>>(c->cols, c->d_tables, c->snicar_scratch, nwindows);
      if (split) { cudaEventRecord(t1.t1, c->stream); c->timed.push_back(t1);
                   t2 = {"albedo:rest", 0u, take_event(c), take_event(c)}; cudaEventRecord(t2.t0, c->stream); }
      L.fn<<<grid_for(L), L.block, 0, c->stream>>>(c->cols, c->d_tables, A);
      if (split) { cudaEventRecord(t2.t1, c->stream); c->timed.push_back(t2); }
      c->launches += 2;
    } else if (want == L.mask) {
      TimedScope ts(c, L.name, L.mask);
#ifdef ELMK_BULK_PREFETCH
      StepArgs Ap = A;
      if (L.mask == ELMK_G_SOIL_TEMPERATURE && !std::getenv("ELMK_NO_PREFETCH")) {
        if (!c->pf_rows) {
          std::vector<const char*> rows;
          const Cols& S = c->cols;
#define PF(field, lev0, nlev) for (int l = 0; l < (nlev); ++l) rows.push_back((const char*)(S.field + (long long)((lev0) + l) * S.np));
          PF(t_soisno, 0, 20) PF(zsoi, 0, 20) PF(h2osoi_liq, 0, 20) PF(h2osoi_ice, 0, 20) PF(dz, 0, 20) PF(zisoi, 0, 21)
          PF(watsat, 0, 15) PF(tkdry, 0, 15) PF(tkmg, 0, 15) PF(csol, 5, 15) PF(sucsat, 0, 15) PF(bsw, 0, 15) PF(sabg_lyr, 0, 6)
          PF(frac_sno, 0, 1) PF(frac_sno_eff, 0, 1) PF(frac_h2osfc, 0, 1) PF(h2osfc, 0, 1) PF(h2osno, 0, 1) PF(t_h2osfc, 0, 1)
          PF(dlrad, 0, 1) PF(emg, 0, 1) PF(forc_lwrad, 0, 1) PF(htvp, 0, 1) PF(sabg_soil, 0, 1) PF(sabg_snow, 0, 1)
          PF(eflx_sh_soil, 0, 1) PF(qflx_ev_soil, 0, 1) PF(eflx_sh_h2osfc, 0, 1) PF(qflx_ev_h2osfc, 0, 1) PF(eflx_sh_snow, 0, 1)
          PF(qflx_ev_snow, 0, 1) PF(cgrnd, 0, 1) PF(t_grnd, 0, 1) PF(int_snow, 0, 1) PF(snow_depth, 0, 1)
#undef PF
          c->pf_nrows = (int)rows.size();
          CU(cudaMalloc(&c->pf_rows, sizeof(char*) * rows.size()));
          CU(cudaMemcpy(c->pf_rows, rows.data(), sizeof(char*) * rows.size(), cudaMemcpyHostToDevice));
        }
        Ap.pf_rows = c->pf_rows;
        Ap.pf_nrows = c->pf_nrows;
      }
      L.fn<<<grid_for(L), L.block, 0, c->stream>>>(c->cols, c->d_tables, Ap);
#else
      L.fn<<<grid_for(L), L.block, 0, c->stream>>>(c->cols, c->d_tables, A);
#endif
      c->launches += 1;
    } else {
      for (const Launch& G : kSplit) {
        if (G.mask & want) {
          TimedScope ts(c, G.name, G.mask);
          G.fn<<<grid_for(G), G.block, 0, c->stream>>>(c->cols, c->d_tables, A);
          c->launches += 1;
        }
      }
    }
    if (c->timed.size() > 4096) {
      if (int rc = drain_timing(c)) return rc;
    }
  }
  CU(cudaGetLastError());
  return ELMK_OK;
}

int elmk_set_plan(elmk_handle h, int plan) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (plan == ELMK_PLAN_FUSED) {
    c->plan = kFused;
    c->plan_len = sizeof(kFused) / sizeof(kFused[0]);
  } else if (plan == ELMK_PLAN_SPLIT) {
    c->plan = kSplit;
    c->plan_len = sizeof(kSplit) / sizeof(kSplit[0]);
  } else {
    c->last_error = "elmk_set_plan: unknown plan";
    return ELMK_EINVAL;
  }
  return ELMK_OK;
}

int elmk_sync(elmk_handle h) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return ELMK_OK;
}
int64_t elmk_launch_count(elmk_handle h) { return h ? ctx(h)->launches : 0; }

int elmk_timing_enable(elmk_handle h, int on) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  if (int rc = drain_timing(c)) return rc;
  c->acc.clear();
  c->timing = on != 0;
  c->timing_detail = on > 1;   // 2: also the sub-launches of the composite launches (SNICAR / rest, begin / iterate / end)
  return ELMK_OK;
}
int elmk_timing_read(elmk_handle h, int max, const char** names, double* total_ms, int64_t* launches,
                     uint32_t* group_masks) {
  Ctx* c = ctx(h);
  if (!c || max < 0) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  if (int rc = drain_timing(c)) return rc;
  const int n = std::min<int>(max, (int)c->acc.size());
  for (int i = 0; i < n; ++i) {
    if (names) names[i] = c->acc[i].name;
    if (total_ms) total_ms[i] = c->acc[i].ms;
    if (launches) launches[i] = c->acc[i].n;
    if (group_masks) group_masks[i] = c->acc[i].mask;
  }
  return (int)c->acc.size();
}

int elmk_errors(elmk_handle h, uint32_t* any, int64_t* first_col) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  struct { unsigned int any; unsigned int pad; long long first; } init = {0u, 0u, 0x7fffffffffffffffLL}, *res;
  res = static_cast<decltype(res)>(c->h_pinned);
  *res = init;
  CU(cudaMemcpyAsync(c->d_err, res, 16, cudaMemcpyHostToDevice, c->stream));
  const unsigned grid = (unsigned)std::min<int64_t>((c->ncols + 255) / 256, 148 * 8);
  k_errors<<<grid, 256, 0, c->stream>>>(c->cols.errmask, (int)c->ncols, c->d_err, reinterpret_cast<long long*>(c->d_err + 2));
  c->launches += 1;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(res, c->d_err, 16, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (any) *any = res->any;
  if (first_col) *first_col = res->any ? res->first : -1;
  return ELMK_OK;
}
int elmk_clear_errors(elmk_handle h) {
  Ctx* c = ctx(h);
  if (!c) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  CU(cudaMemsetAsync(c->cols.errmask, 0, (size_t)c->np * sizeof(int), c->stream));
  return ELMK_OK;
}
const char* elmk_error_text(uint32_t bit) {
  switch (bit) {
    case ELMK_ERR_CANOPY_LAYER: return "ELM ERROR: multi-layer canopy not implemented";
    case ELMK_ERR_SNICAR_RADIUS: return "ELM ERROR: SNICAR snow grain radius out of bounds";
    case ELMK_ERR_SNICAR_NEGABS: return "ELM ERROR: SNICAR negative absoption";
    case ELMK_ERR_SNICAR_ENERGY: return "ELM ERROR: SNICAR Energy conservation error";
    case ELMK_ERR_SNICAR_ALBEDO: return "ELM ERROR: SNICAR Albedo > 1.0";
    case ELMK_ERR_SABG_LAYERS: return "surface_radiation: absorbed solar radiation of the snow layers does not sum to sabg_snow";
    case ELMK_ERR_FORC_HEIGHT: return "canopy_fluxes: forcing height is below the canopy displacement height";
    case ELMK_ERR_QUADRATIC: return "ELM ERROR: quadratic solution a == 0";
    case ELMK_ERR_BRENT_BRACKET: return "ELM ERROR: root must be bracketed for brent";
    case ELMK_ERR_NEG_STOMATAL: return "ELM ERROR: Negative stomatal conductance";
    case ELMK_ERR_SNOWAGE_DR: return "ELM ERROR: SnowAge dr_fresh < 0.0.";
    case ELMK_ERR_DIVIDE_RADIUS: return "ELM ERROR: snow radius out of bounds in snow::divide_layers.";
    default: return "unknown error bit";
  }
}

int elmk_diag_reduce(elmk_handle h, double out[24]) {
  Ctx* c = ctx(h);
  if (!c || !out) return ELMK_EINVAL;
  if (int rc = bind(c)) return rc;
  k_diag<<<dim3(kDiagSlices, 8), 256, 0, c->stream>>>(c->cols, c->d_diag);
  c->launches += 1;
  CU(cudaGetLastError());
  double* hp = static_cast<double*>(c->h_pinned);
  CU(cudaMemcpyAsync(hp, c->d_diag, sizeof(double) * 8 * kDiagSlices * 3, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 8; ++k) {
    double s = 0.0, lo = INFINITY, hi = -INFINITY;
    for (int b = 0; b < kDiagSlices; ++b) {
      const double* p = hp + ((size_t)k * kDiagSlices + b) * 3;
      s += p[0];
      lo = std::min(lo, p[1]);
      hi = std::max(hi, p[2]);
    }
    out[k] = s; out[8 + k] = lo; out[16 + k] = hi;
  }
  return ELMK_OK;
}

int elmk_device_ptr(elmk_handle h, int field, void** ptr, int64_t* level_stride) {
  Ctx* c = ctx(h);
  if (!c || field < 0 || field >= kNumFields) return ELMK_EINVAL;
  if (ptr) *ptr = c->base[field];
  if (level_stride) *level_stride = c->np;
  return ELMK_OK;
}

int elmk_stream(elmk_handle h, void** stream) {
  Ctx* c = ctx(h);
  if (!c || !stream) return ELMK_EINVAL;
  *stream = static_cast<void*>(c->stream);
  return ELMK_OK;
}

} // extern "C"
