/* elmk_b200.h - C ABI of the B200-native ELM column-timestep library (libelmk_b200.so).
 *
 * This is the drop-in boundary below the reference's kernel-group wrappers
 * (reference driver/kokkos/<group>_kokkos.{hh,cc}, called from ELMInterface::advance,
 * driver/kokkos/elm_kokkos_interface.cc:269-322).  Plain pointers and sizes only.
 *
 * Every entry point returns 0 on success or a negative ELMK_E* code; the text of the
 * last failure on a handle is available from elmk_last_error().
 *
 * Threading: a handle owns one CUDA stream on one device.  Calls on one handle are not
 * thread-safe; different handles are independent (reference: single host thread,
 * SURVEY.md section 8(b)).
 */
#ifndef ELMK_B200_H_
#define ELMK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ELMK_ABI_VERSION 6

/* ---- dimensions (reference src/data/elm_constants.h:84-98) ---- */
#define ELMK_NLEVSNO 5
#define ELMK_NLEVGRND 15
#define ELMK_NLEVTOT 20
#define ELMK_NUMRAD 2
#define ELMK_NUMPFT 17
#define ELMK_NUMRAD_SNW 5
#define ELMK_SNO_NBR_AER 8
#define ELMK_MIE_SNW 1471
#define ELMK_BC_NCLRDS 10
#define ELMK_BCINT_ICERDS 8
#define ELMK_SNOWAGE_T 11
#define ELMK_SNOWAGE_TGRD 31
#define ELMK_SNOWAGE_RHOS 8
#define ELMK_NSOILCOL 20
#define ELMK_NPFT_TABLES 40

/* ---- error codes ---- */
#define ELMK_OK 0
#define ELMK_EINVAL (-1)   /* bad argument                                           */
#define ELMK_ECUDA (-2)    /* CUDA runtime failure (text in elmk_last_error)          */
#define ELMK_ENOTABLES (-3)/* elmk_step before elmk_set_tables                        */
#define ELMK_EUNSUPPORTED (-4) /* land-unit configuration outside the hot path's scope */
#define ELMK_ECOLUMN (-5)  /* a column raised one of the reference's throw/assert sites */

/* ---- element types of per-column fields ---- */
#define ELMK_F64 0
#define ELMK_I32 1
#define ELMK_U8 2

/* ---- host memory layouts for upload/download ---- */
#define ELMK_COL_OUTER 0 /* host[(col)*nlev + lev]  : the reference's layout (array.hh:176-183) */
#define ELMK_COL_INNER 1 /* host[(lev)*n + col]     : the device layout                         */

/* ---- kernel groups = the reference's wrapper calls, in chain order
 *      (elm_kokkos_interface.cc:289-318); a step runs the selected groups in this order ---- */
#define ELMK_G_FRAC_WET (1u << 0)           /* kokkos_frac_wet            canopy_hydrology_kokkos.cc:98  */
#define ELMK_G_ALBEDO (1u << 1)             /* kokkos_albedo_snicar       albedo_kokkos.cc:10            */
#define ELMK_G_CANOPY_HYDROLOGY (1u << 2)   /* kokkos_canopy_hydrology    canopy_hydrology_kokkos.cc:7   */
#define ELMK_G_SURFACE_RADIATION (1u << 3)  /* kokkos_surface_radiation   surface_radiation_kokkos.cc:7  */
#define ELMK_G_CANOPY_TEMPERATURE (1u << 4) /* kokkos_canopy_temperature  canopy_temperature_kokkos.cc:6 */
#define ELMK_G_BAREGROUND_FLUXES (1u << 5)  /* kokkos_bareground_fluxes   bareground_fluxes_kokkos.cc:7  */
#define ELMK_G_CANOPY_FLUXES (1u << 6)      /* kokkos_canopy_fluxes       canopy_fluxes_kokkos.cc:6      */
#define ELMK_G_SOIL_TEMPERATURE (1u << 7)   /* kokkos_soil_temperature    soil_temperature_kokkos.cc:6   */
#define ELMK_G_SNOW_HYDROLOGY (1u << 8)     /* kokkos_snow_hydrology      snow_hydrology_kokkos.cc:23    */
#define ELMK_G_SURFACE_FLUXES (1u << 9)     /* kokkos_surface_fluxes      surface_fluxes_kokkos.cc:6     */
#define ELMK_G_CONSERVATION (1u << 10)      /* kokkos_evaluate_conservation conserved_quantity_kokkos.cc:8 */
#define ELMK_G_ALL 0x7FFu
#define ELMK_NGROUPS 11

/* ---- per-column error bits (one per reference throw/assert site) ---- */
#define ELMK_ERR_CANOPY_LAYER (1u << 0)     /* surface_albedo_impl.hh:270               */
#define ELMK_ERR_SNICAR_RADIUS (1u << 1)    /* snow_snicar_impl.hh:76                   */
#define ELMK_ERR_SNICAR_NEGABS (1u << 2)    /* snow_snicar_impl.hh:618                  */
#define ELMK_ERR_SNICAR_ENERGY (1u << 3)    /* snow_snicar_impl.hh:658                  */
#define ELMK_ERR_SNICAR_ALBEDO (1u << 4)    /* snow_snicar_impl.hh:664                  */
#define ELMK_ERR_SABG_LAYERS (1u << 5)      /* surface_radiation_impl.hh:173 (assert)   */
#define ELMK_ERR_FORC_HEIGHT (1u << 6)      /* canopy_fluxes_impl.hh:178 (assert)       */
#define ELMK_ERR_QUADRATIC (1u << 7)        /* photosynthesis_impl.hh:289               */
#define ELMK_ERR_BRENT_BRACKET (1u << 8)    /* photosynthesis_impl.hh:439               */
#define ELMK_ERR_NEG_STOMATAL (1u << 9)     /* photosynthesis_impl.hh:232               */
#define ELMK_ERR_SNOWAGE_DR (1u << 10)      /* snow_hydrology_impl.hh:146               */
#define ELMK_ERR_DIVIDE_RADIUS (1u << 11)   /* snow_hydrology_impl.hh:1022,1100,1176,1253 */

typedef struct elmk_ctx* elmk_handle;

/* Global (not per-column) inputs.  All pointers are host pointers, copied by elmk_set_tables.
 *
 * pft[k] follows the member order of the reference's PFTData (src/data/pft_data.h:38-77):
 *   0 fnr 1 act25 2 kcha 3 koha 4 cpha 5 vcmaxha 6 jmaxha 7 tpuha 8 lmrha 9 vcmaxhd 10 jmaxhd
 *   11 tpuhd 12 lmrhd 13 lmrse 14 qe 15 theta_cj 16 bbbopt 17 mbbopt 18 c3psn 19 slatop
 *   20 leafcn 21 flnr 22 fnitr 23 dleaf 24 smpso 25 smpsc 26 tc_stress 27 z0mr 28 displar 29 xl
 *   30 roota_par 31 rootb_par 32 rholvis 33 rholnir 34 rhosvis 35 rhosnir 36 taulvis 37 taulnir
 *   38 tausvis 39 tausnir
 * each ELMK_NUMPFT doubles, except tc_stress which is one double (pft_data_impl.hh:54).
 *
 * snicar_band[k] ([ELMK_NUMRAD_SNW] each) in the member order of SnicarData
 * (src/data/snicar_data.h:40-57): ss_alb/asm_prm/ext_cff_mss for oc1, oc2, dst1..dst4.
 * snicar_snow[k] ([ELMK_NUMRAD_SNW][ELMK_MIE_SNW] each): ss_alb_snw_drc, asm_prm_snw_drc,
 * ext_cff_mss_snw_drc, ss_alb_snw_dfs, asm_prm_snw_dfs, ext_cff_mss_snw_dfs (:58-63).
 * snicar_bc[k] ([ELMK_BC_NCLRDS][ELMK_NUMRAD_SNW] each): ss_alb_bc1, asm_prm_bc1,
 * ext_cff_mss_bc1, ss_alb_bc2, asm_prm_bc2, ext_cff_mss_bc2 (:64-69).
 * bcenh [ELMK_BCINT_ICERDS][ELMK_BC_NCLRDS][ELMK_NUMRAD_SNW] (:70).
 * snowage[k] ([ELMK_SNOWAGE_T][ELMK_SNOWAGE_TGRD][ELMK_SNOWAGE_RHOS] each): tau, kappa, drdt0 (:80-82).
 * albsat/albdry [ELMK_NSOILCOL][ELMK_NUMRAD] (elm_state_impl.hh:106-107).
 */
typedef struct elmk_tables {
  int32_t ltype, ctype, vtype, urbpoi, lakpoi; /* LandType, src/data/land_data.h:36-44 */
  int32_t oldfflag;                            /* elm_state.h:224 */
  double dewmx;                                /* elm_state.h:223 */
  const double* pft[ELMK_NPFT_TABLES];
  const double* albsat;
  const double* albdry;
  const double* snicar_band[18];
  const double* snicar_snow[6];
  const double* snicar_bc[6];
  const double* bcenh;
  const double* snowage[3];
} elmk_tables;

/* ---- introspection of the per-column field table (include/elmk_fields.def) ---- */
int elmk_abi_version(void);
const char* elmk_backend(void); /* "cuda-sm100a" for the product library */
int elmk_field_count(void);
int elmk_field_id(const char* name);                                   /* -1 if unknown */
int elmk_field_info(int field, const char** name, int* dtype, int* nlev);

/* ---- lifetime: replaces ELMInterface::ELMInterface(ncols) / ELMState allocation
 *      (elm_kokkos_interface.cc:38-56, elm_state_impl.hh:369-403) ---- */
int elmk_create(elmk_handle* out, int device, int64_t ncols);
int elmk_destroy(elmk_handle h);
const char* elmk_last_error(elmk_handle h);
int64_t elmk_ncols(elmk_handle h);

/* ---- tables: replaces the table part of initialize_kokkos_elm
 *      (initialize_elm_kokkos.cc:267-366) ---- */
int elmk_set_tables(elmk_handle h, const elmk_tables* t);

/* ---- state movement: replaces Kokkos::deep_copy of ELMState arrays
 *      (elm_kokkos_interface.cc:145-255, copyPrimaryVars :324-347).
 *      host holds n columns [col0, col0+n) of one field, element type per elmk_field_info ---- */
int elmk_upload(elmk_handle h, int field, const void* host, int64_t col0, int64_t n, int layout);
int elmk_download(elmk_handle h, int field, void* host, int64_t col0, int64_t n, int layout);
/* set every element of a field (Utils::assign, helper_functions.hh:18-19) */
int elmk_fill(elmk_handle h, int field, double value);

/* several fields in one call: hosts[i] is the host buffer of fields[i] (same col0, n, layout).
 * Used for the per-step forcing refresh and the per-step result read-back. */
int elmk_upload_many(elmk_handle h, int nfields, const int* fields, const void* const* hosts,
                     int64_t col0, int64_t n, int layout);
int elmk_download_many(elmk_handle h, int nfields, const int* fields, void* const* hosts,
                       int64_t col0, int64_t n, int layout);

/* ---- overlapped per-step exchange.  The reference's driver refreshes the forcing of ELMState and copies the
 *      primary variables out once per step from the single host thread that also runs the kernels
 *      (elm_kokkos_interface.cc:269-288 and copyPrimaryVars :324-347): movement and compute are serial.  An
 *      exchange owns two copy streams and double-buffered device staging for a fixed set of input fields and a
 *      fixed set of output fields (reference host layout, ELMK_COL_OUTER, all columns of the handle), so that the
 *      host->device copy of step k+1's inputs and the device->host copy of step k's outputs run while step k+1
 *      computes:
 *        elmk_exchange_post(x, in_hosts)    start the asynchronous copy of the next inputs into staging;
 *        elmk_exchange_commit(x)            the inputs posted last become the fields' contents, ordered after
 *                                           everything issued on the handle's stream so far;
 *        elmk_exchange_fetch(x, out_hosts)  snapshot the output fields (ordered after everything issued so far)
 *                                           and start their asynchronous copy to out_hosts;
 *        elmk_exchange_wait(x)              block until the oldest unfinished fetch has arrived on the host.
 *        elmk_exchange_post_wait(x)         block until the oldest post whose copy may still be reading its host
 *                                           buffers has left the host (the buffers may then be refilled).
 *      At most two posts and two fetches can be in flight.  Host buffers should be pinned
 *      (cudaHostAlloc / cudaHostRegister); pageable memory works but does not overlap.
 *      LIFETIME OF HOST BUFFERS: with pinned memory every copy of this interface is asynchronous.  The buffers given
 *      to elmk_exchange_post must stay untouched until elmk_exchange_post_wait has returned for that post (or
 *      elmk_sync after the matching commit); those given to elmk_upload / elmk_upload_many / elmk_atm_series /
 *      elmk_phen_series until elmk_sync (or any blocking call: elmk_download*, elmk_errors, elmk_diag_reduce);
 *      elmk_set_tables and elmk_init_columns block before they return.  Pageable host memory is consumed before
 *      the call returns (the driver stages it). ---- */
typedef struct elmk_exchange_s* elmk_exchange;
int elmk_exchange_create(elmk_handle h, int n_in, const int* in_fields, int n_out, const int* out_fields,
                         elmk_exchange* out);
int elmk_exchange_destroy(elmk_exchange x);
int elmk_exchange_post(elmk_exchange x, const void* const* in_hosts);
int elmk_exchange_commit(elmk_exchange x);
int elmk_exchange_fetch(elmk_exchange x, void* const* out_hosts);
int elmk_exchange_wait(elmk_exchange x);
int elmk_exchange_post_wait(elmk_exchange x);

/* ---- one-time cold-start initialisation of every column: the per-column lambda of initialize_kokkos_elm
 *      (driver/kokkos/initialize_elm_kokkos.cc:374-431) - psn_pft from vtype, init_topo_slope, init_melt_factor,
 *      init_micro_sigma, init_snow_layers, init_soil_hydraulics, init_vegrootfr, init_soil_temp, init_snow_state,
 *      init_soilh2o_state.  Needs elmk_set_tables and, in the state, vtype, the raw topo_slope, topo_std and the soil
 *      part of dz / zsoi / zisoi (elm_kokkos_interface.cc:137-184).  pct_sand, pct_clay, organic: host[col * 15 + lev]
 *      (ViewD2(ncols, nlevgrnd)); snow_depth: host[col], the depth the snow layers are built from. ---- */
int elmk_init_columns(elmk_handle h, const double* pct_sand, const double* pct_clay, const double* organic,
                      double organic_max, const double* snow_depth);

/* ---- solar geometry: the first lines of kokkos_init_timestep (init_timestep_kokkos.cc:27-35).
 *      elmk_set_coordinates: latitude and longitude [rad] of the columns; n == 1 is the reference's single site
 *        (S.lat_r, S.lon_r apply to every column), n == ncols gives every column its own.  The time-invariant
 *        sin / cos / tan of the latitudes are evaluated here, once, on the host.
 *      elmk_solar_step: incident_shortwave::average_cosz(lat, lon, dtime, decday) of every column into the coszen
 *        field (device), with decday = Utils::decimal_doy(current) + 1; *dayl = ELM::daylength(lat,
 *        declination_angle_sin(doy1)) with doy1 = current.doy + 1 and *max_dayl = ELM::max_daylength(lat), both for
 *        the first coordinate, to be handed to elmk_step. ---- */
int elmk_set_coordinates(elmk_handle h, const double* lat_r, const double* lon_r, int64_t n);
int elmk_solar_step(elmk_handle h, double dtime, double decday, int doy1, double* dayl, double* max_dayl);

/* ---- CO2 and O2 partial pressures [Pa] of every column for the photosynthesis of group a7.  The reference's wrapper
 *      derives them from constants (355 ppmv CO2, 0.209 mol/mol O2: canopy_fluxes_kokkos.cc:49-51,
 *      atm_physics_impl.hh derive_forc_pco2 / derive_forc_po2), its library functions take them as arguments
 *      (canopy_fluxes::stability_iteration, canopy_fluxes_impl.hh:187-202) - a host model with its own CO2 supplies
 *      them here: host arrays of ncols values, both or neither; NULL, NULL returns to the constants.
 *      ELMK_EUNSUPPORTED on the reference checker (its wrapper has no such input). ---- */
int elmk_set_gas_pressures(elmk_handle h, const double* forc_pco2, const double* forc_po2);

/* ---- producers of the per-step inputs, on the device (the step before the chain in kokkos_init_timestep,
 *      init_timestep_kokkos.cc:39-47).
 *      elmk_atm_series: one raw forcing series, host[t * ncols + col] - the layout of AtmDataManager::data
 *        (ntimes, ncells), src/data/atm_data.h; it stays resident on the device until replaced
 *        (replaces read_atm_data's deep_copy, atm_forcing_kokkos.cc:15-27).
 *      elmk_atm_forcing: the eight forcing functors of ELM::get_forcing (atm_forcing_kokkos.cc:48-63,
 *        src/physics/atm_physics_impl.hh:40-211) in one pass: forc_tbot/thbot, forc_pbot, forc_qbot, forc_lwrad,
 *        forc_solad/solai (reads the coszen field), forc_rain/snow, forc_u/v, forc_hgt*.  t_idx, wt1, wt2 are what
 *        AtmDataManager::forc_t_idx_check_bounds / forcing_time_weights (atm_data_impl.hh:147-199) return for the
 *        step's centred time; qbot_is_rh selects AtmForcType::RH (series in percent) over AtmForcType::QBOT.
 *      elmk_phen_series / elmk_phenology: monthly LAI, SAI, canopy top and bottom height, host[m * ncols + col]
 *        (PhenologyDataManager::mlai..mhbot), and ComputePhenology (src/physics/phenology_physics_impl.hh:20-69):
 *        tlai, tsai, htop, hbot, elai, esai, frac_veg_nosno_alb from snow_depth, frac_sno, vtype; start_idx, wt1,
 *        wt2 as from monthly_data::first_month_idx / monthly_data_weights (phenology_data_impl.hh:46-63). ---- */
#define ELMK_ATM_TBOT 0
#define ELMK_ATM_PBOT 1
#define ELMK_ATM_QBOT 2
#define ELMK_ATM_FLDS 3
#define ELMK_ATM_FSDS 4
#define ELMK_ATM_PREC 5
#define ELMK_ATM_WIND 6
#define ELMK_ATM_NVARS 7
#define ELMK_PHEN_MLAI 0
#define ELMK_PHEN_MSAI 1
#define ELMK_PHEN_MHTOP 2
#define ELMK_PHEN_MHBOT 3
#define ELMK_PHEN_NVARS 4
int elmk_atm_series(elmk_handle h, int var, const double* host, int ntimes);
int elmk_atm_forcing(elmk_handle h, int t_idx, double wt1, double wt2, int qbot_is_rh);
/* one time level of a resident series replaced from the host (the reference's read_atm_data refreshes its window of
 * records while the run proceeds, atm_forcing_kokkos.cc:15-27): asynchronous on a copy stream of the handle, ordered
 * after the forcing kernel that may still read the row and before the next elmk_atm_forcing; pinned host memory
 * overlaps with the step.  host holds ncols values and must stay untouched until the next elmk_atm_forcing has been
 * followed by elmk_sync (or any blocking call). */
int elmk_atm_series_row(elmk_handle h, int var, int t, const double* host);
int elmk_phen_series(elmk_handle h, int var, const double* host, int nmonths);
int elmk_phenology(elmk_handle h, int start_idx, double wt1, double wt2);

/* ---- the per-column part of kokkos_init_timestep (init_timestep_kokkos.cc:53-72):
 *      h2osno_old, dtbegin_column_h2o, ELM::init_timestep; resets forc_hgt_*_patch to
 *      forc_hgt (atm_physics_impl.hh:197-202) when reset_forc_hgt != 0 ---- */
int elmk_init_timestep(elmk_handle h, int reset_forc_hgt);

/* ---- one pass of the selected kernel groups over all columns, asynchronous on the
 *      handle's stream: replaces the 11 wrapper calls of ELMInterface::advance ---- */
int elmk_step(elmk_handle h, double dtime, double dayl, double max_dayl, uint32_t group_mask);

/* Launch plan of elmk_step.  ELMK_PLAN_FUSED (default): the production plan - the SNICAR kernel (a warp solves one
 * band of one flux type for 32 sunlit snow columns), bare-ground fluxes on compacted columns, the re-packed
 * CanopyFluxes iteration, snow hydrology + surface fluxes + conservation in one launch; the closed-form groups as
 * launches of their own (measured faster than fused).  ELMK_PLAN_SPLIT: one plain launch per kernel group with one
 * thread per column, the reference's wrapper granularity (23 parallel_for -> 11 launches); the two plans give
 * identical bits (tests/test_gpu_parity.py).  Takes effect from the next elmk_step. */
#define ELMK_PLAN_FUSED 0
#define ELMK_PLAN_SPLIT 1
int elmk_set_plan(elmk_handle h, int plan);
int elmk_sync(elmk_handle h);

/* number of kernel launches issued by this handle since creation (for bench accounting) */
int64_t elmk_launch_count(elmk_handle h);

/* ---- per-launch device timing (CUDA events on the handle's stream around every kernel of elmk_step
 *      and elmk_init_timestep).  Counterpart of the kernel name strings the reference passes to
 *      Kokkos::parallel_for (e.g. canopy_fluxes_kokkos.cc:264) for the Kokkos profiling tools.
 *      elmk_timing_read synchronises, accumulates the recorded intervals and returns, for up to `max`
 *      distinct launch names, the name, the total milliseconds and the number of launches; the return
 *      value is the number of names (or a negative error).  elmk_timing_enable(h, 0|1|2) also resets; 2 times the
 *      sub-launches of the composite launches as well (group mask 0: "albedo:snicar", "canopy_fluxes:iterate" ...). ---- */
int elmk_timing_enable(elmk_handle h, int on);
int elmk_timing_read(elmk_handle h, int max, const char** names, double* total_ms, int64_t* launches,
                     uint32_t* group_masks);

/* ---- error convention: replaces C++ exceptions thrown inside kernels.  any = OR of all
 *      columns' errmask words, first_col = lowest column index with a non-zero word (-1 if none).
 *      Synchronises the stream. ---- */
int elmk_errors(elmk_handle h, uint32_t* any, int64_t* first_col);
int elmk_clear_errors(elmk_handle h);
const char* elmk_error_text(uint32_t bit); /* the reference's message for one error bit */

/* ---- optional global balance diagnostic: sum/min/max over all columns of this handle of the
 *      8 fields dtend_column_h2o, errh2o, errh2osno, dwb, errsol, errlon, errseb, netrad
 *      (conserved_quantity_kokkos.cc:13-20), out[0..7]=sum, out[8..15]=min, out[16..23]=max.
 *      Counterpart of ELMKokkos::min_max_sum (src/utils/kokkos_utils.hh:13-58); the cross-rank
 *      reduction of the 24 doubles is done by the caller (NCCL all-reduce in the Python host). ---- */
int elmk_diag_reduce(elmk_handle h, double out[24]);

/* ---- pass counts of the CanopyFluxes stability iteration (canopy_fluxes_impl.hh:215-451, itlef) of the last elmk_step
 *      that ran group a7 in the fused plan: hist[k] = number of vegetated columns that took k passes (k = 3..41).
 *      The reference keeps the count in a wrapper-local View; here it is read from the iteration scratch.
 *      ELMK_EUNSUPPORTED on the CPU checkers. ---- */
int elmk_canflux_pass_histogram(elmk_handle h, int64_t hist[42]);

/* ---- the library's transcendental functions, evaluated where the library computes (the product: on the device):
 *      out[i] = fn(x[i]) or fn(x[i], y[i]).  exp, log, log10, pow, atan, cos, tanh, erf and acos return the bits of
 *      the libm the reference is built against (csrc/elmk_libm.h); this entry point exists so that a parity test
 *      can check that on the device itself.  y may be NULL for one-argument functions. ---- */
#define ELMK_MATH_EXP 0
#define ELMK_MATH_LOG 1
#define ELMK_MATH_LOG10 2
#define ELMK_MATH_POW 3
#define ELMK_MATH_ATAN 4
#define ELMK_MATH_COS 5
#define ELMK_MATH_TANH 6
#define ELMK_MATH_ERF 7
#define ELMK_MATH_ACOS 8
#define ELMK_MATH_DIV 9 /* x / y through the library's division */
#define ELMK_MATH_COUNT 10
int elmk_math_eval(elmk_handle h, int fn, int64_t n, const double* x, const double* y, double* out);

/* ---- library-level physics functions (the reference's ELM::<namespace>::<function> free functions, e.g.
 *      src/physics/canopy_hydrology.h) executed on the device for ONE column: args holds the function's arguments in
 *      the reference's order without the LandType - scalars (ints and bools as doubles) one slot each, per-column rows
 *      expanded in place with their full extent - and is updated in place with everything the function writes.
 *      Synchronous, no handle (device = CUDA device ordinal).  The C++ headers of include/elm/ are the typed front end;
 *      the device code is the one the fused kernels of elmk_step run.  ELMK_EUNSUPPORTED on the CPU checkers. ---- */
#define ELMK_FN_INTERCEPTION 0      /* canopy_hydrology::interception     canopy_hydrology_impl.hh:8    13 slots */
#define ELMK_FN_GROUND_FLUX 1       /* canopy_hydrology::ground_flux      :83    14 slots */
#define ELMK_FN_FRACTION_WET 2      /* canopy_hydrology::fraction_wet     :123    7 slots */
#define ELMK_FN_SNOW_INIT 3         /* canopy_hydrology::snow_init        :146  150 slots */
#define ELMK_FN_FRACTION_H2OSFC 4   /* canopy_hydrology::fraction_h2osfc  :312   26 slots */
#define ELMK_FN_RAD_INITIALIZE_FLUX 5 /* surface_radiation::initialize_flux           surface_radiation_impl.hh:9    11 slots */
#define ELMK_FN_RAD_TOTAL_ABSORBED 6  /* surface_radiation::total_absorbed_radiation  :30   36 slots */
#define ELMK_FN_RAD_LAYER_ABSORBED 7  /* surface_radiation::layer_absorbed_radiation  :77   39 slots (last: 0 where the reference asserts) */
#define ELMK_FN_RAD_REFLECTED 8       /* surface_radiation::reflected_radiation       :179   9 slots */
#define ELMK_FN_RAD_SUNSHADE 9        /* surface_radiation::canopy_sunshade_fractions :202  18 slots */
#define ELMK_FN_TMP_OLD_GROUND_TEMP 10   /* canopy_temperature::old_ground_temp     canopy_temperature_impl.hh:9   42 slots */
#define ELMK_FN_TMP_GROUND_TEMP 11       /* canopy_temperature::ground_temp         :32   25 slots */
#define ELMK_FN_TMP_CALC_SOILALPHA 12    /* canopy_temperature::calc_soilalpha      :51  160 slots */
#define ELMK_FN_TMP_CALC_SOILBETA 13     /* canopy_temperature::calc_soilbeta       :133  93 slots */
#define ELMK_FN_TMP_HUMIDITIES 14        /* canopy_temperature::humidities          :143  35 slots */
#define ELMK_FN_TMP_GROUND_PROPERTIES 15 /* canopy_temperature::ground_properties   :205  61 slots (displar, z0mr: the Land.vtype entries) */
#define ELMK_FN_TMP_FORCING_HEIGHT 16    /* canopy_temperature::forcing_height      :260  10 slots */
#define ELMK_FN_TMP_INIT_ENERGY_FLUXES 17 /* canopy_temperature::init_energy_fluxes :299   6 slots */
#define ELMK_FN_BGF_INITIALIZE_FLUX 18     /* bareground_fluxes::initialize_flux      bareground_fluxes_impl.hh:7   20 slots */
#define ELMK_FN_BGF_STABILITY_ITERATION 19 /* bareground_fluxes::stability_iteration  :30   22 slots */
#define ELMK_FN_BGF_COMPUTE_FLUX 20        /* bareground_fluxes::compute_flux         :82   56 slots */
#define ELMK_FN_CF_INITIALIZE_FLUX 21     /* canopy_fluxes::initialize_flux      canopy_fluxes_impl.hh:95  217 slots */
#define ELMK_FN_CF_STABILITY_ITERATION 22 /* canopy_fluxes::stability_iteration  :187  127 slots (slot 0: Land.vtype, psn_pft as its 27 members) */
#define ELMK_FN_CF_COMPUTE_FLUX 23        /* canopy_fluxes::compute_flux         :456   83 slots */
#define ELMK_FN_COUNT 24
int elmk_fn_call(int device, int fn, double* args, int64_t nargs);

/* raw device pointer + level stride of a field (for zero-copy interop with torch tensors) */
int elmk_device_ptr(elmk_handle h, int field, void** ptr, int64_t* level_stride);
/* the handle's CUDA stream (a cudaStream_t), so that a caller can record its own events on it or make
 * other streams wait for the step */
int elmk_stream(elmk_handle h, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* ELMK_B200_H_ */
