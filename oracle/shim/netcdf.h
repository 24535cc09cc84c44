// TEST INFRASTRUCTURE (oracle) - not part of the product path.
//
// Compile-only stand-in for <netcdf.h>.  The reference's data headers include its NetCDF
// readers (src/utils/read_netcdf.hh:17), but libnetcdf is not in this image and no file IO
// happens on the hot path: oracle/ref_capi.cc feeds every table through memory.  Every
// call reports failure so that an accidental file read is loud.
#pragma once
#include <cstddef>
#define NC_NOERR 0
#define NC_NOWRITE 0
#define NC_WRITE 1
#define NC_CLOBBER 0
#define NC_DOUBLE 6
#define NC_MAX_VAR_DIMS 1024
#define NC_MAX_NAME 256
#define ELMK_NC_FAIL (-61)
inline int nc_open(const char*, int, int*) { return ELMK_NC_FAIL; }
inline int nc_close(int) { return ELMK_NC_FAIL; }
inline int nc_create(const char*, int, int*) { return ELMK_NC_FAIL; }
inline int nc_inq_varid(int, const char*, int*) { return ELMK_NC_FAIL; }
inline int nc_inq_var(int, int, char*, int*, int*, int*, int*) { return ELMK_NC_FAIL; }
inline int nc_inq_dimlen(int, int, size_t*) { return ELMK_NC_FAIL; }
inline int nc_inq_dimid(int, const char*, int*) { return ELMK_NC_FAIL; }
inline int nc_inq_vardimid(int, int, int*) { return ELMK_NC_FAIL; }
inline int nc_get_att(int, int, const char*, void*) { return ELMK_NC_FAIL; }
inline int nc_get_vara_double(int, int, const size_t*, const size_t*, double*) { return ELMK_NC_FAIL; }
inline int nc_get_vara_int(int, int, const size_t*, const size_t*, int*) { return ELMK_NC_FAIL; }
inline int nc_get_vara_text(int, int, const size_t*, const size_t*, char*) { return ELMK_NC_FAIL; }
inline int nc_def_dim(int, const char*, size_t, int*) { return ELMK_NC_FAIL; }
inline int nc_def_var(int, const char*, int, int, const int*, int*) { return ELMK_NC_FAIL; }
inline int nc_enddef(int) { return ELMK_NC_FAIL; }
inline int nc_put_vara_double(int, int, const size_t*, const size_t*, const double*) { return ELMK_NC_FAIL; }
inline const char* nc_strerror(int) { return "netcdf is stubbed out in the oracle build"; }
