// TEST INFRASTRUCTURE (oracle).  Stand-in for <netcdf.h> in the serial builds of the reference's unit tests
// (oracle/build_ref.py --tests): libnetcdf is not in this image, and test_SurfAlb / test_CanFlux read the PFT constants
// of test/data/clm_params_c180524.nc through ELM::IO::read_pft_var / read_names (src/utils/read_input.hh:200-240).
// This header serves exactly those reads from text dumps of the file's variables (one file per variable under
// $ELMK_NC_DUMP, default oracle/_ref/clm_params/, written by oracle/dump_params.py with scipy.io.netcdf_file):
// doubles one per line, the character variable `pftname` as its raw bytes.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#define NC_NOERR 0
#define NC_NOWRITE 0
#define NC_WRITE 1
#define NC_CLOBBER 0
#define NC_DOUBLE 6
#define NC_MAX_VAR_DIMS 1024
#define NC_MAX_NAME 256
#define ELMK_NC_FAIL (-61)
#ifndef ELMK_NC_DUMP_DEFAULT
#define ELMK_NC_DUMP_DEFAULT "."
#endif
namespace elmk_nc {
inline std::string dir() { const char* e = std::getenv("ELMK_NC_DUMP"); return e ? e : ELMK_NC_DUMP_DEFAULT; }
inline std::vector<std::string>& names() { static std::vector<std::string> n; return n; }
} // namespace elmk_nc
inline int nc_open(const char*, int, int* id) { *id = 1; return NC_NOERR; }
inline int nc_close(int) { return NC_NOERR; }
inline int nc_create(const char*, int, int*) { return ELMK_NC_FAIL; }
inline int nc_inq_varid(int, const char* name, int* id) {
  FILE* f = std::fopen((elmk_nc::dir() + "/" + name + ".txt").c_str(), "rb");
  if (!f) return ELMK_NC_FAIL;
  std::fclose(f);
  elmk_nc::names().push_back(name);
  *id = (int)elmk_nc::names().size() - 1;
  return NC_NOERR;
}
inline int nc_inq_var(int, int, char*, int*, int*, int*, int*) { return ELMK_NC_FAIL; }
inline int nc_inq_dimlen(int, int, size_t*) { return ELMK_NC_FAIL; }
inline int nc_inq_dimid(int, const char*, int*) { return ELMK_NC_FAIL; }
inline int nc_inq_vardimid(int, int, int*) { return ELMK_NC_FAIL; }
inline int nc_get_att(int, int, const char*, void*) { return ELMK_NC_FAIL; }
inline int nc_get_vara_double(int, int id, const size_t* start, const size_t* count, double* out) {
  if (id < 0 || id >= (int)elmk_nc::names().size()) return ELMK_NC_FAIL;
  FILE* f = std::fopen((elmk_nc::dir() + "/" + elmk_nc::names()[id] + ".txt").c_str(), "r");
  if (!f) return ELMK_NC_FAIL;
  double v;
  size_t i = 0, got = 0;
  while (got < count[0] && std::fscanf(f, "%la", &v) == 1) {
    if (i >= start[0]) out[got++] = v;
    ++i;
  }
  std::fclose(f);
  return got == count[0] ? NC_NOERR : ELMK_NC_FAIL;
}
inline int nc_get_vara_int(int, int, const size_t*, const size_t*, int*) { return ELMK_NC_FAIL; }
inline int nc_get_vara_text(int, int id, const size_t*, const size_t* count, char* out) {
  if (id < 0 || id >= (int)elmk_nc::names().size()) return ELMK_NC_FAIL;
  FILE* f = std::fopen((elmk_nc::dir() + "/" + elmk_nc::names()[id] + ".txt").c_str(), "rb");
  if (!f) return ELMK_NC_FAIL;
  const size_t n = count[0] * count[1];
  const size_t got = std::fread(out, 1, n, f);
  std::fclose(f);
  out[n] = '\0';   // read_names (read_input.hh:222-236) tokenises the buffer with strtok and allocates one byte more for this
  return got == n ? NC_NOERR : ELMK_NC_FAIL;
}
inline int nc_def_dim(int, const char*, size_t, int*) { return ELMK_NC_FAIL; }
inline int nc_def_var(int, const char*, int, int, const int*, int*) { return ELMK_NC_FAIL; }
inline int nc_enddef(int) { return ELMK_NC_FAIL; }
inline int nc_put_vara_double(int, int, const size_t*, const size_t*, const double*) { return ELMK_NC_FAIL; }
inline const char* nc_strerror(int) { return "variable not in the text dump served by oracle/shim_serial/netcdf.h"; }
