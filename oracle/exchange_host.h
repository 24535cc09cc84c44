// exchange_host.h - the elmk_exchange_* entry points of include/elmk_b200.h for the two CPU checker libraries
// (oracle/_ref, oracle/port): same call sequence and results as the product library, no overlap (host memory on
// both sides).  Test infrastructure only.  Included at the end of ref_capi.cc / port_capi.cc, inside extern "C".
#pragma once
#include <cstring>
#include <vector>

struct HostExchange {
  elmk_handle h;
  std::vector<int> in_fields, out_fields;
  std::vector<std::vector<char>> staged[2];
  long posts = 0, commits = 0;
};

static size_t hx_bytes(elmk_handle h, int field) {
  const char* nm; int dt, nlev;
  elmk_field_info(field, &nm, &dt, &nlev);
  return (size_t)elmk_ncols(h) * nlev * (dt == ELMK_F64 ? 8 : dt == ELMK_I32 ? 4 : 1);
}

int elmk_exchange_create(elmk_handle h, int n_in, const int* in_fields, int n_out, const int* out_fields, elmk_exchange* out) {
  if (!h || !out) return ELMK_EINVAL;
  auto* x = new HostExchange();
  x->h = h;
  x->in_fields.assign(in_fields, in_fields + n_in);
  x->out_fields.assign(out_fields, out_fields + n_out);
  for (auto& s : x->staged) s.resize(n_in);
  *out = reinterpret_cast<elmk_exchange>(x);
  return ELMK_OK;
}
int elmk_exchange_destroy(elmk_exchange xh) { delete reinterpret_cast<HostExchange*>(xh); return ELMK_OK; }
int elmk_exchange_post(elmk_exchange xh, const void* const* in_hosts) {
  auto* x = reinterpret_cast<HostExchange*>(xh);
  if (!x || x->posts - x->commits >= 2) return ELMK_EINVAL;
  auto& slot = x->staged[x->posts & 1];
  for (size_t i = 0; i < x->in_fields.size(); ++i) {
    const size_t nb = hx_bytes(x->h, x->in_fields[i]);
    slot[i].resize(nb);
    std::memcpy(slot[i].data(), in_hosts[i], nb);
  }
  x->posts += 1;
  return ELMK_OK;
}
int elmk_exchange_commit(elmk_exchange xh) {
  auto* x = reinterpret_cast<HostExchange*>(xh);
  if (!x || x->commits >= x->posts) return ELMK_EINVAL;
  auto& slot = x->staged[x->commits & 1];
  for (size_t i = 0; i < x->in_fields.size(); ++i)
    if (int rc = elmk_upload(x->h, x->in_fields[i], slot[i].data(), 0, elmk_ncols(x->h), ELMK_COL_OUTER)) return rc;
  x->commits += 1;
  return ELMK_OK;
}
int elmk_exchange_fetch(elmk_exchange xh, void* const* out_hosts) {
  auto* x = reinterpret_cast<HostExchange*>(xh);
  if (!x) return ELMK_EINVAL;
  for (size_t i = 0; i < x->out_fields.size(); ++i)
    if (int rc = elmk_download(x->h, x->out_fields[i], out_hosts[i], 0, elmk_ncols(x->h), ELMK_COL_OUTER)) return rc;
  return ELMK_OK;
}
int elmk_exchange_wait(elmk_exchange) { return ELMK_OK; }
int elmk_exchange_post_wait(elmk_exchange) { return ELMK_OK; }   // posts copy synchronously here

// elmk_math_eval for the CPU checkers: the host's libm, i.e. what the reference itself calls
#include <cmath>
int elmk_math_eval(elmk_handle h, int fn, int64_t n, const double* x, const double* y, double* out) {
  if (!h || !x || !out || fn < 0 || fn >= ELMK_MATH_COUNT || n < 0) return ELMK_EINVAL;
  if ((fn == ELMK_MATH_POW || fn == ELMK_MATH_DIV) && !y) return ELMK_EINVAL;
  for (int64_t i = 0; i < n; ++i) {
    volatile double a = x[i];
    switch (fn) {
      case ELMK_MATH_EXP: out[i] = std::exp(a); break;
      case ELMK_MATH_LOG: out[i] = std::log(a); break;
      case ELMK_MATH_LOG10: out[i] = std::log10(a); break;
      case ELMK_MATH_POW: { volatile double b = y[i]; out[i] = std::pow(a, b); } break;
      case ELMK_MATH_ATAN: out[i] = std::atan(a); break;
      case ELMK_MATH_COS: out[i] = std::cos(a); break;
      case ELMK_MATH_TANH: out[i] = std::tanh(a); break;
      case ELMK_MATH_ERF: out[i] = std::erf(a); break;
      case ELMK_MATH_ACOS: out[i] = std::acos(a); break;
      default: { volatile double b = y[i]; out[i] = a / b; } break;
    }
  }
  return ELMK_OK;
}
int elmk_set_plan(elmk_handle h, int plan) { return (h && (plan == ELMK_PLAN_FUSED || plan == ELMK_PLAN_SPLIT)) ? ELMK_OK : ELMK_EINVAL; }
int elmk_canflux_pass_histogram(elmk_handle, int64_t*) { return ELMK_EUNSUPPORTED; }
int elmk_fn_call(int, int, double*, int64_t) { return ELMK_EUNSUPPORTED; }   // the checkers ARE the function library
