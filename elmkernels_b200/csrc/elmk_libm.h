// elmk_libm.h - double-precision exp / log / log10 / pow / atan / cos / acos / tanh (expm1) / erf that return, bit for
// bit, what glibc 2.39's x86-64 libm returns on a CPU with FMA (the __exp_fma / __log_fma / __pow_fma / __atan_fma /
// __cos_fma / __acos_fma / __expm1_fma ifunc variants and the generic __ieee754_log10, tanh, erf) - the library the
// reference's g++ build calls on this container and on the GPU box's host.  These are all the libm functions the
// column chain calls (sqrt, fabs, round, copysign are exact by IEEE 754).
//
// Why: the iterative kernel groups (CanopyFluxes, SoilTemperature) stop on thresholds; a transcendental that differs
// from the reference's in the last bit is amplified by cancellation (canopy-air temperature minus air temperature ...),
// flips a regime or convergence test and leaves the column ~1e-4 away from the reference.  With the same bits out of
// every library call - divisions and square roots are IEEE on both sides, FMA contraction is off - the CUDA step
// follows the reference's path exactly.
//
// What is restated: the algorithms of glibc sysdeps/ieee754/dbl-64/{e_exp,e_log,e_pow,s_atan,s_sin,e_asin,e_log10,
// s_expm1,s_tanh,s_erf}.c (exp, log, pow: Szabolcs Nagy's table-driven routines, 2018; atan, cos, acos: the IBM Accurate
// Mathematical Library routines as simplified in glibc 2.35; log10, expm1, tanh, erf: the fdlibm routines), with every
// multiply-add fused exactly where the FMA build of libm.so.6 fuses it (read off the
// disassembly of the shipped binary: the contraction pattern is the compiler's choice and decides the last bit).
// Tables: elmk_libm_tables.h (tools/gen_libm_tables.py).  Pinned to libm on 10^8 arguments per function by
// tests/test_libm_cpu.py (host build of this header: g++ -mfma -ffp-contract=off, so that only the explicit fma()
// calls below fuse); tests/test_gpu_libm.py runs the device build against libm values computed on the host.
//
// Arguments outside the range the column physics can produce (x <= 0 or non-finite for log / pow, |y| beyond
// 2^-65..2^63 for pow) fall back to the platform's function: defined results, not pinned.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "elmk_libm_tables.h"

#if defined(__CUDACC__)
#define ELMK_LM_HD __host__ __device__ __forceinline__
#else
#define ELMK_LM_HD inline
#endif

namespace elmk {
namespace lm {

#if defined(__CUDACC__)
// device copies (global memory, read through L1: the indices differ from lane to lane, which constant memory would
// serialise); declared in both compilation passes so that the host pass can register them
static __device__ const uint64_t exp_tab_dev[256] = {ELMK_EXP_TAB_INIT};
static __device__ const uint64_t log_tab_dev[256] = {ELMK_LOG_TAB_INIT};
static __device__ const uint64_t powlog_tab_dev[384] = {ELMK_POWLOG_TAB_INIT};
static __device__ const uint64_t atan_tab_dev[241 * 7] = {ELMK_ATAN_TAB_INIT};
static __device__ const uint64_t sincos_tab_dev[440] = {ELMK_SINCOS_TAB_INIT};
static __device__ const uint64_t asncs_tab_dev[2568] = {ELMK_ASNCS_TAB_INIT};
static __device__ const uint64_t inroot_tab_dev[128] = {ELMK_INROOT_TAB_INIT};
static __device__ const uint64_t powtwo_tab_dev[28] = {ELMK_POWTWO_TAB_INIT};
// host pass of the CUDA translation unit: never executed (the product library has no host path)
static const uint64_t exp_tab_host[2] = {0, 0}, log_tab_host[2] = {0, 0}, powlog_tab_host[3] = {0, 0, 0}, atan_tab_host[7] = {0},
                      sincos_tab_host[4] = {0}, asncs_tab_host[1] = {0}, inroot_tab_host[1] = {0}, powtwo_tab_host[1] = {0};
#else
static const uint64_t exp_tab_host[256] = {ELMK_EXP_TAB_INIT};
static const uint64_t log_tab_host[256] = {ELMK_LOG_TAB_INIT};
static const uint64_t powlog_tab_host[384] = {ELMK_POWLOG_TAB_INIT};
static const uint64_t atan_tab_host[241 * 7] = {ELMK_ATAN_TAB_INIT};
static const uint64_t sincos_tab_host[440] = {ELMK_SINCOS_TAB_INIT};
static const uint64_t asncs_tab_host[2568] = {ELMK_ASNCS_TAB_INIT};
static const uint64_t inroot_tab_host[128] = {ELMK_INROOT_TAB_INIT};
static const uint64_t powtwo_tab_host[28] = {ELMK_POWTWO_TAB_INIT};
#endif
#if defined(__CUDA_ARCH__)
#define ELMK_LM_TABLE(name) name##_dev
#else
#define ELMK_LM_TABLE(name) name##_host
#endif

ELMK_LM_HD uint64_t as_u64(const double x)
{
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
ELMK_LM_HD double as_f64(const uint64_t u)
{
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}
ELMK_LM_HD double tab_f64(const uint64_t* t, const int i) { return as_f64(t[i]); }

// ---- exp ------------------------------------------------------------------------------------------------------
// exp(x) = 2^(k/128) * exp(r), k = round(x * 128/ln2), |r| <= ln2/256; degree-5 polynomial for exp(r) - 1.
constexpr double kInvLn2N = 0x1.71547652b82fep0 * 128, kShift = 0x1.8p52;
constexpr double kNegLn2hiN = -0x1.62e42fefa0000p-8, kNegLn2loN = -0x1.cf79abc9e3b3ap-47;
constexpr double kExpC2 = 0x1.ffffffffffdbdp-2, kExpC3 = 0x1.555555555543cp-3, kExpC4 = 0x1.55555cf172b91p-5,
                 kExpC5 = 0x1.1111167a4d017p-7;

// result when 2^(k/128) alone over- or underflows (e_exp.c specialcase)
ELMK_LM_HD double exp_special(const double tmp, uint64_t sbits, const uint64_t ki)
{
  if ((ki & 0x80000000ull) == 0) {
    sbits -= 1009ull << 52;
    const double scale = as_f64(sbits);
    return 0x1p1009 * fma(scale, tmp, scale);
  }
  sbits += 1022ull << 52;
  const double scale = as_f64(sbits);
  const double st = scale * tmp;
  double y = scale + st;
  if (fabs(y) < 1.0) {
    const double one = (y < 0.0) ? -1.0 : 1.0;
    double lo = scale - y + st;
    const double hi = one + y;
    lo = one - hi + y + lo;
    y = (hi + lo) - one;
    if (y == 0.0) y = as_f64(sbits & 0x8000000000000000ull);
  }
  return 0x1p-1022 * y;
}

// exp(x + xtail) * (-1 if sign_bias), the core shared by exp and pow; xtail = 0, sign_bias = 0 for exp
ELMK_LM_HD double exp_core(const double x, const double xtail, const bool has_tail, const uint64_t sign_bias, uint32_t abstop)
{
  const uint64_t* T = ELMK_LM_TABLE(exp_tab);
  double kd = fma(x, kInvLn2N, kShift);
  const uint64_t ki = as_u64(kd);
  kd -= kShift;
  double r = fma(kd, kNegLn2hiN, x);
  r = fma(kd, kNegLn2loN, r);
  if (has_tail) r = xtail + r;
  const int idx = 2 * (int)(ki & 127u);
  const uint64_t top = (ki + sign_bias) << 45;
  const double tail = tab_f64(T, idx);
  const uint64_t sbits = T[idx + 1] + top;
  const double r2 = r * r;
  const double p23 = fma(r, kExpC3, kExpC2);
  const double tr = r + tail;
  const double p45 = fma(r, kExpC5, kExpC4);
  double tmp = fma(p23, r2, tr);
  const double r4 = r2 * r2;
  tmp = fma(p45, r4, tmp);
  if (abstop == 0) return exp_special(tmp, sbits, ki);
  const double scale = as_f64(sbits);
  return fma(scale, tmp, scale);
}

ELMK_LM_HD double g_exp(const double x)
{
  const uint64_t ix = as_u64(x);
  uint32_t abstop = (uint32_t)(ix >> 52) & 0x7ffu;
  if (abstop - 0x3c9u > 0x3eu) {
    if ((int32_t)(abstop - 0x3c9u) < 0) return 1.0 + x;   // |x| < 2^-54
    if (abstop >= 0x409u) {                               // |x| >= 1024, inf, nan
      if (ix == 0xfff0000000000000ull) return 0.0;
      if (abstop >= 0x7ffu) return 1.0 + x;
      return (ix >> 63) ? 0.0 : INFINITY;
    }
    abstop = 0;   // 512 <= |x| < 1024: through exp_special
  }
  return exp_core(x, 0.0, false, 0, abstop);
}

// ---- log ------------------------------------------------------------------------------------------------------
constexpr double kLn2hi = 0x1.62e42fefa3800p-1, kLn2lo = 0x1.ef35793c76730p-45;
constexpr double kLogA0 = -0x1.0000000000001p-1, kLogA1 = 0x1.555555551305bp-2, kLogA2 = -0x1.fffffffeb4590p-3,
                 kLogA3 = 0x1.999b324f10111p-3, kLogA4 = -0x1.55575e506c89fp-3;
constexpr double kLogB0 = -0x1p-1, kLogB1 = 0x1.5555555555577p-2, kLogB2 = -0x1.ffffffffffdcbp-3, kLogB3 = 0x1.999999995dd0cp-3,
                 kLogB4 = -0x1.55555556745a7p-3, kLogB5 = 0x1.24924a344de30p-3, kLogB6 = -0x1.fffffa4423d65p-4,
                 kLogB7 = 0x1.c7184282ad6cap-4, kLogB8 = -0x1.999eb43b068ffp-4, kLogB9 = 0x1.78182f7afd085p-4,
                 kLogB10 = -0x1.5521375d145cdp-4;

ELMK_LM_HD double g_log(const double x)
{
  uint64_t ix = as_u64(x);
  if (ix - 0x3fee000000000000ull <= 0x308ffffffffffull) {
    // 1 - 2^-4 <= x < 1 + 0x1.09p-4: polynomial in r = x - 1 with a split leading term
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double r = x - 1.0;
    const double r2 = r * r;
    const double r3 = r * r2;
    const double q1 = fma(r2, kLogB3, fma(r, kLogB2, kLogB1));
    const double q2 = fma(r2, kLogB6, fma(r, kLogB5, kLogB4));
    double q3 = fma(r2, kLogB9, fma(r, kLogB8, kLogB7));
    q3 = fma(r3, kLogB10, q3);
    q3 = fma(q3, r3, q2);
    const double p = fma(q3, r3, q1);
    const double w = fma(r, 0x1p27, r);
    const double rhi = fma(-0x1p27, r, w);   // -(2^27 r) + w
    const double rhi2 = rhi * rhi;
    const double rlo = r - rhi;
    const double hi = fma(rhi2, kLogB0, r);
    double lo = fma(rhi2, kLogB0, r - hi);
    lo = fma(kLogB0 * rlo, r + rhi, lo);
    const double y = fma(p, r3, lo);
    return y + hi;
  }
  const uint32_t top = (uint32_t)(ix >> 48);
  if (top - 0x0010u > 0x7fdfu) {
    if (ix * 2 == 0) return -INFINITY;
    if (ix == 0x7ff0000000000000ull) return x;
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return as_f64(0x7ff8000000000000ull);   // negative or nan
    ix = as_u64(x * 0x1p52);   // subnormal
    ix -= 52ull << 52;
  }
  const uint64_t* T = ELMK_LM_TABLE(log_tab);
  const uint64_t tmp = ix - 0x3fe6000000000000ull;
  const int i = (int)((tmp >> 45) & 127u);
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
  const double invc = tab_f64(T, 2 * i), logc = tab_f64(T, 2 * i + 1);
  const double z = as_f64(iz);
  const double kd = (double)k;
  const double w = fma(kd, kLn2hi, logc);
  const double r = fma(z, invc, -1.0);
  const double p12 = fma(r, kLogA2, kLogA1);
  const double hi = r + w;
  const double r2 = r * r;
  double lo = (w - hi) + r;
  lo = fma(kd, kLn2lo, lo);
  const double r3 = r * r2;
  const double p34 = fma(r, kLogA4, kLogA3);
  lo = fma(r2, kLogA0, lo);
  const double p = fma(p34, r2, p12);
  return fma(r3, p, lo) + hi;
}

// __ieee754_log10 (e_log10.c): no fused operations, calls log
ELMK_LM_HD double g_log10(double x)
{
  constexpr double ivln10 = 0x1.bcb7b1526e50ep-2, log10_2hi = 0x1.34413509f6000p-2, log10_2lo = 0x1.9fef311f12b36p-42;
  uint64_t hx = as_u64(x);
  int k = 0;
  if ((int64_t)hx < 0x0010000000000000ll) {
    if ((hx & 0x7fffffffffffffffull) == 0) return -INFINITY;
    if ((int64_t)hx < 0) return as_f64(0x7ff8000000000000ull);
    k -= 54;
    x *= 0x1p54;
    hx = as_u64(x);
  }
  if (hx >= 0x7ff0000000000000ull) return x + x;
  k += (int)(hx >> 52) - 1023;
  const int i = (k < 0) ? 1 : 0;
  hx = (hx & 0x000fffffffffffffull) | ((uint64_t)(0x3ff - i) << 52);
  const double y = (double)(k + i);
  const double z = y * log10_2lo + ivln10 * g_log(as_f64(hx));
  return z + y * log10_2hi;
}

// ---- pow ------------------------------------------------------------------------------------------------------
constexpr double kPowA0 = -0x1p-1, kPowA1 = -0x1.5555555555560p-1, kPowA2 = 0x1.0000000000006p-1, kPowA3 = 0x1.999999959554ep-1,
                 kPowA4 = -0x1.555555529a47ap-1, kPowA5 = -0x1.2495b9b4845e9p0, kPowA6 = 0x1.0002b8b263fc3p0;

// log(x) as lhi + llo with ~68 bits (e_pow.c log_inline), x positive and normal
ELMK_LM_HD void pow_log(const uint64_t ix, double& lhi, double& llo)
{
  const uint64_t* T = ELMK_LM_TABLE(powlog_tab);
  const uint64_t tmp = ix - 0x3fe6955500000000ull;
  const int i = (int)((tmp >> 45) & 127u);
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
  const double z = as_f64(iz);
  const double kd = (double)k;
  const double invc = tab_f64(T, 3 * i), logc = tab_f64(T, 3 * i + 1), logctail = tab_f64(T, 3 * i + 2);
  const double t1 = fma(kd, kLn2hi, logc);
  const double lo1 = fma(kd, kLn2lo, logctail);
  const double r = fma(z, invc, -1.0);
  const double ar = r * kPowA0;
  const double p12 = fma(r, kPowA2, kPowA1);
  const double p34 = fma(r, kPowA4, kPowA3);
  const double t2 = r + t1;
  const double lo2 = (t1 - t2) + r;
  const double ar2 = r * ar;
  const double ar3 = r * ar2;
  const double lo3 = fma(ar, r, -ar2);
  const double hi = t2 + ar2;
  double p = fma(r, kPowA6, kPowA5);
  p = fma(p, ar2, p34);
  const double lo4 = (t2 - hi) + ar2;
  p = fma(ar2, p, p12);
  double lo = lo1 + lo2;
  lo = lo + lo3;
  lo = lo + lo4;
  lo = fma(ar3, p, lo);
  lhi = hi + lo;
  llo = (hi - lhi) + lo;
}
// exp(y * (lhi + llo)) (e_pow.c exp_inline after the product)
ELMK_LM_HD double pow_exp(const double y, const double lhi, const double llo)
{
  const double ehi = y * lhi;
  const double elo = fma(y, llo, fma(lhi, y, -ehi));
  uint32_t abstop = (uint32_t)(as_u64(ehi) >> 52) & 0x7ffu;
  if (abstop - 0x3c9u > 0x3eu) {
    if ((int32_t)(abstop - 0x3c9u) < 0) return 1.0 + ehi;   // |y log x| < 2^-54
    if (abstop >= 0x409u) return (as_u64(ehi) >> 63) ? 0.0 : INFINITY;
    abstop = 0;
  }
  return exp_core(ehi, elo, true, 0, abstop);
}

ELMK_LM_HD double g_pow(const double x, const double y)
{
  const uint64_t ix = as_u64(x), iy = as_u64(y);
  const uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52) & 0x7ffu;
  if (topx - 1u > 0x7fdu || topy - 0x3beu > 0x7fu) {
    if (iy * 2 == 0) return 1.0;
    if (ix == 0x3ff0000000000000ull) return 1.0;
    return pow(x, y);   // zero, negative, subnormal or non-finite base, tiny or huge exponent: not on the column path
  }
  double lhi, llo;
  pow_log(ix, lhi, llo);
  return pow_exp(y, lhi, llo);
}

// pow(c, y) for a constant base c whose (lhi, llo) = pow_log(c) are given as literals: the first half of pow is the
// same for every call, so it is folded; the value is pow's own (tests/libm/libm_check.cc pins pow_cbase against libm).
ELMK_LM_HD double g_pow_cbase(const double base, const double lhi, const double llo, const double y)
{
  const uint32_t topy = (uint32_t)(as_u64(y) >> 52) & 0x7ffu;
  if (topy - 0x3beu > 0x7fu) return g_pow(base, y);
  return pow_exp(y, lhi, llo);
}
// pow_log of 0.57, 2.29 and 2.0 (tests/libm/libm_check.cc recomputes them)
#define ELMK_POWLOG_0_57 -0x1.1fce0d03dd5e6p-1, 0x1.a4ee52p-57
#define ELMK_POWLOG_2_29 0x1.a837f19ef9d69p-1, 0x1.2e2417cp-55
#define ELMK_POWLOG_2_0 0x1.62e42fefa39efp-1, 0x1.abc9e3b398p-56

// ---- atan -----------------------------------------------------------------------------------------------------
ELMK_LM_HD double g_atan(const double x)
{
  constexpr double d3 = -0x1.5555555555555p-2, d5 = 0x1.99999999997fdp-3, d7 = -0x1.24924923f7603p-3, d9 = 0x1.c71c6e5129a3bp-4,
                   d11 = -0x1.7458022b13c25p-4, d13 = 0x1.375f08b31cbcep-4;
  constexpr double hpi = 0x1.921fb54442d18p0, hpi1 = 0x1.1a62633145c07p-54;
  const uint64_t* T = ELMK_LM_TABLE(atan_tab);
  const uint64_t ix = as_u64(x);
  if (((ix >> 52) & 0x7ffu) == 0x7ffu && (ix & 0x000fffffffffffffull)) return x + x;   // nan
  const double u = fabs(x);
  const uint64_t sign = ix & 0x8000000000000000ull;
  if (u < 1.0) {
    if (u < 0.0625) {
      if (u < 0x1.bb67ap-27) return x;
      const double v = x * x;
      double p = fma(v, d13, d11);
      p = fma(v, p, d9);
      p = fma(v, p, d7);
      p = fma(v, p, d5);
      p = fma(v, p, d3);
      return fma(x * v, p, x);
    }
    const int i = (int)(fma(u, 256.0, 0x1p52) - 0x1p52) - 16;
    const uint64_t* c = T + 7 * i;
    const double z = u - as_f64(c[0]);
    double p = fma(z, as_f64(c[6]), as_f64(c[5]));
    p = fma(z, p, as_f64(c[4]));
    p = fma(z, p, as_f64(c[3]));
    p = fma(z, p, as_f64(c[2]));
    const double r = fma(p, z, as_f64(c[1]));
    return as_f64((as_u64(r) & 0x7fffffffffffffffull) | sign);
  }
  if (u < 16.0) {
    const double w = 1.0 / u;
    const double t1 = w * u;
    const double t2 = fma(u, w, -t1);
    const double e = (1.0 - t1) - t2;
    const int i = (int)(fma(w, 256.0, 0x1p52) - 0x1p52) - 16;
    const uint64_t* c = T + 7 * i;
    const double z = fma(e, w, w - as_f64(c[0]));
    double p = fma(z, as_f64(c[6]), as_f64(c[5]));
    p = fma(z, p, as_f64(c[4]));
    p = fma(z, p, as_f64(c[3]));
    p = fma(z, p, as_f64(c[2]));
    const double t3 = fma(-p, z, hpi1);
    const double r = (hpi - as_f64(c[1])) + t3;
    return as_f64((as_u64(r) & 0x7fffffffffffffffull) | sign);
  }
  if (u < 0x1.49ff2p52) {
    const double w = 1.0 / u;
    const double t1 = w * u;
    const double t3 = hpi - w;
    const double v = w * w;
    double p = fma(v, d13, d11);
    p = fma(v, p, d9);
    p = fma(v, p, d7);
    p = fma(v, p, d5);
    p = fma(v, p, d3);
    const double t2 = fma(u, w, -t1);
    double cor = ((hpi - t3) - w) + hpi1;
    const double wv = w * v;
    const double e = (1.0 - t1) - t2;
    cor = fma(-e, w, cor);
    const double s = fma(-wv, p, cor);
    const double r = t3 + s;
    return as_f64((as_u64(r) & 0x7fffffffffffffffull) | sign);
  }
  return as_f64(as_u64(hpi) | sign);
}


// ---- cos ------------------------------------------------------------------------------------------------------
// s_sin.c (__cos): table of sin / cos at k/128 plus short polynomials; arguments up to 1e8 by a three-part
// Cody-Waite reduction.  Larger arguments (Payne-Hanek in glibc) fall back to the platform's cos.
namespace sc {
constexpr double big = 0x1.8p45, sn3 = -0x1.5555555555515p-3, sn5 = 0x1.11110e829872fp-7, cs2 = 0.5, cs4 = -0x1.5555555555535p-5,
                 cs6 = 0x1.6c16bedd9e239p-10;
constexpr double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ecep-7, s3 = -0x1.a01a019db08b8p-13, s4 = 0x1.71de27b9a7ed9p-19,
                 s5 = -0x1.addffc2fcdf59p-26;
constexpr double hp0 = 0x1.921fb54442d18p0, hp1 = 0x1.1a62633145c07p-54;
constexpr double hpinv = 0x1.45f306dc9c883p-1, toint = 0x1.8p52, mp1 = 0x1.921fb58000000p0, mp2 = -0x1.dde973c000000p-27,
                 pp3 = -0x1.cb3b398000000p-55, pp4 = -0x1.d747f23e32ed7p-83;

// cos(x + dx) for |x| < 0.855469 + a little
ELMK_LM_HD double do_cos(double x, double dx)
{
  const uint64_t* T = ELMK_LM_TABLE(sincos_tab);
  if (x < 0.0) dx = -dx;
  const double ax = fabs(x);
  const double u = ax + big;
  const int k = (int)((uint32_t)as_u64(u) << 2);
  x = (ax - (u - big)) + dx;
  const double xx = x * x;
  const double s = fma(x * xx, fma(xx, sn5, sn3), x);
  const double c = xx * fma(xx, fma(xx, cs6, cs4), cs2);
  const double sn = tab_f64(T, k), ssn = tab_f64(T, k + 1), cs = tab_f64(T, k + 2), ccs = tab_f64(T, k + 3);
  double cor = fma(-s, ssn, ccs);
  cor = fma(-c, cs, cor);
  cor = fma(-s, sn, cor);
  return cs + cor;
}
// sin(x + dx), same range
ELMK_LM_HD double do_sin(const double x, double dx)
{
  const uint64_t* T = ELMK_LM_TABLE(sincos_tab);
  const double xold = x;
  if (fabs(x) < 0.126) {
    const double xx = x * x;
    double p = fma(xx, s5, s4);
    p = fma(xx, p, s3);
    p = fma(xx, p, s2);
    p = fma(xx, p, s1);
    const double t = fma(xx, fma(p, x, -(dx * 0.5)), dx);
    return x + t;
  }
  if (x <= 0.0) dx = -dx;
  const double ax = fabs(x);
  const double u = ax + big;
  const int k = (int)((uint32_t)as_u64(u) << 2);
  const double xr = ax - (u - big);
  const double xx = xr * xr;
  const double s = xr + fma(xr * xx, fma(xx, sn5, sn3), dx);
  const double c = fma(xr, dx, xx * fma(xx, fma(xx, cs6, cs4), cs2));
  const double sn = tab_f64(T, k), ssn = tab_f64(T, k + 1), cs = tab_f64(T, k + 2), ccs = tab_f64(T, k + 3);
  double cor = fma(s, ccs, ssn);
  cor = fma(-c, sn, cor);
  cor = fma(s, cs, cor);
  const double r = sn + cor;
  return as_f64((as_u64(r) & 0x7fffffffffffffffull) | (as_u64(xold) & 0x8000000000000000ull));
}
} // namespace sc

ELMK_LM_HD double g_cos(const double x)
{
  using namespace sc;
  const uint32_t k = (uint32_t)(as_u64(x) >> 32) & 0x7fffffffu;
  if (k <= 0x3e3fffffu) return 1.0;
  if (k <= 0x3feb5fffu) return do_cos(x, 0.0);
  if (k <= 0x400368fcu) {
    const double y = hp0 - fabs(x);
    const double a = y + hp1;
    const double da = (y - a) + hp1;
    return do_sin(a, da);
  }
  if (k <= 0x419921fau) {
    const double t = fma(x, hpinv, toint);
    const double xn = t - toint;
    const uint32_t n = (uint32_t)as_u64(t) & 3u;
    double y = fma(-xn, mp1, x);
    y = fma(-xn, mp2, y);
    const double b = fma(-xn, pp3, y);
    const double db = fma(-xn, pp3, y - b);
    const double a = fma(-xn, pp4, b);
    const double da = db + fma(-xn, pp4, b - a);
    double r;
    if (n & 1u) {
      r = do_sin(a, da);
    } else {
      r = do_cos(a, da);
    }
    return ((n + 1u) & 2u) ? -r : r;
  }
  return cos(x);
}

// sin, same file (__sin): used by the per-column solar geometry (phys_forcing.h)
ELMK_LM_HD double g_sin(const double x)
{
  using namespace sc;
  const uint32_t k = (uint32_t)(as_u64(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e500000u) return x;
  if (k < 0x3feb6000u) return do_sin(x, 0.0);
  if (k < 0x400368fdu) {
    const double t = hp0 - fabs(x);
    const double r = do_cos(t, hp1);
    return as_f64((as_u64(r) & 0x7fffffffffffffffull) | (as_u64(x) & 0x8000000000000000ull));
  }
  if (k < 0x419921fbu) {
    const double t = fma(x, hpinv, toint);
    const double xn = t - toint;
    const uint32_t n = (uint32_t)as_u64(t) & 3u;
    double y = fma(-xn, mp1, x);
    y = fma(-xn, mp2, y);
    const double b = fma(-xn, pp3, y);
    const double db = fma(-xn, pp3, y - b);
    const double a = fma(-xn, pp4, b);
    const double da = db + fma(-xn, pp4, b - a);
    const double r = (n & 1u) ? do_cos(a, da) : do_sin(a, da);
    return (n & 2u) ? -r : r;
  }
  return sin(x);
}

// ---- tanh (s_tanh.c, not fused) over expm1 (s_expm1.c, the __expm1_fma variant) -----------------------------------
ELMK_LM_HD double hi_add(const double y, const int k)   // add k to the binary exponent of y
{
  return as_f64(as_u64(y) + ((uint64_t)(uint32_t)(k << 20) << 32));
}
ELMK_LM_HD double g_expm1(double x)
{
  constexpr double ln2_hi = 0x1.62e42fee00000p-1, ln2_lo = 0x1.a39ef35793c76p-33, invln2 = 0x1.71547652b82fep0;
  constexpr double Q1 = -0x1.11111111110f4p-5, Q2 = 0x1.a01a019fe5585p-10, Q3 = -0x1.4ce199eaadbb7p-14, Q4 = 0x1.0cfca86e65239p-18,
                   Q5 = -0x1.afdb76e09c32dp-23;
  const uint32_t hx0 = (uint32_t)(as_u64(x) >> 32);
  const bool neg = (hx0 & 0x80000000u) != 0;
  const uint32_t hx = hx0 & 0x7fffffffu;
  double hi, lo, c = 0.0;
  int k;
  if (hx > 0x40436879u) {   // |x| >= 56 ln2
    if (hx > 0x40862e41u) {
      if (hx > 0x7fefffffu) return ((hx & 0xfffffu) | (uint32_t)as_u64(x)) ? x + x : (neg ? -1.0 : x);
      if (x > 0x1.62e42fefa39efp9) return INFINITY;
    }
    if (neg) return -1.0;
  }
  if (hx > 0x3fd62e42u) {   // |x| > 0.5 ln2
    if (hx <= 0x3ff0a2b1u) {
      if (!neg) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
      else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
    } else {
      k = (int)((neg ? -0.5 : 0.5) + invln2 * x);
      const double t = (double)k;
      hi = fma(-t, ln2_hi, x);
      lo = t * ln2_lo;
    }
    x = hi - lo;
    c = (hi - x) - lo;
  } else if (hx <= 0x3c8fffffu) {
    return x;
  } else {
    k = 0;
  }
  const double hfx = 0.5 * x;
  const double hxs = x * hfx;
  const double R2 = fma(hxs, Q3, Q2);
  const double R3 = fma(hxs, Q5, Q4);
  const double h2 = hxs * hxs;
  const double R1 = fma(hxs, Q1, 1.0);
  const double h4 = h2 * h2;
  const double r1 = fma(h4, R3, fma(h2, R2, R1));
  const double t = fma(-r1, hfx, 3.0);
  double e = hxs * ((r1 - t) / fma(-x, t, 6.0));
  if (k == 0) return x - fma(e, x, -hxs);
  e = fma(e - c, x, -c);
  e -= hxs;
  if (k == -1) return fma(0.5, x - e, -0.5);
  if (k == 1) {
    if (x < -0.25) return -2.0 * (e - (x + 0.5));
    return fma(x - e, 2.0, 1.0);
  }
  if (k <= -2 || k > 56) {
    const double y = 1.0 - (e - x);
    return hi_add(y, k) - 1.0;
  }
  if (k < 20) {
    const double tt = as_f64((uint64_t)(0x3ff00000u - (0x200000u >> k)) << 32);
    return hi_add(tt - (e - x), k);
  }
  const double tt = as_f64((uint64_t)((uint32_t)(0x3ff - k) << 20) << 32);
  double y = x - (e + tt);
  y += 1.0;
  return hi_add(y, k);
}

ELMK_LM_HD double g_tanh(const double x)
{
  const uint64_t ux = as_u64(x);
  const uint32_t ix = (uint32_t)(ux >> 32) & 0x7fffffffu;
  const bool neg = (ux >> 63) != 0;
  if (ix > 0x7fefffffu) return neg ? 1.0 / x - 1.0 : 1.0 / x + 1.0;
  double z;
  if (ix <= 0x4035ffffu) {
    if ((ux << 1) == 0) return x;
    if (ix <= 0x3c7fffffu) return x * (1.0 + x);
    const double ax = fabs(x);
    if (ix > 0x3fefffffu) {
      const double t = g_expm1(ax + ax);
      z = 1.0 - 2.0 / (t + 2.0);
    } else {
      const double t = g_expm1(ax * -2.0);
      z = -t / (t + 2.0);
    }
  } else {
    z = 1.0;
  }
  return neg ? -z : z;
}

// ---- erf (s_erf.c, not fused; its two exponentials are libm's exp) ------------------------------------------------
ELMK_LM_HD double g_erf(const double x)
{
  constexpr double efx = 0x1.06eba8214db69p-3, erx = 0x1.b0ac160000000p-1;
  constexpr double pp0 = 0x1.06eba8214db68p-3, pp1 = -0x1.4cd7d691cb913p-2, pp2 = -0x1.d2a51dbd7194fp-6, pp3 = -0x1.7a291236668e4p-8,
                   pp4 = -0x1.8ead6120016acp-16;
  constexpr double qq1 = 0x1.97779cddadc09p-2, qq2 = 0x1.0a54c5536cebap-4, qq3 = 0x1.4d022c4d36b0fp-8, qq4 = 0x1.15dc9221c1a10p-13,
                   qq5 = -0x1.09c4342a26120p-18;
  constexpr double pa0 = -0x1.359b8bef77538p-9, pa1 = 0x1.a8d00ad92b34dp-2, pa2 = -0x1.7d240fbb8c3f1p-2, pa3 = 0x1.45fca805120e4p-2,
                   pa4 = -0x1.c63983d3e28ecp-4, pa5 = 0x1.22a36599795ebp-5, pa6 = -0x1.1bf380a96073fp-9;
  constexpr double qa1 = 0x1.b3e6618eee323p-4, qa2 = 0x1.14af092eb6f33p-1, qa3 = 0x1.2635cd99fe9a7p-4, qa4 = 0x1.02660e763351fp-3,
                   qa5 = 0x1.bedc26b51dd1cp-7, qa6 = 0x1.88b545735151dp-7;
  constexpr double ra0 = -0x1.43412600d6435p-7, ra1 = -0x1.63416e4ba7360p-1, ra2 = -0x1.51e0441b0e726p3, ra3 = -0x1.f300ae4cba38dp5,
                   ra4 = -0x1.44cb184282266p7, ra5 = -0x1.7135cebccabb2p7, ra6 = -0x1.4526557e4d2f2p6, ra7 = -0x1.3a0efc69ac25cp3;
  constexpr double sa1 = 0x1.3a6b9bd707687p4, sa2 = 0x1.1350c526ae721p7, sa3 = 0x1.b290dd58a1a71p8, sa4 = 0x1.42b1921ec2868p9,
                   sa5 = 0x1.ad02157700314p8, sa6 = 0x1.b28a3ee48ae2cp6, sa7 = 0x1.a47ef8e484a93p2, sa8 = -0x1.eeff2ee749a62p-5;
  constexpr double rb0 = -0x1.4341239e86f4ap-7, rb1 = -0x1.993ba70c285dep-1, rb2 = -0x1.1c209555f995ap4, rb3 = -0x1.4145d43c5ed98p7,
                   rb4 = -0x1.3ec881375f228p9, rb5 = -0x1.004616a2e5992p10, rb6 = -0x1.e384e9bdc383fp8;
  constexpr double sb1 = 0x1.e568b261d5190p4, sb2 = 0x1.45cae221b9f0ap8, sb3 = 0x1.802eb189d5118p10, sb4 = 0x1.8ffb7688c246ap11,
                   sb5 = 0x1.3f219cedf3be6p11, sb6 = 0x1.da874e79fe763p8, sb7 = -0x1.670e242712d62p4;
  const uint64_t ux = as_u64(x);
  const uint32_t ix = (uint32_t)(ux >> 32) & 0x7fffffffu;
  const bool neg = (ux >> 63) != 0;
  if (ix > 0x7fefffffu) return (double)(1 - 2 * (int)neg) + 1.0 / x;   // erf(nan) = nan, erf(+-inf) = +-1
  if (ix <= 0x3feaffffu) {   // |x| < 0.84375
    if (ix <= 0x3e2fffffu) {
      if (ix < 0x00800000u) return 0.0625 * (16.0 * x + (16.0 * efx) * x);
      return x + efx * x;
    }
    const double z = x * x;
    const double r1 = pp0 + z * pp1, z2 = z * z;
    const double r2 = pp2 + z * pp3, z4 = z2 * z2;
    const double s1 = 1.0 + z * qq1;
    const double s2 = qq2 + z * qq3;
    const double s3 = qq4 + z * qq5;
    const double r = r1 + z2 * r2 + z4 * pp4;
    const double s = s1 + z2 * s2 + z4 * s3;
    const double y = r / s;
    return x + x * y;
  }
  if (ix <= 0x3ff3ffffu) {   // 0.84375 <= |x| < 1.25
    const double s = fabs(x) - 1.0;
    const double P1 = pa0 + s * pa1, s2 = s * s;
    const double Q1 = 1.0 + s * qa1, s4 = s2 * s2;
    const double P2 = pa2 + s * pa3, s6 = s4 * s2;
    const double Q2 = qa2 + s * qa3;
    const double P3 = pa4 + s * pa5;
    const double Q3 = qa4 + s * qa5;
    const double P = P1 + s2 * P2 + s4 * P3 + s6 * pa6;
    const double Q = Q1 + s2 * Q2 + s4 * Q3 + s6 * qa6;
    return neg ? -erx - P / Q : erx + P / Q;
  }
  if (ix > 0x4017ffffu) return neg ? -1.0 : 1.0;   // |x| >= 6
  const double ax = fabs(x);
  const double s = 1.0 / (ax * ax);
  double R, S;
  if (ix <= 0x4006db6du) {   // |x| < 1/0.35
    const double R1 = ra0 + s * ra1, s2 = s * s;
    const double S1 = 1.0 + s * sa1, s4 = s2 * s2;
    const double R2 = ra2 + s * ra3, s6 = s4 * s2;
    const double S2 = sa2 + s * sa3, s8 = s4 * s4;
    const double R3 = ra4 + s * ra5;
    const double S3 = sa4 + s * sa5;
    const double R4 = ra6 + s * ra7;
    const double S4 = sa6 + s * sa7;
    R = R1 + s2 * R2 + s4 * R3 + s6 * R4;
    S = S1 + s2 * S2 + s4 * S3 + s6 * S4 + s8 * sa8;
  } else {
    const double R1 = rb0 + s * rb1, s2 = s * s;
    const double S1 = 1.0 + s * sb1, s4 = s2 * s2;
    const double R2 = rb2 + s * rb3, s6 = s4 * s2;
    const double S2 = sb2 + s * sb3;
    const double R3 = rb4 + s * rb5;
    const double S3 = sb4 + s * sb5;
    const double S4 = sb6 + s * sb7;
    R = R1 + s2 * R2 + s4 * R3 + s6 * rb6;
    S = S1 + s2 * S2 + s4 * S3 + s6 * S4;
  }
  const double z = as_f64(as_u64(ax) & 0xffffffff00000000ull);
  const double r = g_exp(-z * z - 0.5625) * g_exp((z - ax) * (z + ax) + R / S);
  return neg ? r / ax - 1.0 : 1.0 - r / ax;
}

// ---- acos (e_asin.c __ieee754_acos, the __acos_fma variant) --------------------------------------------------------
namespace ac {
constexpr double hp0 = 0x1.921fb54442d18p0, hp1 = 0x1.1a62633145c07p-54;
// one row of the asncs table: Taylor-like expansion of acos around T[b]; TOP = index of the leading coefficient
template <int TOP> ELMK_LM_HD double row(const uint64_t* T, const int b, const double ax, const bool pos)
{
  const double xx = ax - as_f64(T[b]);
  double p = as_f64(T[b + TOP]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = TOP - 1; j >= 2; --j) p = fma(xx, p, as_f64(T[b + j]));
  p = fma(xx * xx, p, as_f64(T[b + TOP + 1]));
  const double t = fma(xx, as_f64(T[b + 1]), p);
  const double y0 = as_f64(T[b + TOP + 2]);
  return pos ? (hp1 - t) + (hp0 - y0) : (t + hp1) + (y0 + hp0);
}
} // namespace ac

ELMK_LM_HD double g_acos(const double x)
{
  using namespace ac;
  constexpr double f1 = 0x1.55555555554f9p-3, f2 = 0x1.333333336127dp-4, f3 = 0x1.6db6dae42c0e4p-5, f4 = 0x1.f1c7e04f4ad99p-6,
                   f5 = 0x1.6e442c822d419p-6, f6 = 0x1.292d80f453c72p-6;
  constexpr double rt0 = 0x1.fffffffecc1ddp-1, rt1 = 0x1.fffffff757304p-2, rt2 = 0x1.800496769c91ap-2, rt3 = 0x1.4006318d1dab9p-2;
  const uint64_t* T = ELMK_LM_TABLE(asncs_tab);
  const uint64_t ux = as_u64(x);
  const uint32_t k = (uint32_t)(ux >> 32) & 0x7fffffffu;
  const bool pos = (int32_t)(ux >> 32) > 0;
  if (k <= 0x3c87ffffu) return hp0;
  const double ax = fabs(x);
  if (k <= 0x3fbfffffu) {
    const double x2 = x * x;
    double p = fma(x2, f6, f5);
    p = fma(x2, p, f4);
    const double r = hp0 - x;
    p = fma(x2, p, f3);
    p = fma(x2, p, f2);
    p = fma(x2, p, f1);
    const double cor = fma(-p, x * x2, ((hp0 - r) - x) + hp1);
    return r + cor;
  }
  if (k <= 0x3fcfffffu) return row<6>(T, (int)((k >> 15) & 0x1fu) * 11, ax, pos);
  if (k <= 0x3fdfffffu) return row<6>(T, (int)((k >> 14) & 0x3fu) * 11 + 352, ax, pos);
  if (k <= 0x3fe7ffffu) return row<7>(T, (int)((k >> 13) & 0x7fu) * 12 + 1056, ax, pos);
  if (k <= 0x3fed7fffu) return row<8>(T, (int)((k >> 13) & 0x7fu) * 13 + 992, ax, pos);
  if (k <= 0x3fee7fffu) return row<9>(T, (int)((k >> 13) & 0x7fu) * 14 + 884, ax, pos);
  if (k <= 0x3feeffffu) return row<10>(T, (int)((k >> 13) & 0x7fu) * 15 + 768, ax, pos);
  if (k <= 0x3fefffffu) {
    const double z = (pos ? 1.0 - x : x + 1.0) * 0.5;
    const uint64_t uz = as_u64(z);
    const int i1 = (int)(((int64_t)uz >> 46) & 0x7f), i2 = 0x1ff - (int)((int64_t)uz >> 53);
    double t = tab_f64(ELMK_LM_TABLE(inroot_tab), i1) * tab_f64(ELMK_LM_TABLE(powtwo_tab), i2);
    const double r = fma(-(t * t), z, 1.0);
    double q = fma(r, rt3, rt2);
    q = fma(r, q, rt1);
    q = fma(r, q, rt0);
    t = q * t;
    const double c = z * t;
    const double h = fma(-c, t * 0.5, 1.5);
    const double w = fma(c, 0x1p27, c);
    const double cc = fma(-0x1p27, c, w);
    const double den = fma(h, c, cc);
    const double cor = fma(-cc, cc, z) / den;
    double p = fma(z, f6, f5);
    p = fma(z, p, f4);
    p = fma(z, p, f3);
    p = fma(z, p, f2);
    p = fma(z, p, f1);
    p = (p * z) * (cc + cor);
    if (pos) {
      const double s = (cor + p) + cc;
      return s + s;
    }
    const double s = ((hp1 - cor) - p) + (hp0 - cc);
    return s + s;
  }
  if (k == 0x3ff00000u && (uint32_t)ux == 0) return pos ? 0.0 : 0x1.921fb54442d18p1;
  return as_f64(0x7ff8000000000000ull);
}

} // namespace lm
} // namespace elmk
