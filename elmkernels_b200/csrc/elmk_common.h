// elmk_common.h - shared definitions of the column-physics core.
//
// The physics core (phys_*.h) is single-source: the CUDA kernels in k_*.cu instantiate it for
// sm_100a; oracle/port compiles the same headers with g++ as a CPU checker of the source itself
// (test infrastructure only - the product library has no host path).
#pragma once
#include <math.h>
#include <stdint.h>

#include "elmk_libm.h"

#if defined(__CUDACC__)
#define ELMK_HD __host__ __device__ __forceinline__
#define ELMK_HD_NOINLINE __host__ __device__ __noinline__
#else
#define ELMK_HD inline
#define ELMK_HD_NOINLINE inline
#endif

// Transcendentals are CALLED, not inlined, on the device: CUDA's double-precision pow/exp/log expand to
// 60-350 SASS instructions per call site, and the column kernels have ~70 call sites inside divergent
// loops.  Inlined, the CanopyFluxes iteration body was 390 KB of SASS against a 32 KB L1.5 instruction
// cache and the kernel spent >20 issue slots per instruction waiting for instruction fetch (ncu
// "no_instruction" stall, profiles/r1_baseline_raw.csv).  One shared copy per function keeps the loop
// bodies cache-resident.
namespace elmk {
// On the device exp / log / log10 / pow / atan / cos / acos / tanh / erf are the restatements of glibc's routines in
// elmk_libm.h: the same bits as the libm the reference calls (pinned by tests/test_libm_cpu.py), so that threshold tests downstream take the
// reference's branch.  The host checker build (oracle/port) calls libm itself.
#if defined(__CUDA_ARCH__)
ELMK_HD_NOINLINE double m_exp(double x) { return lm::g_exp(x); }
ELMK_HD_NOINLINE double m_log(double x) { return lm::g_log(x); }
ELMK_HD_NOINLINE double m_log10(double x) { return lm::g_log10(x); }
ELMK_HD_NOINLINE double m_pow(double x, double y) { return lm::g_pow(x, y); }
ELMK_HD_NOINLINE double m_atan(double x) { return lm::g_atan(x); }
ELMK_HD_NOINLINE double m_cos(double x) { return lm::g_cos(x); }
ELMK_HD_NOINLINE double m_acos(double x) { return lm::g_acos(x); }
ELMK_HD_NOINLINE double m_tanh(double x) { return lm::g_tanh(x); }
ELMK_HD_NOINLINE double m_erf(double x) { return lm::g_erf(x); }
#else
ELMK_HD_NOINLINE double m_exp(double x) { return exp(x); }
ELMK_HD_NOINLINE double m_log(double x) { return log(x); }
ELMK_HD_NOINLINE double m_log10(double x) { return log10(x); }
ELMK_HD_NOINLINE double m_pow(double x, double y) { return pow(x, y); }
ELMK_HD_NOINLINE double m_atan(double x) { return atan(x); }
ELMK_HD_NOINLINE double m_cos(double x) { return cos(x); }
ELMK_HD_NOINLINE double m_acos(double x) { return acos(x); }
ELMK_HD_NOINLINE double m_tanh(double x) { return tanh(x); }
ELMK_HD_NOINLINE double m_erf(double x) { return erf(x); }
#endif
// the same, in line (for the called functions named *_inl, whose independent transcendentals ptxas interleaves)
#if defined(__CUDA_ARCH__) && defined(ELMK_INLINE_LIBM_CALLED)   // experiment: one shared copy after all
ELMK_HD double i_exp(double x) { return m_exp(x); }
ELMK_HD double i_log(double x) { return m_log(x); }
ELMK_HD double i_atan(double x) { return m_atan(x); }
#elif defined(__CUDA_ARCH__)
ELMK_HD double i_exp(double x) { return lm::g_exp(x); }
ELMK_HD double i_log(double x) { return lm::g_log(x); }
ELMK_HD double i_atan(double x) { return lm::g_atan(x); }
#else
ELMK_HD double i_exp(double x) { return exp(x); }
ELMK_HD double i_log(double x) { return log(x); }
ELMK_HD double i_atan(double x) { return atan(x); }
#endif

// IEEE double division with a short cut for a zero numerator.  ptxas expands every `a / b` inline into a
// Newton sequence whose fast path excludes zero and subnormal numerators; those go through a ~70-instruction
// slow path, and column state is full of exact zeros (no snow, dry canopy, frozen soil, night).  For a == 0 and
// a finite non-zero b the quotient is the zero with the sign of the product, i.e. exactly a * b.  The build
// (elmkernels_b200/ptx_rewrite.py) routes every `/` of the device code through this one function, so the
// sources keep the plain operator; the host checker build never sees the short cut.
ELMK_HD_NOINLINE double m_div(double a, double b)
{
#if defined(__CUDA_ARCH__)
  if (__builtin_expect(a == 0.0, 0)) {
    const double m = fabs(b);
    if (m > 0.0 && m <= 1.7976931348623157e308) return a * b;
  }
#endif
  return a / b;
}

// Two independent divisions in one call.  As separate calls of m_div two divisions run strictly one after the other
// - ~9 dependent FP64 operations each - although neither needs the other's result; inside one function ptxas
// interleaves the two Newton sequences, so the pair costs about the latency of one division and one call overhead.
// The build pairs neighbouring independent `/` of the device code (ptx_rewrite.py); results are those of a / b.
struct Div2 {
  double q0, q1;
};
ELMK_HD_NOINLINE Div2 m_div2(double a0, double b0, double a1, double b1)
{
#if defined(__CUDA_ARCH__)
  if (__builtin_expect(a0 == 0.0 || a1 == 0.0, 0)) return {m_div(a0, b0), m_div(a1, b1)};
#endif
  return {a0 / b0, a1 / b1};
}
} // namespace elmk

namespace elmk {

// ---- dimensions (reference src/data/elm_constants.h:84-98) ----
constexpr int NLEVSNO = 5;       // snow layer slots
constexpr int NLEVGRND = 15;     // soil layers
constexpr int NLEVTOT = 20;      // snow + soil
constexpr int NLEVSOI = 10;      // hydrologically active soil layers
constexpr int NLEVBED = 15;      // layers to bedrock
constexpr int NUMRAD = 2;        // VIS, NIR
constexpr int NBND_SNW = 5;      // SNICAR spectral bands
constexpr int NAER = 8;          // aerosol species in snow
constexpr int NROWS = 21;        // unknowns of the temperature system: 5 snow + surface water + 15 soil

// ---- physical constants (reference src/data/elm_constants.h:18-52) ----
constexpr double TFRZ = 273.15;
constexpr double PI = 3.14159265358979323846;
constexpr double BOLTZ = 1.38065e-23;
constexpr double AVOGAD = 6.02214e26;
constexpr double MWWV = 18.016;
constexpr double RGAS = AVOGAD * BOLTZ;
constexpr double RWV = RGAS / MWWV;
constexpr double STEBOL = 5.67e-8;
constexpr double MWDAIR = 28.966;
constexpr double RAIR = RGAS / MWDAIR;
constexpr double GRAV = 9.80616;
constexpr double ROVERG = RWV / GRAV * 1000.;
constexpr double O2_MOLAR_CONST = 0.209;
constexpr double CO2_PPMV = 355.0;
constexpr double DENICE = 0.917e3;
constexpr double DENH2O = 1.000e3;
constexpr double HVAP = 2.501e6;
constexpr double HFUS = 3.337e5;
constexpr double HSUB = HVAP + HFUS;
constexpr double VKC = 0.4;
constexpr double CPAIR = 1.00464e3;
constexpr double CPICE = 2.11727e3;
constexpr double CPWAT = 4.188e3;
constexpr double CSOILC = 0.004;
constexpr double ZLND = 0.01;
constexpr double ZSNO = 0.0024;
constexpr double SNW_RDS_MIN = 54.526;
constexpr double H2OSNO_MAX = 1000.0;

// land-unit keys (reference src/data/land_data.h:7-20)
constexpr int ISTSOIL = 1;
constexpr int ISTCROP = 2;
// PFT keys used by the hot path (elm_constants.h:56-80)
constexpr int PFT_SOYBEAN = 23;
constexpr int PFT_SOYBEAN_IRRIG = 24;

// ---- per-column error bits (values mirror include/elmk_b200.h) ----
constexpr uint32_t ERR_CANOPY_LAYER = 1u << 0;
constexpr uint32_t ERR_SNICAR_RADIUS = 1u << 1;
constexpr uint32_t ERR_SNICAR_NEGABS = 1u << 2;
constexpr uint32_t ERR_SNICAR_ENERGY = 1u << 3;
constexpr uint32_t ERR_SNICAR_ALBEDO = 1u << 4;
constexpr uint32_t ERR_SABG_LAYERS = 1u << 5;
constexpr uint32_t ERR_FORC_HEIGHT = 1u << 6;
constexpr uint32_t ERR_QUADRATIC = 1u << 7;
constexpr uint32_t ERR_BRENT_BRACKET = 1u << 8;
constexpr uint32_t ERR_NEG_STOMATAL = 1u << 9;
constexpr uint32_t ERR_SNOWAGE_DR = 1u << 10;
constexpr uint32_t ERR_DIVIDE_RADIUS = 1u << 11;

// base^y for a constant base: pow with its logarithm half folded (elmk_libm.h, g_pow_cbase) - the value is pow's own.
// The host checker build keeps the reference's pow call.
ELMK_HD double pow_cbase(const double base, const double lhi, const double llo, const double y)
{
#if defined(__CUDA_ARCH__)
  return lm::g_pow_cbase(base, lhi, llo, y);
#else
  (void)lhi; (void)llo;
  return m_pow(base, y);
#endif
}
#define ELMK_LN_TKWAT ELMK_POWLOG_0_57
#define ELMK_LN_TKICE ELMK_POWLOG_2_29
#define ELMK_LN_2 ELMK_POWLOG_2_0

// Re-alignment point for the warps of a block inside a long straight-line kernel body (device only, and only in the
// instantiation whose launch keeps every thread of the block alive to the end): the soil-temperature body is ~270 KB
// of SASS executed once from top to bottom, and warps that walk it together share each instruction-cache line.
#if defined(__CUDA_ARCH__) && !defined(ELMK_NO_REALIGN)
#define ELMK_REALIGN(on) do { if (on) __syncthreads(); } while (0)
#else
#define ELMK_REALIGN(on) ((void)0)
#endif

// ---- arithmetic helpers ----
// The reference is written with std::min/std::max; their NaN and signed-zero behaviour
// ((b < a) ? b : a and (a < b) ? b : a) differs from fmin/fmax, so it is spelled out.
ELMK_HD double dmin(double a, double b) { return (b < a) ? b : a; }
ELMK_HD double dmax(double a, double b) { return (a < b) ? b : a; }
ELMK_HD int imin(int a, int b) { return (b < a) ? b : a; }
ELMK_HD int imax(int a, int b) { return (a < b) ? b : a; }

// Integer powers.  The reference calls pow(x, 2.0|3.0|4.0); glibc's pow is correctly rounded in
// practice, so x*x is bit-identical to pow(x, 2.0), while products for the cube and the fourth
// power can differ from it in the last bit.  ELMK_EXACT_POW (set by the CPU checker build) keeps
// the libm call so that the checker can be compared bit-for-bit with the reference; the CUDA
// build uses multiplications (3 DMUL instead of a ~100-instruction pow), inside the 1e-12 bar.
ELMK_HD double sq(double x) { return x * x; }
#if defined(ELMK_EXACT_POW) || !defined(ELMK_FAST_POW)
ELMK_HD double cube(double x) { return m_pow(x, 3.0); }
ELMK_HD double pow4(double x) { return m_pow(x, 4.0); }
#else
ELMK_HD double cube(double x) { return x * x * x; }
ELMK_HD double pow4(double x) { const double x2 = x * x; return x2 * x2; }
#endif

} // namespace elmk
