// libm_check.cc - pins csrc/elmk_libm.h (host build) to the system libm bit for bit.
// Build: g++ -O2 -std=c++17 -mfma -ffp-contract=off libm_check.cc -o libm_check -lm     Run: libm_check [N per range]
// Prints one line per function / argument range: "<fn> <range> n=<count> mismatches=<count> [first bad argument]".
// Exit status 0 when every result equals libm's.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../elmkernels_b200/csrc/elmk_libm.h"

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static inline uint64_t rnd() {   // splitmix64
  uint64_t z = (rng_state += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static inline double uni(double lo, double hi) { return lo + (hi - lo) * ((rnd() >> 11) * 0x1p-53); }
// log-uniform magnitude in [2^e0, 2^e1) with random mantissa, optional random sign
static inline double logu(int e0, int e1, bool sign) {
  const uint64_t m = rnd() & 0x000fffffffffffffull;
  const uint64_t e = (uint64_t)(1023 + e0 + (int)(rnd() % (uint64_t)(e1 - e0)));
  uint64_t u = (e << 52) | m;
  if (sign && (rnd() & 1)) u |= 0x8000000000000000ull;
  double x;
  memcpy(&x, &u, 8);
  return x;
}
static inline bool same(double a, double b) {
  uint64_t x, y;
  memcpy(&x, &a, 8);
  memcpy(&y, &b, 8);
  return x == y || (a != a && b != b);
}
static int failures = 0;
template <class G, class R, class A> void run1(const char* fn, const char* range, long n, G g, R ref, A arg) {
  long bad = 0;
  double first = 0;
  for (long i = 0; i < n; ++i) {
    const volatile double x = arg();
    if (!same(g(x), ref(x))) { if (!bad) first = x; ++bad; }
  }
  printf("%-6s %-28s n=%ld mismatches=%ld", fn, range, n, bad);
  if (bad) printf("  first x=%a  got %a  want %a", first, g(first), ref(first));
  printf("\n");
  failures += bad != 0;
}
template <class A> void run2(const char* range, long n, A arg) {
  long bad = 0;
  double fx = 0, fy = 0;
  for (long i = 0; i < n; ++i) {
    volatile double x, y;
    { double a, b; arg(a, b); x = a; y = b; }
    if (!same(elmk::lm::g_pow(x, y), pow(x, y))) { if (!bad) { fx = x; fy = y; } ++bad; }
  }
  printf("%-6s %-28s n=%ld mismatches=%ld", "pow", range, n, bad);
  if (bad) printf("  first x=%a y=%a  got %a  want %a", fx, fy, elmk::lm::g_pow(fx, fy), pow(fx, fy));
  printf("\n");
  failures += bad != 0;
}

int main(int argc, char** argv) {
  const long N = argc > 1 ? atol(argv[1]) : 2000000;
  using namespace elmk::lm;
  auto rexp = [](double x) { return exp(x); };
  auto rlog = [](double x) { return log(x); };
  auto rlog10 = [](double x) { return log10(x); };
  auto ratan = [](double x) { return atan(x); };
  auto gexp = [](double x) { return g_exp(x); };
  auto glog = [](double x) { return g_log(x); };
  auto glog10 = [](double x) { return g_log10(x); };
  auto gatan = [](double x) { return g_atan(x); };
  run1("exp", "uniform [-40, 40]", N, gexp, rexp, [] { return uni(-40, 40); });
  run1("exp", "uniform [-1, 1]", N, gexp, rexp, [] { return uni(-1, 1); });
  run1("exp", "uniform [-745, 710]", N, gexp, rexp, [] { return uni(-745, 710); });
  run1("exp", "log-uniform 2^-60..2^10 +-", N, gexp, rexp, [] { return logu(-60, 10, true); });
  run1("log", "uniform (0, 4]", N, glog, rlog, [] { return uni(1e-300, 4); });
  run1("log", "uniform [0.9, 1.1]", N, glog, rlog, [] { return uni(0.9, 1.1); });
  run1("log", "log-uniform 2^-1022..2^1023", N, glog, rlog, [] { return logu(-1022, 1023, false); });
  run1("log", "subnormal", N / 10, glog, rlog, [] { return uni(0, 2.2e-308); });
  run1("log10", "uniform (0, 2000]", N, glog10, rlog10, [] { return uni(1e-300, 2000); });
  run1("log10", "log-uniform 2^-300..2^300", N, glog10, rlog10, [] { return logu(-300, 300, false); });
  run1("atan", "uniform [-1, 1]", N, gatan, ratan, [] { return uni(-1, 1); });
  run1("atan", "uniform [-16, 16]", N, gatan, ratan, [] { return uni(-16, 16); });
  run1("atan", "uniform [-1000, 1000]", N, gatan, ratan, [] { return uni(-1000, 1000); });
  run1("atan", "log-uniform 2^-40..2^60 +-", N, gatan, ratan, [] { return logu(-40, 60, true); });
  auto rcos = [](double x) { return cos(x); };
  auto gcos = [](double x) { return g_cos(x); };
  run1("cos", "uniform [-pi, pi]", N, gcos, rcos, [] { return uni(-3.14159265358979, 3.14159265358979); });
  run1("cos", "uniform [-0.9, 0.9]", N, gcos, rcos, [] { return uni(-0.9, 0.9); });
  run1("cos", "uniform [-1000, 1000]", N, gcos, rcos, [] { return uni(-1000, 1000); });
  run1("cos", "log-uniform 2^-40..2^26 +-", N, gcos, rcos, [] { return logu(-40, 26, true); });
  auto rtanh = [](double x) { return tanh(x); };
  auto gtanh = [](double x) { return g_tanh(x); };
  auto rerf = [](double x) { return erf(x); };
  auto gerf = [](double x) { return g_erf(x); };
  auto rexpm1 = [](double x) { return expm1(x); };
  auto gexpm1 = [](double x) { return g_expm1(x); };
  run1("expm1", "uniform [-50, 50]", N, gexpm1, rexpm1, [] { return uni(-50, 50); });
  run1("expm1", "uniform [-2, 2]", N, gexpm1, rexpm1, [] { return uni(-2, 2); });
  run1("expm1", "log-uniform 2^-60..2^9 +-", N, gexpm1, rexpm1, [] { return logu(-60, 9, true); });
  run1("tanh", "uniform [-3, 3]", N, gtanh, rtanh, [] { return uni(-3, 3); });
  run1("tanh", "uniform [-25, 25]", N, gtanh, rtanh, [] { return uni(-25, 25); });
  run1("tanh", "log-uniform 2^-60..2^6 +-", N, gtanh, rtanh, [] { return logu(-60, 6, true); });
  run1("erf", "uniform [-7, 7]", N, gerf, rerf, [] { return uni(-7, 7); });
  run1("erf", "uniform [-1.3, 1.3]", N, gerf, rerf, [] { return uni(-1.3, 1.3); });
  run1("erf", "log-uniform 2^-60..2^4 +-", N, gerf, rerf, [] { return logu(-60, 4, true); });
  auto rsin = [](double x) { return sin(x); };
  auto gsin = [](double x) { return g_sin(x); };
  run1("sin", "uniform [-7, 7]", N, gsin, rsin, [] { return uni(-7, 7); });
  run1("sin", "uniform [-1000, 1000]", N, gsin, rsin, [] { return uni(-1000, 1000); });
  run1("sin", "log-uniform 2^-40..2^26 +-", N, gsin, rsin, [] { return logu(-40, 26, true); });
  auto racos = [](double x) { return acos(x); };
  auto gacos = [](double x) { return g_acos(x); };
  run1("acos", "uniform [-1, 1]", N, gacos, racos, [] { return uni(-1, 1); });
  run1("acos", "uniform [0.9, 1]", N, gacos, racos, [] { return uni(0.9, 1); });
  run1("acos", "uniform [-1, -0.9]", N, gacos, racos, [] { return uni(-1, -0.9); });
  run1("acos", "1 - log-uniform 2^-52..2^-3", N, gacos, racos, [] { return 1.0 - logu(-52, -3, false); });
  run1("acos", "log-uniform 2^-60..2^0 +-", N, gacos, racos, [] { return logu(-60, 0, true); });
  run2("x in (0,2], y in [-3,3]", N, [](double& x, double& y) { x = uni(1e-6, 2); y = uni(-3, 3); });
  run2("x in (0,1e3], y in [0,1]", N, [](double& x, double& y) { x = uni(1e-9, 1e3); y = uni(0, 1); });
  run2("x log-u 2^-200..2^200, y +-8", N, [](double& x, double& y) { x = logu(-200, 200, false); y = uni(-8, 8); });
  run2("y in {3,4,0.333,0.45,0.25,1.5}", N, [](double& x, double& y) {
    static const double ys[] = {3.0, 4.0, 0.333, 0.45, 0.25, 1.5, 0.666666666666, -0.5, 0.5};
    x = uni(1e-8, 400);
    y = ys[rnd() % 9];
  });
  run2("x = 2, y in [-12, 12]", N, [](double& x, double& y) { x = 2.0; y = uni(-12, 12); });
  run2("overflow / underflow edge", N / 4, [](double& x, double& y) { x = uni(0.5, 40); y = uni(-400, 400); });
  // pow with a constant base and its logarithm half folded (soil conductivities 0.57^y, 2.29^y; 2^y of photosynthesis)
  {
    const double bases[3] = {0.57, 2.29, 2.0};
    const double folded[3][2] = {{ELMK_POWLOG_0_57}, {ELMK_POWLOG_2_29}, {ELMK_POWLOG_2_0}};
    for (int k = 0; k < 3; ++k) {
      double lhi, llo;
      pow_log(as_u64(bases[k]), lhi, llo);
      const bool same_consts = same(lhi, folded[k][0]) && same(llo, folded[k][1]);
      long bad = 0;
      for (long i = 0; i < N; ++i) {
        const volatile double y = (i & 1) ? uni(-8, 8) : uni(0, 1);
        if (!same(g_pow_cbase(bases[k], folded[k][0], folded[k][1], y), pow(bases[k], y))) ++bad;
      }
      printf("%-6s base %-4g folded log %-10s n=%ld mismatches=%ld\n", "powc", bases[k], same_consts ? "ok" : "STALE", N, bad + !same_consts);
      failures += (bad != 0) || !same_consts;
    }
  }
  return failures ? 1 : 0;
}
