/* Exhaustive-style check of the constant-divisor sequence used by elmkernels_b200/ptx_rewrite.py:
 *     q = a * y;  r = fma(-c, q, a);  q' = fma(r, y, q)      with y = RN(1 / c)
 * against IEEE division a / c, for every constant divisor the device code contains (argv: hex doubles as PTX
 * writes them, 0dXXXXXXXXXXXXXXXX) and N numerators each: random mantissas across the guarded exponent range,
 * numerators q * c perturbed by a few ulps, and numerators whose quotient lies next to a midpoint between
 * two doubles (the hard cases for rounding; with y off by one ulp the check does report mismatches there).
 * Markstein's theorem (y correctly rounded, q faithful) says the result is the correctly rounded quotient; this
 * program is the empirical side of that argument.  Build: gcc -O2 -mfma -o check_constdiv check_constdiv.c -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static uint64_t rnd(void) { uint64_t a = s[0], b = s[1]; s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26); return s[1] + b; }
static double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static uint64_t to_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

int main(int argc, char** argv)
{
  long n = 20000000, bad = 0, total = 0;
  for (int k = 1; k < argc; ++k) {
    if (!strncmp(argv[k], "-n", 2)) { n = atol(argv[k] + 2); continue; }
    const double c = from_bits(strtoull(argv[k] + 2, NULL, 16));
    const double y = 1.0 / c;
    long b = 0;
    for (long i = 0; i < n; ++i) {
      double a;
      const uint64_t r = rnd();
      if (i % 3 == 2) {
        /* hard cases: numerators whose quotient lies next to a midpoint between two doubles:
           a = RN((q + ulp/2) * c), nudged by -1..1 ulp */
        const uint64_t e = 1023 - 200 + (r >> 52) % 401;
        const double q = from_bits((r & 0x000FFFFFFFFFFFFFull) | (e << 52));
        const long double m = (long double)q + 0.5L * ((long double)from_bits(to_bits(q) + 1) - (long double)q);
        a = from_bits(to_bits((double)(m * (long double)c)) + (int64_t)((r >> 40) % 3) - 1);
      } else if (i & 1) {
        /* random sign, mantissa, exponent in the guarded range [2^-500, 2^500] */
        const uint64_t e = 1023 - 500 + (r >> 52) % 1001;
        a = from_bits((r & 0x800FFFFFFFFFFFFFull) | (e << 52));
      } else {
        /* a quotient with random mantissa, mapped back: a = (q * c) nudged by -2..2 ulps -> a / c lands close to q or
           to the midpoint between neighbours of q */
        const uint64_t e = 1023 - 200 + (r >> 52) % 401;
        const double q = from_bits((r & 0x000FFFFFFFFFFFFFull) | (e << 52));
        a = from_bits(to_bits(q * c) + (int64_t)((r >> 40) % 5) - 2);
      }
      const volatile double q0 = a * y;
      const double rr = fma(-c, q0, a);
      const double q1 = fma(rr, y, q0);
      const double ref = a / c;
      if (to_bits(q1) != to_bits(ref)) { if (b < 3) fprintf(stderr, "MISMATCH c=%a a=%a got %a want %a\n", c, a, q1, ref); ++b; }
    }
    printf("%s c=%.17g: %ld numerators, %ld mismatches\n", argv[k], c, n, b);
    bad += b; total += n;
  }
  printf("total %ld numerators, %ld mismatches\n", total, bad);
  return bad != 0;
}
