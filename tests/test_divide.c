/* exhaustive check of gca_div_const_f32 / gca_div_const_f32_1 (compiled and run by tests/test_divide.py) */
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include "gca_math.h"
int main(void) {
  const float ds[3] = {800.0f, (float)((8.0 / 3.0) * 2), 1234.5f};
  int rc = 0;
  for (int k = 0; k < 3; k++) {
    const float d = ds[k], inv = 1.0f / d;
    uint64_t bad2 = 0, bad1 = 0, n = 0;
    /* every significand at 2 exponents (the computation is scale invariant), then a sweep of exponents */
    for (uint32_t e = 120; e <= 135; e += 15)
      for (uint32_t m = 0; m < (1u << 23); m++) {
        uint32_t u = (e << 23) | m; float x; memcpy(&x, &u, 4);
        n++; if (gca_div_const_f32(x, d, inv) != x / d) bad2++; if (gca_div_const_f32_1(x, d, inv) != x / d) bad1++;
      }
    for (uint32_t e = 40; e < 215; e++)
      for (uint32_t m = 0; m < (1u << 23); m += 1021) {
        uint32_t u = (e << 23) | m; float x; memcpy(&x, &u, 4);
        n++; if (gca_div_const_f32(x, d, inv) != x / d) bad2++; if (gca_div_const_f32(-x, d, inv) != -x / d) bad2++;
      }
    printf("d=%.9g n=%llu bad_two_step=%llu bad_one_step=%llu is_exact1=%d\n", d, (unsigned long long)n,
           (unsigned long long)bad2, (unsigned long long)bad1, gca_div1_is_exact(d));
    if (bad2) rc = 1;
    if ((bad1 == 0) != (gca_div1_is_exact(d) != 0) && k < 2) rc = 2;
  }
  /* f64: gca_div_const_f64 against IEEE division - dense random significands over many exponents, values
     next to representable quotients' midpoints excluded by nothing: any mismatch is a failure */
  {
    const double dd[6] = {800.0, 1.0, 2.0 * 3.141592653589793, (80.0 / 30.0) * 2.0, 1200.0, 777.123456789};
    uint64_t st = 0x9E3779B97F4A7C15ull, badd = 0, nd = 0;
    for (int k = 0; k < 6; k++) {
      const double d = dd[k], inv = 1.0 / d;
      if (!gca_div_f64_divisor_ok(d)) { printf("divisor %g rejected\n", d); rc = 3; }
      for (int it = 0; it < 4000000; it++) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        uint64_t mant = st & 0xFFFFFFFFFFFFFull;
        uint64_t expo = 1023 - 40 + (uint64_t)(it % 80);
        uint64_t u = (expo << 52) | mant;
        if (it & 1) u |= 0x8000000000000000ull;
        double x; memcpy(&x, &u, 8);
        nd++; if (gca_div_const_f64(x, d, inv) != x / d) badd++;
      }
      /* exact multiples, their neighbours, zeros */
      for (int m = -2000; m <= 2000; m++) {
        const double base = (double)m * d;
        const double xs[3] = {base, nextafter(base, 1e300), nextafter(base, -1e300)};
        for (int j = 0; j < 3; j++) { nd++; if (gca_div_const_f64(xs[j], d, inv) != xs[j] / d) badd++; }
      }
      if (gca_div_const_f64(0.0, d, inv) != 0.0 || !signbit(gca_div_const_f64(-0.0, d, inv))) badd++;
    }
    printf("f64 n=%llu bad=%llu\n", (unsigned long long)nd, (unsigned long long)badd);
    if (badd) rc = 4;
  }
  return rc;
}
