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
  return rc;
}
