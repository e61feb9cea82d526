/*
 * sweep_libm.c -- exhaustive check: restated glibc logf/expf == the libm the
 * reference links, over the whole input domain of the path.
 * TEST INFRASTRUCTURE ONLY.   make -C oracle sweep && oracle/_build/sweep_libm
 *
 *   logf: h = rnd_real() in {0} U [2^-63, 1]  (src/layer.cpp:136-137); swept
 *         over every float in [2^-126, 1] plus 0.
 *   expf: argument -sig_a*di in [-inf, +0]    (src/layer.cpp:175); swept over
 *         every negative float and +0; what the path consumes is 1 - expf().
 */
#include "mc_oracle.h"
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(void) {
  long n = 0, bad_log = 0, bad_exp = 0, bad_one_minus = 0;
#pragma omp parallel for reduction(+ : n, bad_log)
  for (uint32_t u = 0x00800000u; u <= 0x3f800000u; u++) {
    float x = u2f(u);
    if (f2u(logf(x)) != f2u(orc_logf_restated(x))) bad_log++;
    n++;
  }
  if (f2u(logf(0.0f)) != f2u(orc_logf_restated(0.0f))) bad_log++;
  printf("logf: %ld inputs, %ld mismatches\n", n + 1, bad_log);
  n = 0;
#pragma omp parallel for reduction(+ : n, bad_exp, bad_one_minus)
  for (uint32_t u = 0x80000000u; u <= 0xff800000u; u++) {
    float x = u2f(u);
    float g = expf(x), r = orc_expf_restated(x);
    if (f2u(g) != f2u(r)) {
      bad_exp++;
      if (f2u(1.0f - g) != f2u(1.0f - r)) bad_one_minus++;
    }
    n++;
  }
  if (f2u(expf(0.0f)) != f2u(orc_expf_restated(0.0f))) bad_exp++;
  printf("expf: %ld inputs, %ld mismatches, %ld mismatches of 1-expf\n", n + 1,
         bad_exp, bad_one_minus);
  return (bad_log || bad_one_minus) ? 1 : 0;
}
