// TEST-ONLY: compares aprilgrid-rs_b200/csrc/ag_libm.h (host build) with the machine's libm.
// usage: libm_port_check <acos stride> <atan stride> <atan2 pairs>; prints three difference counts.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ag_libm.h"

int main(int argc, char** argv) {
  const uint32_t s_acos = argc > 1 ? (uint32_t)atoi(argv[1]) : 997, s_atan = argc > 2 ? (uint32_t)atoi(argv[2]) : 997;
  const long pairs = argc > 3 ? atol(argv[3]) : 5000000;
  long bad_acos = 0, bad_atan = 0, bad_atan2 = 0;
  for (uint64_t u = 0; u <= 0x3f800000u; u += s_acos)
    for (int sg = 0; sg < 2; ++sg) {
      const float x = lm_float((int32_t)((uint32_t)u | (sg ? 0x80000000u : 0u)));
      if (lm_bits(acosf(x)) != lm_bits(lm_acosf(x))) ++bad_acos;
    }
  for (uint64_t u = 0; u < 0x7f800000u; u += s_atan)
    for (int sg = 0; sg < 2; ++sg) {
      const float x = lm_float((int32_t)((uint32_t)u | (sg ? 0x80000000u : 0u)));
      if (lm_bits(atanf(x)) != lm_bits(lm_atanf(x))) ++bad_atan;
    }
  uint32_t s = 777;
  for (long i = 0; i < pairs; ++i) {
    s = s * 1664525u + 1013904223u;
    const uint32_t a = s;
    s = s * 1664525u + 1013904223u;
    const uint32_t b = s;
    float y, x;
    if (i & 1) {  // the magnitudes the detector sees
      y = (float)((int32_t)a) / 2147483648.0f * 3.0f;
      x = (float)((int32_t)b) / 2147483648.0f * 3.0f;
    } else {  // any bit patterns (zeros, infinities, denormals, huge ratios)
      y = lm_float((int32_t)a);
      x = lm_float((int32_t)b);
      if (y != y || x != x) continue;
    }
    if (lm_bits(atan2f(y, x)) != lm_bits(lm_atan2f(y, x))) ++bad_atan2;
  }
  printf("%ld %ld %ld\n", bad_acos, bad_atan, bad_atan2);
  return 0;
}
