// acosf / atanf / atan2f exactly as glibc 2.39 computes them (sysdeps/ieee754/flt-32/e_acosf.c,
// s_atanf.c, e_atan2f.c -- the fdlibm single-precision routines: f32 arithmetic only, no FMA on
// x86-64, no ifunc variants), restated operation for operation so that host and device produce
// the SAME BITS as the `acosf` / `atan2f` a Rust binary of the reference calls on this platform
// (f32::acos / f32::atan2, src/detector.rs:348-349; math_util.rs:31).  theta and phi of a saddle
// feed discontinuous gates (phi in [30, 60], |theta - theta'| < 5 / > 80, round(theta)), so
// "within an ulp" is not good enough.
//
// Pinned by tests/test_libm_port.py against the machine's libm: atanf over every positive float,
// acosf and atan2f over hundreds of millions of arguments -- zero differences.  (glibc >= 2.41
// ships correctly rounded versions instead; the oracle calls the platform libm, so on such a
// platform that test fails loudly and this header is the thing to update.)
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LM_FN __host__ __device__ __forceinline__
#else
#define LM_FN static inline
#endif

// every operation rounded separately (no contraction), in both builds
#if defined(__CUDA_ARCH__)
LM_FN float lm_mul(float a, float b) { return __fmul_rn(a, b); }
LM_FN float lm_add(float a, float b) { return __fadd_rn(a, b); }
LM_FN float lm_sub(float a, float b) { return __fsub_rn(a, b); }
LM_FN float lm_div(float a, float b) { return __fdiv_rn(a, b); }
LM_FN float lm_sqrt(float a) { return __fsqrt_rn(a); }
LM_FN int32_t lm_bits(float f) { return __float_as_int(f); }
LM_FN float lm_float(int32_t i) { return __int_as_float(i); }
#else
LM_FN float lm_mul(float a, float b) { volatile float r = a * b; return r; }
LM_FN float lm_add(float a, float b) { volatile float r = a + b; return r; }
LM_FN float lm_sub(float a, float b) { volatile float r = a - b; return r; }
LM_FN float lm_div(float a, float b) { volatile float r = a / b; return r; }
LM_FN float lm_sqrt(float a) { volatile float r = __builtin_sqrtf(a); return r; }
LM_FN int32_t lm_bits(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
LM_FN float lm_float(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
#endif

// e_acosf.c
LM_FN float lm_acosf(float x) {
  const float one = 1.0f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f,
              pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
              pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
              qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
  const int32_t hx = lm_bits(x), ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) return hx > 0 ? 0.0f : lm_add(pi, lm_mul(2.0f, pio2_lo));
  if (ix > 0x3f800000) return lm_div(lm_sub(x, x), lm_sub(x, x));
  float z, p, q, r, w, s, c, df;
#define LM_P(z) lm_mul(z, lm_add(pS0, lm_mul(z, lm_add(pS1, lm_mul(z, lm_add(pS2, lm_mul(z, lm_add(pS3, lm_mul(z, lm_add(pS4, lm_mul(z, pS5)))))))))))
#define LM_Q(z) lm_add(one, lm_mul(z, lm_add(qS1, lm_mul(z, lm_add(qS2, lm_mul(z, lm_add(qS3, lm_mul(z, qS4))))))))
  if (ix < 0x3f000000) {
    if (ix <= 0x23000000) return lm_add(pio2_hi, pio2_lo);
    z = lm_mul(x, x);
    p = LM_P(z); q = LM_Q(z); r = lm_div(p, q);
    return lm_sub(pio2_hi, lm_sub(x, lm_sub(pio2_lo, lm_mul(x, r))));
  } else if (hx < 0) {
    z = lm_mul(lm_add(one, x), 0.5f);
    p = LM_P(z); q = LM_Q(z);
    s = lm_sqrt(z);
    r = lm_div(p, q);
    w = lm_sub(lm_mul(r, s), pio2_lo);
    return lm_sub(pi, lm_mul(2.0f, lm_add(s, w)));
  } else {
    z = lm_mul(lm_sub(one, x), 0.5f);
    s = lm_sqrt(z);
    df = lm_float(lm_bits(s) & (int32_t)0xfffff000);
    c = lm_div(lm_sub(z, lm_mul(df, df)), lm_add(s, df));
    p = LM_P(z); q = LM_Q(z);
    r = lm_div(p, q);
    w = lm_add(lm_mul(r, s), c);
    return lm_mul(2.0f, lm_add(df, w));
  }
}

#undef LM_P
#undef LM_Q

// s_atanf.c (glibc returns the pi/2 constant from |x| >= 2^25 on)
LM_FN float lm_atanf(float x) {
  // atanhi / atanlo: atan(0.5), atan(1), atan(1.5), atan(inf) split in two floats
  const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
  const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
              aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
              aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
  const float one = 1.0f;
  const int32_t hx = lm_bits(x), ix = hx & 0x7fffffff;
  int reduced = 1;
  float hi = 0.0f, lo = 0.0f;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return lm_add(x, x);
    return hx > 0 ? lm_add(hi3, lo3) : lm_sub(-hi3, lo3);
  }
  if (ix < 0x3ee00000) {  // |x| < 0.4375
    if (ix < 0x31000000) return x;
    reduced = 0;
  } else {
    x = lm_float(ix);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) { hi = hi0; lo = lo0; x = lm_div(lm_sub(lm_mul(2.0f, x), one), lm_add(2.0f, x)); }
      else { hi = hi1; lo = lo1; x = lm_div(lm_sub(x, one), lm_add(x, one)); }
    } else {
      if (ix < 0x401c0000) { hi = hi2; lo = lo2; x = lm_div(lm_sub(x, 1.5f), lm_add(one, lm_mul(1.5f, x))); }
      else { hi = hi3; lo = lo3; x = lm_div(-1.0f, x); }
    }
  }
  const float z = lm_mul(x, x), w = lm_mul(z, z);
  const float s1 = lm_mul(z, lm_add(aT0, lm_mul(w, lm_add(aT2, lm_mul(w, lm_add(aT4, lm_mul(w, lm_add(aT6, lm_mul(w, lm_add(aT8, lm_mul(w, aT10)))))))))));
  const float s2 = lm_mul(w, lm_add(aT1, lm_mul(w, lm_add(aT3, lm_mul(w, lm_add(aT5, lm_mul(w, lm_add(aT7, lm_mul(w, aT9)))))))));
  if (!reduced) return lm_sub(x, lm_mul(x, lm_add(s1, s2)));
  const float zz = lm_sub(hi, lm_sub(lm_sub(lm_mul(x, lm_add(s1, s2)), lo), x));
  return hx < 0 ? -zz : zz;
}

// e_atan2f.c
LM_FN float lm_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  const int32_t hx = lm_bits(x), ix = hx & 0x7fffffff, hy = lm_bits(y), iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return lm_add(x, y);
  if (hx == 0x3f800000) return lm_atanf(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0: case 1: return y;
      case 2: return lm_add(pi, tiny);
      default: return lm_sub(-pi, tiny);
    }
  }
  if (ix == 0) return hy < 0 ? lm_sub(-pi_o_2, tiny) : lm_add(pi_o_2, tiny);
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return lm_add(pi_o_4, tiny);
        case 1: return lm_sub(-pi_o_4, tiny);
        case 2: return lm_add(lm_mul(3.0f, pi_o_4), tiny);
        default: return lm_sub(lm_mul(-3.0f, pi_o_4), tiny);
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return lm_add(pi, tiny);
        default: return lm_sub(-pi, tiny);
      }
    }
  }
  if (iy == 0x7f800000) return hy < 0 ? lm_sub(-pi_o_2, tiny) : lm_add(pi_o_2, tiny);
  const int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = lm_add(pi_o_2, lm_mul(0.5f, pi_lo));
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = lm_atanf(lm_float(lm_bits(lm_div(y, x)) & 0x7fffffff));
  switch (m) {
    case 0: return z;
    case 1: return lm_float(lm_bits(z) ^ (int32_t)0x80000000);
    case 2: return lm_sub(pi, lm_sub(z, pi_lo));
    default: return lm_sub(lm_sub(z, pi_lo), pi);
  }
}
