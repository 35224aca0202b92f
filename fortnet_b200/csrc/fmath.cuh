// fmath.cuh -- short FP64 exp / log for the ACSF pair loops.
//
// The angular sums evaluate (1 + lam cos)^xi = exp(xi * log(1 + lam cos)) for every neighbour
// pair (the reference calls `**`, lib_descriptors/acsf.F90:1431,1487), so the kernels are bound
// by the instruction count of these two functions.  CUDA's exp()/log() spend about 40 / 70
// SASS instructions per call (sub-ulp rounding, denormals, NaN/inf plumbing); the versions
// below need about 20 / 35 and are accurate to ~2 ulp on the domain the kernels use:
//   fnet_exp(x): x <= 700 (x < -708 and -inf flush to 0; no overflow handling)
//   fnet_log(x): x >= 0   (0 and denormals return -inf; no negative / NaN handling)
// The file also compiles as plain C++ (tests/cpp/fmath_check.cpp checks both against libm).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define FNET_HD __host__ __device__ __forceinline__
#else
#define FNET_HD static inline
#endif
#ifdef __CUDA_ARCH__
#define FNET_UNROLL _Pragma("unroll")
#else
#define FNET_UNROLL
#endif


// Polynomial coefficients live in constant memory on the device: a DFMA then takes the
// coefficient straight from the constant bank (one issue slot per Horner step) instead of two
// UMOVs that rebuild a 64-bit immediate in a uniform register.
#define FNET_EXP_COEFFS { 1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, \
  2.755731922398589e-07, 2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, \
  1.388888888888889e-03, 8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, \
  1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10 }
#define FNET_LOG_COEFFS { 4.7619047619047616e-02, 5.2631578947368418e-02, 5.8823529411764705e-02, \
  6.6666666666666666e-02, 7.6923076923076927e-02, 9.0909090909090912e-02, 1.1111111111111110e-01, \
  1.4285714285714285e-01, 2.0000000000000001e-01, 3.3333333333333331e-01, \
  1.90821492927058770002e-10, 6.93147180369123816490e-01 }
#ifdef __CUDACC__
__constant__ double fnet_exp_cd[14] = FNET_EXP_COEFFS;
__constant__ double fnet_log_cd[12] = FNET_LOG_COEFFS;
#endif
static const double fnet_exp_ch[14] = FNET_EXP_COEFFS;
static const double fnet_log_ch[12] = FNET_LOG_COEFFS;
#ifdef __CUDA_ARCH__
#define FNET_EC(i) fnet_exp_cd[i]
#define FNET_LC(i) fnet_log_cd[i]
#else
#define FNET_EC(i) fnet_exp_ch[i]
#define FNET_LC(i) fnet_log_ch[i]
#endif

FNET_HD double fnet_mk_double(int hi, int lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, lo);
#else
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d; memcpy(&d, &u, 8); return d;
#endif
}
FNET_HD int fnet_hi(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32);
#endif
}
FNET_HD int fnet_lo(double x) {
#ifdef __CUDA_ARCH__
  return __double2loint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(u & 0xffffffffu);
#endif
}

// exp(x) = 2^k * e^r, k = rint(x log2 e), r = x - k ln2 (two-part ln2), e^r by its degree-13
// Taylor polynomial on |r| <= 0.3466 (truncation 4e-18 relative).
FNET_HD double fnet_exp(double x) {
  const double magic = 6755399441055744.0;                 // 1.5 * 2^52: rint() through the adder
  const double tk = fma(x, FNET_EC(11), magic);
  const int k = fnet_lo(tk);
  const double kd = tk - magic;
  double r = fma(kd, FNET_EC(12), x);                      // ln2 high part: 32 significant bits
  r = fma(kd, FNET_EC(13), r);                             // ln2 low part
  double p = FNET_EC(0);                                   // 1/13!
FNET_UNROLL
  for (int i = 1; i <= 10; i++) p = fma(p, r, FNET_EC(i)); // 1/12! ... 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double s = fnet_mk_double((k + 1023) << 20, 0);    // 2^k, k in [-1021, 1010]
  return (x < -708.0) ? 0.0 : p * s;
}

// log(x) = e ln2 + 2 atanh(s), x = 2^e m, m in [sqrt(1/2), sqrt(2)), s = (m-1)/(m+1),
// 2 atanh(s) = 2s (1 + z/3 + z^2/5 + ... + z^10/21), z = s^2 <= 0.02944 (truncation 6e-19 rel.).
FNET_HD double fnet_log(double x) {
  int hi = fnet_hi(x);
  const int lo = fnet_lo(x);
  int e = (hi >> 20) - 1023;
  const bool tiny = hi < 0x00100000;                        // zero / denormal (x >= 0 assumed)
  hi = (hi & 0x000fffff) | 0x3ff00000;                      // m in [1, 2)
  if (hi >= 0x3ff6a09f) { hi -= 0x00100000; e += 1; }       // m >= sqrt(2): halve
  const double m = fnet_mk_double(hi, lo);
  const double f = m - 1.0, d = m + 1.0;
#ifdef __CUDA_ARCH__
  double rc;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(d));   // ~2^-23
  double er = fma(-d, rc, 1.0);
  rc = fma(rc, er, rc);
  er = fma(-d, rc, 1.0);
  rc = fma(rc, er, rc);                                     // ~2^-92
  double s = f * rc;
  s = fma(fma(-d, s, f), rc, s);                            // correctly rounded up to the last bit or so
#else
  const double s = f / d;
#endif
  const double z = s * s;
  double q = FNET_LC(0);                                   // 1/21
FNET_UNROLL
  for (int i = 1; i <= 9; i++) q = fma(q, z, FNET_LC(i));  // 1/19 ... 1/3
  const double s2 = s + s;
  const double ed = (double)e;
  // e*ln2_hi is exact (ln2_hi has 32 significant bits, |e| < 2^11)
  double res = fma(ed, FNET_LC(10), (s2 * z) * q);
  res = res + s2;
  res = fma(ed, FNET_LC(11), res);
  return tiny ? -INFINITY : res;
}

// 1/d for d >= 1 (no zero / inf / denormal handling): hardware seed + two Newton steps + a
// residual correction; within 1 ulp.
FNET_HD double fnet_rcp(double d) {
#ifdef __CUDA_ARCH__
  double rc;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(d));
  double er = fma(-d, rc, 1.0);
  rc = fma(rc, er, rc);
  er = fma(-d, rc, 1.0);
  rc = fma(rc, er, rc);
  er = fma(-d, rc, 1.0);
  return fma(rc, er, rc);
#else
  return 1.0 / d;
#endif
}

// tanh(x) = 1 - 2 / (exp(2x) + 1); |x| < 2^-9 uses x - x^3/3 (relative error < 2e-12 at the
// switch-over from the truncation, ~1e-13 from the cancellation in the closed form above it).
FNET_HD double fnet_tanh(double x) {
  const double e2 = fnet_exp(fmin(x + x, 700.0));
  const double big = 1.0 - 2.0 * fnet_rcp(e2 + 1.0);
  const double x2 = x * x;
  const double small = fma(x * x2, fma(x2, 0.13333333333333333, -0.33333333333333331), x);
  return (fabs(x) < 0.001953125) ? small : big;
}
