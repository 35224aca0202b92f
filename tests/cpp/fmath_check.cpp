// Checks fortnet_b200/csrc/fmath.cuh (host build) against libm on the domains the ACSF kernels use.
// Prints "FMATH_OK <max rel err exp> <max rel err log>"; exits non-zero beyond 1e-14.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../../fortnet_b200/csrc/fmath.cuh"

static double urand(unsigned long long &s) {
  s = s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)(s >> 11) / 9007199254740992.0;
}

int main() {
  unsigned long long seed = 12345;
  double me = 0.0, ml = 0.0;
  for (int i = 0; i < 2000000; i++) {
    double u = urand(seed);
    double x = (i % 4 == 0) ? -700.0 * u : (i % 4 == 1 ? 700.0 * u : (i % 4 == 2 ? -40.0 * u : 3.0 * (u - 0.5)));
    double a = fnet_exp(x), b = exp(x);
    double re = fabs(a - b) / b;
    if (re > me) me = re;
  }
  if (fnet_exp(-709.0) != 0.0 || fnet_exp(-INFINITY) != 0.0 || fnet_exp(0.0) != 1.0) { printf("exp edge cases failed\n"); return 1; }
  for (int i = 0; i < 2000000; i++) {
    double u = urand(seed);
    double x;
    switch (i % 5) {
      case 0: x = 2.0 * u; break;                       // 1 + lam cos
      case 1: x = ldexp(1.0 + u, -(int)(60 * urand(seed))); break;   // small b
      case 2: x = 1.0 + 1e-6 * (u - 0.5); break;        // near 1
      case 3: x = exp(-700.0 * u); break;
      default: x = 0.70710678 + 0.0000001 * (u - 0.5) + (i & 8 ? 0.70710678 : 0.0); break;   // branch points
    }
    if (x <= 0.0) continue;
    double a = fnet_log(x), b = log(x);
    double re = (b == 0.0) ? fabs(a) : fabs(a - b) / fabs(b);
    // near x = 1 the absolute error relative to |x-1| is what matters
    if (fabs(b) < 1e-300) continue;
    if (re > ml) ml = re;
  }
  if (fnet_log(0.0) != -INFINITY || fnet_log(1.0) != 0.0) { printf("log edge cases failed\n"); return 1; }
  double mt = 0.0;
  for (int i = 0; i < 2000000; i++) {
    double u = urand(seed);
    double x = (i % 3 == 0) ? 40.0 * (u - 0.5) : (i % 3 == 1 ? 2.0 * (u - 0.5) : 0.01 * (u - 0.5));
    if (i % 1000 == 7) x = 800.0 * (u - 0.5);
    double a = fnet_tanh(x), b = tanh(x);
    double re = (b == 0.0) ? fabs(a) : fabs(a - b) / fabs(b);
    if (re > mt) mt = re;
  }
  if (fnet_tanh(0.0) != 0.0 || fnet_tanh(1e3) != 1.0 || fnet_tanh(-1e3) != -1.0) { printf("tanh edge cases failed\n"); return 1; }
  // table-driven variants (ACSF pair loops): exp relative, log absolute + relative mix (the power
  // (1 + lam cos)^xi = exp(xi log b) needs a small ABSOLUTE error of the logarithm)
  static double tab[FNET_TAB_DOUBLES];
  for (int i = 0; i < FNET_EXP_TAB_N; i++) tab[i] = fnet_exp_tab_h[i];
  for (int i = 0; i < 2 * FNET_LOG_TAB_N; i++) tab[FNET_EXP_TAB_N + i] = fnet_log_tab_h[i];
  double met = 0.0, mlt = 0.0;
  for (int i = 0; i < 4000000; i++) {
    double u = urand(seed);
    double x = (i % 4 == 0) ? -700.0 * u : (i % 4 == 1 ? 700.0 * u : (i % 4 == 2 ? -40.0 * u : 3.0 * (u - 0.5)));
    double a = fnet_exp_tab(x, tab), b = exp(x);
    double re = fabs(a - b) / b;
    if (re > met) met = re;
  }
  if (fnet_exp_tab(-709.0, tab) != 0.0 || fnet_exp_tab(-INFINITY, tab) != 0.0 || fnet_exp_tab(0.0, tab) != 1.0) { printf("exp_tab edge cases failed\n"); return 1; }
  for (int i = 0; i < 4000000; i++) {
    double u = urand(seed);
    double x;
    switch (i % 6) {
      case 0: x = 2.0 * u; break;
      case 1: x = ldexp(1.0 + u, -(int)(60 * urand(seed))); break;
      case 2: x = 1.0 + 1e-6 * (u - 0.5); break;
      case 3: x = exp(-700.0 * u); break;
      case 4: x = 0.6875 + 0.6875 * u; break;                  // the whole table range
      default: x = 0.70710678 + 0.0000001 * (u - 0.5) + (i & 8 ? 0.70710678 : 0.0); break;
    }
    if (x <= 0.0) continue;
    double a = fnet_log_tab(x, tab), b = log(x);
    double err = fabs(a - b) / (4e-16 + 4e-16 * fabs(b));       // in units of the bound
    if (err > mlt) mlt = err;
  }
  if (fnet_log_tab(0.0, tab) != -INFINITY || fnet_log_tab(1.0, tab) != 0.0) { printf("log_tab edge cases failed\n"); return 1; }
  double mtt = 0.0;
  for (int i = 0; i < 2000000; i++) {
    double u = urand(seed);
    double x = (i % 3 == 0) ? 40.0 * (u - 0.5) : (i % 3 == 1 ? 2.0 * (u - 0.5) : 0.01 * (u - 0.5));
    if (i % 1000 == 7) x = 800.0 * (u - 0.5);
    double a = fnet_tanh_tab(x, tab), b = tanh(x);
    double re = (b == 0.0) ? fabs(a) : fabs(a - b) / fabs(b);
    if (re > mtt) mtt = re;
  }
  if (fnet_tanh_tab(0.0, tab) != 0.0 || fnet_tanh_tab(1e3, tab) != 1.0 || fnet_tanh_tab(-1e3, tab) != -1.0) { printf("tanh_tab edge cases failed\n"); return 1; }
  // expm1-form tanh of the DMMA subnetwork kernels: relative error on the whole axis, incl. tiny arguments
  double mte = 0.0;
  for (int i = 0; i < 4000000; i++) {
    double u = urand(seed);
    double x;
    switch (i % 5) {
      case 0: x = 40.0 * (u - 0.5); break;
      case 1: x = 2.0 * (u - 0.5); break;
      case 2: x = 0.02 * (u - 0.5); break;                       // around the k = 0 / j = 0, 1 table cells
      case 3: x = ldexp(u - 0.5, -(int)(300 * urand(seed))); break;
      default: x = 0.34657359 * (1.0 + 1e-6 * (u - 0.5)) * (double)(1 + (i % 7)); break;   // k boundaries (2x = m ln2)
    }
    if (i % 1000 == 7) x = 800.0 * (u - 0.5);
    double a = fnet_tanh_em1(x, tab), b = tanh(x);
    double re = (b == 0.0) ? fabs(a) : fabs(a - b) / fabs(b);
    if (re > mte) mte = re;
  }
  if (fnet_tanh_em1(0.0, tab) != 0.0 || fnet_tanh_em1(1e3, tab) != 1.0 || fnet_tanh_em1(-1e3, tab) != -1.0 ||
      fnet_tanh_em1(INFINITY, tab) != 1.0 || fnet_tanh_em1(-INFINITY, tab) != -1.0 || fnet_tanh_em1(-1e-310, tab) != -1e-310) {
    printf("tanh_em1 edge cases failed\n"); return 1;
  }
  printf("FMATH_OK %.3e %.3e %.3e tab: %.3e %.3f %.3e em1: %.3e\n", me, ml, mt, met, mlt, mtt, mte);
  return (me < 1e-14 && ml < 1e-14 && mt < 5e-13 && met < 1e-14 && mlt < 1.0 && mtt < 5e-13 && mte < 1e-14) ? 0 : 2;
}
