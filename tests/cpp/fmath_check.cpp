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
  printf("FMATH_OK %.3e %.3e %.3e\n", me, ml, mt);
  return (me < 1e-14 && ml < 1e-14 && mt < 5e-13) ? 0 : 2;
}
