// Host-side check of every register codelet in dft.cuh against a naive double DFT.
// Built by `make selftest`, run by tests/test_codelets.py (no GPU needed: the codelets are
// __host__ __device__).
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../dft.cuh"

using namespace b200fft;

template <int R, bool INV>
double check() {
  float2 x[R];
  double xr[R], xi[R];
  unsigned s = 12345u + R * 7 + INV;
  for (int i = 0; i < R; ++i) {
    s = s * 1664525u + 1013904223u;
    xr[i] = ((s >> 8) & 0xffff) / 65536.0 - 0.5;
    s = s * 1664525u + 1013904223u;
    xi[i] = ((s >> 8) & 0xffff) / 65536.0 - 0.5;
    x[i] = make_float2((float)xr[i], (float)xi[i]);
    xr[i] = x[i].x;
    xi[i] = x[i].y;
  }
  Dft<R, INV>::run(x);
  double worst = 0, scale = 0;
  for (int k = 0; k < R; ++k) {
    double re = 0, im = 0;
    for (int n = 0; n < R; ++n) {
      const double th = (INV ? 2.0 : -2.0) * M_PI * (double)((long long)n * k % R) / R;
      re += xr[n] * cos(th) - xi[n] * sin(th);
      im += xr[n] * sin(th) + xi[n] * cos(th);
    }
    worst = fmax(worst, fmax(fabs(re - x[k].x), fabs(im - x[k].y)));
    scale = fmax(scale, fmax(fabs(re), fabs(im)));
  }
  return worst / scale;
}

// the real-input form of the odd-prime codelets (stage 0 of real-input transforms) against the general form
template <int R, bool INV>
double check_real() {
  float2 x[R], y[R];
  unsigned s = 999u + R;
  for (int i = 0; i < R; ++i) {
    s = s * 1664525u + 1013904223u;
    x[i] = make_float2(((s >> 8) & 0xffff) / 65536.0f - 0.5f, 0.f);
    y[i] = x[i];
  }
  Dft<R, INV>::run(x);
  Dft<R, INV>::run_real(y);
  double worst = 0;
  for (int k = 0; k < R; ++k) worst = fmax(worst, fmax(fabs((double)x[k].x - y[k].x), fabs((double)x[k].y - y[k].y)));
  return worst;
}
template <int R>
int run_real_form() {
  static_assert(has_run_real<Dft<R, false>>::value, "odd primes carry a real-input form");
  const double e0 = check_real<R, false>(), e1 = check_real<R, true>();
  const int bad = !(e0 < 2e-6 && e1 < 2e-6);
  printf("radix %3d  real-input form vs general: fwd %.2e  inv %.2e %s\n", R, e0, e1, bad ? "FAIL" : "ok");
  return bad;
}

template <int R>
int run() {
  const double e0 = check<R, false>(), e1 = check<R, true>();
  const int bad = !(e0 < 2e-6 && e1 < 2e-6);
  printf("radix %3d  fwd %.2e  inv %.2e %s\n", R, e0, e1, bad ? "FAIL" : "ok");
  return bad;
}

int main() {
  int bad = 0;
  bad += run<2>(); bad += run<3>(); bad += run<4>(); bad += run<5>(); bad += run<6>(); bad += run<7>();
  bad += run<8>(); bad += run<9>(); bad += run<10>(); bad += run<11>(); bad += run<12>(); bad += run<13>();
  bad += run<15>(); bad += run<16>(); bad += run<17>(); bad += run<20>(); bad += run<25>(); bad += run<31>();
  bad += run<32>();
  bad += run_real_form<3>(); bad += run_real_form<5>(); bad += run_real_form<7>(); bad += run_real_form<13>();
  bad += run_real_form<31>();
  static_assert(!has_run_real<Dft<8, false>>::value && !has_run_real<Dft<15, false>>::value, "");
  // exactness of the compile-time trig at the octants
  static_assert(Tw<0, 8, false>::re == 1.0f && Tw<0, 8, false>::im == 0.0f, "");
  static_assert(Tw<2, 8, false>::re == 0.0f && Tw<2, 8, false>::im == -1.0f, "");
  static_assert(Tw<4, 8, false>::re == -1.0f && Tw<4, 8, false>::im == 0.0f, "");
  static_assert(Tw<6, 8, true>::re == 0.0f && Tw<6, 8, true>::im == -1.0f, "");
  printf(bad ? "FAILED\n" : "all codelets ok\n");
  return bad;
}
