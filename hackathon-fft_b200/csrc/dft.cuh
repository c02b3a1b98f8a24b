// Register-resident radix-R butterflies ("codelets") with compile-time twiddles.
//
// A Stockham stage of radix R (fft/fft/_fft.mojo:228-296) applies an R-point DFT to R
// strided inputs. The reference evaluates it per output point as r-1 complex FMAs
// (O(r^2) per butterfly). Here the whole butterfly lives in one thread's registers:
//   * R = 2, 4           hand-written
//   * odd R (any prime)  symmetric-pair form: a_j = x_j + x_{R-j}, b_j = x_j - x_{R-j},
//                        X_k / X_{R-k} = (x_0 + sum c_jk a_j) -/+ i (sum s_jk b_j):
//                        half the multiplies of the naive form, no Rader/Bluestein
//   * composite R        Cooley-Tukey R = R1*R2 on registers, inner twiddles W_R^{n2*k1}
//                        are compile-time constants (trivial ones strength-reduced like
//                        the reference's _unit_phasor_mul, _utils.mojo:291-317)
// Several reference stages fuse into one codelet: Stockham stages with radices r1..rm at
// processed = P compose to ONE radix-(r1*..*rm) stage at P, so a user's [2,2,2,2] is run
// as a radix-16 codelet.
//
// Everything is __host__ __device__ so tests/test_codelets.py can check every
// instantiation against a naive DFT on the CPU (csrc/codelet_selftest.cu).
#pragma once
#include "rtc_prelude.cuh"

namespace b200fft {

#define B200_HD __host__ __device__ __forceinline__

// ---- compile-time loops -------------------------------------------------------------
template <class F, int... Is>
B200_HD void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
B200_HD void static_for(F&& f) {
  static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

// ---- compile-time trigonometry --------------------------------------------------------
constexpr double kPiOver4 = 0.78539816339744830961566084581987572;

constexpr double cx_sin_small(double x) {  // |x| <= pi/4
  const double x2 = x * x;
  double term = x, sum = x;
  for (int k = 1; k <= 12; ++k) {
    term *= -x2 / double((2 * k) * (2 * k + 1));
    sum += term;
  }
  return sum;
}
constexpr double cx_cos_small(double x) {  // |x| <= pi/4
  const double x2 = x * x;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k <= 12; ++k) {
    term *= -x2 / double((2 * k - 1) * (2 * k));
    sum += term;
  }
  return sum;
}

struct CxD {
  double re, im;
};

// exp(+2*pi*i*k/n) with exact octant reduction (exact at multiples of 1/8 turn up to sqrt(1/2))
constexpr CxD cx_unit(long long k, long long n) {
  k %= n;
  if (k < 0) k += n;
  const long long oct = (8 * k) / n;
  const long long r = 8 * k - oct * n;  // 0 <= r < n, angle = oct*pi/4 + (pi/4)*(r/n)
  const double t = kPiOver4 * (double(r) / double(n));
  const double tc = kPiOver4 * (double(n - r) / double(n));
  const double ct = cx_cos_small(t), st = cx_sin_small(t);
  const double cc = cx_cos_small(tc), sc = cx_sin_small(tc);
  switch (oct) {
    case 0: return {ct, st};
    case 1: return {sc, cc};
    case 2: return {-st, ct};
    case 3: return {-cc, sc};
    case 4: return {-ct, -st};
    case 5: return {-sc, -cc};
    case 6: return {st, -ct};
    default: return {cc, -sc};
  }
}

// W_N^K = exp(-2*pi*i*K/N) (forward) or its conjugate (inverse), as float constants
template <int K, int N, bool INV>
struct Tw {
  static constexpr CxD v = cx_unit(INV ? K : -K, N);
  static constexpr float re = float(v.re);
  static constexpr float im = float(v.im);
};

// ---- complex helpers --------------------------------------------------------------------
// With B200FFT_PACKED (defined per translation unit, before this header): complex add / subtract as ONE
// packed instruction on sm_100a (add.rn.f32x2 -> FADD2 on a 64-bit register pair; same IEEE rounding as
// two scalar adds). Radix-2^k butterflies are ~2/3 complex adds, so this removes about a quarter of all
// issued instructions: FADD 27.7k -> 2.0k + 13.4k FADD2 in the pow2 row kernels' SASS. Measured on B200
// (profiles/r1_packed_fadd2.md) it HURTS the HBM-bound contiguous pow2 kernels (128: 0.151 -> 0.191 ms,
// 1024: 0.237 -> 0.306 ms: FADD2 issues slower than the two FADDs it replaces, which dual-issue across the
// fma/alu pipes) and helps the issue-bound strided / mixed-radix ones (512^3: 1.11 -> 1.005 ms), so it is
// enabled only in the translation units where it wins.
B200_HD float2 cadd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && defined(B200FFT_PACKED)
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
B200_HD float2 csub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && defined(B200FFT_PACKED)
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
B200_HD float2 cmulf(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
// multiply by -i (forward direction) / +i (inverse): the quarter-turn twiddle
template <bool INV>
B200_HD float2 rot90(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// a * W_N^K with the trivial cases strength-reduced
template <int K, int N, bool INV>
B200_HD float2 twiddle_mul(float2 a) {
  constexpr int k = ((K % N) + N) % N;
  if constexpr (k == 0) {
    return a;
  } else if constexpr (4 * k == N) {
    return rot90<INV>(a);
  } else if constexpr (2 * k == N) {
    return make_float2(-a.x, -a.y);
  } else if constexpr (4 * k == 3 * N) {
    return rot90<!INV>(a);
  } else if constexpr ((8 * k) % N == 0) {
    // odd multiples of 1/8 turn: (+-1 +- i)/sqrt(2)
    constexpr float h = (float)0.70710678118654752440;
    constexpr float sr = Tw<k, N, INV>::re > 0 ? h : -h;
    constexpr float si = Tw<k, N, INV>::im > 0 ? h : -h;
    return make_float2(sr * a.x - si * a.y, si * a.x + sr * a.y);
  } else {
    constexpr float wr = Tw<k, N, INV>::re, wi = Tw<k, N, INV>::im;
    return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
  }
}

// ---- codelets ---------------------------------------------------------------------------
constexpr int smallest_factor(int r) {
  for (int f = 2; f * f <= r; ++f)
    if (r % f == 0) return f;
  return r;
}
// how a composite radix is split into R1 x R2 (R1 = first sub-DFT)
constexpr int ct_split(int r) {
  if (r % 4 == 0 && r > 4) return 4;
  return smallest_factor(r);
}

template <int R, bool INV, class Enable = void>
struct Dft;

template <bool INV>
struct Dft<1, INV> {
  static B200_HD void run(float2 (&)[1]) {}
};

template <bool INV>
struct Dft<2, INV> {
  static B200_HD void run(float2 (&x)[2]) {
    const float2 a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
  }
};

template <bool INV>
struct Dft<4, INV> {
  static B200_HD void run(float2 (&x)[4]) {
    const float2 t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
    const float2 t2 = cadd(x[1], x[3]), t3 = rot90<INV>(csub(x[1], x[3]));
    x[0] = cadd(t0, t2);
    x[1] = cadd(t1, t3);
    x[2] = csub(t0, t2);
    x[3] = csub(t1, t3);
  }
};

// Primes above this run the symmetric-pair form as LOOPS over a constant table instead of fully unrolled straight-line
// code: a radix-97 butterfly unrolls to ~9000 FMAs with compile-time coefficients, which NVRTC needs 20 s to schedule
// (127: 60 s); the looped form compiles in about a second and is only ever built by the plan-time specialisation tier
// (csrc/jit.cu), where every namespace-scope entity is a device entity (-default-device).
constexpr int kLoopedPrimeMin = 64;
constexpr bool looped_prime(int r) {
#ifdef __CUDACC_RTC__
  return r > kLoopedPrimeMin && r % 2 == 1 && smallest_factor(r) == r;
#else
  return (void)r, false;
#endif
}

#ifdef __CUDACC_RTC__
template <int R>
struct PrimeTable {
  float c[R], s[R];  // cos / sin of 2*pi*k/R
};
template <int R>
constexpr PrimeTable<R> make_prime_table() {
  PrimeTable<R> t{};
  for (int k = 0; k < R; ++k) {
    const CxD v = cx_unit(k, R);
    t.c[k] = float(v.re);
    t.s[k] = float(v.im);
  }
  return t;
}
template <int R>
struct PrimeTw {
  static constexpr PrimeTable<R> tab = make_prime_table<R>();
};

template <int R, bool INV>
struct Dft<R, INV, std::enable_if_t<looped_prime(R)>> {
  static B200_HD void run(float2 (&x)[R]) {
    constexpr int H = (R - 1) / 2;
    float2 a[H], b[H];
#pragma unroll 4
    for (int j = 1; j <= H; ++j) {
      a[j - 1] = cadd(x[j], x[R - j]);
      b[j - 1] = csub(x[j], x[R - j]);
    }
    const float2 x0 = x[0];
    float2 s0 = x0;
#pragma unroll 4
    for (int j = 0; j < H; ++j) s0 = cadd(s0, a[j]);
    x[0] = s0;
#pragma unroll 1
    for (int k = 1; k <= H; ++k) {
      float2 A = x0, B = make_float2(0.f, 0.f);
      int idx = 0;  // (j * k) mod R
#pragma unroll 4
      for (int j = 1; j <= H; ++j) {
        idx += k;
        if (idx >= R) idx -= R;
        const float c = PrimeTw<R>::tab.c[idx], s = PrimeTw<R>::tab.s[idx];
        A.x = fmaf(c, a[j - 1].x, A.x);
        A.y = fmaf(c, a[j - 1].y, A.y);
        B.x = fmaf(s, b[j - 1].x, B.x);
        B.y = fmaf(s, b[j - 1].y, B.y);
      }
      // forward: X_k = A - iB, X_{R-k} = A + iB ; inverse: swapped
      const float2 lo = make_float2(A.x + B.y, A.y - B.x);
      const float2 hi = make_float2(A.x - B.y, A.y + B.x);
      x[k] = INV ? hi : lo;
      x[R - k] = INV ? lo : hi;
    }
  }
};
#endif

// any odd radix that is prime (or that we choose not to split): symmetric-pair form
template <int R, bool INV>
struct Dft<R, INV, std::enable_if_t<(R > 2) && (R % 2 == 1) && smallest_factor(R) == R && !looped_prime(R)>> {
  static constexpr bool real_form = true;
  static B200_HD void run(float2 (&x)[R]) {
    constexpr int H = (R - 1) / 2;
    float2 a[H], b[H];
    static_for<H>([&](auto jc) {
      constexpr int j = decltype(jc)::value + 1;
      a[j - 1] = cadd(x[j], x[R - j]);
      b[j - 1] = csub(x[j], x[R - j]);
    });
    const float2 x0 = x[0];
    float2 s0 = x0;
    static_for<H>([&](auto jc) { s0 = cadd(s0, a[decltype(jc)::value]); });
    x[0] = s0;
    static_for<H>([&](auto kc) {
      constexpr int k = decltype(kc)::value + 1;
      float2 A = x0, B = make_float2(0.f, 0.f);
      static_for<H>([&](auto jc) {
        constexpr int j = decltype(jc)::value + 1;
        // cos / sin of 2*pi*j*k/R (direction-independent; the sign is applied below)
        constexpr float c = Tw<(j * k) % R, R, true>::re;
        constexpr float s = Tw<(j * k) % R, R, true>::im;
        A.x = fmaf(c, a[j - 1].x, A.x);
        A.y = fmaf(c, a[j - 1].y, A.y);
        B.x = fmaf(s, b[j - 1].x, B.x);
        B.y = fmaf(s, b[j - 1].y, B.y);
      });
      // forward: X_k = A - iB, X_{R-k} = A + iB ; inverse: swapped
      const float2 lo = make_float2(A.x + B.y, A.y - B.x);
      const float2 hi = make_float2(A.x - B.y, A.y + B.x);
      x[k] = INV ? hi : lo;
      x[R - k] = INV ? lo : hi;
    });
  }
  // The same butterfly on REAL inputs (every x[j].y == 0: stage 0 of a real-input transform): the pair sums a_j and
  // differences b_j are real, so each output pair costs 2 FMAs per j instead of 4 and X_{R-k} = conj X_k.
  static B200_HD void run_real(float2 (&x)[R]) {
    constexpr int H = (R - 1) / 2;
    float a[H], b[H];
    static_for<H>([&](auto jc) {
      constexpr int j = decltype(jc)::value + 1;
      a[j - 1] = x[j].x + x[R - j].x;
      b[j - 1] = x[j].x - x[R - j].x;
    });
    const float x0 = x[0].x;
    float s0 = x0;
    static_for<H>([&](auto jc) { s0 += a[decltype(jc)::value]; });
    x[0] = make_float2(s0, 0.f);
    static_for<H>([&](auto kc) {
      constexpr int k = decltype(kc)::value + 1;
      float A = x0, B = 0.f;
      static_for<H>([&](auto jc) {
        constexpr int j = decltype(jc)::value + 1;
        constexpr float c = Tw<(j * k) % R, R, true>::re;
        constexpr float s = Tw<(j * k) % R, R, true>::im;
        A = fmaf(c, a[j - 1], A);
        B = fmaf(s, b[j - 1], B);
      });
      // forward: X_k = A - iB, X_{R-k} = A + iB ; inverse: swapped
      x[k] = make_float2(A, INV ? B : -B);
      x[R - k] = make_float2(A, INV ? -B : B);
    });
  }
};

// run_real where a codelet has one (the unrolled odd primes), the general butterfly otherwise
template <class D, class = void>
struct has_run_real : std::false_type {};
template <class D>
struct has_run_real<D, std::enable_if_t<D::real_form>> : std::true_type {};

// composite radix: Cooley-Tukey on registers, R = R1 * R2, n = R2*n1 + n2, k = k1 + R1*k2
template <int R, bool INV>
struct Dft<R, INV, std::enable_if_t<(R > 4) && smallest_factor(R) != R>> {
  static constexpr int R1 = ct_split(R), R2 = R / R1;
  static B200_HD void run(float2 (&x)[R]) {
    float2 y[R];  // y[n2*R1 + k1]
    static_for<R2>([&](auto n2c) {
      constexpr int n2 = decltype(n2c)::value;
      float2 t[R1];
      static_for<R1>([&](auto n1c) { t[decltype(n1c)::value] = x[R2 * decltype(n1c)::value + n2]; });
      Dft<R1, INV>::run(t);
      static_for<R1>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        y[n2 * R1 + k1] = twiddle_mul<n2 * k1, R, INV>(t[k1]);
      });
    });
    static_for<R1>([&](auto k1c) {
      constexpr int k1 = decltype(k1c)::value;
      float2 t[R2];
      static_for<R2>([&](auto n2c) { t[decltype(n2c)::value] = y[decltype(n2c)::value * R1 + k1]; });
      Dft<R2, INV>::run(t);
      static_for<R2>([&](auto k2c) { x[k1 + R1 * decltype(k2c)::value] = t[decltype(k2c)::value]; });
    });
  }
};

}  // namespace b200fft
