// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of martinvuyk/hackathon-fft's radix-n Stockham FFT (the Mojo
// CPU path). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library, and only as the checker or the
// timed CPU baseline. The CUDA product (hackathon-fft_b200/csrc) never links,
// loads or calls anything in this directory.
//
// Parity status: PINNED. tests/test_oracle.py checks this file against every
// golden vector the reference's own tests hold for the path
// (fft/_test_values.mojo, fft/tests.mojo:422-458,613-905; transcribed by
// tests/golden/make_golden.py) at the reference's atol=1e-2 / rtol=1e-5, and
// against numpy float64. The reference itself is Mojo (pin: mojo
// 0.26.3.0.dev2026032521, fft/pixi.toml:49-50) and cannot be compiled or run in
// this image, so oracle/_ref does not exist; rounding inside the Mojo stdlib
// (cos/sin, ComplexSIMD.fma) is restated, not reproduced bit-for-bit.
//
// What follows what (reference file:line):
//   ref_times_divisible / ref_ordered_bases  fft/fft/_utils.mojo:125-183
//   ref_default_bases                        fft/fft/fft.mojo:49-104
//   twiddle()                                fft/fft/_utils.mojo:63-104
//   unit_phasor_fma*()                       fft/fft/_utils.mojo:320-372
//   stage_point()                            fft/fft/_fft.mojo:228-296 (runtime form)
//                                            fft/fft/_fft.mojo:331-391 (unrolled form, N<=128)
//   run_1d()                                 fft/fft/_ndim_fft_cpu.mojo:147-241
//   transpose()                              fft/fft/_ndim_fft_cpu.mojo:63-93, _utils.mojo:400-419
//   Plan::exec() (run_batch)                 fft/fft/_ndim_fft_cpu.mojo:96-323
//   Plan::init()                             fft/fft/_ndim_fft_cpu.mojo:28-60 (_CPUPlan), fft/fft/fft.mojo:122-157
//   Pool / parallel_for                      std.algorithm.parallelize as used at _ndim_fft_cpu.mojo:306-308,323
//   ref_plan_create / _exec / _destroy       the plan_fft(...) / fft(out, x, plan=plan) split timed by fft/bench.mojo:83-90
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace {

typedef uint64_t u64;

// ---- base ordering ---------------------------------------------------------

// _div_by (_utils.mojo:125-129)
u64 div_by(u64 x, u64 base) {
  if (base == x) return 1;
  if (base > x || x % base != 0) return 0;
  return div_by(x / base, base) + 1;
}

// _times_divisible_by (_utils.mojo:132-152): power-of-two bases use
// ctz(length) / log2(base); everything else the recursive _div_by.
u64 times_divisible(u64 length, u64 base) {
  if ((base & (base - 1)) == 0) {
    u64 ctz = length ? (u64)__builtin_ctzll(length) : 64;
    u64 lg = (u64)__builtin_ctzll(base);
    return ctz / lg;
  }
  return div_by(length, base);
}

// _build_ordered_bases (_utils.mojo:163-183)
std::vector<u64> ordered_bases(u64 length, std::vector<u64> bases) {
  std::sort(bases.begin(), bases.end());
  u64 prod = 1;
  for (u64 b : bases) prod *= b;
  if (prod == length) {
    std::reverse(bases.begin(), bases.end());
    return bases;
  }
  std::vector<u64> out;
  u64 processed = 1;
  for (size_t k = bases.size(); k-- > 0;) {
    u64 base = bases[k];
    u64 times = times_divisible(length, base);
    for (u64 t = 0; t < times; ++t) {
      out.push_back(base);
      processed *= base;
    }
    if (processed == length) break;
  }
  return out;
}

// validity asserts of _get_ordered_bases_processed_list (_utils.mojo:186-221)
bool bases_valid(u64 length, const std::vector<u64>& ordered) {
  if (ordered.empty()) return false;
  u64 prod = 1;
  for (u64 b : ordered) {
    if (b == 1 || b == 0) return false;
    prod *= b;
  }
  return prod == length;
}

// _estimate_best_bases (fft.mojo:49-104)
std::vector<u64> default_bases(u64 length, bool gpu_target) {
  const u64 max_radix = 32, block = 1024;
  if (gpu_target && length / max_radix <= block) {
    u64 lo = std::max<u64>((length + block - 1) / block, 2);
    std::vector<u64> pot;
    u64 processed = 1;
    for (u64 r = lo; r <= max_radix; ++r) {
      u64 times = times_divisible(length / processed, r);
      for (u64 t = 0; t < times; ++t) {
        pot.push_back(r);
        processed *= r;
      }
      if (processed == length) {
        std::reverse(pot.begin(), pot.end());
        return pot;
      }
    }
  }
  static const u64 primes[25] = {97, 89, 83, 79, 73, 71, 67, 61, 59, 53, 47, 43, 41,
                                 37, 31, 29, 23, 19, 17, 13, 11, 7,  5,  3,  2};
  std::vector<u64> out;
  u64 processed = 1;
  for (u64 p : primes) {
    u64 times = times_divisible(length / processed, p);
    for (u64 t = 0; t < times; ++t) {
      out.push_back(p);
      processed *= p;
    }
    if (processed == length) {
      std::reverse(out.begin(), out.end());
      return out;
    }
  }
  return out;  // product != length: caller reports the error
}

// ---- arithmetic ------------------------------------------------------------

template <class T>
struct Cx {
  T re, im;
};

// _get_twiddle_factor (_utils.mojo:63-104). theta is formed in the working
// dtype; `exact_quarters` is the compile-time-interpreter branch (:73-82) the
// reference takes for its inline tables (CPU lengths <= 128).
template <class T>
Cx<T> twiddle(u64 n, u64 N, bool inverse, bool exact_quarters) {
  const T c = (T)(-2.0 * M_PI) / (T)N;
  const T theta = c * (T)n;
  Cx<T> w;
  bool done = false;
  if (exact_quarters) {
    T factor = (T)2 * (T)n / (T)N;
    if (factor < (T)1e-9) { w = {1, 0}; done = true; }
    else if (factor == (T)0.5) { w = {0, -1}; done = true; }
    else if (factor == (T)1) { w = {-1, 0}; done = true; }
    else if (factor == (T)1.5) { w = {0, 1}; done = true; }
  }
  if (!done) w = {std::cos(theta), std::sin(theta)};
  if (inverse) w.im = -w.im;
  return w;
}

// ComplexSIMD.fma: w * x + acc with fused multiply-adds
template <class T>
inline Cx<T> cfma(Cx<T> w, Cx<T> x, Cx<T> acc) {
  Cx<T> r;
  r.re = std::fma(w.re, x.re, std::fma(-w.im, x.im, acc.re));
  r.im = std::fma(w.re, x.im, std::fma(w.im, x.re, acc.im));
  return r;
}

// _unit_phasor_fma, complex input (_utils.mojo:320-346)
template <class T>
inline Cx<T> unit_phasor_fma(Cx<T> w, Cx<T> x, Cx<T> acc) {
  if (w.re == 1) return {acc.re + x.re, acc.im + x.im};
  if (w.im == -1) return {acc.re + x.im, acc.im - x.re};
  if (w.re == -1) return {acc.re - x.re, acc.im - x.im};
  if (w.im == 1) return {acc.re - x.im, acc.im + x.re};
  if (std::fabs(w.re) == std::fabs(w.im)) {
    const T f = std::fabs(w.re);
    const T re = x.re, im = x.im;
    if (w.re > 0 && w.im > 0) return {acc.re + f * (re - im), acc.im + f * (re + im)};
    if (w.re < 0 && w.im > 0) return {acc.re + f * (-re - im), acc.im + f * (re - im)};
    if (w.re < 0 && w.im < 0) return {acc.re + f * (-re + im), acc.im + f * (-re - im)};
    return {acc.re + f * (re + im), acc.im + f * (-re + im)};
  }
  return cfma(w, x, acc);
}

// _unit_phasor_fma, real input (_utils.mojo:349-372)
template <class T>
inline Cx<T> unit_phasor_fma_real(Cx<T> w, T x, Cx<T> acc, bool accum_is_real) {
  if (w.re == 1) return {acc.re + x, acc.im};
  if (w.im == -1 && accum_is_real) return {acc.re, -x};
  if (w.im == -1) return {acc.re, acc.im - x};
  if (w.re == -1) return {acc.re - x, acc.im};
  if (w.im == 1 && accum_is_real) return {acc.re, x};
  if (w.im == 1) return {acc.re, acc.im + x};
  if (accum_is_real) return {std::fma(w.re, x, acc.re), w.im * x};
  return {std::fma(w.re, x, acc.re), std::fma(w.im, x, acc.im)};
}

enum InDType { IN_U8 = 0, IN_F32 = 1, IN_F64 = 2 };

template <class T>
inline T load_scalar(const void* p, int dt, int64_t idx) {
  switch (dt) {
    case IN_U8: return (T)((const uint8_t*)p)[idx];
    case IN_F32: return (T)((const float*)p)[idx];
    default: return (T)((const double*)p)[idx];
  }
}

// One axis worth of plan data
template <class T>
struct AxisPlan {
  u64 N;
  std::vector<u64> radix, processed;
  std::vector<Cx<T>> tw;  // W_N^n, n in [0, N)
  bool unrolled;          // N <= MAX_STACK_SEQ_LEN (128): comptime path
  // per-stage copy of the twiddles in the order the stage walks them, stw[s][(j-1)*Q + u] = W[((j*u) mod Q)*rho]:
  // the same values as the gather from `tw`, laid out contiguously so the vector path below can load them
  std::vector<std::vector<Cx<T>>> stw;
};

// Source of a stage: either the typed user input (stage 0 of the last axis) or
// a working-dtype complex buffer.
template <class T>
struct Src {
  const Cx<T>* c;    // complex working buffer, or null
  const void* raw;   // user input row
  int dt, comps;     // dtype code, 1 (real) or 2 (complex)
  inline Cx<T> get(u64 i) const {
    if (c) return c[i];
    if (comps == 1) return {load_scalar<T>(raw, dt, (int64_t)i), 0};
    return {load_scalar<T>(raw, dt, 2 * (int64_t)i), load_scalar<T>(raw, dt, 2 * (int64_t)i + 1)};
  }
};

// One Stockham stage over a whole row (_fft.mojo:228-296 / :331-391):
//   u = i mod Q, n = (i div Q)*P + (u mod P)
//   acc = X[n]; for j in 1..r-1: acc = fma(W[((j*u) mod Q)*rho], X[n + j*N/r], acc)
// The output index is walked as i = q*Q + k*P + p (q < N/Q, k < r, p < P), which gives
// u = k*P + p and n = q*P + p without a division per point; the arithmetic per output
// point (operands, order of the FMAs) is exactly the reference's. R > 0 fixes the radix
// at compile time so the j loop unrolls like the reference's `comptime for`.
template <class T, int R, bool UNROLLED, bool REAL_IN, class Getter>
void run_stage_impl(const AxisPlan<T>& ax, size_t s, bool inverse, Getter get, Cx<T>* out) {
  const u64 N = ax.N, r = R > 0 ? (u64)R : ax.radix[s], P = ax.processed[s], Q = P * r, rho = N / Q,
            step = N / r;
  const bool scale = inverse && (Q == N);
  const T inv_n = (T)(1.0 / (double)N);
  const Cx<T>* tw = ax.tw.data();
  for (u64 q = 0; q < N / Q; ++q) {
    for (u64 k = 0; k < r; ++k) {
      const u64 i0 = q * Q + k * P, n0 = q * P;
      for (u64 p = 0; p < P; ++p) {
        const u64 u = k * P + p, n = n0 + p;
        Cx<T> acc = get(n);
        u64 ju = 0;
        for (u64 j = 1; j < r; ++j) {
          ju += u;
          if (ju >= Q) ju -= Q;  // (j*u) mod Q, since u < Q
          const Cx<T> w = tw[ju * rho];
          if (UNROLLED) {
            if (REAL_IN) acc = unit_phasor_fma_real(w, get(n + j * step).re, acc, j == 1);
            else acc = unit_phasor_fma(w, get(n + j * step), acc);
          } else {
            acc = cfma(w, get(n + j * step), acc);
          }
        }
        if (scale) { acc.re *= inv_n; acc.im *= inv_n; }
        out[i0 + p] = acc;
      }
    }
  }
}

#if defined(__AVX2__) && defined(__FMA__)
}  // namespace
#include <immintrin.h>
namespace {
// Vector form of one runtime-path stage (float, complex source, P >= 4): four output points per iteration.
// Every lane performs exactly the scalar sequence of cfma() — fma(w.re, x.re, fma(-w.im, x.im, acc.re)) and
// fma(w.re, x.im, fma(w.im, x.re, acc.im)) — so the results are bit-identical to run_stage_impl; it only makes the
// CPU baseline run at the speed a SIMD implementation like the reference's reaches.
inline bool run_stage_f32_avx2(const AxisPlan<float>& ax, size_t s, bool inverse, const Cx<float>* x, Cx<float>* out) {
  const u64 N = ax.N, r = ax.radix[s], P = ax.processed[s], Q = P * r, step = N / r;
  if (ax.unrolled || s >= ax.stw.size() || ax.stw[s].empty()) return false;
  const bool scale = inverse && (Q == N);
  if (P < 4) {
    // P = 1 or 2 (the first stages): the inputs x[m + j*step], m = q*P + p, are contiguous in m, the twiddle
    // pattern has period P, and the four results go to out[(m / P)*Q + k*P + m % P]: P-element pieces, Q apart.
    if ((P != 1 && P != 2) || (N / r) % 4 != 0) return false;
    const __m256 vinv1 = _mm256_set1_ps((float)(1.0 / (double)N));
    const __m256 sgn = _mm256_castsi256_ps(_mm256_set_epi32(0, (int)0x80000000, 0, (int)0x80000000, 0, (int)0x80000000, 0,
                                                            (int)0x80000000));
    const Cx<float>* stw1 = ax.stw[s].data();
    for (u64 k = 0; k < r; ++k) {
      // 4-wide twiddle vectors for this k: lanes m % P select u = k*P + p
      alignas(32) Cx<float> wl[32][4];
      for (u64 j = 1; j < r && j < 32; ++j)
        for (int l = 0; l < 4; ++l) wl[j][l] = stw1[(j - 1) * Q + k * P + (u64)l % P];
      if (r > 32) return false;
      for (u64 m = 0; m < N / r; m += 4) {
        __m256 acc = _mm256_loadu_ps(reinterpret_cast<const float*>(x + m));
        for (u64 j = 1; j < r; ++j) {
          const __m256 w = _mm256_load_ps(reinterpret_cast<const float*>(wl[j]));
          const __m256 xv = _mm256_loadu_ps(reinterpret_cast<const float*>(x + m + j * step));
          const __m256 wre = _mm256_moveldup_ps(w), wim = _mm256_movehdup_ps(w);
          const __m256 xs = _mm256_permute_ps(xv, 0xB1);
          const __m256 t = _mm256_fmadd_ps(_mm256_xor_ps(wim, sgn), xs, acc);
          acc = _mm256_fmadd_ps(wre, xv, t);
        }
        if (scale) acc = _mm256_mul_ps(acc, vinv1);
        // out[(mm / P)*Q + k*P + mm % P] for mm = m..m+3, without a division per point
        if (P == 1) {
          alignas(32) Cx<float> res[4];
          _mm256_store_ps(reinterpret_cast<float*>(res), acc);
          Cx<float>* o = out + m * Q + k;
          o[0] = res[0]; o[Q] = res[1]; o[2 * Q] = res[2]; o[3 * Q] = res[3];
        } else {  // P == 2: two adjacent points per Q-block
          float* o = reinterpret_cast<float*>(out + (m >> 1) * Q + k * 2);
          _mm_storeu_ps(o, _mm256_castps256_ps128(acc));
          _mm_storeu_ps(o + 2 * Q, _mm256_extractf128_ps(acc, 1));
        }
      }
    }
    return true;
  }
  const float inv_n = (float)(1.0 / (double)N);
  const __m256 vinv = _mm256_set1_ps(inv_n);
  const __m256 sign_even = _mm256_castsi256_ps(_mm256_set_epi32(0, (int)0x80000000, 0, (int)0x80000000, 0, (int)0x80000000, 0,
                                                                (int)0x80000000));
  const Cx<float>* stw = ax.stw[s].data();
  const u64 P4 = P & ~(u64)3;
  for (u64 q = 0; q < N / Q; ++q) {
    for (u64 k = 0; k < r; ++k) {
      const u64 i0 = q * Q + k * P, n0 = q * P, u0 = k * P;
      for (u64 p = 0; p < P4; p += 4) {
        __m256 acc = _mm256_loadu_ps(reinterpret_cast<const float*>(x + n0 + p));
        for (u64 j = 1; j < r; ++j) {
          const __m256 w = _mm256_loadu_ps(reinterpret_cast<const float*>(stw + (j - 1) * Q + u0 + p));
          const __m256 xv = _mm256_loadu_ps(reinterpret_cast<const float*>(x + n0 + p + j * step));
          const __m256 wre = _mm256_moveldup_ps(w), wim = _mm256_movehdup_ps(w);
          const __m256 xs = _mm256_permute_ps(xv, 0xB1);  // (im, re) pairs
          const __m256 t = _mm256_fmadd_ps(_mm256_xor_ps(wim, sign_even), xs, acc);
          acc = _mm256_fmadd_ps(wre, xv, t);
        }
        if (scale) acc = _mm256_mul_ps(acc, vinv);
        _mm256_storeu_ps(reinterpret_cast<float*>(out + i0 + p), acc);
      }
      for (u64 p = P4; p < P; ++p) {
        Cx<float> acc = x[n0 + p];
        for (u64 j = 1; j < r; ++j) acc = cfma(stw[(j - 1) * Q + u0 + p], x[n0 + p + j * step], acc);
        if (scale) { acc.re *= inv_n; acc.im *= inv_n; }
        out[i0 + p] = acc;
      }
    }
  }
  return true;
}
template <class T>
inline bool run_stage_simd(const AxisPlan<T>&, size_t, bool, const Cx<T>*, Cx<T>*) { return false; }
template <>
inline bool run_stage_simd<float>(const AxisPlan<float>& ax, size_t s, bool inverse, const Cx<float>* x, Cx<float>* out) {
  return run_stage_f32_avx2(ax, s, inverse, x, out);
}
#else
template <class T>
inline bool run_stage_simd(const AxisPlan<T>&, size_t, bool, const Cx<T>*, Cx<T>*) { return false; }
#endif

template <class T, bool UNROLLED, bool REAL_IN, class Getter>
void run_stage_radix(const AxisPlan<T>& ax, size_t s, bool inverse, Getter get, Cx<T>* out) {
  switch (ax.radix[s]) {
    case 2: return run_stage_impl<T, 2, UNROLLED, REAL_IN>(ax, s, inverse, get, out);
    case 3: return run_stage_impl<T, 3, UNROLLED, REAL_IN>(ax, s, inverse, get, out);
    case 4: return run_stage_impl<T, 4, UNROLLED, REAL_IN>(ax, s, inverse, get, out);
    case 5: return run_stage_impl<T, 5, UNROLLED, REAL_IN>(ax, s, inverse, get, out);
    case 7: return run_stage_impl<T, 7, UNROLLED, REAL_IN>(ax, s, inverse, get, out);
    case 8: return run_stage_impl<T, 8, UNROLLED, REAL_IN>(ax, s, inverse, get, out);
    default: return run_stage_impl<T, 0, UNROLLED, REAL_IN>(ax, s, inverse, get, out);
  }
}

template <class T>
void run_stage(const AxisPlan<T>& ax, size_t s, bool inverse, const Src<T>& x, Cx<T>* out) {
  if (x.c) {  // working-dtype complex buffer: the common case
    const Cx<T>* c = x.c;
    if (run_stage_simd<T>(ax, s, inverse, c, out)) return;
    auto get = [c](u64 i) { return c[i]; };
    if (ax.unrolled) run_stage_radix<T, true, false>(ax, s, inverse, get, out);
    else run_stage_radix<T, false, false>(ax, s, inverse, get, out);
    return;
  }
  if (x.comps == 2 && x.dt == (sizeof(T) == 4 ? IN_F32 : IN_F64)) {  // complex input already in T
    const Cx<T>* c = reinterpret_cast<const Cx<T>*>(x.raw);
    if (run_stage_simd<T>(ax, s, inverse, c, out)) return;
    auto get = [c](u64 i) { return c[i]; };
    if (ax.unrolled) run_stage_radix<T, true, false>(ax, s, inverse, get, out);
    else run_stage_radix<T, false, false>(ax, s, inverse, get, out);
    return;
  }
  auto get = [x](u64 i) { return x.get(i); };
  if (x.comps == 1) {
    // real input: the unrolled path uses the real-operand FMA forms (_utils.mojo:349-372);
    // the runtime path widens to (x, 0) and uses the complex FMA (_fft.mojo:254-255)
    if (ax.unrolled) run_stage_radix<T, true, true>(ax, s, inverse, get, out);
    else run_stage_radix<T, false, false>(ax, s, inverse, get, out);
  } else {
    if (ax.unrolled) run_stage_radix<T, true, false>(ax, s, inverse, get, out);
    else run_stage_radix<T, false, false>(ax, s, inverse, get, out);
  }
}

// _transpose (_ndim_fft_cpu.mojo:63-93): [b][M][N] -> [b][N][M], tiled
template <class T>
void transpose(Cx<T>* dst, const Cx<T>* src, u64 b, u64 M, u64 N) {
  const u64 TILE = 64 / sizeof(T);
  for (u64 k = 0; k < b; ++k) {
    const Cx<T>* s = src + k * M * N;
    Cx<T>* d = dst + k * M * N;
    for (u64 i = 0; i < M; i += TILE)
      for (u64 j = 0; j < N; j += TILE)
        for (u64 ii = i; ii < std::min(i + TILE, M); ++ii)
          for (u64 jj = j; jj < std::min(j + TILE, N); ++jj) d[jj * M + ii] = s[ii * N + jj];
  }
}

// parallelize[func](n, workers): the reference hands its closures to the Mojo runtime's persistent worker pool
// (std.algorithm.parallelize, used at _ndim_fft_cpu.mojo:306-308,323), so no thread is created per call. The same
// here: one process-wide pool of sleeping workers; a job is a shared atomic index counter over [0, n).
class Pool {
 public:
  static Pool& get() {
    static Pool p;
    return p;
  }
  template <class F>
  void run(u64 n, u64 workers, F&& f) {
    workers = std::max<u64>(1, std::min(workers, n));
    if (workers == 1 || in_worker_) {  // nested parallelize from inside a job: run inline (see exec: the split makes
      for (u64 i = 0; i < n; ++i) f(i);  // one of the two levels serial whenever the other has more than one worker)
      return;
    }
    std::lock_guard<std::mutex> job_lock(job_mu_);  // one job at a time
    grow(workers - 1);
    std::atomic<u64> next(0);
    const u64 grain = std::max<u64>(1, n / (workers * 8));
    std::function<void()> body = [&]() {
      for (;;) {
        const u64 lo = next.fetch_add(grain);
        if (lo >= n) break;
        const u64 hi = std::min(n, lo + grain);
        for (u64 i = lo; i < hi; ++i) f(i);
      }
    };
    {
      std::lock_guard<std::mutex> lk(mu_);
      body_ = &body;
      want_ = workers - 1;
      started_ = 0;
      done_ = 0;
      ++gen_;
    }
    cv_.notify_all();
    in_worker_ = true;  // the caller takes part in the job: a nested parallelize inside f runs inline
    body();
    in_worker_ = false;
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return done_ == want_; });
    body_ = nullptr;
  }

 private:
  Pool() {}
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void grow(u64 n) {
    while (th_.size() < n) {
      const u64 seen = gen_;
      th_.emplace_back([this, seen] { loop(seen); });
    }
  }
  void loop(u64 seen) {
    in_worker_ = true;
    for (;;) {
      std::function<void()>* body = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        if (started_ >= want_) continue;  // this job needs fewer workers than the pool holds
        ++started_;
        body = body_;
      }
      (*body)();
      {
        std::lock_guard<std::mutex> lk(mu_);
        ++done_;
      }
      done_cv_.notify_one();
    }
  }
  std::mutex job_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> th_;
  std::function<void()>* body_ = nullptr;
  u64 gen_ = 0, want_ = 0, started_ = 0, done_ = 0;
  bool stop_ = false;
  static thread_local bool in_worker_;
};
thread_local bool Pool::in_worker_ = false;

template <class F>
void parallel_for(u64 n, u64 workers, F&& f) {
  Pool::get().run(n, workers, static_cast<F&&>(f));
}

// _CPUPlan (_ndim_fft_cpu.mojo:28-60) + the compile-time facts plan_fft fixes (fft.mojo:122-157): stage lists,
// twiddle tables and the calc_buf scratch are built ONCE; exec() below is what the reference's bench times
// (fft/bench.mojo:83-90: plan_fft outside, fft(out, x, plan=plan) inside the timed loop).
struct PlanBase {
  virtual ~PlanBase() {}
  virtual int exec(const void* x, void* out, int workers) = 0;
  bool f64 = false;
};

template <class T>
struct Plan : PlanBase {
  std::vector<AxisPlan<T>> axes;
  int in_dtype = IN_F32, in_comps = 2, ndim = 0, inverse = 0;
  int64_t batches = 0;
  u64 prod = 1, total_stages = 0, max_batch_prod = 0;
  std::vector<Cx<T>> calc;  // plan.calc_buf: one full copy of the output (_ndim_fft_cpu.mojo:44-45)

  int init(int in_dtype_, int in_comps_, int64_t batches_, int ndim_, const int64_t* dims, const uint32_t* bases_flat,
           const int32_t* bases_counts, int inverse_) {
    if (ndim_ < 1 || batches_ < 1 || (in_comps_ != 1 && in_comps_ != 2)) return 1;
    in_dtype = in_dtype_; in_comps = in_comps_; batches = batches_; ndim = ndim_; inverse = inverse_;
    f64 = sizeof(T) == 8;
    axes.resize(ndim);
    const uint32_t* bp = bases_flat;
    for (int a = 0; a < ndim; ++a) {
      AxisPlan<T>& ax = axes[a];
      ax.N = (u64)dims[a];
      if (ax.N < 2) return 2;  // "no inner dimension should be of size 1"
      std::vector<u64> user;
      if (bases_counts && bases_counts[a] > 0) {
        user.assign(bp, bp + bases_counts[a]);
        bp += bases_counts[a];
      } else {
        user = default_bases(ax.N, /*gpu_target=*/false);
      }
      for (u64 b : user)
        if (b < 2) return 3;
      ax.radix = ordered_bases(ax.N, user);
      if (!bases_valid(ax.N, ax.radix)) return 3;
      u64 p = 1;
      for (u64 r : ax.radix) { ax.processed.push_back(p); p *= r; }
      ax.unrolled = ax.N <= 128;  // _CPUPlan.MAX_STACK_SEQ_LEN
      ax.tw.resize(ax.N);
      for (u64 n = 0; n < ax.N; ++n) ax.tw[n] = twiddle<T>(n, ax.N, inverse != 0, ax.unrolled);
      if (!ax.unrolled) {
        ax.stw.resize(ax.radix.size());
        for (size_t st = 0; st < ax.radix.size(); ++st) {
          const u64 rr = ax.radix[st], PP = ax.processed[st], QQ = PP * rr, rho = ax.N / QQ;
          if (PP == 3) continue;  // no vector path for P = 3
          ax.stw[st].resize((size_t)((rr - 1) * QQ));
          for (u64 j = 1; j < rr; ++j)
            for (u64 u = 0; u < QQ; ++u) ax.stw[st][(size_t)((j - 1) * QQ + u)] = ax.tw[((j * u) % QQ) * rho];
        }
      }
      prod *= ax.N;
      total_stages += ax.radix.size();
      u64 mn = *std::min_element(user.begin(), user.end());
      max_batch_prod = std::max(max_batch_prod, ax.N / mn);
    }
    total_stages += 2 * (u64)(ndim - 1);
    calc.resize((size_t)batches * prod);
    return 0;
  }

  // stages of axes to the right of `a` (_num_stages_end_of)
  u64 stages_from(int a) const {
    u64 s = 0;
    for (int k = a; k < ndim; ++k) s += axes[k].radix.size();
    return s;
  }

  void run_1d(int a, Cx<T>* lhs, Cx<T>* rhs, const void* xin) const {
    const int last_axis = ndim - 1;
    const AxisPlan<T>& ax = axes[a];
    const u64 prev = stages_from(a + 1) + (u64)(last_axis - a);
    for (size_t b = 0; b < ax.radix.size(); ++b) {
      const u64 s = prev + b;
      const bool write_lhs = (total_stages - (s + 1)) % 2 == 0;
      Src<T> src;
      if (b == 0 && a == last_axis) src = {nullptr, xin, in_dtype, in_comps};
      else src = {write_lhs ? rhs : lhs, nullptr, 0, 2};
      run_stage(ax, b, inverse != 0, src, write_lhs ? lhs : rhs);
    }
  }

  int exec(const void* x, void* out_raw, int workers) override {
    const int last_axis = ndim - 1;
    // worker split (_ndim_fft_cpu.mojo:125-140)
    u64 threads = workers > 0 ? (u64)workers : std::max(1u, std::thread::hardware_concurrency());
    u64 per_batch_workers = ndim > 1 ? std::min(threads, max_batch_prod) : 1;
    u64 parallel_batches =
        std::min<u64>(std::max<int64_t>((int64_t)threads - ((int64_t)per_batch_workers - 1), 1), (u64)batches);
    Cx<T>* out_all = reinterpret_cast<Cx<T>*>(out_raw);
    const size_t in_elem = (in_dtype == IN_U8 ? 1 : in_dtype == IN_F32 ? 4 : 8);

    auto run_batch = [&](u64 bi) {
      Cx<T>* base_out = out_all + bi * prod;
      Cx<T>* base_calc = calc.data() + bi * prod;
      const char* base_x = (const char*)x + (size_t)bi * prod * in_comps * in_elem;
      if (ndim == 1) {
        run_1d(0, base_out, base_calc, base_x);
        return;
      }
      for (int a = last_axis; a >= 0; --a) {
        const u64 dim = axes[a].N;
        if (a != last_axis) {
          // "into" transpose: [prod d<a][d_a][prod d>a] -> [prod d<a][prod d>a][d_a]
          const u64 s = stages_from(a + 1) + (u64)(last_axis - (a + 1));
          const bool write_lhs = (total_stages - (s + 1)) % 2 == 0;
          u64 b = 1, n = 1;
          for (int k = 0; k < a; ++k) b *= axes[k].N;
          for (int k = a + 1; k < ndim; ++k) n *= axes[k].N;
          if (write_lhs) transpose(base_out, base_calc, b, dim, n);
          else transpose(base_calc, base_out, b, dim, n);
        }
        const u64 rows = prod / dim;
        parallel_for(rows, per_batch_workers, [&](u64 r) {
          run_1d(a, base_out + r * dim, base_calc + r * dim, base_x + (size_t)r * dim * in_comps * in_elem);
        });
      }
      const u64 fft_stages = stages_from(0);
      for (int a = 0; a < last_axis; ++a) {
        // "restore" transpose: [prod d<a][prod d>a][d_a] -> [prod d<a][d_a][prod d>a]
        const u64 s = fft_stages + (u64)last_axis + (u64)a;
        const bool write_lhs = (total_stages - (s + 1)) % 2 == 0;
        u64 b = 1, n = 1;
        for (int k = 0; k < a; ++k) b *= axes[k].N;
        for (int k = a + 1; k < ndim; ++k) n *= axes[k].N;
        if (write_lhs) transpose(base_out, base_calc, b, n, axes[a].N);
        else transpose(base_calc, base_out, b, n, axes[a].N);
      }
    };
    parallel_for((u64)batches, parallel_batches, run_batch);
    return 0;
  }
};

template <class T>
int exec(const void* x, int in_dtype, int in_comps, T* out_raw, int64_t batches, int ndim,
         const int64_t* dims, const uint32_t* bases_flat, const int32_t* bases_counts, int inverse,
         int workers) {
  Plan<T> plan;
  int rc = plan.init(in_dtype, in_comps, batches, ndim, dims, bases_flat, bases_counts, inverse);
  if (rc) return rc;
  return plan.exec(x, out_raw, workers);
}

int copy_bases(const std::vector<u64>& v, uint32_t* out, int cap) {
  if ((int)v.size() > cap) return -1;
  for (size_t i = 0; i < v.size(); ++i) out[i] = (uint32_t)v[i];
  return (int)v.size();
}

}  // namespace

extern "C" {

// Ordered stage list for `length` from user bases. Returns the count, or -1 when
// the reference would reject the bases (product != length, or a base of 1).
int ref_ordered_bases(uint64_t length, const uint32_t* bases, int nbases, uint32_t* out, int cap) {
  std::vector<u64> user(bases, bases + nbases);
  for (u64 b : user)
    if (b < 2) return -1;
  std::vector<u64> o = ordered_bases(length, user);
  if (!bases_valid(length, o)) return -1;
  return copy_bases(o, out, cap);
}

// Default user bases (before ordering) for a CPU (gpu_target=0) or GPU target.
int ref_default_bases(uint64_t length, int gpu_target, uint32_t* out, int cap) {
  return copy_bases(default_bases(length, gpu_target != 0), out, cap);
}

// Batched N-d transform over all non-batch axes, fp32 working dtype.
//   x: dense row-major (batches, dims..., in_comps) of in_dtype (0=u8,1=f32,2=f64)
//   out: dense row-major (batches, dims..., 2) float
//   bases_flat/bases_counts: per-axis user bases, or bases_counts==NULL / 0 for the CPU default
//   workers: 0 = hardware_concurrency
int ref_fft_exec_f32(const void* x, int in_dtype, int in_comps, float* out, int64_t batches,
                     int ndim, const int64_t* dims, const uint32_t* bases_flat,
                     const int32_t* bases_counts, int inverse, int workers) {
  return exec<float>(x, in_dtype, in_comps, out, batches, ndim, dims, bases_flat, bases_counts,
                     inverse, workers);
}

int ref_fft_exec_f64(const void* x, int in_dtype, int in_comps, double* out, int64_t batches,
                     int ndim, const int64_t* dims, const uint32_t* bases_flat,
                     const int32_t* bases_counts, int inverse, int workers) {
  return exec<double>(x, in_dtype, in_comps, out, batches, ndim, dims, bases_flat, bases_counts,
                      inverse, workers);
}

// Plan handle: everything plan_fft builds once (stage lists, twiddles, calc_buf). out_f64 selects the working dtype.
int ref_plan_create(void** plan, int out_f64, int in_dtype, int in_comps, int64_t batches, int ndim, const int64_t* dims,
                    const uint32_t* bases_flat, const int32_t* bases_counts, int inverse) {
  if (!plan) return 1;
  *plan = nullptr;
  int rc;
  if (out_f64) {
    auto* p = new Plan<double>();
    rc = p->init(in_dtype, in_comps, batches, ndim, dims, bases_flat, bases_counts, inverse);
    if (rc) delete p; else *plan = static_cast<PlanBase*>(p);
  } else {
    auto* p = new Plan<float>();
    rc = p->init(in_dtype, in_comps, batches, ndim, dims, bases_flat, bases_counts, inverse);
    if (rc) delete p; else *plan = static_cast<PlanBase*>(p);
  }
  return rc;
}
// fft(out, x, plan=plan, cpu_workers=workers): writes every element of the caller's `out`; nothing else is touched.
int ref_plan_exec(void* plan, const void* x, void* out, int workers) {
  if (!plan || !x || !out) return 1;
  return static_cast<PlanBase*>(plan)->exec(x, out, workers);
}
void ref_plan_destroy(void* plan) { delete static_cast<PlanBase*>(plan); }

int ref_hardware_threads(void) { return (int)std::max(1u, std::thread::hardware_concurrency()); }

}  // extern "C"
