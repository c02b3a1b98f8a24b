// Device helpers shared by all kernels: complex types, FMA, typed input loads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200fft.h"

namespace b200fft {

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

__device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ double2 mk2(double a, double b) { return make_double2(a, b); }

// acc + w * x (complex), four FMAs
template <typename T2>
__device__ __forceinline__ T2 cfma(T2 w, T2 x, T2 acc) {
  T2 r;
  r.x = fma(w.x, x.x, fma(-w.y, x.y, acc.x));
  r.y = fma(w.x, x.y, fma(w.y, x.x, acc.y));
  return r;
}

template <typename T2>
__device__ __forceinline__ T2 cmul(T2 a, T2 b) {
  T2 r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}

// Element `g` of the user's input as a working-dtype complex: the reference's
// cast-on-load (_fft.mojo:250-257); real input becomes (x, 0).
template <typename T>
__device__ __forceinline__ typename Vec2<T>::type load_any(const void* __restrict__ src, int dtype, int comps,
                                                           long long g) {
  if (comps == 2) {
    switch (dtype) {
      case B200FFT_F32: {
        const float2 v = __ldg(reinterpret_cast<const float2*>(src) + g);
        return mk2((T)v.x, (T)v.y);
      }
      case B200FFT_F64: {
        const double2 v = __ldg(reinterpret_cast<const double2*>(src) + g);
        return mk2((T)v.x, (T)v.y);
      }
      default: {
        const uchar2 v = __ldg(reinterpret_cast<const uchar2*>(src) + g);
        return mk2((T)v.x, (T)v.y);
      }
    }
  }
  switch (dtype) {
    case B200FFT_F32: return mk2((T)__ldg(reinterpret_cast<const float*>(src) + g), (T)0);
    case B200FFT_F64: return mk2((T)__ldg(reinterpret_cast<const double*>(src) + g), (T)0);
    default: return mk2((T)__ldg(reinterpret_cast<const unsigned char*>(src) + g), (T)0);
  }
}

}  // namespace b200fft
