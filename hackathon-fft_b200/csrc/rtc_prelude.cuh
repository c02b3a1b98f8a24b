// What the device headers (dft.cuh, tma.cuh, fast.cuh) need from the host toolchain's headers, restated for NVRTC:
// the plan-time specialisation tier (jit.cu) compiles those same headers at run time, where no host include path
// exists. Under nvcc this file only includes the real headers.
//
// Two more things are decided here per translation unit, both only ever set by jit.cu:
//   * the scalar type of the INPUT array the first pass reads (cast on load like the reference, _fft.mojo:257):
//     B200FFT_JIT_IN_U8 / B200FFT_JIT_IN_F64, default fp32 -> b200fft::in_scalar / in_vec2;
//   * B200FFT_JIT_F64: the working precision. The kernels are written for fp32 complex (float2); their fp64 build is
//     the SAME text with the scalar type swapped (float -> double, float2 -> double2, fmaf -> fma), which is what
//     the reference's `out_dtype` parameter does to its kernels (_fft.mojo:190-227). The swap is a set of macros that
//     exists only inside that run-time compiled unit, after the input typedefs (which therefore keep their own type).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include <utility>
#else
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef long long int64_t;
typedef unsigned long size_t;
struct alignas(64) CUtensorMap_st {
  unsigned long long opaque[16];
};
typedef CUtensorMap_st CUtensorMap;
namespace std {
template <class T, T v>
struct integral_constant {
  static constexpr T value = v;
  using value_type = T;
  using type = integral_constant;
  constexpr operator value_type() const noexcept { return value; }
  constexpr value_type operator()() const noexcept { return value; }
};
using true_type = integral_constant<bool, true>;
using false_type = integral_constant<bool, false>;
template <bool B, class T = void>
struct enable_if {};
template <class T>
struct enable_if<true, T> {
  using type = T;
};
template <bool B, class T = void>
using enable_if_t = typename enable_if<B, T>::type;
template <class T, T... Is>
struct integer_sequence {
  static constexpr size_t size() noexcept { return sizeof...(Is); }
};
namespace rtc_detail {
template <class T, class A, class B>
struct concat;
template <class T, T... A, T... B>
struct concat<T, integer_sequence<T, A...>, integer_sequence<T, B...>> {
  using type = integer_sequence<T, A..., (T)(sizeof...(A) + B)...>;
};
template <class T, int N>
struct make_seq {
  using type = typename concat<T, typename make_seq<T, N / 2>::type, typename make_seq<T, N - N / 2>::type>::type;
};
template <class T>
struct make_seq<T, 0> {
  using type = integer_sequence<T>;
};
template <class T>
struct make_seq<T, 1> {
  using type = integer_sequence<T, (T)0>;
};
}  // namespace rtc_detail
template <class T, T N>
using make_integer_sequence = typename rtc_detail::make_seq<T, (int)N>::type;
}  // namespace std
#endif

namespace b200fft {
#if defined(B200FFT_JIT_IN_U8)
typedef unsigned char in_scalar;
typedef uchar2 in_vec2;
#elif defined(B200FFT_JIT_IN_F64)
typedef double in_scalar;
typedef double2 in_vec2;
#else
typedef float in_scalar;
typedef float2 in_vec2;
#endif
}  // namespace b200fft

#ifdef B200FFT_JIT_F64
#define float double
#define float2 double2
#define float4 double4
#define make_float2 make_double2
#define make_float4 make_double4
#define fmaf fma
#endif
