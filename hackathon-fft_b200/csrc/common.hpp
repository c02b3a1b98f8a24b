// Shared host-side helpers: status codes, thread-local error text, launch counter.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/b200fft.h"

namespace b200fft {

// thread-local detail string behind b200fft_last_error()
std::string& last_error();
int fail(int status, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

extern std::atomic<uint64_t> g_launch_count;

}  // namespace b200fft

#define B200_CUDA_CHECK(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::b200fft::fail(B200FFT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                       \
  } while (0)
