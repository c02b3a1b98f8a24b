// Strided-axis variants that move their tiles with TMA (cp.async.bulk.tensor.3d):
// <N, columns per CTA, threads, super-stages...>. Registered after the LDG variants; the
// order below / B200FFT_PREFER decides which is tried first (see profiles/ for the comparison).
#include "fast_registry.hpp"
namespace b200fft {
void register_cols_tma() {
  reg_cols_tma<64, 16, 128, 8, 8>();
  reg_cols_tma<128, 16, 128, 16, 8>();
  reg_cols_tma<256, 16, 256, 16, 16>();
  reg_cols_tma<512, 16, 256, 32, 16>();
  reg_cols_tma<512, 8, 128, 32, 16>();
  reg_cols_tma<640, 16, 320, 32, 20>();
  reg_cols_tma<640, 8, 160, 32, 20>();
  reg_cols_tma<480, 16, 320, 24, 20>();
}
}  // namespace b200fft
