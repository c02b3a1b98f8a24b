// Strided-axis variants, mixed-radix lengths: <N, columns per CTA, threads, FULL, super-stages...>
#define B200FFT_PACKED 1  // packed FADD2 complex adds: measured win for these kernels (dft.cuh)
#include "fast_registry.hpp"
namespace b200fft {
void register_cols_mixed() {
  reg_cols<640, 16, 320, true, 32, 20>();
  reg_cols<640, 8, 160, true, 10, 8, 8>();
  reg_cols<640, 8, 160, false, 32, 20>();
  reg_cols<640, 16, 640, false, 32, 20>();
  reg_cols<480, 16, 320, true, 24, 20>();
  reg_cols<160, 16, 160, true, 16, 10>();   // 5-D (25,160,160,48), bench.mojo:121
  reg_cols<25, 64, 64, true, 25>();
  reg_cols<48, 32, 96, true, 16, 3>();
  reg_cols<480, 8, 192, true, 10, 8, 6>();
}
}  // namespace b200fft
