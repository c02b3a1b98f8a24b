// Strided-axis variants, power-of-two lengths: <N, columns per CTA, threads, FULL, super-stages...>
#define B200FFT_PACKED 1  // packed FADD2 complex adds: measured win for these kernels (dft.cuh)
#include "fast_registry.hpp"
namespace b200fft {
void register_cols_pow2() {
  reg_cols<64, 16, 128, true, 8, 8>();
  reg_cols<64, 32, 256, true, 8, 8>();   // wider tiles for the z pass behind a plane pass (B200FFT_PREFER=cols64_8x8_w32 / _w64)
  reg_cols<64, 64, 512, true, 8, 8>();
  reg_cols<64, 32, 128, true, 8, 8>();
  reg_cols<128, 16, 128, true, 16, 8>();
  reg_cols<256, 16, 256, true, 16, 16>();
  reg_cols<512, 16, 256, true, 32, 16>();
  reg_cols<512, 8, 256, true, 8, 8, 8>();
  reg_cols<512, 8, 128, false, 32, 16>();
  reg_cols<512, 16, 512, false, 32, 16>();
  reg_cols<512, 32, 512, true, 32, 16>();  // 256-byte runs for the slab decomposition's remote stores (B200FFT_PREFER=_w32; r2_slab.md)
  reg_cols<1024, 8, 256, true, 32, 32>();
  // Tried and dropped: a 33-column tile for the middle pass of a 64^3 R2C (inner = 33 = the half spectrum of a 64-point real
  // axis, one contiguous 16.9 KB block per tile instead of 2 full 16-column tiles + 1 column): 100 x 64^3 R2C 0.1727 ms vs
  // 0.1717 ms with the 16-column tiles (profiles/r1_r2c.md) - the ragged tile is not what bounds that shape.
  // Tried and dropped: single-stage register columns (one thread = one strided 64-point transform, radix-64
  // codelet, 154 registers, no shared memory, no barrier; 17 instructions per point instead of 35):
  // 100 x 64^3 0.2099 ms vs 0.1972 ms with the two-stage 8x8 tiles, 64^4 0.1810 vs 0.1673 ms (gpurun_out/sweep_r64.log).
}
}  // namespace b200fft
