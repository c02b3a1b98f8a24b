// Contiguous-row variants, mixed-radix lengths: <N, rows per CTA, threads, FULL, super-stages...>
#define B200FFT_PACKED 1  // packed FADD2 complex adds: measured win for these kernels (dft.cuh)
#include "fast_registry.hpp"
namespace b200fft {
void register_rows_mixed() {
  reg_rows<93, 32, 96, true, 31, 3>();
  reg_rows<93, 16, 48, false, 31, 3>();
  reg_rows<93, 64, 96, false, 31, 3>();
  reg_rows<48, 64, 192, true, 16, 3>();     // 5-D (25,160,160,48), bench.mojo:121
  reg_rows<160, 32, 320, true, 16, 10>();
  reg_rows<240, 16, 256, true, 16, 15>();
  reg_rows<320, 16, 256, true, 20, 16>();
  reg_rows<480, 8, 192, true, 24, 20>();
  reg_rows<480, 8, 192, true, 10, 8, 6>();
  reg_rows<480, 16, 384, false, 24, 20>();
  reg_rows<480, 8, 256, false, 30, 16>();
  reg_rows<640, 8, 256, true, 32, 20>();
  // contiguous axes of the reference's 1080p / 4K / 8K shapes (bench.mojo:112-115)
  reg_rows<1080, 4, 144, true, 36, 30>();
  // complex or real input, full spectrum: one shared buffer, the middle stage exchanged in place (rows_ip_kernel; profiles/r2_long_rows.md:
  // 12000 x 2160 0.107 -> 0.075 ms, 6000 x 4320 0.104 -> 0.076); the two-buffer kernels serve R2C / C2R
  reg_rows_inplace<2160, 144, 16, 15, 9>();
  reg_rows_inplace<4320, 288, 16, 18, 15>();
  reg_rows<2160, 2, 288, true, 16, 15, 9>();
  reg_rows<4320, 1, 288, true, 16, 18, 15>();
  reg_rows<640, 4, 256, true, 10, 8, 8>();
}
}  // namespace b200fft
