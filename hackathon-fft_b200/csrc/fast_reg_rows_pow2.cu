// Contiguous-row variants, power-of-two lengths: <N, rows per CTA, threads, FULL, super-stages...>
// Order = preference (first variant whose super-stages can be grouped from the user's stage list).
// Measured on B200 (tools/sweep.py, profiles/r1_sweep.md): one shared-memory exchange with two
// large register codelets beats three smaller stages by 20-45 %.
#include "fast_registry.hpp"
namespace b200fft {
void register_rows_pow2() {
  reg_rows<8, 256, 256, true, 8>();
  reg_rows<16, 128, 128, true, 16>();
  reg_rows<32, 128, 128, true, 32>();
  reg_rows<64, 32, 256, true, 8, 8>();
  reg_rows<128, 32, 256, true, 16, 8>();
  reg_rows<128, 32, 256, true, 8, 16>();
  reg_rows<256, 16, 256, true, 16, 16>();
  reg_rows<512, 8, 256, true, 32, 16>();
  reg_rows<512, 8, 256, true, 8, 8, 8>();
  reg_rows<1024, 8, 256, true, 32, 32>();
  reg_rows<1024, 4, 256, true, 16, 16, 4>();
  reg_rows<2048, 4, 256, true, 32, 8, 8>();
  reg_rows<4096, 2, 256, true, 16, 16, 16>();
}
}  // namespace b200fft
