// Contiguous-row variants, power-of-two lengths: <N, rows per CTA, threads, FULL, super-stages...>
// Order = preference (first variant whose super-stages can be grouped from the user's stage list).
// Measured on B200 (tools/sweep.py, profiles/r1_sweep.md): one shared-memory exchange with two
// large register codelets beats three smaller stages by 20-45 %.
// reg_rows_v4 = the same kernel with 128-bit global loads/stores and a two-lane shuffle exchange
// (fast.cuh: GlobalSrcV4): equal speed for 64/128/256 (0.1518 vs 0.1514 ms on 500000 x 128), so it is the
// default there; with a radix-32 first stage the extra selects/shuffles cost 20 % (1024: 0.290 vs 0.237 ms,
// 512^3: 1.22 vs 1.03 ms), so 512/1024 keep the 64-bit form (profiles/r1_vec128.md).
#include "fast_registry.hpp"
namespace b200fft {
void register_rows_pow2() {
  reg_rows<8, 256, 256, true, 8>();
  reg_rows<16, 128, 128, true, 16>();
  reg_rows<32, 128, 128, true, 32>();
  reg_rows_v4<64, 32, 256, true, 8, 8>();
  reg_rows<64, 32, 256, true, 8, 8>();
  reg_rows_v4<128, 32, 256, true, 16, 8>();
  reg_rows<128, 32, 256, true, 16, 8>();
  reg_rows<128, 32, 256, true, 8, 16>();
  reg_rows_v4<256, 16, 256, true, 16, 16>();
  reg_rows<256, 16, 256, true, 16, 16>();
  reg_rows<512, 8, 256, true, 32, 16>();
  reg_rows_v4<512, 8, 256, true, 32, 16>();
  reg_rows<512, 8, 256, true, 8, 8, 8>();
  reg_rows<512, 16, 256, true, 32, 16>();  // 16 rows: every thread busy in the radix-32 stage (R2C of 1024-point rows)
  reg_rows<1024, 8, 256, true, 32, 32>();
  reg_rows_v4<1024, 8, 256, true, 32, 32>();
  reg_rows<1024, 4, 256, true, 16, 16, 4>();
  reg_rows_inplace<2048, 128, 16, 16, 8>();    // 17 KB per row (12800 x 2048: 0.100 -> 0.070 ms)
  reg_rows<2048, 4, 256, true, 32, 8, 8>();
  reg_rows_inplace<4096, 256, 16, 16, 16>();   // 35 KB per row, six CTAs per SM (25000 x 4096: 0.50 -> 0.26 ms)
  reg_rows<4096, 2, 256, true, 16, 16, 16>();  // the H-point core of 8192-point R2C / C2R
  reg_rows_inplace<16384, 512, 32, 32, 16>();  // (100, 16384), fft/bench.mojo:111: one launch instead of two split passes
  reg_rows_inplace<8192, 256, 32, 16, 16>();
}
}  // namespace b200fft
