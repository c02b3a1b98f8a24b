// Fused 2-D variants, version 2 (fused2.cuh): rows (last axis) then the strided axis.
#include "fused_registry.hpp"
namespace b200fft {
void register_fused_async_2d() {
  using R24x20 = Radices<24, 20>;
  using R32x20 = Radices<32, 20>;
  using R16x15 = Radices<16, 15>;
  // Tried and dropped: exchanging IN PLACE in the staging slot (no exchange buffer -> two CTAs per SM): every value
  // has to stay in registers across a barrier (96 registers + 504 B of spills), 0.32 ms vs 0.23 ms (gpurun_out/inplace2d.log).
  // Also tried: 6-column tiles (48-byte TMA box rows, 3 x 30.7 KB -> two CTAs per SM, 192 consumer threads): 0.202 ms
  // vs 0.229 ms below and 0.174 ms for the per-axis kernels (gpurun_out/w6.log) — 2-D stays on the per-axis path.
  reg_fused_async<320, 1, ARows<480, R24x20, 8, false, false>, ACols<640, R32x20, 8, false>>({640, 480}, 0);
  reg_fused_async<320, 1, ARows<480, R24x20, 8, true, false>, ACols<640, R32x20, 8, true>>({640, 480}, 0);
  reg_fused_async<320, 1, ARows<480, R24x20, 8, false, true>, ACols<640, R32x20, 8, false>>({640, 480}, 1);
  // half spectrum: 241 columns are not a whole number of tiles -> no v2 variant (v1 covers it)
}
}  // namespace b200fft
