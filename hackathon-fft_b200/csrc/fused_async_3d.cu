// Fused 3-D variants, version 2 (fused2.cuh: producer warp + bulk-async staging):
// <consumer threads, min CTAs/SM, phases...>(dims, mode). Order = preference.
#include "fused_registry.hpp"
namespace b200fft {
template <bool INV>
static void reg3d_async() {
  using R8x8 = Radices<8, 8>;
  using R16x8 = Radices<16, 8>;
  using R16x16 = Radices<16, 16>;
  using R32x16 = Radices<32, 16>;
  reg_fused_async<256, 2, APlane<64, 64, R8x8, R8x8, INV>, ACols<64, R8x8, 64, INV>>({64, 64, 64}, 0, "z64", 1);
  reg_fused_async<256, 2, APlane<64, 64, R8x8, R8x8, INV>, ACols<64, R8x8, 32, INV>>({64, 64, 64}, 0);
  reg_fused_async<256, 2, ARows<64, R8x8, 32, INV, false>, ACols<64, R8x8, 32, INV>, ACols<64, R8x8, 32, INV>>({64, 64, 64}, 0);
  reg_fused_async<256, 2, ARows<128, R16x8, 32, INV, false>, ACols<128, R16x8, 32, INV>, ACols<128, R16x8, 32, INV>>(
      {128, 128, 128}, 0, "", 4);
  reg_fused_async<256, 2, ARows<256, R16x16, 16, INV, false>, ACols<256, R16x16, 16, INV>, ACols<256, R16x16, 16, INV>>(
      {256, 256, 256}, 0);
  reg_fused_async<256, 2, ARows<512, R32x16, 8, INV, false>, ACols<512, R32x16, 8, INV>, ACols<512, R32x16, 8, INV>>(
      {512, 512, 512}, 0);
}
// half-spectrum R2C: (y, x) planes with the Hermitian unpack and the y pass in shared memory, then the strided z phase
// over inner = NY * (n/2 + 1) (whole 64-column tiles)
static void reg3d_async_r2c() {
  using R8x8 = Radices<8, 8>;
  using R8x4 = Radices<8, 4>;
  reg_fused_async<288, 2, AR2CPlane<64, 32, R8x8, R8x4>, ACols<64, R8x8, 32, false>>({64, 64, 64}, 2, "z32", 1);
  reg_fused_async<288, 2, AR2CPlane<64, 32, R8x8, R8x4>, ACols<64, R8x8, 64, false>>({64, 64, 64}, 2, "z64");
}
void register_fused_async_3d() {
  reg3d_async<false>();
  reg3d_async<true>();
  reg3d_async_r2c();
}
}  // namespace b200fft
