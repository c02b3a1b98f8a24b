// Fused 3-D variants: <threads, min CTAs/SM, phases...>(dims, mode). Order = preference.
//   plane + cols   : (y, x) in shared memory per z plane, then the z axis       (2 L2-level passes)
//   rows+cols+cols : x, y, z as three phases (planes too big for shared memory) (3 L2-level passes)
#include "fused_registry.hpp"
namespace b200fft {
template <bool INV>
static void reg3d() {
  using R8x8 = Radices<8, 8>;
  using R16x8 = Radices<16, 8>;
  using R16x16 = Radices<16, 16>;
  using R32x16 = Radices<32, 16>;
  reg_fused<256, 4, NdPlane<64, 64, R8x8, R8x8, 32, 32, INV, false>, NdCols<64, R8x8, 32, INV>>({64, 64, 64}, 0);
  reg_fused<256, 4, NdRows<64, R8x8, 32, INV, false>, NdCols<64, R8x8, 32, INV>, NdCols<64, R8x8, 32, INV>>({64, 64, 64}, 0);
  reg_fused<512, 1, NdPlane<128, 128, R16x8, R16x8, 64, 64, INV, false>, NdCols<128, R16x8, 64, INV>>({128, 128, 128}, 0);
  reg_fused<256, 2, NdRows<128, R16x8, 32, INV, false>, NdCols<128, R16x8, 32, INV>, NdCols<128, R16x8, 32, INV>>(
      {128, 128, 128}, 0);
  reg_fused<256, 3, NdRows<256, R16x16, 16, INV, false>, NdCols<256, R16x16, 16, INV>, NdCols<256, R16x16, 16, INV>>(
      {256, 256, 256}, 0);
  reg_fused<256, 2, NdRows<512, R32x16, 8, INV, false>, NdCols<512, R32x16, 16, INV>, NdCols<512, R32x16, 16, INV>>(
      {512, 512, 512}, 0);
}
void register_fused_3d() {
  reg3d<false>();
  reg3d<true>();
}
}  // namespace b200fft
