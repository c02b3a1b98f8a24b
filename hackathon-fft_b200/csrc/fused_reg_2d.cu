// Fused 2-D variants: rows (last axis) then the strided axis, one kernel, intermediate in L2.
#include "fused_registry.hpp"
namespace b200fft {
void register_fused_2d() {
  using R24x20 = Radices<24, 20>;
  using R32x20 = Radices<32, 20>;
  using R16x15 = Radices<16, 15>;
  reg_fused<320, 2, NdRows<480, R24x20, 16, false, false>, NdCols<640, R32x20, 16, false>>({640, 480}, 0);
  reg_fused<320, 2, NdRows<480, R24x20, 16, true, false>, NdCols<640, R32x20, 16, true>>({640, 480}, 0);
  reg_fused<320, 2, NdRows<480, R24x20, 16, false, true>, NdCols<640, R32x20, 16, false>>({640, 480}, 1);
  reg_fused<320, 2, NdR2C<240, R16x15, 16>, NdCols<640, R32x20, 16, false>>({640, 480}, 2);
}
}  // namespace b200fft
