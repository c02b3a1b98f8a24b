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
  // ring depth 1: 32 KB slot + 37 KB exchange per CTA -> three / four smaller CTAs per SM cover each other's load latency
  reg_fused_async<128, 3, APlane<64, 64, R8x8, R8x8, INV>, ACols<64, R8x8, 32, INV>, ANone, 1>({64, 64, 64}, 0, "z32_t128_r1");
  reg_fused_async<256, 3, APlane<64, 64, R8x8, R8x8, INV>, ACols<64, R8x8, 64, INV>, ANone, 1>({64, 64, 64}, 0, "z64_t256_r1");
  reg_fused_async<192, 3, APlane<64, 64, R8x8, R8x8, INV>, ACols<64, R8x8, 32, INV>, ANone, 1>({64, 64, 64}, 0, "z32_t192_r1");
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
  using R4x8 = Radices<4, 8>;
  // measured on 100 x 64^3 (profiles/r2_r2c.md): t288 x 2 CTAs 0.113 ms, x 3 CTAs (64 regs) 0.104, t160 x 3 0.099, Z in the
  // staging slot -> 4 CTAs: t160 0.095, t128 0.092 ms (per-axis plan: 0.171, cuFFT: 0.153). x as 4x8: 0.114 (dropped).
  reg_fused_async<128, 4, AR2CPlane<64, 32, R8x8, R8x4, true>, ACols<64, R8x8, 32, false>>({64, 64, 64}, 2, "z32_t128_zslot", 1);
  reg_fused_async<160, 4, AR2CPlane<64, 32, R8x8, R8x4, true>, ACols<64, R8x8, 32, false>>({64, 64, 64}, 2, "z32_t160_zslot");
  reg_fused_async<288, 2, AR2CPlane<64, 32, R8x8, R8x4>, ACols<64, R8x8, 32, false>>({64, 64, 64}, 2, "z32");
  reg_fused_async<288, 2, AR2CPlane<64, 32, R8x8, R8x4>, ACols<64, R8x8, 64, false>>({64, 64, 64}, 2, "z64");
  // 128^3: the real plane alone is 64 KB -> one CTA per SM: 0.0883 ms vs 0.0897 ms for the per-axis plan: opt-in only
  using R16x8 = Radices<16, 8>;
  reg_fused_async<384, 1, AR2CPlane<128, 64, R16x8, R8x8, true>, ACols<128, R16x8, 32, false>>({128, 128, 128}, 2, "z32_t384");
}
void register_fused_async_3d() {
  reg3d_async<false>();
  reg3d_async<true>();
  reg3d_async_r2c();
}
}  // namespace b200fft
