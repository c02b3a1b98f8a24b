// Half-spectrum INVERSE of the two innermost axes in one tile: a (y, x) block of NY x (H + 1) complex bins — contiguous
// in the half spectrum — becomes NY real rows of n = 2H points.
//   y: NY-point inverse over each of the H + 1 (odd!) bin columns, in shared memory;
//   x: Hermitian pack Z[k] = (X[k] + conj X[H-k]) + i W_n^{-k} (X[k] - conj X[H-k]) formed on the fly from the y result,
//      H-point inverse, stored as interleaved pairs = the real row, scaled 1 / (NY n).
// The per-axis plan needs a strided pass over the ragged 33-column extent plus a row pass for these two axes; here HBM
// sees one read of the block and one write of the real plane (the mirror image of AR2CPlane in fused2.cuh, as a plain
// one-CTA-per-plane kernel: no persistent schedule, 6 CTAs of 128 threads per SM for 64 x 64).
#pragma once
#include "fast.cuh"

namespace b200fft {

struct PlaneArgs {
  const float2* in;   // [planes][NY][H+1]
  float2* out;        // [planes][NY][H] float2 = [planes][NY][n] real
  const float2* twy;  // stage twiddles of the y transform (inverse)
  const float2* twx;  // stage twiddles of the H-point x transform (inverse)
  const float2* tw2;  // W_n^{-k}, k = 0..H
  long long planes;
  float scale;
};

// stage-0 source of the x pass: Z[k] of row o from the y result staged in shared memory as [row][H+1]
template <int H>
struct HermSmemSrc {
  const float2* xb;
  const float2* __restrict__ tw2;
  __device__ __forceinline__ float2 load(int o, int i, int) const {
    const float2 xk = xb[o * (H + 1) + i];
    float2 xm = xb[o * (H + 1) + H - i];
    xm.y = -xm.y;
    const float2 s = make_float2(xk.x + xm.x, xk.y + xm.y), d = make_float2(xk.x - xm.x, xk.y - xm.y);
    const float2 t = cmulf(d, __ldg(&tw2[i]));  // W_n^{-i} * (X[i] - conj(X[H-i]))
    return make_float2(s.x - t.y, s.y + t.x);   // s + i t
  }
};

template <int NY, int H, class RLY, class RLX>
constexpr int c2r_plane_exchange_elems() {
  constexpr int ex = max_exchange_elems<RLX, NY, RowLayoutN<H>::template type>();
  return ex > NY * (H + 1) ? ex : NY * (H + 1);
}
template <int NY, int H, class RLY, class RLX>
constexpr size_t c2r_plane_smem_bytes() {
  return sizeof(float2) * (size_t)(NY * (H + 1) + c2r_plane_exchange_elems<NY, H, RLY, RLX>());
}

template <int NY, int H, class RLY, class RLX, int NT>
__global__ void __launch_bounds__(NT) c2r_plane_kernel(const __grid_constant__ PlaneArgs a) {
  static_assert(RLY::count == 2 && RLX::count == 2, "plane tiles: two super-stages per axis");
  static_assert(RLY::product() == NY && RLX::product() == H, "radices must multiply to the axis lengths");
  constexpr int HB = H + 1;
  extern __shared__ __align__(16) float2 smem_f2[];
  float2* S = smem_f2;             // the block, later the y result [row][HB]
  float2* R1 = smem_f2 + NY * HB;  // exchange
  const long long p = blockIdx.x;
  const float2* __restrict__ in = a.in + p * (long long)(NY * HB);
  constexpr int TOTAL = NY * HB;
  constexpr int ROUNDS = (TOTAL + NT - 1) / NT;
#pragma unroll 8
  for (int it = 0; it < ROUNDS; ++it) {
    const int idx = (int)threadIdx.x + it * NT;
    if (idx < TOTAL) S[idx] = __ldg(&in[idx]);
  }
  __syncthreads();
  using LY = DenseLayout<NY, HB>;
  run_stage<RLY::r[0], 1, NY, 1, HB, NT, true>(SmemSrc<LY>{S}, SmemDst<LY>{R1}, a.twy, 1.f, false);
  __syncthreads();
  run_stage<RLY::r[1], RLY::r[0], NY, 1, HB, NT, true>(SmemSrc<LY>{R1}, SmemDst<LY>{S}, a.twy + RLY::tw_offset(1), 1.f, false);
  __syncthreads();
  using LX = typename RowLayoutN<H>::template type<RLX::r[0], 1>;
  run_stage<RLX::r[0], 1, H, NY, 1, NT, true>(HermSmemSrc<H>{S, a.tw2}, SmemDst<LX>{R1}, a.twx, 1.f, false);
  __syncthreads();
  GlobalDst dst{a.out + p * (long long)(NY * H), H, 1, NY, 1};
  run_stage<RLX::r[1], RLX::r[0], H, NY, 1, NT, true>(SmemSrc<LX>{R1}, dst, a.twx + RLX::tw_offset(1), a.scale, true);
}

}  // namespace b200fft
