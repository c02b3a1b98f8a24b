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

// ---- forward / complex planes: the two innermost axes of a C2C (or real-input, or half-spectrum R2C) transform in one
// tile per (y, x) plane, as plain one-CTA-per-plane kernels with several CTAs per SM. Together with ONE strided pass over
// the remaining axis a 64^3 transform costs two HBM passes of simple, HBM-bound kernels instead of three — and, measured,
// less time than the single-pass persistent fused kernel, whose tile code runs at ~0.4 of either roofline.
struct PlaneFwdArgs {
  const void* in;     // C2C: [planes][NY][NX] complex (or real, REAL); R2C: [planes][NY][2H] real
  float2* out;        // C2C: [planes][NY][NX]; R2C: [planes][NY][H+1]
  const float2* twx;  // stage twiddles of the x transform (C2C: NX points; R2C: H points)
  const float2* twy;
  const float2* tw2;  // R2C: W_n^k, k = 0..H
  long long planes;
  float scale;        // inverse C2C: 1 / (NY NX)
  int do_scale;
};

// the x result handed to the y pass: dense rows with a pitch of NX + 8. With a pitch of NX the four rows a warp writes in
// the last x stage (lanes = 8 butterflies x 4 rows) start on the same bank: 2 excess wavefronts per store, a quarter of
// all STS wavefronts of the tile (ncu source page); a pitch = 8 (mod 16) complex puts them on disjoint halves of a wavefront.
template <int NX>
struct PlanePadRows {  // (o = y, i = x)
  static __device__ __forceinline__ int off(int o, int i, int) { return o * (NX + 8) + i; }
};
template <int NX>
struct PlanePadCols {  // the same buffer for the y stages: (i = y, c = x)
  static __device__ __forceinline__ int off(int, int i, int c) { return i * (NX + 8) + c; }
};

template <int NY, int NX, class RLY, class RLX>
constexpr int c2c_plane_exchange_elems() {
  constexpr int ex = max_exchange_elems<RLX, NY, RowLayoutN<NX>::template type>();
  return ex > NY * NX ? ex : NY * NX;
}
template <int NY, int NX, class RLY, class RLX>
constexpr size_t c2c_plane_smem_bytes() {
  return sizeof(float2) * (size_t)(NY * (NX + 8) + c2c_plane_exchange_elems<NY, NX, RLY, RLX>());
}

// rows of NI points, one pad element after every Q = first x radix (the RowLayout padding of the first stage: the scatter
// of butterfly n to Q n + k lands on (Q + 1) n + k, distinct banks), pitch NI + NI / Q
template <int NI, int Q>
struct PitchLayout {
  static constexpr int pitch = NI + NI / Q;
  static __device__ __forceinline__ int off(int o, int i, int) { return o * pitch + i + i / Q; }
};
// the same buffer seen by the y stages: element (i = y, c = x)
template <int NI, int Q>
struct PitchColsLayout {
  static __device__ __forceinline__ int off(int, int i, int c) { return i * (NI + NI / Q) + c + c / Q; }
};

template <int NY, int NX, class RLX>
constexpr size_t c2c_plane_ip_smem_bytes() {
  return sizeof(float2) * (size_t)NY * (NX + NX / RLX::r[0]);
}

// In-place variant of c2c_plane_kernel: ONE buffer. x stage 0 global -> B, x stage 1 in place, y stage 0 in place, y stage 1
// B -> global.
template <int NY, int NX, class RLY, class RLX, int NT, bool INV, bool REAL>
__global__ void __launch_bounds__(NT) c2c_plane_ip_kernel(const __grid_constant__ PlaneFwdArgs a) {
  static_assert(RLY::count == 2 && RLX::count == 2, "plane tiles: two super-stages per axis");
  static_assert(RLY::product() == NY && RLX::product() == NX, "radices must multiply to the axis lengths");
  extern __shared__ __align__(16) float2 smem_f2[];
  float2* B = smem_f2;
  const long long p = blockIdx.x;
  const void* in = REAL ? (const void*)(reinterpret_cast<const in_scalar*>(a.in) + p * (long long)(NY * NX))
                        : (const void*)(reinterpret_cast<const in_vec2*>(a.in) + p * (long long)(NY * NX));
  using LR = PitchLayout<NX, RLX::r[0]>;
  using LC = PitchColsLayout<NX, RLX::r[0]>;
  run_stage<RLX::r[0], 1, NX, NY, 1, NT, INV>(GlobalSrc<REAL>{in, NX, 1, NY, 1}, SmemDst<LR>{B}, a.twx, 1.f, false);
  __syncthreads();
  run_stage_inplace<RLX::r[1], RLX::r[0], NX, NY, 1, NT, INV>(SmemSrc<LR>{B}, SmemDst<LR>{B}, a.twx + RLX::tw_offset(1));
  __syncthreads();
  run_stage_inplace<RLY::r[0], 1, NY, 1, NX, NT, INV>(SmemSrc<LC>{B}, SmemDst<LC>{B}, a.twy);
  __syncthreads();
  GlobalDst dst{a.out + p * (long long)(NY * NX), 0, NX, 1, NX};
  run_stage<RLY::r[1], RLY::r[0], NY, 1, NX, NT, INV>(SmemSrc<LC>{B}, dst, a.twy + RLY::tw_offset(1), a.scale, a.do_scale != 0);
}

template <int NY, int NX, class RLY, class RLX, int NT, bool INV, bool REAL>
__global__ void __launch_bounds__(NT) c2c_plane_kernel(const __grid_constant__ PlaneFwdArgs a) {
  static_assert(RLY::count == 2 && RLX::count == 2, "plane tiles: two super-stages per axis");
  static_assert(RLY::product() == NY && RLX::product() == NX, "radices must multiply to the axis lengths");
  extern __shared__ __align__(16) float2 smem_f2[];
  float2* S = smem_f2;                   // x result, rows of pitch NX + 8
  float2* R1 = smem_f2 + NY * (NX + 8);  // exchange
  const long long p = blockIdx.x;
  const void* in = REAL ? (const void*)(reinterpret_cast<const in_scalar*>(a.in) + p * (long long)(NY * NX))
                        : (const void*)(reinterpret_cast<const in_vec2*>(a.in) + p * (long long)(NY * NX));
  using LX = typename RowLayoutN<NX>::template type<RLX::r[0], 1>;
  run_stage<RLX::r[0], 1, NX, NY, 1, NT, INV>(GlobalSrc<REAL>{in, NX, 1, NY, 1}, SmemDst<LX>{R1}, a.twx, 1.f, false);
  __syncthreads();
  run_stage<RLX::r[1], RLX::r[0], NX, NY, 1, NT, INV>(SmemSrc<LX>{R1}, SmemDst<PlanePadRows<NX>>{S}, a.twx + RLX::tw_offset(1), 1.f, false);
  __syncthreads();
  using LY = DenseLayout<NY, NX>;
  run_stage<RLY::r[0], 1, NY, 1, NX, NT, INV>(SmemSrc<PlanePadCols<NX>>{S}, SmemDst<LY>{R1}, a.twy, 1.f, false);
  __syncthreads();
  GlobalDst dst{a.out + p * (long long)(NY * NX), 0, NX, 1, NX};
  run_stage<RLY::r[1], RLY::r[0], NY, 1, NX, NT, INV>(SmemSrc<LY>{R1}, dst, a.twy + RLY::tw_offset(1), a.scale, a.do_scale != 0);
}

// bin c of row i of the half spectrum, formed on the fly from the H-point result Z[i][0..H) in shared memory
template <int H>
struct UnpackSmemSrc {
  const float2* z;
  const float2* __restrict__ w;  // W_n^k in global memory (L1-resident)
  __device__ __forceinline__ float2 load(int, int i, int c) const {
    const float2 zk = z[i * H + (c == H ? 0 : c)];
    float2 zm = z[i * H + (c == 0 ? 0 : H - c)];
    zm.y = -zm.y;
    const float2 s = make_float2(zk.x + zm.x, zk.y + zm.y), d = make_float2(zk.x - zm.x, zk.y - zm.y);
    const float2 t = cmulf(d, __ldg(&w[c]));
    return make_float2(0.5f * (s.x + t.y), 0.5f * (s.y - t.x));
  }
};

template <int NY, int H, class RLY, class RLX>
constexpr size_t r2c_plane_smem_bytes() {  // Z = NY x H, exchange = max(x exchange, NY x (H + 1))
  return sizeof(float2) * (size_t)(NY * H + c2r_plane_exchange_elems<NY, H, RLY, RLX>());
}

template <int NY, int H, class RLY, class RLX, int NT>
__global__ void __launch_bounds__(NT) r2c_plane_kernel(const __grid_constant__ PlaneFwdArgs a) {
  static_assert(RLY::count == 2 && RLX::count == 2, "plane tiles: two super-stages per axis");
  static_assert(RLY::product() == NY && RLX::product() == H, "radices must multiply to the axis lengths");
  constexpr int HB = H + 1;
  extern __shared__ __align__(16) float2 smem_f2[];
  float2* Z = smem_f2;            // H-point x result of every row, [row][H]
  float2* R1 = smem_f2 + NY * H;  // exchange
  const long long p = blockIdx.x;
  const in_vec2* in = reinterpret_cast<const in_vec2*>(a.in) + p * (long long)(NY * H);  // the real plane as NY x H complex
  using LX = typename RowLayoutN<H>::template type<RLX::r[0], 1>;
  run_stage<RLX::r[0], 1, H, NY, 1, NT, false>(GlobalSrc<false>{in, H, 1, NY, 1}, SmemDst<LX>{R1}, a.twx, 1.f, false);
  __syncthreads();
  run_stage<RLX::r[1], RLX::r[0], H, NY, 1, NT, false>(SmemSrc<LX>{R1}, SmemDst<PlaneLayout<H>>{Z}, a.twx + RLX::tw_offset(1), 1.f, false);
  __syncthreads();
  using LY = DenseLayout<NY, HB>;
  run_stage<RLY::r[0], 1, NY, 1, HB, NT, false>(UnpackSmemSrc<H>{Z, a.tw2}, SmemDst<LY>{R1}, a.twy, 1.f, false);
  __syncthreads();
  GlobalDst dst{a.out + p * (long long)(NY * HB), 0, HB, 1, HB};
  run_stage<RLY::r[1], RLY::r[0], NY, 1, HB, NT, false>(SmemSrc<LY>{R1}, dst, a.twy + RLY::tw_offset(1), 1.f, false);
}

// In-place half-spectrum forward plane: ONE buffer of NY rows with a pitch of H + H / r0 >= H + 1 complex. x stage 0 global
// -> B, x stage 1 in place (Z = the H-point result of every row), y stage 0 in place with the Hermitian unpack formed on
// load (the H + 1 bins of a row fit its pitch), y stage 1 B -> global. 128 x 128 real planes: 74 KB instead of 131 KB.
template <int H, int PX>
struct UnpackPitchSrc {  // bin c of row i from Z[i][0..H) stored with pitch PX
  const float2* z;
  const float2* __restrict__ w;
  __device__ __forceinline__ float2 load(int, int i, int c) const {
    const float2 zk = z[i * PX + (c == H ? 0 : c)];
    float2 zm = z[i * PX + (c == 0 ? 0 : H - c)];
    zm.y = -zm.y;
    const float2 s = make_float2(zk.x + zm.x, zk.y + zm.y), d = make_float2(zk.x - zm.x, zk.y - zm.y);
    const float2 t = cmulf(d, __ldg(&w[c]));
    return make_float2(0.5f * (s.x + t.y), 0.5f * (s.y - t.x));
  }
};
template <int PX>
struct PitchRowsDense {  // (o = row, i) at o * PX + i
  static __device__ __forceinline__ int off(int o, int i, int) { return o * PX + i; }
};
template <int PX>
struct PitchColsDense {  // (i = row, c) at i * PX + c
  static __device__ __forceinline__ int off(int, int i, int c) { return i * PX + c; }
};
template <int NY, int H, class RLX>
constexpr size_t r2c_plane_ip_smem_bytes() {
  return sizeof(float2) * (size_t)NY * (H + H / RLX::r[0]);
}

template <int NY, int H, class RLY, class RLX, int NT>
__global__ void __launch_bounds__(NT) r2c_plane_ip_kernel(const __grid_constant__ PlaneFwdArgs a) {
  static_assert(RLY::count == 2 && RLX::count == 2, "plane tiles: two super-stages per axis");
  static_assert(RLY::product() == NY && RLX::product() == H, "radices must multiply to the axis lengths");
  constexpr int HB = H + 1;
  constexpr int PX = H + H / RLX::r[0];
  static_assert(PX >= HB, "the H + 1 bins of a row must fit its pitch");
  extern __shared__ __align__(16) float2 smem_f2[];
  float2* B = smem_f2;
  const long long p = blockIdx.x;
  const in_vec2* in = reinterpret_cast<const in_vec2*>(a.in) + p * (long long)(NY * H);
  using LR = PitchLayout<H, RLX::r[0]>;
  run_stage<RLX::r[0], 1, H, NY, 1, NT, false>(GlobalSrc<false>{in, H, 1, NY, 1}, SmemDst<LR>{B}, a.twx, 1.f, false);
  __syncthreads();
  run_stage_inplace<RLX::r[1], RLX::r[0], H, NY, 1, NT, false>(SmemSrc<LR>{B}, SmemDst<PitchRowsDense<PX>>{B}, a.twx + RLX::tw_offset(1));
  __syncthreads();
  run_stage_inplace<RLY::r[0], 1, NY, 1, HB, NT, false>(UnpackPitchSrc<H, PX>{B, a.tw2}, SmemDst<PitchColsDense<PX>>{B}, a.twy);
  __syncthreads();
  GlobalDst dst{a.out + p * (long long)(NY * HB), 0, HB, 1, HB};
  run_stage<RLY::r[1], RLY::r[0], NY, 1, HB, NT, false>(SmemSrc<PitchColsDense<PX>>{B}, dst, a.twy + RLY::tw_offset(1), 1.f, false);
}

// In-place half-spectrum INVERSE plane (mirror of r2c_plane_ip_kernel): ONE buffer of NY rows with pitch PX = H + H / r0.
// The block's NY x (H + 1) bins are loaded into it, both y stages run in place over the H + 1 columns, x stage 0 forms
// Z[k] from the row's bins (Hermitian pack) and writes the H-point exchange layout in place, x stage 1 stores the real rows.
template <int H, int PX>
struct HermPitchSrc {
  const float2* xb;  // rows of pitch PX holding H + 1 bins
  const float2* __restrict__ tw2;
  __device__ __forceinline__ float2 load(int o, int i, int) const {
    const float2 xk = xb[o * PX + i];
    float2 xm = xb[o * PX + H - i];
    xm.y = -xm.y;
    const float2 s = make_float2(xk.x + xm.x, xk.y + xm.y), d = make_float2(xk.x - xm.x, xk.y - xm.y);
    const float2 t = cmulf(d, __ldg(&tw2[i]));
    return make_float2(s.x - t.y, s.y + t.x);
  }
};

template <int NY, int H, class RLY, class RLX, int NT>
__global__ void __launch_bounds__(NT) c2r_plane_ip_kernel(const __grid_constant__ PlaneArgs a) {
  static_assert(RLY::count == 2 && RLX::count == 2, "plane tiles: two super-stages per axis");
  static_assert(RLY::product() == NY && RLX::product() == H, "radices must multiply to the axis lengths");
  constexpr int HB = H + 1;
  constexpr int PX = H + H / RLX::r[0];
  static_assert(PX >= HB, "the H + 1 bins of a row must fit its pitch");
  extern __shared__ __align__(16) float2 smem_f2[];
  float2* B = smem_f2;
  const long long p = blockIdx.x;
  const float2* __restrict__ in = a.in + p * (long long)(NY * HB);
  constexpr int TOTAL = NY * HB;
  constexpr int ROUNDS = (TOTAL + NT - 1) / NT;
#pragma unroll 8
  for (int it = 0; it < ROUNDS; ++it) {
    const int idx = (int)threadIdx.x + it * NT;
    if (idx < TOTAL) {
      const int y = idx / HB, c = idx - y * HB;
      B[y * PX + c] = __ldg(&in[idx]);
    }
  }
  __syncthreads();
  using LC = PitchColsDense<PX>;
  run_stage_inplace<RLY::r[0], 1, NY, 1, HB, NT, true>(SmemSrc<LC>{B}, SmemDst<LC>{B}, a.twy);
  __syncthreads();
  run_stage_inplace<RLY::r[1], RLY::r[0], NY, 1, HB, NT, true>(SmemSrc<LC>{B}, SmemDst<LC>{B}, a.twy + RLY::tw_offset(1));
  __syncthreads();
  using LR = PitchLayout<H, RLX::r[0]>;
  run_stage_inplace<RLX::r[0], 1, H, NY, 1, NT, true>(HermPitchSrc<H, PX>{B, a.tw2}, SmemDst<LR>{B}, a.twx);
  __syncthreads();
  GlobalDst dst{a.out + p * (long long)(NY * H), H, 1, NY, 1};
  run_stage<RLX::r[1], RLX::r[0], H, NY, 1, NT, true>(SmemSrc<LR>{B}, dst, a.twx + RLX::tw_offset(1), a.scale, true);
}

}  // namespace b200fft
