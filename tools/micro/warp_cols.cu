// Microbenchmark (round 2): is the SM-side cost of a strided 64-point pass set by CTA barriers / instruction count?
// Compares, in place on the same array, L2-resident (32 MB) and HBM-resident (1 GB):
//   A  cols_kernel<64, Radices<8,8>, 16, 128>   the shipped CTA-tile kernel (two stages, __syncthreads between them)
//   B  wcols64<false>   warp-private tiles: each warp owns 8 adjacent columns, two butterflies (adjacent columns) per
//                       thread, LDG.128 / STS.128 / LDS.128 / STG.128, stage twiddles in registers, __syncwarp only
//   C  wcols64<true>    the same with the two butterflies packed SoA into f32x2 registers (FADD2 / FMUL2 / FFMA2)
//   nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo \
//        -I hackathon-fft_b200/csrc -o tools/build/warp_cols tools/micro/warp_cols.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <vector>

#include "fast.cuh"

using namespace b200fft;

// ---- packed pairs: one 64-bit register pair = the same component of two butterflies -------------------------------
typedef unsigned long long p2;
__device__ __forceinline__ p2 pk(float a, float b) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(p2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ p2 padd(p2 a, p2 b) { p2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 psub(p2 a, p2 b) { p2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 pmul(p2 a, p2 b) { p2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 pfma(p2 a, p2 b, p2 c) { p2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
struct pc { p2 re, im; };
__device__ __forceinline__ pc operator+(pc a, pc b) { return {padd(a.re, b.re), padd(a.im, b.im)}; }
__device__ __forceinline__ pc operator-(pc a, pc b) { return {psub(a.re, b.re), psub(a.im, b.im)}; }
// forward direction: multiply by -i
__device__ __forceinline__ pc mul_mi(pc a) { return {a.im, a.re}; }  // (im, -re): the sign is folded by the caller
// packed forward DFT8 (DIT-ordered like Dft<8>: same outputs), exploiting add/sub folding for the -i rotations
__device__ __forceinline__ void pdft8(pc (&x)[8]) {
  const p2 H = pk(0.70710678118654752440f, 0.70710678118654752440f);
  // radix-2 layer over (j, j+4)
  pc a0 = x[0] + x[4], b0 = x[0] - x[4];
  pc a1 = x[1] + x[5], b1 = x[1] - x[5];
  pc a2 = x[2] + x[6], b2 = x[2] - x[6];
  pc a3 = x[3] + x[7], b3 = x[3] - x[7];
  // even outputs: DFT4 of a0..a3
  pc c0 = a0 + a2, c1 = a0 - a2, c2 = a1 + a3, d3 = a1 - a3;  // c3 = -i * d3 = (d3.im, -d3.re)
  x[0] = c0 + c2;
  x[4] = c0 - c2;
  x[2] = {padd(c1.re, d3.im), psub(c1.im, d3.re)};
  x[6] = {psub(c1.re, d3.im), padd(c1.im, d3.re)};
  // odd outputs: b_k * W8^k then DFT4
  // b1 * W8 = ((re + im), (im - re)) * h ; b2 * (-i) = (im, -re) ; b3 * W8^3 = ((im - re), -(re + im)) * h
  pc t1 = {pmul(padd(b1.re, b1.im), H), pmul(psub(b1.im, b1.re), H)};
  pc t3 = {pmul(psub(b3.im, b3.re), H), pmul(padd(b3.re, b3.im), H)};  // im carries the opposite sign: use (re, -im)
  // e0 = b0 + (-i b2), e1 = b0 - (-i b2); f0 = t1 + t3', f1 = t1 - t3' with t3' = (t3.re, -t3.im)
  pc e0 = {padd(b0.re, b2.im), psub(b0.im, b2.re)};
  pc e1 = {psub(b0.re, b2.im), padd(b0.im, b2.re)};
  pc f0 = {padd(t1.re, t3.re), psub(t1.im, t3.im)};
  pc g1 = {psub(t1.re, t3.re), padd(t1.im, t3.im)};  // f1; -i * f1 = (f1.im, -f1.re)
  x[1] = e0 + f0;
  x[5] = e0 - f0;
  x[3] = {padd(e1.re, g1.im), psub(e1.im, g1.re)};
  x[7] = {psub(e1.re, g1.im), padd(e1.im, g1.re)};
}

constexpr int WROWS = 64, WPITCH = 4;                  // float4 per row (4 column pairs)
constexpr int WSMEM4 = WROWS * WPITCH + 8 * WPITCH;    // + one row of padding per 8 rows
__device__ __forceinline__ int wrow(int i) { return (i + (i >> 3)) * WPITCH; }

struct WArgs {
  float2* data;
  const float2* tw;  // stage-1 table of Radices<8,8>: tw[(j-1)*8 + p] = W_64^{j p}
  long long inner;
  int tiles_per_outer;  // inner / 8
  long long ntiles;
};

template <bool PACKED>
__global__ void __launch_bounds__(256) wcols64(const __grid_constant__ WArgs a) {
  __shared__ __align__(16) float4 sm[8][WSMEM4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cp = lane & 3, n = lane >> 2;
  float4* ex = sm[warp];
  float2 w[7];
#pragma unroll
  for (int j = 1; j < 8; ++j) w[j - 1] = __ldg(a.tw + (j - 1) * 8 + n);
  const long long wstride = (long long)gridDim.x * 8;
  for (long long t = (long long)blockIdx.x * 8 + warp; t < a.ntiles; t += wstride) {
    const long long o = t / a.tiles_per_outer;
    const long long c0 = (t - o * a.tiles_per_outer) * 8 + 2 * cp;
    float2* base = a.data + o * 64 * a.inner + c0;
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldcg(reinterpret_cast<const float4*>(base + (long long)(n + 8 * j) * a.inner));
    if constexpr (PACKED) {
      pc x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = {pk(v[j].x, v[j].z), pk(v[j].y, v[j].w)};
      pdft8(x);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float4 s;
        upk(x[k].re, s.x, s.y);
        upk(x[k].im, s.z, s.w);
        ex[wrow(n * 8 + k) + cp] = s;  // SoA: (re_a, re_b, im_a, im_b)
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 s = ex[wrow(n + 8 * j) + cp];
        x[j] = {pk(s.x, s.y), pk(s.z, s.w)};
      }
#pragma unroll
      for (int j = 1; j < 8; ++j) {
        const p2 wr = pk(w[j - 1].x, w[j - 1].x), wi = pk(w[j - 1].y, w[j - 1].y), nwi = pk(-w[j - 1].y, -w[j - 1].y);
        const pc y = {pfma(x[j].im, nwi, pmul(x[j].re, wr)), pfma(x[j].im, wr, pmul(x[j].re, wi))};
        x[j] = y;
      }
      pdft8(x);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float4 s;
        upk(x[k].re, s.x, s.z);
        upk(x[k].im, s.y, s.w);
        *reinterpret_cast<float4*>(base + (long long)(n + 8 * k) * a.inner) = s;
      }
    } else {
      float2 xa[8], xb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { xa[j] = make_float2(v[j].x, v[j].y); xb[j] = make_float2(v[j].z, v[j].w); }
      Dft<8, false>::run(xa);
      Dft<8, false>::run(xb);
#pragma unroll
      for (int k = 0; k < 8; ++k) ex[wrow(n * 8 + k) + cp] = make_float4(xa[k].x, xa[k].y, xb[k].x, xb[k].y);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 s = ex[wrow(n + 8 * j) + cp];
        xa[j] = make_float2(s.x, s.y);
        xb[j] = make_float2(s.z, s.w);
      }
#pragma unroll
      for (int j = 1; j < 8; ++j) { xa[j] = cmulf(xa[j], w[j - 1]); xb[j] = cmulf(xb[j], w[j - 1]); }
      Dft<8, false>::run(xa);
      Dft<8, false>::run(xb);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        *reinterpret_cast<float4*>(base + (long long)(n + 8 * k) * a.inner) = make_float4(xa[k].x, xa[k].y, xb[k].x, xb[k].y);
    }
  }
}

static std::vector<float2> stage_tw() {
  std::vector<float2> t(7 * 8);
  for (int j = 1; j < 8; ++j)
    for (int p = 0; p < 8; ++p) {
      const double th = -2.0 * M_PI * (double)(j * p) / 64.0;
      t[(j - 1) * 8 + p] = make_float2((float)cos(th), (float)sin(th));
    }
  return t;
}

int main() {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  std::vector<float2> tw = stage_tw();
  float2* d_tw;
  cudaMalloc(&d_tw, tw.size() * sizeof(float2));
  cudaMemcpy(d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice);
  using RL = Radices<8, 8>;
  auto kA = cols_kernel<64, RL, 16, 128, false, false>;
  const size_t smemA = cols_smem_bytes<64, RL, 16>();
  cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA);
  for (long long inner : {4096LL, 64LL}) {
    for (long long ntrans : {16LL, 500LL}) {  // x 64^3 complex = 2 MB each: 32 MB (L2) and 1 GB (HBM)
      const long long pts = ntrans * 64 * 64 * 64;
      const long long outer = pts / (64 * inner);
      float2 *d, *ref;
      cudaMalloc(&d, pts * sizeof(float2));
      cudaMalloc(&ref, pts * sizeof(float2));
      std::vector<float2> h((size_t)std::min<long long>(pts, 1 << 22));
      for (size_t i = 0; i < h.size(); ++i) h[i] = make_float2((float)((i * 7919) % 1000) / 500.f - 1.f, (float)((i * 104729) % 1000) / 500.f - 1.f);
      for (long long off = 0; off < pts; off += (long long)h.size())
        cudaMemcpy(d + off, h.data(), std::min<long long>((long long)h.size(), pts - off) * sizeof(float2), cudaMemcpyHostToDevice);
      cudaMemcpy(ref, d, pts * sizeof(float2), cudaMemcpyDeviceToDevice);
      // correctness: one application of A on ref vs B / C on copies
      ColsArgs ca{ref, ref, d_tw, inner, (int)(inner / 16), 1.f, 0};
      kA<<<(unsigned)(outer * ca.tiles_per_outer), 128, smemA>>>(ca);
      WArgs wa{d, d_tw, inner, (int)(inner / 8), outer * (inner / 8)};
      double err[2] = {0, 0};
      for (int variant = 0; variant < 2; ++variant) {
        float2* tmp;
        cudaMalloc(&tmp, pts * sizeof(float2));
        cudaMemcpy(tmp, d, pts * sizeof(float2), cudaMemcpyDeviceToDevice);
        WArgs wt = wa;
        wt.data = tmp;
        if (variant == 0) wcols64<false><<<148 * 4, 256>>>(wt);
        else wcols64<true><<<148 * 4, 256>>>(wt);
        std::vector<float2> g(h.size()), r(h.size());
        cudaMemcpy(g.data(), tmp, g.size() * sizeof(float2), cudaMemcpyDeviceToHost);
        cudaMemcpy(r.data(), ref, r.size() * sizeof(float2), cudaMemcpyDeviceToHost);
        double num = 0, den = 0;
        for (size_t i = 0; i < g.size(); ++i) {
          num += (double)(g[i].x - r[i].x) * (g[i].x - r[i].x) + (double)(g[i].y - r[i].y) * (g[i].y - r[i].y);
          den += (double)r[i].x * r[i].x + (double)r[i].y * r[i].y;
        }
        err[variant] = sqrt(num / den);
        cudaFree(tmp);
      }
      auto timeit = [&](auto&& launch) {
        for (int i = 0; i < 3; ++i) launch();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        return ms / 20;
      };
      ColsArgs cd{d, d, d_tw, inner, (int)(inner / 16), 1.f, 0};
      const float tA = timeit([&] { kA<<<(unsigned)(outer * cd.tiles_per_outer), 128, smemA>>>(cd); });
      float tB[3], tC[3];
      int gi = 0;
      for (int mult : {2, 4, 8}) {
        tB[gi] = timeit([&] { wcols64<false><<<148 * mult, 256>>>(wa); });
        tC[gi] = timeit([&] { wcols64<true><<<148 * mult, 256>>>(wa); });
        ++gi;
      }
      const double gb = 2.0 * pts * 8 / 1e9;
      printf("{\"inner\": %lld, \"MB\": %lld, \"A_cols16_ms\": %.4f, \"A_TBs\": %.2f, \"B_warp_ms\": [%.4f, %.4f, %.4f], \"B_TBs\": %.2f, "
             "\"C_packed_ms\": [%.4f, %.4f, %.4f], \"C_TBs\": %.2f, \"relerr_B\": %.2e, \"relerr_C\": %.2e, \"cuda\": \"%s\"}\n",
             inner, pts * 8 >> 20, tA, gb / tA, tB[0], tB[1], tB[2], gb / fminf(tB[0], fminf(tB[1], tB[2])), tC[0], tC[1], tC[2],
             gb / fminf(tC[0], fminf(tC[1], tC[2])), err[0], err[1], cudaGetErrorString(cudaGetLastError()));
      cudaFree(d);
      cudaFree(ref);
    }
  }
  return 0;
}
