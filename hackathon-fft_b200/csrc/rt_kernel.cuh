// Kernel of the runtime-length tier (see rt.cu). Included by rt.cu (host side) and by the two instantiation
// units rt_inst_{fwd,inv}_{small,large}.cu (compiled in parallel).
#pragma once
#include <cuda_runtime.h>

#include "device_utils.cuh"
#include "dft.cuh"

namespace b200fft {

constexpr int RT_MAX_STAGES = 6;
constexpr int RT_THREADS = 256;
constexpr int RT_MAX_RADIX = 32;

// q / d for 0 <= q < 2^31 with host-prepared constants (d >= 1): one multiply-high and a shift instead of the
// ~25-instruction integer division, three of which were needed per butterfly
struct RtDiv {
  unsigned mul, shift, d;
  __host__ void set(unsigned dd) {
    d = dd;
    if (dd <= 1) { mul = 0; shift = 0; return; }
    unsigned l = 0;
    while ((1ull << l) < dd) ++l;                    // ceil(log2 d)
    mul = (unsigned)(((1ull << 32) * ((1ull << l) - dd)) / dd + 1);
    shift = l;
  }
  __device__ __forceinline__ unsigned div(unsigned q) const {
    if (d <= 1) return q;
    const unsigned t = __umulhi(q, mul);
    return (t + ((q - t) >> 1)) >> (shift - 1);
  }
};

struct RtArgs {
  const void* in;
  float2* out;
  const float2* tw;       // per-stage tables, stage s >= 1 at tw + tw_off[s]: tw[(j-1)*P + p] = W_{P*R}^{j*p}
  long long outer;        // rows: number of rows; cols: number of outer slabs
  long long inner;        // element stride along the axis (1 = rows)
  long long ntiles;
  int tiles_per_outer;    // cols: ceil(inner / tile)
  int n, nstages, tile;   // tile = rows per CTA (rows) / columns per CTA (cols)
  int row;                // 1 = contiguous rows
  int in_dtype, in_comps;
  int half;               // rows only. 1 (R2C): real rows in, bins 0..n/2 out with row pitch n/2+1. 2 (C2R): rows of
                          // n/2+1 bins in, Hermitian-extended on load (X[n-k] = conj X[k]), real rows out.
  int radix[RT_MAX_STAGES];
  int tw_off[RT_MAX_STAGES];
  int stride[RT_MAX_STAGES];  // rows: padded row pitch of the exchange written by stage s
  int padP[RT_MAX_STAGES];    // rows: pad of P elements per Q-block after stage s (0 = dense)
  int buf_elems;              // elements of one exchange buffer
  RtDiv div_nb[RT_MAX_STAGES];  // by NB = n / R_s
  RtDiv div_p[RT_MAX_STAGES];   // by P_s
  RtDiv div_cn;                 // by the number of columns per tile (cols)
  float scale;
  int do_scale;
};

struct RtTile {
  long long gbase;   // first element of the tile in global memory
  long long so, si;  // global strides of (o, i); c has stride 1
  int O, CN, valid_o, valid_c;
};

template <int R, bool INV>
__device__ __forceinline__ void rt_stage(const RtArgs& a, const RtTile& t, int s, int P, const float2* cur, float2* nxt) {
  const int N = a.n, NB = N / R, Q = P * R;
  const bool first = s == 0, last = s == a.nstages - 1;
  const int total = t.O * NB * t.CN;
  const float2* __restrict__ tw = a.tw + a.tw_off[s];
  // exchange layouts: rows = [o][i] with `padP` extra elements per Q-block (bank conflicts of small-P scatters),
  // cols = dense [i][c]
  const int in_stride = first ? 0 : a.stride[s - 1], in_pad = first ? 0 : a.padP[s - 1];
  const int out_stride = last ? 0 : a.stride[s], out_pad = last ? 0 : a.padP[s];
  for (int q = threadIdx.x; q < total; q += RT_THREADS) {
    const int qn = t.CN == 1 ? q : (int)a.div_cn.div((unsigned)q);
    const int c = t.CN == 1 ? 0 : q - qn * t.CN;
    const int o = t.O == 1 ? 0 : (int)a.div_nb[s].div((unsigned)qn);
    const int n = qn - o * NB;
    const int g = (int)a.div_p[s].div((unsigned)n);
    const int p = n - g * P;
    float2 x[R];
    if (first && a.half == 2) {
      const int hb = N / 2 + 1;
      const long long rowb = (t.gbase / N + o) * hb;  // rows: gbase = first row * N
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int i = n + j * NB;
        float2 v = make_float2(0.f, 0.f);
        if (o < t.valid_o) {
          v = load_any<float>(a.in, a.in_dtype, 2, rowb + (i < hb ? i : N - i));
          if (i >= hb) v.y = -v.y;
        }
        x[j] = v;
      }
    } else if (first) {
      const bool ok = o < t.valid_o && c < t.valid_c;
      const long long e0 = t.gbase + o * t.so + (long long)n * t.si + c;
      const long long ej = (long long)NB * t.si;
      if (a.in_dtype == B200FFT_F32 && a.in_comps == 2) {  // the common case without the per-element type switch
        const float2* __restrict__ src = reinterpret_cast<const float2*>(a.in);
#pragma unroll
        for (int j = 0; j < R; ++j) x[j] = ok ? __ldg(src + e0 + j * ej) : make_float2(0.f, 0.f);
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j)
          x[j] = ok ? load_any<float>(a.in, a.in_dtype, a.in_comps, e0 + j * ej) : make_float2(0.f, 0.f);
      }
    } else if (t.CN == 1) {
      // element n + j*NB of row o; its Q-block in the previous exchange is (n + j*NB) / P = g + j*(NB/P)
      const int blocks = in_pad ? (int)a.div_p[s].div((unsigned)NB) : 0;
#pragma unroll
      for (int j = 0; j < R; ++j) x[j] = cur[o * in_stride + n + j * NB + (g + j * blocks) * in_pad];
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j) x[j] = cur[(n + j * NB) * t.CN + c];
    }
    if (P > 1) {
#pragma unroll
      for (int j = 1; j < R; ++j) x[j] = cmulf(x[j], __ldg(&tw[(j - 1) * P + p]));
    }
    Dft<R, INV>::run(x);
    if (last && a.half != 0) {
      if (o < t.valid_o) {
        const int hb = N / 2 + 1;
        const long long row = t.gbase / N + o;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int i = g * Q + p + k * P;
          float2 v = x[k];
          if (a.do_scale) { v.x *= a.scale; v.y *= a.scale; }
          if (a.half == 1) {
            if (i < hb) a.out[row * hb + i] = v;
          } else {
            reinterpret_cast<float*>(a.out)[row * N + i] = v.x;
          }
        }
      }
    } else if (last) {
      if (o < t.valid_o && c < t.valid_c) {
        float2* __restrict__ dst = a.out + t.gbase + o * t.so + c;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          float2 v = x[k];
          if (a.do_scale) { v.x *= a.scale; v.y *= a.scale; }
          dst[(long long)(g * Q + p + k * P) * t.si] = v;
        }
      }
    } else if (t.CN == 1) {
#pragma unroll
      for (int k = 0; k < R; ++k) nxt[o * out_stride + g * Q + p + k * P + g * out_pad] = x[k];
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) nxt[(g * Q + p + k * P) * t.CN + c] = x[k];
    }
  }
}

// RMAXK = 16: codelets up to radix 16 only (128 registers without spills); RMAXK = 32: all codelets, also held to
// 128 registers — the radix-27..32 codelets then spill ~2 KB, which measured FASTER than 246 registers at one CTA
// per SM (50000 x 1000: 0.51 vs 0.82 ms). The planner picks by the largest super-stage of the axis.
template <bool INV, int RMAXK>
__global__ void __launch_bounds__(RT_THREADS, 2) rt_axis_kernel(const __grid_constant__ RtArgs a) {
  extern __shared__ __align__(16) float2 smem_f2[];
  float2* buf0 = smem_f2;
  float2* buf1 = smem_f2 + a.buf_elems;
  for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    RtTile t;
    if (a.row) {
      const long long o0 = tile * a.tile;
      t.gbase = o0 * a.n;
      t.so = a.n;
      t.si = 1;
      t.O = a.tile;
      t.CN = 1;
      t.valid_o = (int)min((long long)a.tile, a.outer - o0);
      t.valid_c = 1;
    } else {
      const long long o = tile / a.tiles_per_outer;
      const long long c0 = (tile - o * a.tiles_per_outer) * a.tile;
      t.gbase = o * a.n * a.inner + c0;
      t.so = 0;
      t.si = a.inner;
      t.O = 1;
      t.CN = a.tile;
      t.valid_o = 1;
      t.valid_c = (int)min((long long)a.tile, a.inner - c0);
    }
    int P = 1;
    for (int s = 0; s < a.nstages; ++s) {
      const float2* cur = (s % 2 == 1) ? buf0 : buf1;  // stage s reads what stage s-1 wrote
      float2* nxt = (s % 2 == 0) ? buf0 : buf1;
      switch (a.radix[s]) {
#define B200_RT_CASE(R) \
  case R:               \
    if constexpr (R <= RMAXK) rt_stage<R, INV>(a, t, s, P, cur, nxt); \
    break;
        B200_RT_CASE(2) B200_RT_CASE(3) B200_RT_CASE(4) B200_RT_CASE(5) B200_RT_CASE(6) B200_RT_CASE(7) B200_RT_CASE(8)
        B200_RT_CASE(9) B200_RT_CASE(10) B200_RT_CASE(11) B200_RT_CASE(12) B200_RT_CASE(13) B200_RT_CASE(14)
        B200_RT_CASE(15) B200_RT_CASE(16) B200_RT_CASE(17) B200_RT_CASE(18) B200_RT_CASE(19) B200_RT_CASE(20)
        B200_RT_CASE(21) B200_RT_CASE(22) B200_RT_CASE(23) B200_RT_CASE(24) B200_RT_CASE(25) B200_RT_CASE(26)
        B200_RT_CASE(27) B200_RT_CASE(28) B200_RT_CASE(29) B200_RT_CASE(30) B200_RT_CASE(31) B200_RT_CASE(32)
#undef B200_RT_CASE
        default: break;
      }
      P *= a.radix[s];
      __syncthreads();
    }
  }
}

// defined in rt_inst_fwd.cu / rt_inst_inv.cu
void rt_launch_fwd_small(const RtArgs& a, unsigned grid, size_t smem, cudaStream_t stream);
void rt_launch_fwd_large(const RtArgs& a, unsigned grid, size_t smem, cudaStream_t stream);
void rt_launch_inv_small(const RtArgs& a, unsigned grid, size_t smem, cudaStream_t stream);
void rt_launch_inv_large(const RtArgs& a, unsigned grid, size_t smem, cudaStream_t stream);
cudaError_t rt_prepare_fwd_small(int max_smem);
cudaError_t rt_prepare_fwd_large(int max_smem);
cudaError_t rt_prepare_inv_small(int max_smem);
cudaError_t rt_prepare_inv_large(int max_smem);
inline void rt_launch(bool inverse, bool small, const RtArgs& a, unsigned grid, size_t smem, cudaStream_t stream) {
  if (inverse) (small ? rt_launch_inv_small : rt_launch_inv_large)(a, grid, smem, stream);
  else (small ? rt_launch_fwd_small : rt_launch_fwd_large)(a, grid, smem, stream);
}
inline cudaError_t rt_prepare(bool inverse, bool small, int max_smem) {
  if (inverse) return (small ? rt_prepare_inv_small : rt_prepare_inv_large)(max_smem);
  return (small ? rt_prepare_fwd_small : rt_prepare_fwd_large)(max_smem);
}

}  // namespace b200fft
