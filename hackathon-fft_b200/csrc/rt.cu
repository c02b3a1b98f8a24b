// Runtime-length axis pass with compile-time codelets: the fast tier for every axis length that has no
// hand-registered variant.
//
// The registered variants (fast_reg_*.cu) fix the axis length at compile time. Everything else used to fall to
// the generic kernel, which follows the reference's per-output-point formula (O(sum r) complex FMAs per point,
// generic.cu) and is 15-25x slower. This kernel keeps the tile structure of fast.cuh — stage 0 reads global
// memory, the last stage writes it, the exchanges in between go through padded shared memory, each butterfly is a
// register codelet Dft<R> — but takes the axis length, the tile shape and the list of super-stage radices at run
// time; only the codelet is selected by a switch over R = 2..32 (all compiled into one kernel). The user's
// ordered stage list (any primes up to 31, any composites, the reference's rules) is grouped into super-stages
// of product <= 32; a stage list with a radix above 32 stays on the generic kernel.
// Serves f32 output with f32 / f64 / u8, real or complex input; rows and strided axes.
#define B200FFT_PACKED 1
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "device_utils.cuh"
#include "dft.cuh"
#include "fast_registry.hpp"
#include "plan.hpp"

namespace b200fft {

constexpr int RT_MAX_STAGES = 6;
constexpr int RT_THREADS = 256;
constexpr int RT_MAX_RADIX = 32;

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
  int radix[RT_MAX_STAGES];
  int tw_off[RT_MAX_STAGES];
  int stride[RT_MAX_STAGES];  // rows: padded row pitch of the exchange written by stage s
  int padP[RT_MAX_STAGES];    // rows: pad of P elements per Q-block after stage s (0 = dense)
  int buf_elems;              // elements of one exchange buffer
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
    const int c = t.CN == 1 ? 0 : q % t.CN;
    const int qn = t.CN == 1 ? q : q / t.CN;
    const int n = qn % NB, o = qn / NB;
    const int p = n % P, g = n / P;
    float2 x[R];
    if (first) {
      const bool ok = o < t.valid_o && c < t.valid_c;
#pragma unroll
      for (int j = 0; j < R; ++j)
        x[j] = ok ? load_any<float>(a.in, a.in_dtype, a.in_comps, t.gbase + o * t.so + (long long)(n + j * NB) * t.si + c)
                  : make_float2(0.f, 0.f);
    } else if (t.CN == 1) {
      // element n + j*NB of row o; its Q-block in the previous exchange is (n + j*NB) / P = g + j*(NB/P)
      const int blocks = NB / P;
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
    if (last) {
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

template <bool INV>
__global__ void __launch_bounds__(RT_THREADS) rt_axis_kernel(const __grid_constant__ RtArgs a) {
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
#define B200_RT_CASE(R) case R: rt_stage<R, INV>(a, t, s, P, cur, nxt); break;
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

namespace {

// contiguous partition of the ordered stage list into the fewest groups of product <= RT_MAX_RADIX, ties broken by
// the smallest maximum group (balanced super-stages)
bool group_stages(const std::vector<uint32_t>& ordered, std::vector<int>* out) {
  const size_t m = ordered.size();
  for (uint32_t r : ordered)
    if (r > (uint32_t)RT_MAX_RADIX) return false;
  std::vector<int> best;
  int best_max = 1 << 30;
  std::vector<int> cur;
  for (int want = 1; want <= RT_MAX_STAGES && best.empty(); ++want) {
    std::function<void(size_t, int)> rec = [&](size_t i, int left) {
      if (i == m) {
        if (left == 0) {
          const int mx = *std::max_element(cur.begin(), cur.end());
          if (mx < best_max) { best_max = mx; best = cur; }
        }
        return;
      }
      if (left == 0) return;
      long long prod = 1;
      for (size_t j = i; j < m; ++j) {
        prod *= ordered[j];
        if (prod > RT_MAX_RADIX) break;
        cur.push_back((int)prod);
        rec(j + 1, left - 1);
        cur.pop_back();
      }
    };
    rec(0, want);
  }
  if (best.empty()) return false;
  *out = best;
  return true;
}

struct RtPass : Pass {
  RtArgs base;
  AxisView view;
  bool inverse = false;
  size_t smem = 0;
  int sm_count = 148;
  std::string text;

  int launch(const void* src, void* dst, int64_t nbatch, cudaStream_t stream) override {
    RtArgs a = base;
    a.in = src;
    a.out = reinterpret_cast<float2*>(dst);
    const long long outer = nbatch * view.outer_per_batch;
    if (a.row) {
      a.outer = outer;
      a.ntiles = (outer + a.tile - 1) / a.tile;
    } else {
      a.outer = outer;
      a.ntiles = outer * a.tiles_per_outer;
    }
    if (a.ntiles <= 0) return B200FFT_OK;
    const unsigned grid = (unsigned)std::min<long long>(a.ntiles, (long long)sm_count * 32);
    if (inverse) rt_axis_kernel<true><<<grid, RT_THREADS, smem, stream>>>(a);
    else rt_axis_kernel<false><<<grid, RT_THREADS, smem, stream>>>(a);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    B200_CUDA_CHECK(cudaGetLastError());
    return B200FFT_OK;
  }
  std::string describe() const override { return text; }
};

}  // namespace

std::unique_ptr<Pass> make_rt_pass(b200fft_plan& plan, int axis, const AxisView& view, const IoSpec& src,
                                   bool scale_inverse, HalfMode half) {
  const Problem& p = plan.prob;
  if (half != HALF_NONE || p.desc.out_dtype != B200FFT_F32 || view.n > (1 << 20)) return nullptr;
  const AxisSpec& ax = p.axes[axis];
  std::vector<int> radices;
  if (!group_stages(ax.ordered, &radices)) return nullptr;
  const int S = (int)radices.size();
  const int n = (int)view.n;

  auto pass = std::make_unique<RtPass>();
  RtArgs& a = pass->base;
  memset(&a, 0, sizeof a);
  a.n = n;
  a.nstages = S;
  a.row = view.inner == 1;
  a.inner = view.inner;
  a.in_dtype = src.dtype;
  a.in_comps = src.comps;
  a.do_scale = scale_inverse ? 1 : 0;
  a.scale = scale_inverse ? (float)(1.0 / (double)n) : 1.f;
  // tile: ~4096 points per CTA for rows (at least enough butterflies for the threads), 16 columns for strided axes
  int tile;
  if (a.row) {
    tile = std::max(1, 4096 / n);
    const int min_radix = *std::min_element(radices.begin(), radices.end());
    (void)min_radix;
  } else {
    tile = (int)std::min<long long>(16, view.inner);
  }
  // exchange geometry
  long long P = 1;
  int max_stride = n;
  int off = 0;
  for (int s = 0; s < S; ++s) {
    const int R = radices[s];
    a.radix[s] = R;
    a.tw_off[s] = off;
    if (s > 0) off += (R - 1) * (int)P;
    const long long Q = P * R;
    const bool padded = a.row && P < 16 && Q % 2 == 0 && Q < n;
    a.padP[s] = padded ? (int)P : 0;
    a.stride[s] = n + (padded ? (int)(n / Q * P) : 0);
    if (s + 1 < S) max_stride = std::max(max_stride, a.stride[s]);
    P = Q;
  }
  const size_t nbuf = S > 2 ? 2 : (S > 1 ? 1 : 0);
  auto smem_for = [&](int t) { return nbuf * (size_t)(a.row ? (size_t)t * max_stride : (size_t)n * t) * sizeof(float2); };
  while (tile > 1 && smem_for(tile) > 160 * 1024) tile /= 2;
  if (smem_for(tile) > 227 * 1024) return nullptr;  // too long for one tile: split_registry / generic decide
  a.tile = tile;
  a.tiles_per_outer = a.row ? 1 : (int)((view.inner + tile - 1) / tile);
  a.buf_elems = (int)(a.row ? (size_t)tile * max_stride : (size_t)n * tile);
  pass->smem = smem_for(tile);
  pass->view = view;
  pass->inverse = p.desc.inverse != 0;
  pass->sm_count = plan.sm_count;
  const void* fn = pass->inverse ? (const void*)rt_axis_kernel<true> : (const void*)rt_axis_kernel<false>;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  std::vector<float2> tw = build_twiddles(radices, pass->inverse);
  float2* d_tw = nullptr;
  if (cudaMalloc(&d_tw, tw.size() * sizeof(float2)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  plan.owned_device.push_back(d_tw);
  if (cudaMemcpy(d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  a.tw = d_tw;
  std::string stages;
  for (uint32_t r : ax.ordered) stages += (stages.empty() ? "" : ",") + std::to_string(r);
  char buf[320];
  snprintf(buf, sizeof buf, "axis %d: rt_%s n=%d inner=%lld tile=%d smem=%zuB user stages=[%s] fused as (%s)%s", axis,
           a.row ? "rows" : "cols", n, (long long)view.inner, tile, pass->smem, stages.c_str(), radix_name(radices).c_str(),
           src.comps == 1 ? " real-in" : "");
  pass->text = buf;
  return pass;
}

}  // namespace b200fft
