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
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "fast_registry.hpp"
#include "rt_kernel.cuh"
#include "plan.hpp"

namespace b200fft {

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
    rt_launch(inverse, a, grid, smem, stream);
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
  if (p.desc.out_dtype != B200FFT_F32 || view.n > (1 << 20)) return nullptr;
  if (half != HALF_NONE && view.inner != 1) return nullptr;  // half-spectrum handling is a row pass
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
  a.half = (int)half;
  a.do_scale = (scale_inverse || half == HALF_C2R) ? 1 : 0;  // the half-spectrum inverse is always normalised
  a.scale = a.do_scale ? (float)(1.0 / (double)n) : 1.f;
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
  if (rt_prepare(pass->inverse, 227 * 1024) != cudaSuccess) {
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
           half == HALF_R2C ? " r2c (n-point, bins 0..n/2 stored)" : half == HALF_C2R ? " c2r (Hermitian load)"
                                                                                       : src.comps == 1 ? " real-in" : "");
  pass->text = buf;
  return pass;
}

}  // namespace b200fft
