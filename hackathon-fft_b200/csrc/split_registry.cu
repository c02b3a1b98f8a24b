// Two-pass ("four-step") treatment of axes that are too long for one tile: N = N1 * N2,
//   pass A  N1-point strided transforms with the W_N^{k1*n2} twiddle fused into the store  (src -> temp)
//   pass B  N2-point transforms whose store puts output k2 at axis index k1 + N1*k2        (temp -> dst)
// so a long axis costs two tile passes and one plan-owned temporary, never a transpose kernel. Used when no
// single-tile variant exists for the axis length (the reference's published shapes (100, 16384), 1920x1080,
// 3840x2160, 7680x4320: fft/bench.mojo:107-122). The user's stage list must be groupable into the two
// variants' super-stages, as for every other fast kernel.
#define B200FFT_PACKED 1  // packed FADD2 complex adds (dft.cuh): strided / mixed-radix kernels
#include "split_registry.hpp"

#include <cstring>

#include "plan.hpp"

namespace b200fft {

std::vector<SplitKernel>& split_registry() {
  static std::vector<SplitKernel> r;
  return r;
}

namespace {

void register_split() {
  static const bool done = [] {  // thread-safe one-time initialisation (see register_all, fast_registry.cu)
    // pass A candidates (N1): <role, N, columns per tile, threads, super-stages...>
    reg_split<SPLIT_A, 128, 16, 128, 16, 8>();
    reg_split<SPLIT_A, 256, 16, 256, 16, 16>();
    reg_split<SPLIT_A, 512, 16, 256, 32, 16>();
    reg_split<SPLIT_A, 512, 16, 256, 8, 8, 8>();  // the reference's default bases for 7680 are [15, 8, 8, 8]
    reg_split<SPLIT_A, 64, 16, 128, 8, 8>();
    // pass B candidates (N2)
    reg_split<SPLIT_B_COLS, 15, 128, 128, 15>();
    reg_split<SPLIT_B_COLS, 30, 64, 64, 30>();
    reg_split<SPLIT_B_COLS, 16, 128, 128, 16>();
    reg_split<SPLIT_B_ROWS, 128, 32, 256, 16, 8>();
    reg_split<SPLIT_B_ROWS, 64, 32, 256, 8, 8>();
    reg_split<SPLIT_B_ROWS, 256, 16, 256, 16, 16>();
    return true;
  }();
  (void)done;
}

struct SplitPass : Pass {
  const SplitKernel *ka = nullptr, *kb = nullptr;
  AxisView view;
  bool inverse = false;
  float scale = 1.f;
  int do_scale = 0;
  float2 *twa = nullptr, *twb = nullptr, *tmp = nullptr;
  const float2* twN = nullptr;
  std::string text;

  int launch(const void* src, void* dst, int64_t nbatch, cudaStream_t stream) override {
    const long long outer = nbatch * view.outer_per_batch;
    if (outer <= 0) return B200FFT_OK;
    SplitArgs a;
    memset(&a, 0, sizeof a);
    a.inner = view.inner;
    a.n1 = ka->n;
    a.n2 = kb->n;
    // pass A: view (outer, n1, n2 * inner)
    a.in = reinterpret_cast<const float2*>(src);
    a.out = tmp;
    a.tw = twa;
    a.twN = twN;
    const long long vinner = (long long)a.n2 * view.inner;
    a.tiles_per_outer = (int)((vinner + ka->tile - 1) / ka->tile);
    long long grid = outer * a.tiles_per_outer;
    if (grid > 0x7fffffffLL) return fail(B200FFT_ERR_UNSUPPORTED, "too many tiles");
    ka->launch(inverse, a, (unsigned)grid, ka->smem, stream);
    // pass B
    a.in = tmp;
    a.out = reinterpret_cast<float2*>(dst);
    a.tw = twb;
    a.scale = scale;
    a.do_scale = do_scale;
    if (kb->role == SPLIT_B_ROWS) {
      a.nrows = outer * a.n1;
      grid = (a.nrows + kb->tile - 1) / kb->tile;
    } else {
      a.tiles_per_outer = (int)((view.inner + kb->tile - 1) / kb->tile);
      grid = outer * a.n1 * a.tiles_per_outer;
    }
    if (grid > 0x7fffffffLL) return fail(B200FFT_ERR_UNSUPPORTED, "too many tiles");
    kb->launch(inverse, a, (unsigned)grid, kb->smem, stream);
    g_launch_count.fetch_add(2, std::memory_order_relaxed);
    B200_CUDA_CHECK(cudaGetLastError());
    return B200FFT_OK;
  }
  std::string describe() const override { return text; }
  int launches() const override { return 2; }
};

}  // namespace

size_t split_variant_count() {
  register_split();
  return split_registry().size();
}

std::unique_ptr<Pass> make_split_pass(b200fft_plan& plan, int axis, const AxisView& view, const IoSpec& src,
                                      bool scale_inverse, HalfMode half) {
  register_split();
  const Problem& p = plan.prob;
  if (half != HALF_NONE || p.desc.out_dtype != B200FFT_F32 || src.dtype != B200FFT_F32 || src.comps != 2) return nullptr;
  if (view.n > (1 << 24) || !plan.tw[axis].ptr) return nullptr;
  const AxisSpec& ax = p.axes[axis];
  const SplitKernel *best_a = nullptr, *best_b = nullptr;
  for (const SplitKernel& ka : split_registry()) {
    if (ka.role != SPLIT_A || view.n % ka.n) continue;
    const long long n2 = view.n / ka.n;
    for (const SplitKernel& kb : split_registry()) {
      if (kb.n != n2) continue;
      if (view.inner == 1 ? kb.role != SPLIT_B_ROWS : kb.role != SPLIT_B_COLS) continue;
      if (kb.role == SPLIT_B_ROWS && ka.n % kb.tile) continue;  // row tiles must not straddle transforms
      std::vector<int> all(ka.radices);
      all.insert(all.end(), kb.radices.begin(), kb.radices.end());
      if (!can_group(ax.ordered, all)) continue;
      if (!best_a) { best_a = &ka; best_b = &kb; }
    }
  }
  if (!best_a) return nullptr;
  if (best_a->prepare(best_a->smem) != cudaSuccess || best_b->prepare(best_b->smem) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  auto pass = std::make_unique<SplitPass>();
  pass->ka = best_a;
  pass->kb = best_b;
  pass->view = view;
  pass->inverse = p.desc.inverse != 0;
  pass->do_scale = scale_inverse ? 1 : 0;
  pass->scale = scale_inverse ? (float)(1.0 / (double)view.n) : 1.f;
  pass->twN = reinterpret_cast<const float2*>(plan.tw[axis].ptr);
  auto upload = [&](const std::vector<float2>& t, float2** out) {
    if (cudaMalloc(out, t.size() * sizeof(float2)) != cudaSuccess) return false;
    plan.owned_device.push_back(*out);
    return cudaMemcpy(*out, t.data(), t.size() * sizeof(float2), cudaMemcpyHostToDevice) == cudaSuccess;
  };
  if (!upload(build_twiddles(best_a->radices, pass->inverse), &pass->twa) ||
      !upload(build_twiddles(best_b->radices, pass->inverse), &pass->twb)) {
    cudaGetLastError();
    return nullptr;
  }
  const size_t tmp_bytes = (size_t)p.batch * view.outer_per_batch * view.n * view.inner * sizeof(float2);
  if (cudaMalloc(&pass->tmp, tmp_bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  plan.owned_device.push_back(pass->tmp);
  plan.workspace_bytes += tmp_bytes;
  std::string stages;
  for (uint32_t r : ax.ordered) stages += (stages.empty() ? "" : ",") + std::to_string(r);
  char buf[400];
  snprintf(buf, sizeof buf, "axis %d: split n=%lld = %d x %d inner=%lld: %s (+W_n twiddle) -> %s (natural-order store); temp=%zuB user stages=[%s]",
           axis, (long long)view.n, best_a->n, best_b->n, (long long)view.inner, best_a->name.c_str(), best_b->name.c_str(),
           tmp_bytes, stages.c_str());
  pass->text = buf;
  return pass;
}

}  // namespace b200fft
