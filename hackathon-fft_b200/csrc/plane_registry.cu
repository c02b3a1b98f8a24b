// Plane kernels (plane.cuh): the two innermost axes of a half-spectrum inverse in one tile per (y, x) plane.
// Registered for the plane sizes of the BASELINE 3-D shapes; the planner (api.cu: build_passes) puts the pass in place of
// the strided pass over the second-to-last axis + the C2R row pass when the stage lists can be grouped into it.
#define B200FFT_PACKED 1  // packed FADD2 complex adds (dft.cuh): strided / shared-memory-resident kernels
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "fast_registry.hpp"
#include "plan.hpp"
#include "plane.cuh"

namespace b200fft {

namespace {

struct PlaneVariant {
  int ny, h;  // NY rows, H = n / 2 complex points per row
  std::vector<int> ry, rx;
  int threads;
  size_t smem;
  void (*launch)(const PlaneArgs&, unsigned, size_t, cudaStream_t);
  const void* func;
  std::string name;
  bool by_default = true;  // false: only with B200FFT_PLANE_C2R=1 (registered, tested, measured slower than the per-axis passes)
};

template <int NY, int H, class RLY, class RLX, int NT>
void launch_plane(const PlaneArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
  c2r_plane_kernel<NY, H, RLY, RLX, NT><<<grid, NT, smem, st>>>(a);
}

template <int NY, int H, class RLY, class RLX, int NT>
PlaneVariant make_variant() {
  PlaneVariant v;
  v.ny = NY;
  v.h = H;
  v.ry = radix_vec<RLY>();
  v.rx = radix_vec<RLX>();
  v.threads = NT;
  v.smem = c2r_plane_smem_bytes<NY, H, RLY, RLX>();
  v.launch = &launch_plane<NY, H, RLY, RLX, NT>;
  v.func = (const void*)c2r_plane_kernel<NY, H, RLY, RLX, NT>;
  v.name = "c2rplane" + std::to_string(NY) + "x" + std::to_string(2 * H) + "(" + radix_name(v.ry) + ";" + radix_name(v.rx) + ";2)_t" +
           std::to_string(NT);
  return v;
}

const std::vector<PlaneVariant>& plane_registry() {
  static const std::vector<PlaneVariant> r = [] {
    std::vector<PlaneVariant> v;
    v.push_back(make_variant<64, 32, Radices<8, 8>, Radices<8, 4>, 128>());
    // 128 x 128: 133 KB of shared memory = one CTA per SM: 10 x 128^3 C2R 0.132 ms vs 0.096 ms per axis (profiles/r2_c2r.md)
    v.push_back(make_variant<128, 64, Radices<16, 8>, Radices<8, 8>, 256>());
    v.back().by_default = false;
    return v;
  }();
  return r;
}

struct PlaneC2RPass : Pass {
  const PlaneVariant* v = nullptr;
  long long planes_per_batch = 1;
  float scale = 1.f;
  float2 *twy = nullptr, *twx = nullptr, *tw2 = nullptr;
  std::string text;
  int launch(const void* src, void* dst, int64_t nbatch, cudaStream_t stream) override {
    PlaneArgs a;
    a.in = reinterpret_cast<const float2*>(src);
    a.out = reinterpret_cast<float2*>(dst);
    a.twy = twy;
    a.twx = twx;
    a.tw2 = tw2;
    a.planes = nbatch * planes_per_batch;
    a.scale = scale;
    if (a.planes <= 0) return B200FFT_OK;
    if (a.planes > 0x7fffffffLL) return fail(B200FFT_ERR_UNSUPPORTED, "too many planes");
    v->launch(a, (unsigned)a.planes, v->smem, stream);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    B200_CUDA_CHECK(cudaGetLastError());
    return B200FFT_OK;
  }
  std::string describe() const override { return text; }
};

}  // namespace

// the two innermost axes of a half-spectrum inverse as one pass, or nullptr when no variant covers them
std::unique_ptr<Pass> make_plane_c2r_pass(b200fft_plan& plan) {
  const Problem& p = plan.prob;
  if (!p.half || !p.desc.inverse || p.rank < 2) return nullptr;
  if (p.desc.out_dtype != B200FFT_F32 || p.desc.in_dtype != B200FFT_F32) return nullptr;
  if (p.desc.flags & (B200FFT_FLAG_FORCE_GENERIC | B200FFT_FLAG_FORCE_RT | B200FFT_FLAG_NO_FUSED)) return nullptr;
  bool force = false;
  if (const char* e = getenv("B200FFT_PLANE_C2R")) {
    if (atoi(e) == 0) return nullptr;
    force = true;
  }
  const int last = p.rank - 1;
  if (!p.axes[last].transformed || !p.axes[last - 1].transformed || p.axes[last].n % 2) return nullptr;
  for (const PlaneVariant& v : plane_registry()) {
    if (p.axes[last - 1].n != v.ny || p.axes[last].n != 2 * v.h || (!v.by_default && !force)) continue;
    if (!can_group(p.axes[last - 1].ordered, v.ry)) continue;
    bool okx = false;
    for (const auto& o : drop_factor_two(p.axes[last].ordered)) okx = okx || can_group(o, v.rx);
    if (!okx) continue;
    if (v.smem > 48 * 1024 &&
        cudaFuncSetAttribute(v.func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    auto pass = std::make_unique<PlaneC2RPass>();
    pass->v = &v;
    pass->planes_per_batch = 1;
    for (int a = 0; a < last - 1; ++a) pass->planes_per_batch *= p.axes[a].n;
    pass->scale = (float)(1.0 / ((double)v.ny * 2.0 * v.h));
    auto upload = [&](const std::vector<float2>& t, float2** d) {
      if (cudaMalloc(d, t.size() * sizeof(float2)) != cudaSuccess) { cudaGetLastError(); return false; }
      plan.owned_device.push_back(*d);
      return cudaMemcpy(*d, t.data(), t.size() * sizeof(float2), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    if (!upload(build_twiddles(v.ry, true), &pass->twy) || !upload(build_twiddles(v.rx, true), &pass->twx) ||
        !upload(build_half_twiddles(2 * v.h, true), &pass->tw2))
      return nullptr;
    char buf[320];
    snprintf(buf, sizeof buf, "axes %d,%d: %s: (y, x) plane of %d x %d bins -> %d x %d reals in one tile (y inverse, Hermitian pack, "
             "x inverse), smem=%zuB", last - 1, last, v.name.c_str(), v.ny, v.h + 1, v.ny, 2 * v.h, v.smem);
    pass->text = buf;
    return pass;
  }
  return nullptr;
}

}  // namespace b200fft
