// Variant table of the fused N-d kernel (fused.cuh), shared by its registration units.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "fast_registry.hpp"
#include "fused.cuh"
#include "fused2.cuh"

namespace b200fft {

struct FusedPhaseInfo {
  int kind = ND_NONE;
  int n = 0, n2 = 0;  // axis length (plane: n = x length, n2 = y length)
  int tile = 0;       // rows per tile / columns per tile
  std::vector<int> radices, radices2;
  std::string text;
};

struct FusedVariant {
  std::string name;
  std::vector<int> dims;
  bool inverse = false;
  int mode = 0;  // 0 complex in, 1 real in -> full spectrum, 2 real in -> half spectrum
  int nphases = 0;
  FusedPhaseInfo ph[ND_MAX_PHASES];
  int threads = 0;
  size_t smem = 0;
  void (*launch)(const NdArgs&, unsigned, size_t, cudaStream_t) = nullptr;
  // v2 (fused2.cuh): producer warp + bulk-async staging; threads = consumers + 32
  int default_min_batch = 0;  // > 0: used without B200FFT_FUSED=1 when the batch is at least this (measured wins only)
  bool async = false;
  cudaError_t (*launch_async)(const NdArgs&, const CUtensorMap&, const CUtensorMap&, unsigned, size_t, cudaStream_t) = nullptr;
  const void* func = nullptr;
};

std::vector<FusedVariant>& fused_registry();

template <int NT, int MINB, class P0, class P1, class P2>
struct FusedV {
  static void launch(const NdArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
    nd_fused_kernel<NT, MINB, P0, P1, P2><<<grid, NT, smem, st>>>(a);
  }
};

// The v2 kernel assigns tiles to CTAs statically and lets CTAs wait on counters other CTAs own: every CTA of the grid
// must be resident at the same time. A cooperative launch is the CUDA guarantee for exactly that (the launch waits until
// the whole grid fits, and is refused with cudaErrorCooperativeLaunchTooLarge when it never can: MPS / green-context
// partitions with fewer SMs than the plan saw).
template <int NT, int MINB, class P0, class P1, class P2, int RING = ND_RING>
struct FusedAsyncV {
  static cudaError_t launch(const NdArgs& a, const CUtensorMap& m1, const CUtensorMap& m2, unsigned grid, size_t smem,
                            cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NT + 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, nd_async_kernel<NT, MINB, P0, P1, P2, RING>, a, m1, m2);
  }
};

template <class P>
FusedPhaseInfo fused_phase_info() {
  FusedPhaseInfo i;
  i.kind = P::kind;
  i.n = P::n;
  i.tile = P::tile;
  if constexpr (P::kind == ND_R2C_PLANE) {
    i.n2 = P::n2;
    i.radices = radix_vec<typename P::RL>();
    i.radices2 = radix_vec<typename P::RLY>();
    i.text = "r2cplane" + std::to_string(P::n2) + "x" + std::to_string(P::n) + "(" + radix_name(i.radices2) + ";2;" +
             radix_name(i.radices) + ")";
  } else if constexpr (P::kind == ND_PLANE) {
    i.n2 = P::n2;
    i.radices = radix_vec<typename P::RL>();
    i.radices2 = radix_vec<typename P::RLY>();
    i.text = "plane" + std::to_string(P::n2) + "x" + std::to_string(P::n) + "(" + radix_name(i.radices2) + ";" +
             radix_name(i.radices) + ")";
  } else if constexpr (P::kind != ND_NONE) {
    i.radices = radix_vec<typename P::RL>();
    const char* k = P::kind == ND_ROWS ? "rows" : P::kind == ND_COLS ? "cols" : "r2c";
    i.text = std::string(k) + std::to_string(P::n) + "(" + (P::kind == ND_R2C ? "2;" : "") + radix_name(i.radices) + ")" +
             (P::kind == ND_COLS ? "w" : "c") + std::to_string(P::tile);
  }
  return i;
}

// <threads, min CTAs per SM (register cap), phase types...>(dims, mode): direction and real-input are
// read off the phase types
template <int NT, int MINB, class P0, class P1, class P2 = NdNone>
void reg_fused(std::vector<int> dims, int mode, const char* tag = "") {
  FusedVariant v;
  v.dims = dims;
  v.inverse = P1::inverse;
  v.mode = mode;
  v.nphases = P2::none ? 2 : 3;
  v.ph[0] = fused_phase_info<P0>();
  v.ph[1] = fused_phase_info<P1>();
  v.ph[2] = fused_phase_info<P2>();
  v.threads = NT;
  v.smem = nd_fused_smem<P0, P1, P2>();
  v.launch = &FusedV<NT, MINB, P0, P1, P2>::launch;
  v.func = (const void*)nd_fused_kernel<NT, MINB, P0, P1, P2>;
  std::string name = "nd";
  for (size_t i = 0; i < dims.size(); ++i) name += (i ? "x" : "") + std::to_string(dims[i]);
  name += std::string(v.inverse ? "_inv" : "") + (mode == 1 ? "_real" : mode == 2 ? "_r2c" : "") + "_t" + std::to_string(NT) +
          "_" + v.ph[0].text + (tag[0] ? std::string("_") + tag : "");
  v.name = name;
  fused_registry().push_back(v);
}

// v2 variants: <consumer threads, min CTAs per SM, async phase types...>(dims, mode)
template <int NT, int MINB, class P0, class P1, class P2 = ANone, int RING = ND_RING>
void reg_fused_async(std::vector<int> dims, int mode, const char* tag = "", int default_min_batch = 0) {
  FusedVariant v;
  v.dims = dims;
  v.inverse = P1::inverse;
  v.mode = mode;
  v.nphases = P2::none ? 2 : 3;
  v.ph[0] = fused_phase_info<P0>();
  v.ph[1] = fused_phase_info<P1>();
  v.ph[2] = fused_phase_info<P2>();
  v.threads = NT + 32;
  v.smem = nd_async_smem<P0, P1, P2, RING>();
  v.async = true;
  v.default_min_batch = default_min_batch;
  v.launch_async = &FusedAsyncV<NT, MINB, P0, P1, P2, RING>::launch;
  v.func = (const void*)nd_async_kernel<NT, MINB, P0, P1, P2, RING>;
  std::string name = "ndA";
  for (size_t i = 0; i < dims.size(); ++i) name += (i ? "x" : "") + std::to_string(dims[i]);
  name += std::string(v.inverse ? "_inv" : "") + (mode == 1 ? "_real" : mode == 2 ? "_r2c" : "") + "_t" + std::to_string(NT) +
          "+32_" + v.ph[0].text + (tag[0] ? std::string("_") + tag : "");
  v.name = name;
  fused_registry().push_back(v);
}

}  // namespace b200fft
