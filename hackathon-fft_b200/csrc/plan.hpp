// Plan = validated problem + device twiddle tables + an ordered list of kernel passes.
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <vector>

#include "planner.hpp"

namespace b200fft {

// The 1-D sub-transforms of one axis inside a dense row-major (batch, dims...) array:
// element (o, n, i) lives at (o*N + n)*inner + i, with outer = batch * prod(dims before
// the axis) and inner = prod(dims after it). inner == 1 is the contiguous-row case.
struct AxisView {
  int64_t outer_per_batch = 1;  // prod(dims before axis)
  int64_t n = 0;                // axis length
  int64_t inner = 1;            // prod(dims after axis) = element stride along the axis
};

// Where a pass reads from: the user's input (any dtype, real or complex) or the
// working complex buffer (out dtype).
struct IoSpec {
  int dtype = B200FFT_F32;  // scalar type
  int comps = 2;            // 1 real, 2 interleaved complex
};

// Peer scatter for the slab exchange fused into the last pass's store.
struct Scatter {
  void* const* peer_out = nullptr;  // HOST array of npeers device pointers (copied into the kernel's arguments)
  int npeers = 0;
  int my_rank = 0;
  long long zbase = -1;  // first z plane of this call inside the peers' slabs; -1 = my_rank * (planes of this call)
};

// which buffer a pass reads / writes
enum BufSel { BUF_INPUT = 0, BUF_OUTPUT = 1, BUF_WORK = 2 };

// Half-spectrum handling of the last axis (B200FFT_REAL_HALF)
enum HalfMode { HALF_NONE = 0, HALF_R2C = 1, HALF_C2R = 2 };

struct Pass {
  virtual ~Pass() {}
  // Run this pass for `nbatch` batch items. src/dst already point at the first item.
  virtual int launch(const void* src, void* dst, int64_t nbatch, cudaStream_t stream) = 0;
  virtual std::string describe() const = 0;
  virtual int launches() const { return 1; }
  // The same pass over `nunits` consecutive group units, a unit being 1 / units_per_batch of a batch item along the
  // outermost dimensions (L2-resident pass groups, api.cu: plan_pass_group). src/dst point at the first unit.
  virtual bool supports_units() const { return false; }
  virtual int launch_units(const void*, void*, int64_t, int64_t, cudaStream_t) {
    return fail(B200FFT_ERR_UNSUPPORTED, "this pass cannot run on a sub-range of a batch item (%s)", describe().c_str());
  }
  // Same pass, but output row i of the transformed axis goes to sc.peer_out[i / rows_per_peer]
  // (slab exchange fused into the store). Only strided fast passes implement it.
  virtual int launch_scatter(const void*, const Scatter&, int64_t, cudaStream_t) {
    return fail(B200FFT_ERR_UNSUPPORTED, "this pass has no scattering store (%s)", describe().c_str());
  }
  int axis = -1;
  BufSel src_sel = BUF_OUTPUT, dst_sel = BUF_OUTPUT;
  // Serpentine pass order: every second pass of a per-axis plan walks its tiles back to front, starting on the part
  // of the array the previous pass wrote last (still in L2) instead of the part it wrote first (long evicted).
  bool reverse_order = false;
  // A pass that stands for ALL axes (the fused N-d kernel) carries the per-axis passes of the same plan: they run
  // in its place when its launch is refused (cooperative launch cannot be satisfied on this context).
  std::vector<std::unique_ptr<Pass>> fallback;
};

struct DeviceTwiddles {
  void* ptr = nullptr;  // out-dtype complex W_N^n, n in [0, N)
  int64_t n = 0;
};

}  // namespace b200fft

struct b200fft_plan {
  b200fft::Problem prob;
  int device = 0;
  int sm_count = 148;
  std::vector<b200fft::DeviceTwiddles> tw;  // one per axis (null for skipped axes)
  std::vector<std::unique_ptr<b200fft::Pass>> passes;
  std::vector<void*> owned_device;          // misc device allocations freed at destroy
  bool building_fallback = false;           // build_passes is collecting the per-axis passes behind a fused pass
  // L2-resident pass group (api.cu: plan_pass_group): passes[0 .. group_passes) run chunk by chunk
  int group_passes = 0;
  int64_t group_mult = 1;         // group units per batch item = prod(dims outside the group)
  int64_t group_chunk_units = 0;
  size_t group_in_stride = 0, group_out_stride = 0;  // bytes per unit
  // pipelined group execution (api.cu: run_passes): the group's first pass runs on the caller's stream, the others on
  // this side stream one chunk behind, so the two kernels overlap and the chunk between them never leaves L2
  cudaStream_t group_stream = nullptr;
  cudaEvent_t group_ev[4] = {};  // fork, first-pass-done, tail[2]
  bool group_pipeline = false;
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
  size_t work_stride = 0;     // workspace bytes per batch item
  // exec_host resources (lazily created)
  cudaStream_t hs[3] = {nullptr, nullptr, nullptr};
  void* h_dev_in[2] = {nullptr, nullptr};
  void* h_dev_out[2] = {nullptr, nullptr};
  cudaEvent_t h_ev[8] = {};
  int64_t host_chunk = 0;
  bool host_ready = false;
};

namespace b200fft {

// kernel families (each returns nullptr when it does not cover the request)
// `half`: HALF_R2C = rows of n reals -> n/2+1 complex bins; HALF_C2R = the inverse (rows only).
std::unique_ptr<Pass> make_generic_pass(const b200fft_plan& plan, int axis, const AxisView& view,
                                        const IoSpec& src, bool scale_inverse, HalfMode half = HALF_NONE);
// compile-time kernels (fast_registry.cu); allocates its stage twiddle table into plan.owned_device
std::unique_ptr<Pass> make_fast_pass(b200fft_plan& plan, int axis, const AxisView& view, const IoSpec& src,
                                     bool scale_inverse, HalfMode half = HALF_NONE);

// long axes as two passes (split_registry.cu); nullptr when no (N1, N2) pair of kernels covers the length
std::unique_ptr<Pass> make_split_pass(b200fft_plan& plan, int axis, const AxisView& view, const IoSpec& src,
                                      bool scale_inverse, HalfMode half = HALF_NONE);
// runtime-length pass with compile-time codelets (rt.cu): any stage list with radices <= 32, f32 output
std::unique_ptr<Pass> make_rt_pass(b200fft_plan& plan, int axis, const AxisView& view, const IoSpec& src,
                                   bool scale_inverse, HalfMode half = HALF_NONE);
// fused N-d kernel (fused_registry.cu): one pass object that replaces ALL per-axis passes, or nullptr
std::unique_ptr<Pass> make_fused_pass(b200fft_plan& plan);
// half-spectrum inverse: the two innermost axes in one tile per plane (plane_registry.cu), or nullptr
std::unique_ptr<Pass> make_plane_c2r_pass(b200fft_plan& plan);
// forward half spectrum / complex / real input: the two innermost axes in one tile per plane, or nullptr
std::unique_ptr<Pass> make_plane_fwd_pass(b200fft_plan& plan);
// ... the same for plane sizes without a registered variant, specialised at plan time (jit.cu)
std::unique_ptr<Pass> make_jit_plane_pass(b200fft_plan& plan);
// plan-time specialisation (jit.cu): the compile-time kernels instantiated through NVRTC for unregistered lengths
std::unique_ptr<Pass> make_jit_pass(b200fft_plan& plan, int axis, const AxisView& view, const IoSpec& src, bool scale_inverse,
                                    HalfMode half);
int jit_probe(int64_t n, int64_t inner, const std::vector<uint32_t>& ordered, bool inverse, bool real_in, int half,
              int in_dtype, int out_dtype, std::string* report);
// number of registered kernel variants per tier (host-only; triggers the one-time registration)
size_t fast_variant_count();
size_t fused_variant_count();
size_t split_variant_count();

}  // namespace b200fft
