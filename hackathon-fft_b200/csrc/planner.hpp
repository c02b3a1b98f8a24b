// Host planner: the reference's compile-time plan rules restated as runtime code.
//   base ordering / validity   fft/fft/_utils.mojo:125-221
//   default bases              fft/fft/fft.mojo:49-104
//   layout conditions          fft/fft/fft.mojo:20-46
// Pure C++ (no CUDA) so it is testable without a GPU.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "common.hpp"

namespace b200fft {

uint64_t times_divisible(uint64_t length, uint64_t base);
std::vector<uint32_t> build_ordered_bases(uint64_t length, std::vector<uint32_t> bases);
bool ordered_bases_valid(uint64_t length, const std::vector<uint32_t>& ordered);
std::vector<uint32_t> estimate_best_bases(uint64_t length, bool gpu_target);

// One transformed (or skipped) axis of a validated descriptor.
struct AxisSpec {
  int64_t n = 0;                  // logical transform length
  bool transformed = true;        // axis_mask bit
  std::vector<uint32_t> user;     // user bases (after defaults)
  std::vector<uint32_t> ordered;  // canonical descending stage list
};

// A validated descriptor: everything the kernel chooser needs.
struct Problem {
  b200fft_desc desc;            // copy (bases pointers cleared)
  int rank = 0;
  int64_t batch = 0;
  std::vector<AxisSpec> axes;   // size rank
  bool half = false;            // B200FFT_REAL_HALF
  // element counts per batch item (scalars of in/out dtype)
  int64_t in_scalars_per_batch = 0, out_scalars_per_batch = 0;
  size_t in_elem = 0, out_elem = 0;  // bytes per scalar
};

// ---- tile schedule of the fused N-d kernel (fused.cuh) ---------------------------------------------
// A run of consecutive work items that are consecutive tiles of one phase.
struct NdSegment {
  int phase;
  int pad;
  long long first_item;  // position of the first item in the global order
  long long first_tile;  // its global tile index: transform * tiles_per_transform + tile
  long long count;
};
struct SchedPhase {
  long long tiles_per_transform = 0;
  long long tiles_per_group = 0;  // tiles of this phase per dependency group it completes (unused for the last phase)
  long long dep_div = 1;          // tile j of this phase needs group j / dep_div of the previous phase (unused for phase 0)
  long long quota = 0;            // max tiles handed out per pipeline round (0 = everything that is ready)
};
// Orders all tiles of `batch` transforms as a software pipeline: round r hands out, for every phase in
// turn, the tiles whose dependency groups were completely handed out in EARLIER rounds (so a consumer
// never precedes, and in steady state never closely follows, its producers). Returns the segments in
// order; every tile of every phase appears exactly once.
std::vector<NdSegment> build_schedule(int nphases, const SchedPhase* phases, long long batch);

// Mirrors _check_layout_conditions_nd + the bases asserts. Returns a status code
// and leaves the detail in last_error().
int validate(const b200fft_desc* d, Problem* out);

}  // namespace b200fft
