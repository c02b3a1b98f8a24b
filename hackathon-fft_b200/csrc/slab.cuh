// Slab-decomposed 3-D transform, one rank's share, as ONE persistent kernel whose tiles overlap the
// NVLink exchange with the butterflies — on both sides of the exchange.
//
// Rank g holds z planes [g*zl, (g+1)*zl) of a (Z, Y, X) volume and ends with the Y-slab
// out_g[Z][yl][X]. cols_scatter_kernel (fast.cuh) already fuses the all-to-all into the Y pass's stores,
// but the three steps (X pass, Y pass + exchange, barrier, Z pass) still run one after the other, so the
// 117 MB per GPU that cross NVLink at 8 ranks are not hidden behind anything. Here the three passes are
// phases of one kernel (tile order, atomic work fetch and local dependency counters as in fused.cuh):
//   phase 0  X rows of the local planes                      in   -> work       (HBM-bound)
//   phase 1  Y columns of plane z, x-block b; output row y   work -> peer[y / yl] (NVLink-bound)
//            is stored straight into the owning rank's slab, then the tile bumps counter[b] ON EVERY RANK
//            (__threadfence_system + atomicAdd_system over NVLink)
//   phase 2  Z columns of x-block b of the local Y-slab, in place; waits until counter[b] shows that all
//            Z planes of that x-block have arrived from all ranks (ld.acquire.sys)
// Phase 1 walks the x-blocks in the same order on every rank and phase 2 follows a few blocks behind, so
// the Z pass of early blocks runs while later blocks are still crossing the switch: no barrier, no
// separate all-to-all, no pack buffer. Counters only grow (target = calls_on_this_buffer * Z); two receive
// slabs alternate per call so a fast peer's next call never lands in a slab that is still being read.
#pragma once
#include "fused.cuh"

namespace b200fft {

constexpr int SLAB_MAX_RANKS = 16;

struct SlabArgs {
  const float2* in;                  // [zl][Y][X] this rank's planes
  float2* work;                      // [zl][Y][X] scratch (X-transformed planes)
  float2* peer[SLAB_MAX_RANKS];      // receive slab [Z][yl][X] of every rank (peer[rank] is local)
  unsigned* peer_ctr[SLAB_MAX_RANKS];// x-block arrival counters of every rank
  const float2 *twx, *twy, *twz;
  int zl, yl, ranks, rank, nb;       // nb = X / CW x-blocks
  unsigned* err;                     // mapped HOST word: sticky "a wait timed out" flag read by the next slab_exec
  unsigned want;                     // counter value that means "x-block complete" for this call
  const NdSegment* segs;
  unsigned total_items;
  unsigned* ctrl;                    // [0] next item, [1] CTAs finished, [2 .. 2+zl) plane counters
  int nwords;
  float scale;
  int do_scale;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// output row i of the Y pass -> rank i / yl, at its place in that rank's [Z][yl][X] slab
struct SlabScatterDst {
  const SlabArgs* a;
  long long tile_off;  // ((rank*zl + z) * yl) * X + c0
  int X;
  __device__ __forceinline__ void store(int, int i, int c, float2 v) const {
    const int h = i / a->yl;
    const int r = i - h * a->yl;
    a->peer[h][tile_off + (long long)r * X + c] = v;
  }
};

// cubic-ish volumes with compile-time axis lengths; C rows per X tile, CW columns per Y / Z tile
template <int NZ, int NY, int NX, class RLZ, class RLY, class RLX, int C, int CW, int NT, bool INV>
__global__ void __launch_bounds__(NT, 2) slab_fused_kernel(const __grid_constant__ SlabArgs a) {
  extern __shared__ __align__(16) float2 smem_f2[];
  __shared__ unsigned s_item;
  __shared__ int s_last;
  __shared__ int s_ready;
  static_assert(NY % C == 0 && NX % CW == 0, "tiles must not straddle planes / x-blocks");
  constexpr int ROW_TILES_PER_PLANE = NY / C;
  // stage twiddles, staged once per CTA (the per-tile fences invalidate L1, so global tables would be
  // re-fetched from L2 for every tile)
  __shared__ float2 s_twx[RLX::tw_total() > 0 ? RLX::tw_total() : 1];
  __shared__ float2 s_twy[RLY::tw_total() > 0 ? RLY::tw_total() : 1];
  __shared__ float2 s_twz[RLZ::tw_total() > 0 ? RLZ::tw_total() : 1];
  for (int i = threadIdx.x; i < RLX::tw_total(); i += NT) s_twx[i] = a.twx[i];
  for (int i = threadIdx.x; i < RLY::tw_total(); i += NT) s_twy[i] = a.twy[i];
  for (int i = threadIdx.x; i < RLZ::tw_total(); i += NT) s_twz[i] = a.twz[i];
  __syncthreads();
  // A tile's completion signal (fence + counter update) is sent one tile LATE: by then its stores — remote
  // ones take microseconds over NVLink — have long drained, so the fence costs nothing and the next tile's
  // loads are never held up behind it. pend_kind: 0 none, 1 local plane counter, 2 x-block counter on every rank.
  int pend_kind = 0, pend_idx = 0;
  auto flush_signal = [&]() {
    if (pend_kind == 1) {
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.ctrl + 2 + pend_idx, 1u);
      }
    } else if (pend_kind == 2) {
      if ((int)threadIdx.x < a.ranks) {
        __threadfence_system();
        atomicAdd_system(a.peer_ctr[threadIdx.x] + pend_idx, 1u);
      }
    }
    pend_kind = 0;
  };
  int seg = 0;
  unsigned next = 0;
  if (threadIdx.x == 0) next = atomicAdd(a.ctrl, 1u);
  while (true) {
    if (threadIdx.x == 0) s_item = next;
    __syncthreads();  // also: every thread's stores of the previous tile are issued (CTA scope)
    const unsigned item = s_item;
    if (item >= a.total_items) break;
    if (threadIdx.x == 0) next = atomicAdd(a.ctrl, 1u);
    while ((long long)item >= a.segs[seg].first_item + a.segs[seg].count) ++seg;
    const int phase = a.segs[seg].phase;
    const int tile = (int)(a.segs[seg].first_tile + ((long long)item - a.segs[seg].first_item));
    static_assert(RLX::count == 2 && RLY::count == 2 && RLZ::count == 2, "slab tiles: two super-stages per axis");
    if (phase == 0) {
      // ---- X rows: C consecutive rows of the local [zl*NY][NX] row array
      using L0 = typename RowLayoutN<NX>::template type<RLX::r[0], 1>;
      const long long row0 = (long long)tile * C;
      GlobalSrc<false, true> s{a.in + row0 * NX, NX, 1, C, 1};
      GlobalDst d{a.work + row0 * NX, NX, 1, C, 1};
      run_stage<RLX::r[0], 1, NX, C, 1, NT, INV, true>(s, SmemDst<L0>{smem_f2}, s_twx, 1.f, false);
      __syncthreads();
      flush_signal();  // the PREVIOUS tile's signal: its stores drained while this tile's loads were in flight
      run_stage<RLX::r[1], RLX::r[0], NX, C, 1, NT, INV, true>(SmemSrc<L0>{smem_f2}, d, s_twx + RLX::tw_offset(1), 1.f, false);
      pend_kind = 1;
      pend_idx = tile / ROW_TILES_PER_PLANE;
    } else if (phase == 1) {
      // ---- Y columns of plane z, x-block b, scattered to the owning ranks
      using LY = DenseLayout<NY, CW>;
      const int b = tile / a.zl, z = tile - b * a.zl;
      if (threadIdx.x == 0) s_ready = ld_acquire_gpu(a.ctrl + 2 + z) >= (unsigned)ROW_TILES_PER_PLANE;
      __syncthreads();
      if (!s_ready) {
        flush_signal();  // never spin while holding a signal somebody may be waiting for
        if (threadIdx.x == 0) wait_counter_gpu(a.ctrl + 2 + z, (unsigned)ROW_TILES_PER_PLANE, 64, a.err);
        __syncthreads();
      }
      const long long base = (long long)z * NY * NX + (long long)b * CW;
      GlobalSrc<false, true> s{a.work + base, 0, NX, 1, CW};
      SlabScatterDst d{&a, ((long long)(a.rank * a.zl + z) * a.yl) * NX + (long long)b * CW, NX};
      run_stage<RLY::r[0], 1, NY, 1, CW, NT, INV, true>(s, SmemDst<LY>{smem_f2}, s_twy, 1.f, false);
      __syncthreads();
      flush_signal();
      run_stage<RLY::r[1], RLY::r[0], NY, 1, CW, NT, INV, true>(SmemSrc<LY>{smem_f2}, d, s_twy + RLY::tw_offset(1), 1.f, false);
      pend_kind = 2;
      pend_idx = b;
    } else {
      // ---- Z columns of x-block b, row y_local of the local Y-slab, in place
      using LZ = DenseLayout<NZ, CW>;
      const int b = tile / a.yl, yloc = tile - b * a.yl;
      const unsigned* cnt = a.peer_ctr[a.rank] + b;
      if (threadIdx.x == 0) s_ready = (int)(ld_acquire_sys(cnt) - a.want) >= 0;
      __syncthreads();
      if (!s_ready) {
        flush_signal();
        if (threadIdx.x == 0) {
          // bounded (~2 s): a peer that never arrives (crashed rank, mismatched call sequence) must not hang
          // the GPU; the time-out is recorded in the word after the counters and the results are then invalid
          unsigned spins = 0;
          while ((int)(ld_acquire_sys(cnt) - a.want) < 0) {
            __nanosleep(100);
            if (++spins > 20000000u) {
              atomicAdd(a.peer_ctr[a.rank] + a.nb, 1u);
              if (a.err) atomicExch_system(a.err, 1u);
              break;
            }
          }
        }
        __syncthreads();
      }
      const long long inner = (long long)a.yl * NX;
      float2* p = a.peer[a.rank] + (long long)yloc * NX + (long long)b * CW;
      GlobalSrc<false, true> s{p, 0, inner, 1, CW};
      GlobalDst d{p, 0, inner, 1, CW};
      run_stage<RLZ::r[0], 1, NZ, 1, CW, NT, INV, true>(s, SmemDst<LZ>{smem_f2}, s_twz, 1.f, false);
      __syncthreads();
      flush_signal();
      run_stage<RLZ::r[1], RLZ::r[0], NZ, 1, CW, NT, INV, true>(SmemSrc<LZ>{smem_f2}, d, s_twz + RLZ::tw_offset(1), a.scale,
                                                                a.do_scale != 0);
    }
  }
  flush_signal();  // the last tile's signal (the loop's final barrier covered its stores)
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(a.ctrl + 1, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    for (int i = threadIdx.x; i < a.nwords; i += NT) a.ctrl[i] = 0u;
  }
}

template <int NZ, int NY, int NX, class RLZ, class RLY, class RLX, int C, int CW>
constexpr size_t slab_fused_smem() {
  size_t m = rows_smem_bytes<NX, RLX, C>();
  const size_t y = cols_smem_bytes<NY, RLY, CW>(), z = cols_smem_bytes<NZ, RLZ, CW>();
  m = y > m ? y : m;
  m = z > m ? z : m;
  return m;
}

}  // namespace b200fft
