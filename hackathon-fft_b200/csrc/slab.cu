// Host side of the fused slab kernel (slab.cuh): plan object, tile schedule, C ABI.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "fast_registry.hpp"
#include "plan.hpp"
#include "slab.cuh"

using namespace b200fft;

struct b200fft_slab {
  int device = 0, ranks = 1, rank = 0, inverse = 0;
  int n = 0;  // cubic volume n^3
  int zl = 0, yl = 0, nb = 0, c = 0, cw = 0, threads = 0;
  size_t smem = 0, slab_bytes = 0, recv_bytes = 0;
  float2 *twx = nullptr, *twy = nullptr, *twz = nullptr;
  NdSegment* d_segs = nullptr;
  unsigned* d_ctrl = nullptr;
  int nwords = 0;
  unsigned total_items = 0;
  int grid = 148;
  unsigned epoch[2] = {0, 0};
  unsigned* h_err = nullptr;  // mapped host word: raised by the kernel when a wait for a peer (or a local plane) gave up
  unsigned* d_err = nullptr;
  bool poisoned = false;      // sticky: a time-out left the epoch counters of the ranks out of step
  void (*launch)(const SlabArgs&, unsigned, size_t, cudaStream_t) = nullptr;
  const void* func = nullptr;
  std::vector<int> radices;
  std::string text;
};

namespace {

struct DeviceGuard {  // run on the plan's device, give the caller its own device back (like api.cu)
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
};

template <int N, class RL, int C, int CW, int NT, bool INV>
void launch_slab(const SlabArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
  slab_fused_kernel<N, N, N, RL, RL, RL, C, CW, NT, INV><<<grid, NT, smem, st>>>(a);
}

template <int N, class RL, int C, int CW, int NT>
bool bind_variant(b200fft_slab* s) {
  if (s->n != N) return false;
  s->c = C;
  s->cw = CW;
  s->threads = NT;
  s->smem = slab_fused_smem<N, N, N, RL, RL, RL, C, CW>();
  s->radices = radix_vec<RL>();
  if (s->inverse) {
    s->launch = &launch_slab<N, RL, C, CW, NT, true>;
    s->func = (const void*)slab_fused_kernel<N, N, N, RL, RL, RL, C, CW, NT, true>;
  } else {
    s->launch = &launch_slab<N, RL, C, CW, NT, false>;
    s->func = (const void*)slab_fused_kernel<N, N, N, RL, RL, RL, C, CW, NT, false>;
  }
  return true;
}

int upload(const std::vector<float2>& t, float2** out) {
  B200_CUDA_CHECK(cudaMalloc(out, t.size() * sizeof(float2)));
  B200_CUDA_CHECK(cudaMemcpy(*out, t.data(), t.size() * sizeof(float2), cudaMemcpyHostToDevice));
  return B200FFT_OK;
}

}  // namespace

extern "C" {

int b200fft_slab_destroy(b200fft_slab* s) {
  if (!s) return B200FFT_OK;
  {
    DeviceGuard guard(s->device);
    for (void* p : {(void*)s->twx, (void*)s->twy, (void*)s->twz, (void*)s->d_segs, (void*)s->d_ctrl})
      if (p) cudaFree(p);
    if (s->h_err) cudaFreeHost(s->h_err);
  }
  delete s;
  return B200FFT_OK;
}

int b200fft_slab_create(b200fft_slab** out, int64_t n, int ranks, int rank, int inverse, int device) {
  if (!out) return fail(B200FFT_ERR_INVALID_ARG, "null slab pointer");
  *out = nullptr;
  if (ranks < 1 || ranks > SLAB_MAX_RANKS || rank < 0 || rank >= ranks)
    return fail(B200FFT_ERR_INVALID_ARG, "rank %d of %d (at most %d ranks)", rank, ranks, SLAB_MAX_RANKS);
  if (n <= 0 || n % ranks) return fail(B200FFT_ERR_INVALID_ARG, "slab decomposition needs n divisible by the number of ranks");
  std::unique_ptr<b200fft_slab> s(new b200fft_slab());
  if (device < 0) B200_CUDA_CHECK(cudaGetDevice(&device));
  cudaDeviceProp prop;
  B200_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(B200FFT_ERR_CUDA, "device %d is sm_%d%d; this library only carries sm_100a code (no fallback)", device,
                prop.major, prop.minor);
  DeviceGuard guard(device);  // the caller's current device is restored on every return path
  s->device = device;
  s->ranks = ranks;
  s->rank = rank;
  s->inverse = inverse != 0;
  s->n = (int)n;
  const bool ok = bind_variant<64, Radices<8, 8>, 32, 16, 256>(s.get()) || bind_variant<128, Radices<16, 8>, 32, 16, 256>(s.get()) ||
                  bind_variant<256, Radices<16, 16>, 16, 16, 256>(s.get()) ||
                  bind_variant<512, Radices<32, 16>, 8, 16, 256>(s.get());
  if (!ok) return fail(B200FFT_ERR_UNSUPPORTED, "fused slab kernel: cubic volumes of 64, 128, 256 or 512 only (got %lld)", (long long)n);
  s->zl = s->yl = (int)(n / ranks);
  s->nb = (int)(n / s->cw);
  s->slab_bytes = (size_t)n * s->yl * n * sizeof(float2);
  // [slab][nb arrival counters][1 time-out counter], rounded up to 256 B
  s->recv_bytes = s->slab_bytes + (((size_t)(s->nb + 1) * sizeof(unsigned) + 255) / 256) * 256;

  int rc = upload(build_twiddles(s->radices, s->inverse), &s->twx);
  if (!rc) rc = upload(build_twiddles(s->radices, s->inverse), &s->twy);
  if (!rc) rc = upload(build_twiddles(s->radices, s->inverse), &s->twz);
  if (rc) { b200fft_slab_destroy(s.release()); return rc; }

  // tile order: all X rows, then per x-block the Y tiles, with the Z tiles following `delay` blocks behind
  int delay = 8;  // measured on 4 GPUs: 2 -> 0.550, 8 -> 0.488, 12 -> 0.498 ms (profiles/r1_slab.md)
  if (const char* e = getenv("B200FFT_SLAB_DELAY")) delay = std::max(0, atoi(e));
  delay = std::min(delay, s->nb);
  std::vector<NdSegment> segs;
  long long item = 0;
  auto add = [&](int phase, long long first_tile, long long count) {
    segs.push_back(NdSegment{phase, 0, item, first_tile, count});
    item += count;
  };
  add(0, 0, (long long)s->zl * n / s->c);
  for (int b = 0; b < s->nb; ++b) {
    add(1, (long long)b * s->zl, s->zl);
    if (b >= delay) add(2, (long long)(b - delay) * s->yl, s->yl);
  }
  for (int b = s->nb - delay; b < s->nb; ++b) add(2, (long long)b * s->yl, s->yl);
  s->total_items = (unsigned)item;
  s->nwords = 2 + s->zl;
  if (cudaMalloc(&s->d_segs, sizeof(NdSegment) * segs.size()) != cudaSuccess ||
      cudaMemcpy(s->d_segs, segs.data(), sizeof(NdSegment) * segs.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMalloc(&s->d_ctrl, sizeof(unsigned) * (size_t)s->nwords) != cudaSuccess ||
      cudaMemset(s->d_ctrl, 0, sizeof(unsigned) * (size_t)s->nwords) != cudaSuccess) {
    cudaGetLastError();
    b200fft_slab_destroy(s.release());
    return fail(B200FFT_ERR_ALLOC, "cannot allocate the slab schedule");
  }
  if (s->smem > 48 * 1024 &&
      cudaFuncSetAttribute(s->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem) != cudaSuccess) {
    cudaGetLastError();
    b200fft_slab_destroy(s.release());
    return fail(B200FFT_ERR_CUDA, "cannot reserve %zu B of shared memory", s->smem);
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s->func, s->threads, s->smem) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    b200fft_slab_destroy(s.release());
    return fail(B200FFT_ERR_CUDA, "fused slab kernel does not fit an SM");
  }
  s->grid = occ * prop.multiProcessorCount;
  if (cudaHostAlloc(&s->h_err, sizeof(unsigned), cudaHostAllocMapped) == cudaSuccess &&
      cudaHostGetDevicePointer(&s->d_err, s->h_err, 0) == cudaSuccess) {
    *s->h_err = 0;
  } else {
    cudaGetLastError();
    if (s->h_err) cudaFreeHost(s->h_err);
    s->h_err = s->d_err = nullptr;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) {  // tables, schedule and zeroed counters are in place before any exec
    cudaGetLastError();
    b200fft_slab_destroy(s.release());
    return fail(B200FFT_ERR_CUDA, "device synchronisation failed while creating the slab plan");
  }
  char buf[256];
  snprintf(buf, sizeof buf, "slab_fused %d^3 rank %d/%d (%s per axis): rows c%d -> cols w%d + scatter -> cols w%d; grid %d x %d, smem=%zuB, z delay=%d blocks",
           s->n, rank, ranks, radix_name(s->radices).c_str(), s->c, s->cw, s->cw, s->grid, s->threads, s->smem, delay);
  s->text = buf;
  *out = s.release();
  return B200FFT_OK;
}

size_t b200fft_slab_recv_bytes(const b200fft_slab* s) { return s ? s->recv_bytes : 0; }

size_t b200fft_slab_timeout_offset(const b200fft_slab* s) { return s ? s->slab_bytes + (size_t)s->nb * sizeof(unsigned) : 0; }

size_t b200fft_slab_describe(const b200fft_slab* s, char* buf, size_t cap) {
  if (!s) return 0;
  if (buf && cap) {
    strncpy(buf, s->text.c_str(), cap - 1);
    buf[cap - 1] = 0;
  }
  return s->text.size() + 1;
}

int b200fft_slab_exec(b200fft_slab* s, const void* d_in, void* d_work, void* const* peer_recv, int buffer, void* cu_stream) {
  if (!s || !d_in || !d_work || !peer_recv) return fail(B200FFT_ERR_INVALID_ARG, "null slab plan or buffer");
  if (buffer < 0 || buffer > 1) return fail(B200FFT_ERR_INVALID_ARG, "receive buffer index must be 0 or 1");
  if (s->poisoned || (s->h_err && *reinterpret_cast<volatile unsigned*>(s->h_err))) {
    // Sticky: after a time-out the arrival counters of the ranks no longer agree on the epoch, so every later call
    // would be wrong too. The caller must destroy the plans, re-zero the receive buffers on every rank and start over.
    s->poisoned = true;
    return fail(B200FFT_ERR_CUDA, "slab plan is poisoned: an earlier call timed out waiting for a peer (a rank died or skipped a "
                                  "call); its result was invalid. Re-zero the receive buffers on every rank and recreate the plans");
  }
  DeviceGuard guard(s->device);
  SlabArgs a;
  memset(&a, 0, sizeof a);
  a.in = reinterpret_cast<const float2*>(d_in);
  a.work = reinterpret_cast<float2*>(d_work);
  for (int h = 0; h < s->ranks; ++h) {
    if (!peer_recv[h]) return fail(B200FFT_ERR_INVALID_ARG, "null receive slab for rank %d", h);
    a.peer[h] = reinterpret_cast<float2*>(peer_recv[h]);
    a.peer_ctr[h] = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(peer_recv[h]) + s->slab_bytes);
  }
  a.twx = s->twx;
  a.twy = s->twy;
  a.twz = s->twz;
  a.zl = s->zl;
  a.yl = s->yl;
  a.ranks = s->ranks;
  a.rank = s->rank;
  a.nb = s->nb;
  a.err = s->d_err;
  a.want = ++s->epoch[buffer] * (unsigned)s->n;  // every rank adds zl per x-block: ranks * zl = n per call
  a.segs = s->d_segs;
  a.total_items = s->total_items;
  a.ctrl = s->d_ctrl;
  a.nwords = s->nwords;
  a.scale = s->inverse ? (float)(1.0 / ((double)s->n * s->n * s->n)) : 1.f;
  a.do_scale = s->inverse;
  const unsigned grid = (unsigned)std::min<long long>(s->total_items, s->grid);
  s->launch(a, grid, s->smem, (cudaStream_t)cu_stream);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(B200FFT_ERR_CUDA, "slab kernel launch failed: %s", cudaGetErrorString(e));
  return B200FFT_OK;
}

}  // extern "C"
