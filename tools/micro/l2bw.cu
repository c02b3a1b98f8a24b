// Microbenchmark: in-place read-modify-write streaming over a buffer of S bytes, repeated, to measure
// the L2-resident bandwidth (S << 126 MB) against the HBM-resident one (S >> 126 MB) on B200.
// Also measures: a dependent 2-kernel chain over chunks (launch-overhead bound) for comparison.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/build/l2bw tools/micro/l2bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void rmw(float4* __restrict__ p, size_t n4, float a) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v0 = __ldcg(p + i), v1 = __ldcg(p + i + stride), v2 = __ldcg(p + i + 2 * stride), v3 = __ldcg(p + i + 3 * stride);
    v0.x += a; v1.x += a; v2.x += a; v3.x += a;
    p[i] = v0; p[i + stride] = v1; p[i + 2 * stride] = v2; p[i + 3 * stride] = v3;
  }
  for (; i < n4; i += stride) { float4 v = __ldcg(p + i); v.x += a; p[i] = v; }
}
__global__ void rd(const float4* __restrict__ p, size_t n4, float* sink) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v0 = __ldcg(p + i), v1 = __ldcg(p + i + stride), v2 = __ldcg(p + i + 2 * stride), v3 = __ldcg(p + i + 3 * stride);
    acc += v0.x + v1.y + v2.z + v3.w;
  }
  for (; i < n4; i += stride) acc += __ldcg(p + i).x;
  if (acc == 123.456f) *sink = acc;
}

int main() {
  const size_t maxb = (size_t)2 << 30;
  float4* buf; float* sink;
  cudaMalloc(&buf, maxb); cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, maxb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t sizes_mb[] = {4, 8, 16, 32, 48, 64, 96, 128, 256, 1024, 2048};
  for (int grid_mult : {4, 8}) {
    for (size_t mb : sizes_mb) {
      const size_t bytes = mb << 20, n4 = bytes / 16;
      const int reps = (int)((8192 + mb - 1) / mb) < 4 ? 4 : (int)((8192 + mb - 1) / mb);
      for (int mode = 0; mode < 2; ++mode) {
        for (int w = 0; w < 3; ++w) { if (mode) rd<<<148 * grid_mult, 256>>>(buf, n4, sink); else rmw<<<148 * grid_mult, 256>>>(buf, n4, 1.f); }
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) { if (mode) rd<<<148 * grid_mult, 256>>>(buf, n4, sink); else rmw<<<148 * grid_mult, 256>>>(buf, n4, 1.f); }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double per = ms / reps;
        printf("{\"grid\": %d, \"mb\": %zu, \"mode\": \"%s\", \"us\": %.2f, \"gbs\": %.1f}\n", 148 * grid_mult, mb,
               mode ? "read" : "rmw", per * 1e3, (mode ? 1.0 : 2.0) * bytes / per / 1e6);
      }
    }
  }
  // launch-overhead probe: 100 back-to-back tiny launches
  for (int w = 0; w < 2; ++w) {
    cudaEventRecord(e0);
    for (int r = 0; r < 100; ++r) rd<<<148, 256>>>(buf, 1024, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"tiny_launch_us\": %.2f}\n", ms * 10);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"status\": \"%s\"}\n", cudaGetErrorString(e));
  return 0;
}
