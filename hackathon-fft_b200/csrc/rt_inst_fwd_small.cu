// Forward, radices <= 16: one instantiation of the runtime-length kernel (rt_kernel.cuh) per translation unit
#define B200FFT_PACKED 1  // packed FADD2 complex adds (dft.cuh)
#include "rt_kernel.cuh"

namespace b200fft {
void rt_launch_fwd_small(const RtArgs& a, unsigned grid, size_t smem, cudaStream_t stream) {
  rt_axis_kernel<false, 16><<<grid, RT_THREADS, smem, stream>>>(a);
}
cudaError_t rt_prepare_fwd_small(int max_smem) {
  return cudaFuncSetAttribute(rt_axis_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}
}  // namespace b200fft
