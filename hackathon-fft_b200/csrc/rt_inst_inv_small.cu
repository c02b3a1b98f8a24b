// Inverse, radices <= 16: one instantiation of the runtime-length kernel (rt_kernel.cuh) per translation unit
#define B200FFT_PACKED 1  // packed FADD2 complex adds (dft.cuh)
#include "rt_kernel.cuh"

namespace b200fft {
void rt_launch_inv_small(const RtArgs& a, unsigned grid, size_t smem, cudaStream_t stream) {
  rt_axis_kernel<true, 16><<<grid, RT_THREADS, smem, stream>>>(a);
}
cudaError_t rt_prepare_inv_small(int max_smem) {
  return cudaFuncSetAttribute(rt_axis_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}
}  // namespace b200fft
