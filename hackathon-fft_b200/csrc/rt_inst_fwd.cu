// Forward instantiation of the runtime-length kernel (rt_kernel.cuh)
#define B200FFT_PACKED 1  // packed FADD2 complex adds (dft.cuh)
#include "rt_kernel.cuh"

namespace b200fft {
void rt_launch_fwd(const RtArgs& a, unsigned grid, size_t smem, cudaStream_t stream) {
  rt_axis_kernel<false><<<grid, RT_THREADS, smem, stream>>>(a);
}
cudaError_t rt_prepare_fwd(int max_smem) {
  return cudaFuncSetAttribute(rt_axis_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}
}  // namespace b200fft
