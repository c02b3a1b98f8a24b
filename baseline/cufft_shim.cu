// cuFFT comparison baseline for bench.py: same plan call the reference's
// cufft-benchmark-main/cufft_benchmark.cu makes (cufftMakePlanMany64, :70-78;
// cufftExecC2C / cufftExecR2C, :86-98), but on caller-provided device buffers and a
// caller-provided stream so bench.py can time it with CUDA events next to our kernels.
// Library calls only; not part of the product.
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

extern "C" {

__attribute__((visibility("default"))) int cufft_shim_create(void** handle, int rank, const long long* dims,
                                                           long long batch, int r2c, size_t* work_size) {
  cufftHandle* h = new cufftHandle;
  if (cufftCreate(h) != CUFFT_SUCCESS) return 1;
  long long n[8];
  long long in_dist = 1, out_dist = 1;
  for (int i = 0; i < rank; ++i) { n[i] = dims[i]; in_dist *= dims[i]; }
  out_dist = r2c ? in_dist / dims[rank - 1] * (dims[rank - 1] / 2 + 1) : in_dist;
  size_t ws = 0;
  cufftResult r = cufftMakePlanMany64(*h, rank, n, NULL, 1, in_dist, NULL, 1, out_dist, r2c ? CUFFT_R2C : CUFFT_C2C,
                                      batch, &ws);
  if (r != CUFFT_SUCCESS) return 100 + (int)r;
  if (work_size) *work_size = ws;
  *handle = h;
  return 0;
}

__attribute__((visibility("default"))) int cufft_shim_exec(void* handle, const void* d_in, void* d_out, void* stream,
                                                         int r2c, int inverse) {
  cufftHandle h = *(cufftHandle*)handle;
  if (cufftSetStream(h, (cudaStream_t)stream) != CUFFT_SUCCESS) return 1;
  cufftResult r = r2c ? cufftExecR2C(h, (cufftReal*)d_in, (cufftComplex*)d_out)
                      : cufftExecC2C(h, (cufftComplex*)d_in, (cufftComplex*)d_out, inverse ? CUFFT_INVERSE : CUFFT_FORWARD);
  return r == CUFFT_SUCCESS ? 0 : 100 + (int)r;
}

__attribute__((visibility("default"))) int cufft_shim_destroy(void* handle) {
  cufftHandle* h = (cufftHandle*)handle;
  cufftDestroy(*h);
  delete h;
  return 0;
}

__attribute__((visibility("default"))) int cufft_shim_version(void) {
  int v = 0;
  cufftGetVersion(&v);
  return v;
}
}
