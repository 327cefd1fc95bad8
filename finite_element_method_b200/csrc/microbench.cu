// FP64 pipe peak of the device, measured ("FP64 peak is not in MEASURED_PEAKS.json -> measure it with
// an FMA micro-benchmark; report FP64 pipe utilisation next to GB/s"). Eight independent DFMA chains per thread,
// 8 x 256-thread CTAs per SM, no memory traffic inside the loop.
#include "common.cuh"

namespace femgpu {

namespace {

__global__ void __launch_bounds__(256)
fp64_fma_kernel(double* __restrict__ out, int iters, double b, double c) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = double(threadIdx.x + k) * 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 1234.5678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the chains alive
}

}  // namespace

}  // namespace femgpu

using namespace femgpu;

extern "C" int32_t femgpu_fp64_fma_peak(femgpu_t* h, double* tflops) {
  if (!h || !tflops) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return h->fail(FEMGPU_ERR_NO_DEVICE, "staging-only handle; femgpu has no CPU fallback");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  if (h->sm_count == 0)
    FEMGPU_CUDA_CHECK(h, cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
  const int iters = 8192, threads = 256, grid = h->sm_count * 8;
  FEMGPU_CUDA_CHECK(h, h->scratch.reserve(size_t(grid) * threads * 8));
  cudaEvent_t e0, e1;
  FEMGPU_CUDA_CHECK(h, cudaEventCreate(&e0));
  FEMGPU_CUDA_CHECK(h, cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {  // the first one warms up
    cudaEventRecord(e0, h->stream);
    fp64_fma_kernel<<<grid, threads, 0, h->stream>>>(reinterpret_cast<double*>(h->scratch.p), iters, 1.0000001, 1e-9);
    cudaEventRecord(e1, h->stream);
    h->launches++;
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * double(iters) * double(grid) * threads / (double(ms) * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  *tflops = best;
  return 0;
}
