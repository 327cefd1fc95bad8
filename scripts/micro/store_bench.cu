// Micro-benchmark behind DESIGN.md §4.2 "store path": how fast can persistent CTAs stream a matrix out of shared
// memory? Every CTA owns `depth` image buffers of `bytes` each and writes slabs k = cta, cta + grid, ... of a
// contiguous output (like assemble_kernel's slab images), by
//   mode 0  cp.async.bulk.global.shared::cta (TMA bulk store), at most `depth` stores in flight per CTA
//   mode 1  plain 16-byte st.global from the image (coalesced), all threads
//   mode 2  like 0, but the image is (re)written with st.shared + fence.proxy.async before every store
// usage: store_bench <mode> <bytes per slab> <threads per CTA> <CTAs per SM> <depth> <byte offset of the output: 0|16|...> [total GB]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int kMode>
__global__ void store_kernel(char* out, uint32_t bytes, uint32_t n_slabs, int depth) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  for (uint32_t i = tid; i < bytes * depth / 16; i += nt) reinterpret_cast<double2*>(smem)[i] = make_double2(double(i), 1.0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  uint32_t it = 0;
  for (uint32_t k = blockIdx.x; k < n_slabs; k += gridDim.x, ++it) {
    unsigned char* img = smem + size_t(it % depth) * bytes;
    char* dst = out + size_t(k) * bytes;
    if (kMode == 1) {
      const double2* s = reinterpret_cast<const double2*>(img);
      double2* d = reinterpret_cast<double2*>(dst);
      for (uint32_t i = tid; i < bytes / 16; i += nt) d[i] = s[i];
      continue;
    }
    // the buffer is free once the store issued `depth` iterations ago has read it
    if (tid == 0) {
      if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    }
    if (kMode == 2) {
      __syncthreads();
      for (uint32_t i = tid; i < bytes / 16; i += nt) reinterpret_cast<double2*>(img)[i] = make_double2(double(k), double(i));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
    if (tid == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(img)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
  if (argc < 7) { printf("usage: store_bench mode bytes threads ctas_per_sm depth offset [GB]\n"); return 1; }
  const int mode = atoi(argv[1]);
  const uint32_t bytes = uint32_t(atoi(argv[2]));
  const int threads = atoi(argv[3]), per_sm = atoi(argv[4]), depth = atoi(argv[5]), offset = atoi(argv[6]);
  const double gb = argc > 7 ? atof(argv[7]) : 10.4;
  const uint32_t n_slabs = uint32_t(gb * 1e9 / bytes);
  char* buf = nullptr;
  CK(cudaMalloc(&buf, size_t(n_slabs) * bytes + 4096));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t smem = size_t(bytes) * depth;
  void (*fn)(char*, uint32_t, uint32_t, int) = mode == 0 ? store_kernel<0> : mode == 1 ? store_kernel<1> : store_kernel<2>;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem));
  const int grid = sms * (per_sm < occ ? per_sm : occ);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(e0));
    fn<<<grid, threads, smem>>>(buf + offset, bytes, n_slabs, depth);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  printf("mode %d bytes %u threads %d CTAs/SM %d (occupancy %d) depth %d offset %d: %.3f ms  %.2f TB/s\n", mode, bytes, threads,
         per_sm, occ, depth, offset, best, double(n_slabs) * bytes / (best * 1e-3) / 1e12);
  return 0;
}
