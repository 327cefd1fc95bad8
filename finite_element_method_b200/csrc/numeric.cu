// Numeric assembly: the hot kernel.
//
// Gather formulation — "owner computes": one thread owns one node-pair block (a, b) of the global
// matrix (6x6, or 3x3 for a truss-only pair). It walks the block's contribution list
// (family, local pair, element) in global insertion order, evaluates each element's global-frame
// block T^T k[la,lb] T from the element record (element_math.cuh) and sums into 36 FP64
// registers. Nothing is ever accumulated through memory: no atomics, no colouring, and the
// summation order is the order the reference's add_* calls would have used, so the result is
// deterministic and every CSR value is written exactly once.
//
// A CTA owns a "slab": a contiguous range of node rows, hence a contiguous range of CSR values.
// Threads drop their block into a shared-memory image of the slab (the 6 row segments of a block
// are 6 doubles wide and a row apart, which would be a poor global store pattern), then the whole
// CTA streams the image to HBM with fully coalesced 16-byte stores. Blocks inside a slab are
// pre-sorted by contribution count (symbolic.cu) so the threads of a warp run the same trip count.
//
// Algorithmic traffic per launch: 8 B x nnz written + element records read (L2-resident re-reads
// across the 4/16 blocks an element touches). See DESIGN.md for the byte model.
#include "common.cuh"
#include "element_math.cuh"

namespace femgpu {

namespace {

struct AsmArgs {
  const SlabDesc* slabs;
  const BlockMeta* meta;
  const uint32_t* cptr;
  const uint32_t* contrib;
  const double4* truss_rec;
  const double* beam_rec;
  const double* plate_rec;
  const double* plate_mat;
  double* values;
  uint32_t n_slabs;
};

__device__ __forceinline__ void load_rec16(const double* __restrict__ base, uint32_t e, double r[16]) {
  const double2* p = reinterpret_cast<const double2*>(base + size_t(e) * 16);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double2 v = __ldg(p + i);
    r[2 * i] = v.x;
    r[2 * i + 1] = v.y;
  }
}

__device__ __forceinline__ void add_contribution(const AsmArgs& A, uint32_t code, double acc[36]) {
  const uint32_t family = code >> 30, pair = (code >> 26) & 15u, e = code & 0x03FFFFFFu;
  if (family == FEMGPU_PLATE) {
    double rec[16], mat[4];
    load_rec16(A.plate_rec, e, rec);
    const double2* mp = reinterpret_cast<const double2*>(A.plate_mat + size_t(e) * 4);
    double2 m0 = __ldg(mp), m1 = __ldg(mp + 1);
    mat[0] = m0.x; mat[1] = m0.y; mat[2] = m1.x; mat[3] = m1.y;
    plate_block(rec, mat, int(pair >> 2), int(pair & 3u), acc);
  } else if (family == FEMGPU_BEAM) {
    double rec[16];
    load_rec16(A.beam_rec, e, rec);
    beam_block(rec, int(pair >> 1), int(pair & 1u), acc);
  } else {
    const double2* tp = reinterpret_cast<const double2*>(A.truss_rec + e);
    double2 t0 = __ldg(tp), t1 = __ldg(tp + 1);
    truss_block(t0.x, t0.y, t1.x, t1.y, int(pair >> 1), int(pair & 1u), acc);
  }
}

// place a 6x6 / 3x3 block into the slab image (shared or global)
__device__ __forceinline__ void store_block(double* __restrict__ img, const BlockMeta& m,
                                            const double acc[36]) {
  const uint32_t s03 = m.strides & 0xFFFFu, s35 = m.strides >> 16;
  const bool full = m.seg3 != 0xFFFFFFFFu;
  if (full) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double* r0 = img + m.seg0 + i * s03;
      double* r3 = img + m.seg3 + i * s35;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        r0[j] = acc[6 * i + j];
        r3[j] = acc[6 * (i + 3) + j];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double* r0 = img + m.seg0 + i * s03;
#pragma unroll
      for (int j = 0; j < 3; ++j) r0[j] = acc[6 * i + j];
    }
  }
}

__global__ void __launch_bounds__(kAsmThreads)
assemble_kernel(const AsmArgs A) {
  extern __shared__ __align__(16) double slab_img[];
  for (uint32_t k = blockIdx.x; k < A.n_slabs; k += gridDim.x) {
    const SlabDesc d = A.slabs[k];
    if (d.blk_count == 0) continue;
    const bool direct = d.flags & 1u;
    double* img = direct ? (A.values + d.val_base) : slab_img;
    for (uint32_t j = threadIdx.x; j < d.blk_count; j += kAsmThreads) {
      const uint32_t p = d.blk_begin + j;
      const BlockMeta m = A.meta[p];
      const uint32_t c1 = A.cptr[p + 1];
      double acc[36];
#pragma unroll
      for (int i = 0; i < 36; ++i) acc[i] = 0.0;
      for (uint32_t c = m.cptr; c < c1; ++c) add_contribution(A, __ldg(A.contrib + c), acc);
      store_block(img, m, acc);
    }
    if (direct) continue;
    __syncthreads();
    // stream the slab image out: 16-byte stores on the aligned body, scalars at the ragged ends
    double* out = A.values + d.val_base;
    const uint32_t n = d.val_count;
    const uint32_t head = uint32_t(d.val_base & 1);  // values[] is 16-byte aligned at index 0
    if (head && threadIdx.x == 0) out[0] = slab_img[0];
    const uint32_t body = (n - head) >> 1;
    if (head == 0) {
      const double2* src = reinterpret_cast<const double2*>(slab_img);
      double2* dst = reinterpret_cast<double2*>(out);
      for (uint32_t i = threadIdx.x; i < body; i += kAsmThreads) dst[i] = src[i];
    } else {
      double2* dst = reinterpret_cast<double2*>(out + 1);
      for (uint32_t i = threadIdx.x; i < body; i += kAsmThreads)
        dst[i] = make_double2(slab_img[1 + 2 * i], slab_img[2 + 2 * i]);
    }
    if (((n - head) & 1u) && threadIdx.x == 0) out[n - 1] = slab_img[n - 1];
    __syncthreads();
  }
}

// test hook: the whole transformed element matrix of one element, built from the same block
// evaluators the assembly uses
__global__ void element_matrix_kernel(int family, uint32_t e, const double4* truss_rec,
                                      const double* beam_rec, const double* plate_rec,
                                      const double* plate_mat, double* out) {
  const int nn = (family == FEMGPU_PLATE) ? 4 : 2;
  const int dof = (family == FEMGPU_TRUSS) ? 3 : 6;
  const int n = nn * dof;
  int pair = threadIdx.x;
  if (pair >= nn * nn) return;
  int la = pair / nn, lb = pair % nn;
  AsmArgs A{};
  A.truss_rec = truss_rec;
  A.beam_rec = beam_rec;
  A.plate_rec = plate_rec;
  A.plate_mat = plate_mat;
  double acc[36];
  for (int i = 0; i < 36; ++i) acc[i] = 0.0;
  add_contribution(A, (uint32_t(family) << 30) | (uint32_t(pair) << 26) | e, acc);
  for (int i = 0; i < dof; ++i)
    for (int j = 0; j < dof; ++j) out[(la * dof + i) * n + lb * dof + j] = acc[6 * i + j];
}

}  // namespace

int32_t run_assembly(Handle* h) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  if (h->n_slabs == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    FEMGPU_CUDA_CHECK(h, cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              kSlabSmemBytes));
    attr_set = true;
  }
  AsmArgs A;
  A.slabs = h->slabs.p;
  A.meta = h->blk_meta.p;
  A.cptr = h->blk_cptr.p;
  A.contrib = h->contrib.p;
  A.truss_rec = reinterpret_cast<const double4*>(h->fd[FEMGPU_TRUSS].rec.p);
  A.beam_rec = h->fd[FEMGPU_BEAM].rec.p;
  A.plate_rec = h->fd[FEMGPU_PLATE].rec.p;
  A.plate_mat = h->fd[FEMGPU_PLATE].mat.p;
  A.values = h->values.p;
  A.n_slabs = h->n_slabs;
  assemble_kernel<<<h->n_slabs, kAsmThreads, kSlabSmemBytes, h->stream>>>(A);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  return 0;
}

int32_t element_matrix(Handle* h, int family, size_t index, double* out_host) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  // make sure records exist for every element (cheap; this is a test hook)
  int32_t st = run_prep(h, /*validate_only=*/false);
  if (st) return st;
  const int nn = kNodesPerElem[family], dof = family == FEMGPU_TRUSS ? 3 : 6, n = nn * dof;
  FEMGPU_CUDA_CHECK(h, h->scratch.reserve(size_t(n) * n * 8 + 64));
  double* d_out = reinterpret_cast<double*>(h->scratch.p);
  element_matrix_kernel<<<1, 32, 0, h->stream>>>(
      family, uint32_t(index), reinterpret_cast<const double4*>(h->fd[FEMGPU_TRUSS].rec.p),
      h->fd[FEMGPU_BEAM].rec.p, h->fd[FEMGPU_PLATE].rec.p, h->fd[FEMGPU_PLATE].mat.p, d_out);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(out_host, d_out, size_t(n) * n * 8, cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

}  // namespace femgpu
