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
  const uint32_t* contrib;
  const WorkItem* items;
  const uint32_t* elist;          // dense [n_slabs][kElistStride]
  const uint32_t* elist_compact;  // compact lists (unstaged fallback only)
  const double4* truss_rec;
  const double* beam_rec;
  const double* plate_rec;
  const double* plate_mat;
  double* values;
  uint32_t n_slabs;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}

// global address of 16-byte chunk `chunk` of an element's record (nullptr past its end):
// plate = 8 chunks of the geometry record + 2 of the material record, beam = 8, truss = 2
__device__ __forceinline__ const void* record_chunk(const AsmArgs& A, uint32_t fe, uint32_t chunk) {
  const uint32_t family = fe >> 26, e = fe & 0x03FFFFFFu;
  if (family == FEMGPU_PLATE) {
    if (chunk < 8) return A.plate_rec + size_t(e) * 16 + chunk * 2;
    return A.plate_mat + size_t(e) * 4 + (chunk - 8) * 2;
  }
  if (family == FEMGPU_BEAM) return chunk < 8 ? A.beam_rec + size_t(e) * 16 + chunk * 2 : nullptr;
  if (family != FEMGPU_TRUSS) return nullptr;  // family 3: placeholder of a remote contribution
  return chunk < 2 ? reinterpret_cast<const double*>(A.truss_rec + e) + chunk * 2 : nullptr;
}

// evaluate one contribution. rec = the element's staged record in shared memory: the prep kernel's
// record for trusses (4 doubles) and beams (16), the per-CTA shared form (element_math.cuh,
// plate_shared_record) for plates
__device__ __forceinline__ void add_contribution(const double* __restrict__ rec, uint32_t code,
                                                 double acc[36]) {
  const uint32_t family = code >> 30, pair = (code >> 26) & 15u;
  if (family == FEMGPU_PLATE) {
    plate_block_shared(rec, int(pair >> 2), int(pair & 3u), acc);
  } else if (family == FEMGPU_BEAM) {
    beam_block(rec, int(pair >> 1), int(pair & 1u), acc);
  } else if (family == FEMGPU_TRUSS) {
    truss_block(rec[0], rec[1], rec[2], rec[3], int(pair >> 1), int(pair & 1u), acc);
  }
  // family 3: slot reserved for another rank's contribution (multi-GPU) — nothing to add here
}

// same from the prep kernel's raw record (unstaged fallback, test hook): plates build their shared
// form on the spot
__device__ __forceinline__ void add_contribution_raw(const double* __restrict__ raw, uint32_t code,
                                                     double acc[36]) {
  if ((code >> 30) == FEMGPU_PLATE) {
    double S[kPlateSharedDoubles];
    plate_shared_record(raw, S);
    add_contribution(S, code, acc);
  } else {
    add_contribution(raw, code, acc);
  }
}

// Place a 6x6 / 3x3 block into the slab image. kShared: image in shared memory (STS) else straight
// into the CSR values (global). When every row segment is 16-byte aligned (true whenever the node
// only has 6-wide blocks) the six doubles of a row go out as three 16-byte stores: at the 48-byte
// lane stride of neighbouring blocks those are bank-conflict free, 8-byte stores are 4-way
// conflicted.
template <bool kShared>
__device__ __forceinline__ void store_block(double* __restrict__ img, const uint4 m,
                                            const double acc[36], bool base_even) {
  const uint32_t seg0 = m.x, seg3 = m.y, s03 = m.z & 0xFFFFu, s35 = m.z >> 16;
  const bool full = seg3 != 0xFFFFFFFFu;
  if (full) {
    const bool al = base_even && (((seg0 | s03 | seg3 | s35) & 1u) == 0);
    if (al) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double2* r0 = reinterpret_cast<double2*>(img + seg0 + i * s03);
        double2* r3 = reinterpret_cast<double2*>(img + seg3 + i * s35);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          r0[j] = make_double2(acc[6 * i + 2 * j], acc[6 * i + 2 * j + 1]);
          r3[j] = make_double2(acc[6 * (i + 3) + 2 * j], acc[6 * (i + 3) + 2 * j + 1]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double* r0 = img + seg0 + i * s03;
        double* r3 = img + seg3 + i * s35;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          r0[j] = acc[6 * i + j];
          r3[j] = acc[6 * (i + 3) + j];
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double* r0 = img + seg0 + i * s03;
#pragma unroll
      for (int j = 0; j < 3; ++j) r0[j] = acc[6 * i + j];
    }
  }
}

// One work item = a run of consecutive thread-ordered blocks, hence of consecutive contributions.
// The loop is flat over contributions (a block boundary is just a predicated store + reset), so
// lanes whose items are one 4-contribution block and lanes whose items are four 1-contribution
// blocks stay converged on the expensive part. Codes and block metadata are loaded one step
// ahead of their use.
struct ItemHead {  // first block metadata / contribution code of a work item, fetched early
  uint4 m, m_next;
  uint32_t code;
};
__device__ __forceinline__ ItemHead load_item_head(const AsmArgs& A, const WorkItem& w) {
  const uint4* meta = reinterpret_cast<const uint4*>(A.meta);
  ItemHead h;
  h.m = __ldg(meta + w.blk_begin);  // meta[] / contrib[] are padded, index 0 is always valid
  h.m_next = __ldg(meta + w.blk_begin + 1);
  h.code = __ldg(A.contrib + w.c_begin);
  return h;
}

template <bool kShared>
__device__ __forceinline__ void run_item(const AsmArgs& A, const SlabDesc& d, const WorkItem& w,
                                         const ItemHead& head, double* img,
                                         const double* __restrict__ recs, bool base_even) {
  // idle lanes (blk_count == 0) run zero iterations but still take part in the warp syncs
  const uint4* meta = reinterpret_cast<const uint4*>(A.meta);
  uint32_t p = w.blk_begin;
  uint4 m = head.m;
  uint4 m_next = head.m_next;
  uint32_t remaining = m.w;
  uint32_t code = head.code;
  const uint32_t c_end = w.c_begin + w.c_count;
  double acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.0;
  if (kShared) __syncwarp();  // phase A's records are visible warp-wide
  for (uint32_t c = w.c_begin; c < c_end; ++c) {
    const uint32_t next = __ldg(A.contrib + c + 1);  // contrib[] is padded by one entry
    if (kShared) {
      add_contribution(recs + (code & 0x03FFFFFFu) * 2u, code, acc);  // offset in 16-byte units
    } else {
      // unstaged fallback (a node with thousands of neighbours): gather the record from global
      double rec[kRecStride];
      const uint32_t fe = __ldg(A.elist_compact + d.el_begin + (code & 0x03FFFFFFu));
#pragma unroll
      for (uint32_t ch = 0; ch < kRecStride / 2; ++ch) {
        const double2* src = reinterpret_cast<const double2*>(record_chunk(A, fe, ch));
        double2 v = src ? __ldg(src) : make_double2(0.0, 0.0);
        rec[2 * ch] = v.x;
        rec[2 * ch + 1] = v.y;
      }
      add_contribution_raw(rec, code, acc);
    }
    code = next;
    if (--remaining == 0) {
      store_block<kShared>(img, m, acc, base_even);
#pragma unroll
      for (int i = 0; i < 36; ++i) acc[i] = 0.0;
      ++p;
      m = m_next;
      remaining = m.w;
      m_next = __ldg(meta + p + 1);
    }
  }
}

// One single-warp CTA per slab; no block barriers anywhere.
//   trip 1   work item, slab descriptor and the slab's element list are fetched together: all three
//            are addressable from the slab id alone (dense tables), nothing waits on anything
//   trip 2   each lane cp.async's the records of "its" elements into shared memory (every record is
//            fetched once per CTA, all requests in flight together) while it also pulls its first
//            block metadata and contribution code
//   phase A  the lane that fetched a plate record turns it into the element's shared form (Jacobians,
//            1/det, adj(J) dh, shear sums — everything the element's 16 node-pair blocks share)
//   phase B  each lane evaluates its work item into the shared-memory image of the slab
//   store    lane 0 hands the image to the TMA engine as one bulk shared->global copy
// Record area: the slab's element list is sorted by family, so records sit in three regions
// [trusses][beams][plates] with slot sizes kTrussSlotDoubles / kBeamSlotDoubles / kPlateSlotDoubles;
// contribution codes carry the record's offset.
__global__ void __maxnreg__(224)
assemble_kernel(const AsmArgs A) {
  extern __shared__ __align__(128) double slab_smem[];
  const uint32_t k = blockIdx.x, lane = threadIdx.x;
  const uint4 wraw = __ldg(reinterpret_cast<const uint4*>(A.items) + size_t(k) * kAsmThreads + lane);
  uint32_t fe[kElistStride / kAsmThreads];
#pragma unroll
  for (int j = 0; j < kElistStride / kAsmThreads; ++j)
    fe[j] = __ldg(A.elist + size_t(k) * kElistStride + j * kAsmThreads + lane);
  const SlabDesc d = A.slabs[k];
  WorkItem w;
  w.blk_begin = wraw.x; w.blk_count = wraw.y; w.c_begin = wraw.z; w.c_count = wraw.w;
  if (d.blk_count == 0) return;
  if (d.flags & 1u) {  // slab larger than the staging buffer: everything straight from/to HBM
    run_item<false>(A, d, w, load_item_head(A, w), A.values + d.val_base, nullptr, (d.val_base & 1) == 0);
    return;
  }
  double* img = slab_smem;
  double* recs = slab_smem + ((d.val_count + 1u) & ~1u);
  const uint32_t n_truss = (d.flags >> 8) & 0xFFFu, n_beam = d.flags >> 20;
  double* my_plate[kElistStride / kAsmThreads];
  {
    const uint32_t recs_s = smem_u32(recs);
#pragma unroll
    for (int j = 0; j < kElistStride / kAsmThreads; ++j) {
      my_plate[j] = nullptr;
      if (fe[j] != 0xFFFFFFFFu) {
        const uint32_t slot = j * kAsmThreads + lane, family = fe[j] >> 26;
        // record offset in doubles; the plate's raw 20 doubles land at the start of its slot
        uint32_t off, chunks;
        if (family == FEMGPU_TRUSS) { off = slot * kTrussSlotDoubles; chunks = 2; }
        else if (family == FEMGPU_BEAM) { off = n_truss * kTrussSlotDoubles + (slot - n_truss) * kBeamSlotDoubles; chunks = 8; }
        else if (family == FEMGPU_PLATE) {
          off = n_truss * kTrussSlotDoubles + n_beam * kBeamSlotDoubles + (slot - n_truss - n_beam) * kPlateSlotDoubles;
          chunks = 10;
          my_plate[j] = recs + off;
        } else { off = 0; chunks = 0; }  // placeholder of a remote contribution: no record
        for (uint32_t ch = 0; ch < chunks; ++ch)
          cp_async16(recs_s + off * 8u + ch * 16u, record_chunk(A, fe[j], ch));
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const ItemHead head = load_item_head(A, w);
  // phase A (a lane only touches records it fetched itself; the warp sync is in run_item)
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int j = 0; j < kElistStride / kAsmThreads; ++j) {
    if (my_plate[j]) {
      double raw[20];
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const double2 v = reinterpret_cast<const double2*>(my_plate[j])[i];
        raw[2 * i] = v.x;
        raw[2 * i + 1] = v.y;
      }
      plate_shared_record(raw, my_plate[j]);
    }
  }
  run_item<true>(A, d, w, head, img, recs, true);
  double* out = A.values + d.val_base;
  const uint32_t n = d.val_count;
  if (((d.val_base | n) & 1) == 0) {
    // 16-byte aligned slab: generic-proxy writes -> async proxy, then one TMA bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out),
                   "r"(smem_u32(img)), "r"(n * 8u)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
  __syncwarp();
  // ragged slab (3-wide truss blocks): coalesced 16-byte stores on the aligned body
  const uint32_t odd = uint32_t(d.val_base & 1);  // values[] is 16-byte aligned at index 0
  if (odd && lane == 0) out[0] = img[0];
  const uint32_t body = (n - odd) >> 1;
  if (odd == 0) {
    const double2* src = reinterpret_cast<const double2*>(img);
    double2* dst = reinterpret_cast<double2*>(out);
    for (uint32_t i = lane; i < body; i += kAsmThreads) dst[i] = src[i];
  } else {
    double2* dst = reinterpret_cast<double2*>(out + 1);
    for (uint32_t i = lane; i < body; i += kAsmThreads)
      dst[i] = make_double2(img[1 + 2 * i], img[2 + 2 * i]);
  }
  if (((n - odd) & 1u) && lane == 0) out[n - 1] = img[n - 1];
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
  double rec[kRecStride];
  for (uint32_t ch = 0; ch < kRecStride / 2; ++ch) {
    const double* src = reinterpret_cast<const double*>(record_chunk(A, (uint32_t(family) << 26) | e, ch));
    rec[2 * ch] = src ? src[0] : 0.0;
    rec[2 * ch + 1] = src ? src[1] : 0.0;
  }
  double acc[36];
  for (int i = 0; i < 36; ++i) acc[i] = 0.0;
  add_contribution_raw(rec, (uint32_t(family) << 30) | (uint32_t(pair) << 26), acc);
  for (int i = 0; i < dof; ++i)
    for (int j = 0; j < dof; ++j) out[(la * dof + i) * n + lb * dof + j] = acc[6 * i + j];
}

}  // namespace

int32_t run_assembly(Handle* h) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  if (h->n_slabs == 0) return 0;
  static int attr_set_for = -1;
  if (attr_set_for != h->device) {
    FEMGPU_CUDA_CHECK(h, cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              kSlabSmemBytes));
    attr_set_for = h->device;
  }
  AsmArgs A;
  A.slabs = h->slabs.p;
  A.meta = h->blk_meta.p;
  A.contrib = h->contrib.p;
  A.items = h->items.p;
  A.elist = h->elist.p;
  A.elist_compact = h->elist_compact.p;
  A.truss_rec = reinterpret_cast<const double4*>(h->fd[FEMGPU_TRUSS].rec.p);
  A.beam_rec = h->fd[FEMGPU_BEAM].rec.p;
  A.plate_rec = h->fd[FEMGPU_PLATE].rec.p;
  A.plate_mat = h->fd[FEMGPU_PLATE].mat.p;
  A.values = h->values.p;
  A.n_slabs = h->n_slabs;
  assemble_kernel<<<h->n_slabs, kAsmThreads, h->slab_smem_bytes, h->stream>>>(A);
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
