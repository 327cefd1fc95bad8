// Numeric assembly: the hot kernel.
//
// Gather formulation — "owner computes": one lane owns one node-pair block (a, b) of the global
// matrix (6x6, or 3x3 for a truss-only pair) at a time. It walks the block's contributions
// (family, local pair, element), evaluates each element's global-frame block T^T k[la,lb] T from
// the element record (element_math.cuh) and sums into 36 FP64 registers. Nothing is accumulated
// through global memory: no atomics, no colouring; the order of summation is fixed by the symbolic
// pass (family-major, insertion order inside a family), so the result is deterministic and every
// CSR value is written to HBM exactly once.
//
// A warp owns a "slab": a contiguous range of node rows, hence a contiguous range of CSR values.
// Lanes drop their blocks into a shared-memory image of the slab (the 6 row segments of a block
// are 6 doubles wide and a row apart, which would be a poor global store pattern) and the image
// leaves as ONE TMA bulk store.
//
// The kernel is persistent (single-warp CTAs, a fixed number per SM, slabs dealt round-robin) and
// software-pipelined: while slab i is evaluated, the element records, block metadata and
// contribution entries of slab i+1 are in flight (cp.async) and the descriptors of slab i+2 are
// being loaded into registers; the bulk store of slab i drains while slab i+1 is evaluated. No
// memory latency sits on the critical path of a warp.
//
// Algorithmic traffic per launch: 8 B x nnz written + element records read. See DESIGN.md.
#include "common.cuh"
#include "element_math.cuh"

namespace femgpu {

namespace {

struct AsmArgs {
  const SlabDesc* slabs;
  const BlockMeta* meta;
  const uint32_t* contrib;
  const WorkItem* items;
  const uint32_t* items_c;        // staged kernel: [n_slabs][kT] packed (first entry - slab's first) | count << 16
  const uint32_t* elist;          // dense [n_slabs][kElistStride]
  const uint32_t* elist_compact;  // compact lists (unstaged kernel only)
  const double* truss_rec;        // records at the stride of their shared-memory slots (kRecDoubles)
  const double* beam_rec;
  const double* plate_rec;        // geometry (16) + material (4)
  double* values;
  uint32_t slab_begin;  // this launch works on the slabs [slab_begin, n_slabs)
  uint32_t n_slabs;
  // shared-memory regions of the staged kernel, bytes: [image][plate forms][raw plates][stage x2]
  uint32_t smem_img, smem_form, smem_rawp, smem_stage;
};

// Per-phase cycle accounting of the staged kernel (profiling builds only: make prof). Warp leaders
// accumulate clock() deltas per phase and add them to g_phase at the end of the kernel.
#ifdef FEMGPU_PHASE_CLOCKS
constexpr int kPhases = 12;
__device__ unsigned long long g_phase[kPhases + 4];
#define PHASE_DECL uint32_t ph_acc[kPhases] = {}; uint32_t ph_last = clock(); uint32_t ph_iters = 0, ph_work = 0, ph_slabs = 0;
#define PHASE_MARK(i) { const uint32_t ph_now = clock(); ph_acc[i] += ph_now - ph_last; ph_last = ph_now; }
#else
#define PHASE_DECL
#define PHASE_MARK(i)
#endif

// Ablation builds (make variant NAME=... DEFS=-DFEMGPU_ABL=<bits>; results are WRONG, only the kernel time means
// something): which part of a slab's work bounds the kernel is found by leaving parts out, one at a time.
//   1 no bulk store of the image     2 no contribution loop (phase B)     4 no staging loads for the next slab
//   8 no block flush into the image  16 no fence.proxy.async before the store   32 no plate forms (phase A)
#ifndef FEMGPU_ABL
#define FEMGPU_ABL 0
#endif
constexpr bool kAblNoStore = (FEMGPU_ABL & 1) != 0, kAblNoLoop = (FEMGPU_ABL & 2) != 0, kAblNoStage = (FEMGPU_ABL & 4) != 0,
               kAblNoFlush = (FEMGPU_ABL & 8) != 0, kAblNoFence = (FEMGPU_ABL & 16) != 0, kAblNoForms = (FEMGPU_ABL & 32) != 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}

// Completion of the cp.async staging is tracked by an mbarrier, not by cp.async groups: the
// commit/wait_group pair and cp.async.bulk.wait_group share one hardware scoreboard (both compile to
// DEPBAR.LE SB0), so waiting for the bulk store of the previous slab would also wait for the loads
// of the next one that were issued a moment earlier.
__device__ __forceinline__ void mbar_init(uint32_t mbar_s, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_s), "r"(count) : "memory");
}
// the mbarrier receives this thread's arrival once all its prior cp.async have landed
__device__ __forceinline__ void cp_async_arrive(uint32_t mbar_s) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mbar_s) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar_s, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(mbar_s), "r"(parity)
      : "memory");
}

// global address of 16-byte chunk `chunk` of an element's record (nullptr past its end):
// plate = 8 chunks of geometry + 2 of material, beam = 8, truss = 2
__device__ __forceinline__ const void* record_chunk(const AsmArgs& A, uint32_t fe, uint32_t chunk) {
  const uint32_t family = fe >> 26, e = fe & 0x03FFFFFFu;
  if (family == FEMGPU_PLATE) return chunk < 10 ? A.plate_rec + size_t(e) * kPlateRawDoubles + chunk * 2 : nullptr;
  if (family == FEMGPU_BEAM) return chunk < 8 ? A.beam_rec + size_t(e) * kBeamSlotDoubles + chunk * 2 : nullptr;
  if (family != FEMGPU_TRUSS) return nullptr;  // family 3: placeholder of a remote contribution
  return chunk < 2 ? A.truss_rec + size_t(e) * kTrussSlotDoubles + chunk * 2 : nullptr;
}

// Place a 6x6 / 3x3 block into the slab image (shared memory) or straight into the CSR values
// (unstaged kernel). kRmw: the block already holds an earlier group's sum, add it first. When every
// row segment is 16-byte aligned (true whenever the node only has 6-wide blocks) the six doubles
// of a row move as three 16-byte accesses: at the 48-byte lane stride of neighbouring blocks those
// are bank-conflict free, 8-byte accesses are 4-way conflicted.
template <bool kRmw>
__device__ __forceinline__ void store_block(double* __restrict__ img, const uint4 m, double acc[36],
                                            bool base_even) {
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
          double2 v0 = make_double2(acc[6 * i + 2 * j], acc[6 * i + 2 * j + 1]);
          double2 v3 = make_double2(acc[6 * (i + 3) + 2 * j], acc[6 * (i + 3) + 2 * j + 1]);
          if (kRmw) {
            const double2 o0 = r0[j], o3 = r3[j];
            v0.x = o0.x + v0.x; v0.y = o0.y + v0.y;
            v3.x = o3.x + v3.x; v3.y = o3.y + v3.y;
          }
          r0[j] = v0;
          r3[j] = v3;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double* r0 = img + seg0 + i * s03;
        double* r3 = img + seg3 + i * s35;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          r0[j] = kRmw ? r0[j] + acc[6 * i + j] : acc[6 * i + j];
          r3[j] = kRmw ? r3[j] + acc[6 * (i + 3) + j] : acc[6 * (i + 3) + j];
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double* r0 = img + seg0 + i * s03;
#pragma unroll
      for (int j = 0; j < 3; ++j) r0[j] = kRmw ? r0[j] + acc[6 * i + j] : acc[6 * i + j];
    }
  }
}

// One half of a block: dof rows 0..2 (half 0: acc[0..17]) or 3..5 (half 1: acc[18..35]). Used when the two
// chunks of a split block each finish one half (phase B, pair merge). A 3x3 block only has a half 0.
__device__ __forceinline__ void store_half(double* __restrict__ img, const uint4 m, const double acc[36], int half) {
  const uint32_t seg0 = m.x, seg3 = m.y, s03 = m.z & 0xFFFFu, s35 = m.z >> 16;
  const bool full = seg3 != 0xFFFFFFFFu;
  if (full) {
    const uint32_t seg = half ? seg3 : seg0, st = half ? s35 : s03;
    const bool al = ((seg | st) & 1u) == 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double* r = img + seg + i * st;
      if (al) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
          reinterpret_cast<double2*>(r)[j] = half ? make_double2(acc[18 + 6 * i + 2 * j], acc[18 + 6 * i + 2 * j + 1])
                                                  : make_double2(acc[6 * i + 2 * j], acc[6 * i + 2 * j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) r[j] = half ? acc[18 + 6 * i + j] : acc[6 * i + j];
      }
    }
  } else if (half == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double* r0 = img + seg0 + i * s03;
#pragma unroll
      for (int j = 0; j < 3; ++j) r0[j] = acc[6 * i + j];
    }
  }
}

// ---- staged, persistent, pipelined kernel --------------------------------------------------------

// kT = threads per CTA (32 or 64), a template parameter of everything below.
// CTA-wide barrier / vote: a warp-level sync when the CTA is a single warp
template <int kT>
__device__ __forceinline__ void cta_sync() {
  if (kT == 32) __syncwarp();
  else __syncthreads();
}
template <int kT>
__device__ __forceinline__ bool cta_all(bool pred) {
  if (kT == 32) return __all_sync(0xFFFFFFFFu, pred) != 0;
  return __syncthreads_and(pred) != 0;
}

// what a thread keeps in registers about a slab that is still to come
template <int kT>
struct SlabRegs {
  // Two adjacent lanes share an element slot and copy alternate 16-byte chunks of its record, so
  // both halves of every 32-byte sector land in shared memory in one wavefront (an LDGSTS costs one
  // wavefront of the shared-memory data pipe per sector it touches).
  static constexpr int kElistPerPair = kElistStride / (kT / 2);
  uint4 d0, d1, d2;             // the SlabDesc as three 16-byte words
  uint32_t fe[kElistPerPair];   // the lane pair's slots of the slab's element list
  uint32_t c_begin, c_count;    // the thread's work item
  __device__ __forceinline__ int64_t val_base() const { return int64_t((uint64_t(d0.y) << 32) | d0.x); }
  __device__ __forceinline__ uint32_t val_count() const { return d0.z; }
  __device__ __forceinline__ uint32_t rounds() const { return (d0.w >> 8) & 3u; }
  __device__ __forceinline__ uint32_t n_plate() const { return d0.w >> 16; }
  __device__ __forceinline__ uint32_t blk_begin() const { return d1.x; }
  __device__ __forceinline__ uint32_t blk_count() const { return d1.y; }
  __device__ __forceinline__ uint32_t slab_c_begin() const { return d1.z; }
  __device__ __forceinline__ uint32_t slab_c_count() const { return d1.w; }
  __device__ __forceinline__ uint32_t n_truss() const { return d2.z; }
  __device__ __forceinline__ uint32_t n_beam() const { return d2.w; }
  // stage layout (bytes): [block metadata][entries][truss records][beam records]
  __device__ __forceinline__ uint32_t ent_off() const { return blk_count() * 16u; }
  __device__ __forceinline__ uint32_t ent_bytes() const {
    return (((slab_c_begin() & 3u) + slab_c_count() + 1u) * 4u + 15u) & ~15u;
  }
  __device__ __forceinline__ uint32_t truss_off() const { return ent_off() + ent_bytes(); }
  __device__ __forceinline__ uint32_t beam_off() const {
    return truss_off() + n_truss() * uint32_t(kTrussSlotDoubles * 8);
  }
  // an oversized slab is done by the unstaged kernel; here it becomes an empty slab. Called when the
  // registers are first used, one iteration after the loads were issued.
  __device__ __forceinline__ void sanitize() {
    if (d0.w & 1u) {
      d0.z = 0;
      d0.w = 1u;
      d1.y = 0;
      d1.w = 0;
      d2.z = d2.w = 0;
      c_count = 0;
#pragma unroll
      for (int j = 0; j < kElistPerPair; ++j) fe[j] = 0xFFFFFFFFu;
    }
  }
};

// Descriptor block of a slab in shared memory: [SlabDesc 48 B][work items kT x 4 B][element
// list kElistStride x 4 B]. Everything is addressable from the slab id alone (dense tables), so it
// is requested two slabs ahead with cp.async; nothing that is in flight lives in registers.
constexpr uint32_t kDescItemsOff = 48;
// alignment of the shared-memory regions after the image (TMA bulk copies and cp.async need 16 bytes)
#ifndef FEMGPU_SMEM_ALIGN
#define FEMGPU_SMEM_ALIGN 16u
#endif
constexpr uint32_t kSmemAlign = FEMGPU_SMEM_ALIGN;
// a lane only needs where its entries start and how many they are: 4 bytes per lane
constexpr uint32_t kDescItemBytes = 4;
template <int kT> __host__ __device__ constexpr uint32_t desc_elist_off() { return kDescItemsOff + kT * kDescItemBytes; }
template <int kT> __host__ __device__ constexpr uint32_t desc_bytes() { return (desc_elist_off<kT>() + kElistStride * 4 + kSmemAlign - 1) & ~(kSmemAlign - 1); }

template <int kT>
__device__ __forceinline__ void issue_desc(const AsmArgs& A, uint32_t k, uint32_t dbuf_s, uint32_t tid) {
  constexpr uint32_t kItemLanes = kT / 4, kElistLanes = kElistStride / 4;
  if (tid < kItemLanes) cp_async16(dbuf_s + kDescItemsOff + tid * 16u, A.items_c + size_t(k) * kT + tid * 4u);
  else if (tid < kItemLanes + kElistLanes)
    cp_async16(dbuf_s + desc_elist_off<kT>() + (tid - kItemLanes) * 16u,
               A.elist + size_t(k) * kElistStride + (tid - kItemLanes) * 4u);
  else if (tid < kItemLanes + kElistLanes + 3)
    cp_async16(dbuf_s + (tid - kItemLanes - kElistLanes) * 16u,
               reinterpret_cast<const uint4*>(A.slabs + k) + (tid - kItemLanes - kElistLanes));
}

// Issuing a TMA bulk copy stalls the issuing warp for a few hundred cycles, and a slab needs five to eight of
// them. In the two-warp shape the warp that is NOT the critical one should issue them: in a slab without plates
// that is warp 1 (warp 0 carries the split blocks and their merges) — it issues the metadata / entry loads and
// heads the element slots 0..31, where most records of a slab sit (B: 0.437 -> 0.430 ms); in a slab with plates
// warp 1 already builds the next slab's forms, and warp 0 keeps the issue work (M: 3.25 ms, against 3.36 the
// other way round). `flip` is uniform over the CTA (it comes from the slab descriptor).
template <int kT> __device__ __forceinline__ uint32_t bulk_slot_lane(uint32_t tid, bool flip) {
  return kT == 64 && flip ? tid ^ 32u : tid;
}

// kBulk: a lane holds the element slots j * kT + bulk_slot_lane(tid) (it heads the run-wise bulk copies); otherwise
// two adjacent lanes share the slots j * kT / 2 + tid / 2 (they copy alternate 16-byte chunks of a record)
template <int kT, bool kBulk>
__device__ __forceinline__ SlabRegs<kT> read_desc(const unsigned char* dbuf, uint32_t tid) {
  SlabRegs<kT> R;
  const uint4* sp = reinterpret_cast<const uint4*>(dbuf);
  R.d0 = sp[0];
  R.d1 = sp[1];
  R.d2 = sp[2];
  const uint32_t* el = reinterpret_cast<const uint32_t*>(dbuf + desc_elist_off<kT>());
#pragma unroll
  for (int j = 0; j < SlabRegs<kT>::kElistPerPair; ++j) {
    if (kBulk) R.fe[j] = j < kElistStride / kT ? el[j * kT + bulk_slot_lane<kT>(tid, R.n_plate() == 0u)] : 0xFFFFFFFFu;
    else R.fe[j] = el[j * (kT / 2) + (tid >> 1)];
  }
  const uint32_t w = reinterpret_cast<const uint32_t*>(dbuf + kDescItemsOff)[tid];
  R.c_begin = R.slab_c_begin() + (w & 0xFFFFu);
  R.c_count = w >> 16;
  R.sanitize();
  return R;
}

// cp.async everything slab R needs into `stage` (+ its raw plate records into `rawp`): block
// metadata, contribution entries, and each lane pair the records of "its" elements — every record
// is fetched once per slab, all requests in flight together
template <int kT, bool kBulk>
__device__ __forceinline__ void issue_stage(const AsmArgs& A, const SlabRegs<kT>& R, uint32_t stage_s,
                                            uint32_t rawp_s, uint32_t mbar_s, uint32_t tid) {
  if (kAblNoStage) return;
  // block metadata and contribution entries are contiguous: two TMA bulk loads by one thread,
  // completing on the same mbarrier as the record copies
  const uint32_t nt = R.n_truss(), nbm = R.n_beam();
  const bool flip = kBulk && kT == 64 && R.n_plate() == 0u;
  // the metadata / entry loads always go out from warp 1 in the two-warp bulk shape: in a slab with plates that
  // splits the issue work about evenly (warp 0 heads five record runs, warp 1 one + these two)
  if (tid == (kBulk && kT == 64 ? 32u : 0u) && R.blk_count()) {
    const uint32_t meta_bytes = R.blk_count() * 16u, ent_bytes = R.slab_c_count() ? R.ent_bytes() : 0u;
    const uint32_t rec_bytes = kBulk ? nt * uint32_t(kTrussSlotDoubles * 8) + nbm * uint32_t(kBeamSlotDoubles * 8) +
                                           R.n_plate() * uint32_t(kPlateRawDoubles * 8)
                                     : 0u;
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(mbar_s),
                 "r"(meta_bytes + ent_bytes + rec_bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(stage_s),
        "l"(A.meta + R.blk_begin()), "r"(meta_bytes), "r"(mbar_s)
        : "memory");
    if (ent_bytes)
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              stage_s + R.ent_off()),
          "l"(A.contrib + (R.slab_c_begin() & ~3u)), "r"(ent_bytes), "r"(mbar_s)
          : "memory");
  }
  if (kBulk) {
    // The slab's element list is sorted by (family, element) and the records lie in global memory exactly as in
    // their shared-memory slots, so a run of consecutively numbered elements — the rule in a mesh numbered along
    // its grid lines — is ONE contiguous TMA bulk copy: a slab of the plate grid needs 2 copies of 9 plate
    // records instead of 180 LDGSTS, none of which passes through the LSU data pipe. The lane at the head of a
    // run (found with a shuffle and two ballots) issues the copy. The symbolic pass selects this variant only
    // when the runs are long enough on average (Handle::asm_bulk): a bulk copy per element is slower than LDGSTS.
    const uint32_t lane = tid & 31u;
#pragma unroll
    for (int j = 0; j < kElistStride / kT; ++j) {
      const uint32_t fe = R.fe[j], family = fe >> 26, e = fe & 0x03FFFFFFu;
      const bool copy = fe != 0xFFFFFFFFu && family < 3u;
      const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, fe, 1);
      const bool head = copy && (lane == 0u || fe != prev + 1u || (prev >> 26) != family);
      const uint32_t heads = __ballot_sync(0xFFFFFFFFu, head), copies = __ballot_sync(0xFFFFFFFFu, copy);
      if (head) {
        const uint32_t after = lane == 31u ? 0u : heads >> (lane + 1u);
        uint32_t len = after ? uint32_t(__ffs(int(after))) : 32u - lane;
        const uint32_t stop = (~copies) >> lane;  // bit 0 = this lane
        if (stop) len = min(len, uint32_t(__ffs(int(stop))) - 1u);
        const uint32_t slot = uint32_t(j) * kT + bulk_slot_lane<kT>(tid, flip);
        uint32_t dst, bytes;
        const double* src;
        if (family == FEMGPU_PLATE) {
          dst = rawp_s + (slot - nt - nbm) * uint32_t(kPlateRawDoubles * 8);
          src = A.plate_rec + size_t(e) * kPlateRawDoubles;
          bytes = len * uint32_t(kPlateRawDoubles * 8);
        } else if (family == FEMGPU_BEAM) {
          dst = stage_s + R.beam_off() + (slot - nt) * uint32_t(kBeamSlotDoubles * 8);
          src = A.beam_rec + size_t(e) * kBeamSlotDoubles;
          bytes = len * uint32_t(kBeamSlotDoubles * 8);
        } else {
          dst = stage_s + R.truss_off() + slot * uint32_t(kTrussSlotDoubles * 8);
          src = A.truss_rec + size_t(e) * kTrussSlotDoubles;
          bytes = len * uint32_t(kTrussSlotDoubles * 8);
        }
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
            "l"(src), "r"(bytes), "r"(mbar_s)
            : "memory");
      }
    }
    return;
  }
  const uint32_t half = tid & 1u;
#pragma unroll
  for (int j = 0; j < SlabRegs<kT>::kElistPerPair; ++j) {
    const uint32_t fe = R.fe[j];
    if (fe == 0xFFFFFFFFu) continue;
    const uint32_t slot = j * (kT / 2) + (tid >> 1), family = fe >> 26, e = fe & 0x03FFFFFFu;
    if (family == FEMGPU_PLATE) {
      // chunks 0..7 = geometry, 8..9 = material; this lane takes every other one
      const uint32_t dst = rawp_s + (slot - nt - nbm) * uint32_t(kPlateRawDoubles * 8) + half * 16u;
      const double* rec = A.plate_rec + size_t(e) * kPlateRawDoubles + half * 2;
#pragma unroll
      for (uint32_t ch = 0; ch < 5; ++ch) cp_async16(dst + ch * 32u, rec + ch * 4);
    } else if (family == FEMGPU_BEAM) {
      const uint32_t dst = stage_s + R.beam_off() + (slot - nt) * uint32_t(kBeamSlotDoubles * 8) + half * 16u;
      const double* rec = A.beam_rec + size_t(e) * kBeamSlotDoubles + half * 2;
#pragma unroll
      for (uint32_t ch = 0; ch < 4; ++ch) cp_async16(dst + ch * 32u, rec + ch * 4);
    } else if (family == FEMGPU_TRUSS) {
      const uint32_t dst = stage_s + R.truss_off() + slot * uint32_t(kTrussSlotDoubles * 8) + half * 16u;
      cp_async16(dst, A.truss_rec + size_t(e) * kTrussSlotDoubles + half * 2);
    }  // family 3: placeholder of a remote contribution, no record
  }
}

// phase A: one thread per plate (alternating between the two warps) turns the raw record into the
// element's shared form (element_math.cuh). Returns whether every plate of the slab has Q == I.
template <int kT>
__device__ __forceinline__ bool phase_a(const SlabRegs<kT>& R, const double* __restrict__ rawp,
                                        double* __restrict__ form, uint32_t tid) {
  bool flat = true;
  const uint32_t np = R.n_plate();
  if (kT == 64) {
    // two adjacent lanes per plate, two Gauss points each: up to 32 plates in one pass
    for (uint32_t idx = tid >> 1; idx < ((np + 31u) & ~31u); idx += kT / 2) {
      const uint32_t pair_mask = __activemask();
      if (idx < np) {
        double raw[20];
        const double2* src = reinterpret_cast<const double2*>(rawp + idx * 20u);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const double2 v = src[i];
          raw[2 * i] = v.x;
          raw[2 * i + 1] = v.y;
        }
        flat = flat && raw[15] != 0.0;
        plate_shared_record_half(raw, form + idx * uint32_t(kPlateSlotDoubles), int(tid & 1u),
                                 3u << (tid & 30u));
      }
      (void)pair_mask;
    }
  } else {
    for (uint32_t idx = tid; idx < np; idx += kT) {
      double raw[20];
      const double2* src = reinterpret_cast<const double2*>(rawp + idx * 20u);
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const double2 v = src[i];
        raw[2 * i] = v.x;
        raw[2 * i + 1] = v.y;
      }
      flat = flat && raw[15] != 0.0;
      plate_shared_record(raw, form + idx * uint32_t(kPlateSlotDoubles));
    }
  }
  // two-warp shape: the image is free once the TMA engine has read the previous slab out of it; the
  // closing barrier publishes that to the CTA together with the forms
  if (kT == 64 && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  return cta_all<kT>(flat);  // also: forms visible CTA-wide, raw plate records free again
}

// phase B: the thread runs its contribution entries. The loop is flat over contributions — a group
// end is a flush of the accumulators into the image, after which `keep` = 0 makes the next
// contribution overwrite them (no zeroing) — so threads with one long group and threads with
// several short ones stay converged on the expensive part. The chunks of a split block keep their
// sums in registers and are merged with warp shuffles after the loop.
// kSplit = false: the model has no split block anywhere (decided by the symbolic pass), the merge-round
// code is compiled out.
template <int kT, bool kSplit>
__device__ __forceinline__ void phase_b(const SlabRegs<kT>& R, const unsigned char* __restrict__ stage,
                                        const double* __restrict__ form, double* __restrict__ img,
                                        const PlatePair* __restrict__ pairs, bool all_flat
#ifdef FEMGPU_PHASE_CLOCKS
                                        , uint32_t* ph_acc, uint32_t& ph_last, uint32_t& ph_iters, uint32_t& ph_work
#endif
                                        ) {
  const uint4* meta = reinterpret_cast<const uint4*>(stage);
  const uint32_t* ent = reinterpret_cast<const uint32_t*>(stage + R.ent_off()) + (R.slab_c_begin() & 3u);
  const double* truss = reinterpret_cast<const double*>(stage + R.truss_off());
  const double* beam = reinterpret_cast<const double*>(stage + R.beam_off());
  uint32_t i = R.c_begin - R.slab_c_begin();
  const uint32_t end = i + R.c_count;
  double acc[36];
#pragma unroll
  for (int q = 0; q < 36; ++q) acc[q] = 0.0;
  double keep = 0.0;
  uint32_t pending = 0;  // merge rounds of the thread's last group (a chunk of a split block)
  bool sender = false;
  uint4 pending_m = make_uint4(0u, 0u, 0u, 0u);
  if (!kAblNoLoop && i < end) {
    uint32_t code = ent[i];
    PlatePair pt = pairs[(code >> 26) & 15u];
    for (; i < end; ++i) {
      const uint32_t next = ent[i + 1];  // the entry area is padded by one
      // the next contribution's pair entry and this one's block metadata are requested now, a whole
      // contribution before they are needed
      const PlatePair pt_next = pairs[(next >> 26) & 15u];
      const uint4 m = meta[(code >> kEntBlkShift) & kEntBlkMask];
      const uint32_t family = code >> 30, pair = (code >> 26) & 15u, rec = (code & kEntRecMask) * 2u;
      if (family == FEMGPU_PLATE) {
        plate_block_shared(form + rec, pt, keep, all_flat, acc);
      } else {
#pragma unroll
        for (int q = 0; q < 36; ++q) acc[q] *= keep;
        if (family == FEMGPU_BEAM) {
          beam_block(beam + rec, int(pair >> 1), int(pair & 1u), acc);
        } else if (family == FEMGPU_TRUSS) {
          const double* t = truss + rec;
          truss_block(t[0], t[1], t[2], t[3], int(pair >> 1), int(pair & 1u), acc);
        }
        // family 3: slot reserved for another rank's contribution (multi-GPU), contributes zero
      }
      keep = 1.0;
      if (code & kEntEnd) {
        const uint32_t defer = (code >> kEntDeferShift) & kEntDeferMask;
        if (kSplit && defer) {  // a chunk of a split block: always the thread's last entry
          pending = defer;
          sender = (code & kEntRmw) != 0;
          pending_m = m;
        } else if (kAblNoFlush) {
          if (acc[0] == 1.2345e300) img[0] = acc[7] + acc[35];  // never true: keeps the sums alive
        } else if (code & kEntRmw) {
          store_block<true>(img, m, acc, true);
        } else {
          store_block<false>(img, m, acc, true);
        }
        keep = 0.0;
      }
      code = next;
      pt = pt_next;
    }
  }
#ifdef FEMGPU_PHASE_CLOCKS
  ph_iters += __reduce_max_sync(0xFFFFFFFFu, R.c_count);
  ph_work += __reduce_add_sync(0xFFFFFFFFu, R.c_count);
#endif
  PHASE_MARK(5)
  // Merge rounds: the lane of chunk 0 collects the partial sums of the block's other chunks from the
  // lanes right above it, in chunk order (so the sum is (chunk 0 + chunk 1) + chunk 2 ...), and stores
  // the block once. No barrier and no read-modify-write through the image.
  const uint32_t rounds = kSplit ? R.rounds() : 0u;  // CTA-uniform
  if (kSplit && rounds == 1u) {
    // Every split block of the slab has two chunks, in lanes l and l + 1. Instead of handing all 36 partial sums
    // to lane l, the two lanes swap halves — l gives its rows 3..5 and gets the partner's rows 0..2, in the same
    // shuffle — and each finishes and stores one half: 18 shuffles and 9 stores per lane instead of 36 and 18 on
    // one. a + b == b + a exactly, so the block is the same (chunk 0) + (chunk 1) as in the general path below.
    const bool part = pending != 0u;
    if (__any_sync(0xFFFFFFFFu, part)) {  // warp-uniform
      const uint32_t lane = threadIdx.x & 31u, partner = part ? (sender ? lane - 1u : lane + 1u) : lane;
#pragma unroll
      for (int q = 0; q < 18; ++q) {
        const double give = sender ? acc[q] : acc[q + 18];
        const double t = __shfl_sync(0xFFFFFFFFu, give, partner);
        if (sender) acc[q + 18] += t;
        else if (part) acc[q] += t;
      }
      if (part) store_half(img, pending_m, acc, sender ? 1 : 0);
    }
  } else if (kSplit && rounds) {
    const uint32_t recv = sender ? 0u : pending;
    for (uint32_t r = 1; r <= rounds; ++r) {
      if (!__any_sync(0xFFFFFFFFu, recv >= r)) continue;  // warp-uniform
#pragma unroll
      for (int q = 0; q < 36; ++q) {
        const double t = __shfl_down_sync(0xFFFFFFFFu, acc[q], r);
        if (recv >= r) acc[q] += t;
      }
    }
    if (recv) store_block<false>(img, pending_m, acc, true);
  }
}

// Two-warp shape: warp 1 finishes its loop trips earlier than warp 0 (which carries the beam trips and the
// merges), so it builds the plate forms of the NEXT slab in that time (the forms are double-buffered) and phase A
// leaves the critical path of the CTA. 0 = both warps build the current slab's forms at the top of the iteration.
#ifndef FEMGPU_AHEAD
#define FEMGPU_AHEAD 1
#endif
// registers per thread of the two-warp shape: 168 = three warps per SM sub-partition (5-6 CTAs per SM when
// shared memory allows; a few bytes of spill), 200 = two (4 CTAs per SM)
#ifndef FEMGPU_MAXNREG64
#define FEMGPU_MAXNREG64 168
#endif
template <int kT, bool kSplit, bool kBulk>
__global__ void __maxnreg__(kT == 64 ? FEMGPU_MAXNREG64 : 255)
assemble_kernel(const AsmArgs A) {
  constexpr bool kAhead = kT == 64 && FEMGPU_AHEAD != 0;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t tid = threadIdx.x, stride = gridDim.x;
  double* img = reinterpret_cast<double*>(smem);
  double* form0 = reinterpret_cast<double*>(smem + A.smem_img);
  const uint32_t form_all = (kAhead ? 2u : 1u) * A.smem_form;
  double* rawp = reinterpret_cast<double*>(smem + A.smem_img + form_all);
  unsigned char* stage0 = smem + A.smem_img + form_all + A.smem_rawp;
  unsigned char* dbuf0 = stage0 + 2 * A.smem_stage;  // three rotating descriptor blocks
  const uint32_t rawp_s = smem_u32(rawp), stage0_s = smem_u32(stage0), dbuf0_s = smem_u32(dbuf0);

  uint32_t k = A.slab_begin + blockIdx.x;
  if (k >= A.n_slabs) return;
  const uint32_t mbar_s = dbuf0_s + 3 * desc_bytes<kT>();
  PlatePair* pairs = reinterpret_cast<PlatePair*>(dbuf0 + 3 * desc_bytes<kT>() + 16);
  volatile uint32_t* flat_flag = reinterpret_cast<volatile uint32_t*>(dbuf0 + 3 * desc_bytes<kT>() + 8);  // [2], kAhead
  if (tid == 0) mbar_init(mbar_s, kT);
  if (tid < 16) pairs[tid] = make_plate_pair(int(tid >> 2), int(tid & 3u));
  cta_sync<kT>();
  // prologue: descriptor of the first slab, then its stage and the descriptor of the second
  issue_desc<kT>(A, k, dbuf0_s, tid);
  cp_async_arrive(mbar_s);
  mbar_wait(mbar_s, 0);
  SlabRegs<kT> cur = read_desc<kT, kBulk>(dbuf0, tid);
  issue_stage<kT, kBulk>(A, cur, stage0_s, rawp_s, mbar_s, tid);
  if (k + stride < A.n_slabs) issue_desc<kT>(A, k + stride, dbuf0_s + desc_bytes<kT>(), tid);
  cp_async_arrive(mbar_s);
  if (kAhead) {  // the first slab's forms: both warps, like the classic phase A
    mbar_wait(mbar_s, 1u);
    const bool flat0 = phase_a<kT>(cur, rawp, form0, tid);
    if (tid == 0) flat_flag[0] = flat0 ? 1u : 0u;
  }

  PHASE_DECL
  uint32_t d_cur = 0;  // descriptor block of `cur`
  for (uint32_t it = 0;; ++it) {
    PHASE_MARK(0)
    const uint32_t buf = it & 1u;
    const unsigned char* stage = stage0 + buf * A.smem_stage;
    double* form = kAhead ? form0 + buf * (A.smem_form / 8u) : form0;
    const uint32_t d_nxt = (d_cur == 2u) ? 0u : d_cur + 1u, d_nn = (d_nxt == 2u) ? 0u : d_nxt + 1u;
    const bool has_next = k + stride < A.n_slabs;
    // slab `cur`: its records, metadata and entries (and the next slab's descriptor) were
    // requested one iteration ago (batch it + 1 of the mbarrier)
    // (each thread waits on the mbarrier itself; the previous slab's phase B ended with a CTA barrier)
    mbar_wait(mbar_s, (it + 1u) & 1u);
    PHASE_MARK(1)
    PHASE_MARK(2)
    bool all_flat;
    if (kAhead) {
      // forms[buf] and their flag were finished before the barrier that closed the previous slab (or in the
      // prologue); this barrier publishes that the TMA engine has read the previous slab out of the image
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      cta_sync<kT>();
      all_flat = flat_flag[buf] != 0u;
    } else {
      all_flat = phase_a<kT>(cur, rawp, form, tid);
    }
    PHASE_MARK(3)
    uint32_t np_next = 0;
    if (has_next) {
      const SlabRegs<kT> nxt = read_desc<kT, kBulk>(dbuf0 + d_nxt * desc_bytes<kT>(), tid);
      np_next = nxt.n_plate();
      issue_stage<kT, kBulk>(A, nxt, stage0_s + (buf ^ 1u) * A.smem_stage, rawp_s, mbar_s, tid);
      if (k + 2 * stride < A.n_slabs) issue_desc<kT>(A, k + 2 * stride, dbuf0_s + d_nn * desc_bytes<kT>(), tid);
      cp_async_arrive(mbar_s);
    }
    // one-warp shape: wait for the previous slab's bulk store to have left the image as late as
    // possible — phase A and the staging above do not touch it
    if (kT == 32) {
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
    PHASE_MARK(4)
#ifdef FEMGPU_PHASE_CLOCKS
    phase_b<kT, kSplit>(cur, stage, form, img, pairs, all_flat, ph_acc, ph_last, ph_iters, ph_work);
    ++ph_slabs;
#else
    phase_b<kT, kSplit>(cur, stage, form, img, pairs, all_flat);
#endif
    PHASE_MARK(6)
    if (kAhead && has_next && tid >= 32u) {
      // warp 1: the next slab's records have been on their way since the top of this iteration
      mbar_wait(mbar_s, it & 1u);  // batch it + 2
      double* form_next = form0 + (buf ^ 1u) * (A.smem_form / 8u);
      bool flat = true;
      for (uint32_t idx = tid - 32u; idx < (kAblNoForms ? 0u : np_next); idx += 32u) {
        double raw[20];
        const double2* src = reinterpret_cast<const double2*>(rawp + idx * 20u);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const double2 v = src[i];
          raw[2 * i] = v.x;
          raw[2 * i + 1] = v.y;
        }
        flat = flat && raw[15] != 0.0;
        plate_shared_record(raw, form_next + idx * uint32_t(kPlateSlotDoubles));
      }
      flat = __all_sync(0xFFFFFFFFu, flat) != 0;
      if (tid == 32u) flat_flag[buf ^ 1u] = flat ? 1u : 0u;
    }

    const uint32_t n = cur.val_count();
    if (kAhead && !n) cta_sync<kT>();  // the barrier below is what orders warp 1's forms before their use
    if (n) {
      double* out = A.values + cur.val_base();
      if (((uint32_t(cur.val_base()) | n) & 1u) == 0) {
        // 16-byte aligned slab: generic-proxy writes -> async proxy, then one TMA bulk store
        if (!kAblNoFence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        cta_sync<kT>();
        if (tid == 0 && !kAblNoStore) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out),
                       "r"(smem_u32(img)), "r"(n * 8u)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        cta_sync<kT>();
        // ragged slab (3-wide truss blocks): coalesced 16-byte stores on the aligned body
        const uint32_t odd = uint32_t(cur.val_base() & 1);  // values[] is 16-byte aligned at index 0
        if (odd && tid == 0) out[0] = img[0];
        const uint32_t body = (n - odd) >> 1;
        if (odd == 0) {
          const double2* src = reinterpret_cast<const double2*>(img);
          double2* dst = reinterpret_cast<double2*>(out);
          for (uint32_t i = tid; i < body; i += kT) dst[i] = src[i];
        } else {
          double2* dst = reinterpret_cast<double2*>(out + 1);
          for (uint32_t i = tid; i < body; i += kT)
            dst[i] = make_double2(img[1 + 2 * i], img[2 + 2 * i]);
        }
        if (((n - odd) & 1u) && tid == 0) out[n - 1] = img[n - 1];
      }
    }
    PHASE_MARK(7)
    if (!has_next) break;
    k += stride;
    d_cur = d_nxt;
    cur = read_desc<kT, kBulk>(dbuf0 + d_cur * desc_bytes<kT>(), tid);
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef FEMGPU_PHASE_CLOCKS
  PHASE_MARK(8)
  if ((tid & 31u) == 0) {
    for (int i = 0; i < kPhases; ++i) atomicAdd(&g_phase[i], (unsigned long long)ph_acc[i]);
    atomicAdd(&g_phase[kPhases], (unsigned long long)ph_iters);
    atomicAdd(&g_phase[kPhases + 1], (unsigned long long)ph_work);
    atomicAdd(&g_phase[kPhases + 2], (unsigned long long)ph_slabs);
    atomicAdd(&g_phase[kPhases + 3], 1ull);
  }
#endif
}

// ---- unstaged kernel: slabs too large for shared memory (a node with hundreds of neighbours) ------
// One warp per oversized slab; records come from global memory per contribution, blocks go straight
// to the CSR values. Entries are block-major: family<<30 | pair<<26 | slot in the slab's element list.
constexpr int kRawDoubles = 20;
__device__ __forceinline__ void add_contribution_raw(const double* __restrict__ raw, uint32_t code,
                                                     double acc[36]) {
  const uint32_t family = code >> 30, pair = (code >> 26) & 15u;
  if (family == FEMGPU_PLATE) {
    __align__(16) double S[kPlateSharedDoubles];
    plate_shared_record(raw, S);
    const PlatePair pt = make_plate_pair(int(pair >> 2), int(pair & 3u));
    plate_block_shared(S, pt, 1.0, raw[15] != 0.0, acc);
  } else if (family == FEMGPU_BEAM) {
    beam_block(raw, int(pair >> 1), int(pair & 1u), acc);
  } else if (family == FEMGPU_TRUSS) {
    truss_block(raw[0], raw[1], raw[2], raw[3], int(pair >> 1), int(pair & 1u), acc);
  }
}

__global__ void __launch_bounds__(kAsmThreadsMax)
assemble_unstaged_kernel(const AsmArgs A) {
  const uint32_t k = blockIdx.x, lane = threadIdx.x, kAsmThreads = blockDim.x;
  const SlabDesc d = A.slabs[k];
  if (!(d.flags & 1u) || d.blk_count == 0) return;
  const WorkItem w = A.items[size_t(k) * kAsmThreads + lane];
  const uint4* meta = reinterpret_cast<const uint4*>(A.meta);
  double* out = A.values + d.val_base;
  const bool base_even = (d.val_base & 1) == 0;
  uint32_t c = w.c_begin;
  for (uint32_t p = w.blk_begin; p < w.blk_begin + (w.blk_count & 0xFFFFu); ++p) {
    const uint4 m = __ldg(meta + p);
    double acc[36];
#pragma unroll
    for (int q = 0; q < 36; ++q) acc[q] = 0.0;
    for (uint32_t j = 0; j < m.w; ++j, ++c) {
      const uint32_t code = __ldg(A.contrib + c);
      if ((code >> 30) == 3u) continue;  // remote placeholder
      const uint32_t fe = __ldg(A.elist_compact + d.el_begin + (code & 0x03FFFFFFu));
      double raw[kRawDoubles];
#pragma unroll
      for (uint32_t ch = 0; ch < uint32_t(kRawDoubles / 2); ++ch) {
        const double2* src = reinterpret_cast<const double2*>(record_chunk(A, fe, ch));
        const double2 v = src ? __ldg(src) : make_double2(0.0, 0.0);
        raw[2 * ch] = v.x;
        raw[2 * ch + 1] = v.y;
      }
      add_contribution_raw(raw, code, acc);
    }
    store_block<false>(out, m, acc, base_even);
  }
}

// (kRawDoubles: the largest per-element record read from global memory: plate 16 + 4)
// test hook: the whole transformed element matrix of one element, built from the same block
// evaluators the assembly uses
__global__ void element_matrix_kernel(int family, uint32_t e, const double* truss_rec,
                                      const double* beam_rec, const double* plate_rec, double* out) {
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
  double rec[kRawDoubles];
  for (uint32_t ch = 0; ch < uint32_t(kRawDoubles / 2); ++ch) {
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

static AsmArgs asm_args(Handle* h) {
  AsmArgs A;
  A.slabs = h->slabs.p;
  A.meta = h->blk_meta.p;
  A.contrib = h->contrib.p;
  A.items = h->items.p;
  A.items_c = h->items_c.p;
  A.elist = h->elist.p;
  A.elist_compact = h->elist_compact.p;
  A.truss_rec = h->fd[FEMGPU_TRUSS].rec.p;
  A.beam_rec = h->fd[FEMGPU_BEAM].rec.p;
  A.plate_rec = h->fd[FEMGPU_PLATE].rec.p;
  A.values = h->values.p;
  A.slab_begin = 0;
  A.n_slabs = h->n_slabs;
  auto up = [](uint32_t b) { return (b + kSmemAlign - 1u) & ~(kSmemAlign - 1u); };
  A.smem_img = up(h->smem_img);
  A.smem_form = up(h->smem_form);
  A.smem_rawp = up(h->smem_rawp);
  A.smem_stage = up(h->smem_stage);
  return A;
}

int32_t run_assembly_unstaged(Handle* h) {
  if (h->n_slabs == 0 || h->n_unstaged == 0) return 0;
  const AsmArgs A = asm_args(h);
  assemble_unstaged_kernel<<<h->n_slabs, h->asm_threads, 0, h->stream>>>(A);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  return 0;
}

int32_t run_assembly(Handle* h, uint32_t slab_begin, uint32_t slab_end) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  if (h->n_slabs == 0 || slab_end <= slab_begin) return 0;
  AsmArgs A = asm_args(h);
  A.slab_begin = slab_begin;
  A.n_slabs = slab_end;
  const int threads = h->asm_threads;
  const uint32_t desc = threads == 64 ? desc_bytes<64>() : desc_bytes<32>();
  const uint32_t form_bufs = (threads == 64 && FEMGPU_AHEAD != 0) ? 2u : 1u;  // kAhead of assemble_kernel
  const uint32_t smem = A.smem_img + form_bufs * A.smem_form + A.smem_rawp + 2 * A.smem_stage + 3 * desc + 16 +
                        16 * uint32_t(sizeof(PlatePair));
  if (h->n_unstaged < h->n_slabs) {
    const bool split = h->asm_split, bulk = h->asm_bulk;
    using Kernel = void (*)(const AsmArgs);
    static const Kernel table[2][2][2] = {
        {{assemble_kernel<32, false, false>, assemble_kernel<32, false, true>},
         {assemble_kernel<32, true, false>, assemble_kernel<32, true, true>}},
        {{assemble_kernel<64, false, false>, assemble_kernel<64, false, true>},
         {assemble_kernel<64, true, false>, assemble_kernel<64, true, true>}}};
    const Kernel kernel = table[threads == 64][split][bulk];
    const void* fn = reinterpret_cast<const void*>(kernel);
    const uint32_t config = smem | (threads == 64 ? 1u << 28 : 0u) | (bulk ? 1u << 29 : 0u) | (split ? 1u << 30 : 0u);
    if (h->asm_smem_set != config) {
      FEMGPU_CUDA_CHECK(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      int per_sm = 0;
      FEMGPU_CUDA_CHECK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
      if (per_sm < 1) return h->fail(FEMGPU_ERR_CUDA, "assemble_kernel does not fit on an SM");
      if (h->sm_count == 0)
        FEMGPU_CUDA_CHECK(h, cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
      h->asm_smem_set = config;
      h->asm_ctas_per_sm = per_sm;
      if (getenv("FEMGPU_ASM_INFO"))  // shared-memory budget of the staged kernel, to stderr
        fprintf(stderr, "[femgpu asm] T=%d split=%d bulk=%d smem=%u B (image %u, forms %u, raw plates %u, stage 2 x %u, descriptors 3 x %u) -> %d CTAs/SM, %u slabs (%u unstaged)\n",
                threads, int(split), int(bulk), smem, A.smem_img, A.smem_form, A.smem_rawp, A.smem_stage, desc, per_sm, h->n_slabs, h->n_unstaged);
    }
    const uint32_t grid = uint32_t(std::min<uint64_t>(slab_end - slab_begin, uint64_t(h->sm_count) * h->asm_ctas_per_sm));
    kernel<<<grid, threads, smem, h->stream>>>(A);
    h->launches++;
    FEMGPU_CUDA_CHECK(h, cudaGetLastError());
#ifdef FEMGPU_PHASE_CLOCKS
    if (getenv("FEMGPU_PHASE_DUMP")) {
      unsigned long long ph[kPhases + 4];
      FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
      FEMGPU_CUDA_CHECK(h, cudaMemcpyFromSymbol(ph, g_phase, sizeof ph));
      unsigned long long zero[kPhases + 4] = {};
      FEMGPU_CUDA_CHECK(h, cudaMemcpyToSymbol(g_phase, zero, sizeof zero));
      const double ws = double(ph[kPhases + 2]);  // warp-slabs
      static const char* names[kPhases] = {"loop top", "wait stage (mbarrier)", "wait image free", "phase A", "issue next stage",
                                           "phase B loop", "merge rounds", "store issue", "tail", "-", "-", "-"};
      fprintf(stderr, "[femgpu phases] T=%d grid=%u warps=%llu slabs/warp=%.1f\n", threads, grid, ph[kPhases + 3],
              ws / double(ph[kPhases + 3]));
      double tot = 0;
      for (int i = 0; i < 9; ++i) tot += double(ph[i]);
      for (int i = 0; i < 9; ++i)
        fprintf(stderr, "[femgpu phases] %-24s %9.1f cycles/slab  %5.1f %%\n", names[i], double(ph[i]) / ws, 100.0 * double(ph[i]) / tot);
      fprintf(stderr, "[femgpu phases] total %.1f cycles/slab; phase-B trips/slab %.2f, lane work/slab %.2f, lane efficiency %.3f\n",
              tot / ws, double(ph[kPhases]) / ws, double(ph[kPhases + 1]) / ws,
              double(ph[kPhases + 1]) / (32.0 * double(ph[kPhases])));
    }
#endif
  }
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
  element_matrix_kernel<<<1, 32, 0, h->stream>>>(family, uint32_t(index), h->fd[FEMGPU_TRUSS].rec.p,
                                                 h->fd[FEMGPU_BEAM].rec.p, h->fd[FEMGPU_PLATE].rec.p, d_out);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(out_host, d_out, size_t(n) * n * 8, cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

}  // namespace femgpu
