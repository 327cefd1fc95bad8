// Shared declarations of the femgpu library: device buffers, the handle, kernel launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/femgpu.h"
#include "host_index.hpp"

namespace femgpu {

// ---- error plumbing ---------------------------------------------------------------------------
struct Status {
  int32_t code = 0;
  std::string text;
};

#define FEMGPU_CUDA_CHECK(h, expr)                                                        \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      return (h)->fail(FEMGPU_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(_e) + \
                                            " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    }                                                                                     \
  } while (0)

// ---- device buffer ----------------------------------------------------------------------------
// Grow-only typed device allocation. The handle owns every byte it allocates; sizes are tracked so
// femgpu_device_bytes() is exact.
struct Handle;
// Stream-ordered allocation from the library's PRIVATE memory pool of the current device (api.cu). The pool keeps
// what is freed (release threshold lifted) so the tens of gigabytes a model needs are mapped once per process; the
// device's default pool — which other cudaMallocAsync users of the process share, PyTorch among them — is left alone.
cudaError_t pool_malloc(void** p, size_t bytes, cudaStream_t s);
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  size_t* tally = nullptr;
  // Buffers tied to a handle come from the library's private stream-ordered pool (pool_malloc on the
  // handle's stream), so the tens of gigabytes a model needs are mapped once per process and re-used by the
  // next symbolic pass / reset / handle instead of going back to the driver. Untied buffers fall back to cudaMalloc.
  cudaStream_t* stream = nullptr;
  bool pooled = false;
  void drop(T* q, bool was_pooled) {
    if (!q) return;
    if (was_pooled && stream && *stream) cudaFreeAsync(q, *stream);
    else cudaFree(q);
  }
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    size_t want = n + n / 8 + 16;
    T* q = nullptr;
    const bool pool = stream && *stream;
    cudaError_t e = pool ? pool_malloc(reinterpret_cast<void**>(&q), want * sizeof(T), *stream)
                         : cudaMalloc(&q, want * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p) {
      drop(p, pooled);
      if (tally) *tally -= cap * sizeof(T);
    }
    p = q;
    pooled = pool;
    cap = want;
    if (tally) *tally += cap * sizeof(T);
    return cudaSuccess;
  }
  void release() {
    if (p) {
      drop(p, pooled);
      if (tally) *tally -= cap * sizeof(T);
    }
    p = nullptr;
    cap = 0;
  }
};

// ---- data layout ------------------------------------------------------------------------------
// Host staging keeps struct-of-arrays copies of everything the caller added (needed for rollback,
// lookups by element number and re-upload after reset); the device holds the same SoA arrays.

struct NodeKey {
  uint64_t x, y, z;
  bool operator==(const NodeKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct NodeKeyHash {
  size_t operator()(const NodeKey& k) const {
    uint64_t h = k.x * 0x9E3779B97F4A7C15ull;
    h ^= (k.y + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.z + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    return size_t(h);
  }
};

// number -> index map with a dense fast path (labels are usually 1..n)
struct NumberMap {
  std::vector<uint32_t> dense;  // value+1, 0 = absent
  std::unordered_map<uint32_t, uint32_t> sparse;
  size_t count = 0;
  uint32_t max_label = 0;  // largest label inserted since clear() (labels erased by a rollback may still count: only a fast path asks)
  static constexpr uint32_t kDenseLimit = 1u << 28;
  // A label may sit in `sparse` although it is below dense.size(): it was inserted while the dense table was
  // still short (labels added out of order, e.g. 100000 before 1..70000) and the table grew past it later. So a
  // miss in the dense table always falls through to `sparse` (empty in the common 1..n case: one branch).
  bool find(uint32_t number, uint32_t* idx) const {
    if (number < dense.size()) {
      uint32_t v = dense[number];
      if (v) {
        *idx = v - 1;
        return true;
      }
    }
    if (sparse.empty()) return false;
    auto it = sparse.find(number);
    if (it == sparse.end()) return false;
    *idx = it->second;
    return true;
  }
  void insert(uint32_t number, uint32_t idx) {
    max_label = std::max(max_label, number);
    if (number < kDenseLimit && number <= 8 * (count + 1024)) {
      if (number >= dense.size()) dense.resize(std::max<size_t>(size_t(number) + 1, dense.size() * 2), 0);
      dense[number] = idx + 1;
      if (!sparse.empty()) sparse.erase(number);  // never two homes for one label
    } else {
      sparse[number] = idx;
      if (number < dense.size()) dense[number] = 0;
    }
    ++count;
  }
  // Bulk form of `for i: if (find(number[i])) stop; insert(number[i], first_idx + i)`: inserts the labels up to the
  // first one that already exists (in the map, or earlier in the batch) and returns its position (n if none).
  // Large batches of table-sized labels are claimed by all host cores: every label takes the smallest index that
  // asks for it (atomic min on its table entry), a second pass finds the first position that did not get its own.
  size_t insert_batch(const uint32_t* number, size_t n, uint32_t first_idx) {
    bool parallel = n >= 65536 && sparse.empty() && host_threads() > 1;
    uint32_t mx = 0;
    bool ascending = false;
    if (parallel) {
      std::atomic<uint32_t> amx(0);
      std::atomic<bool> asc(true);
      parallel_chunks(n, 65536, [&](size_t b, size_t e) {
        uint32_t m = 0;
        bool up = b == 0 || number[b - 1] < number[b];
        for (size_t i = b; i < e; ++i) m = std::max(m, number[i]);
        for (size_t i = b + 1; i < e; ++i) up &= number[i - 1] < number[i];
        if (!up) asc.store(false);
        uint32_t cur = amx.load();
        while (m > cur && !amx.compare_exchange_weak(cur, m)) {
        }
      });
      mx = amx.load();
      ascending = asc.load() && (count == 0 || number[0] > max_label);
      parallel = mx < kDenseLimit && mx <= 8 * (count + n + 1024);
    }
    if (!parallel) {
      for (size_t i = 0; i < n; ++i) {
        uint32_t dummy;
        if (find(number[i], &dummy)) return i;
        insert(number[i], first_idx + uint32_t(i));
      }
      return n;
    }
    if (mx >= dense.size()) dense.resize(std::max<size_t>(size_t(mx) + 1, dense.size() * 2), 0);
    uint32_t* tab = dense.data();
    max_label = std::max(max_label, mx);
    if (ascending) {  // strictly ascending labels above every label of the map: no two can meet, nothing to claim
      parallel_chunks(n, 65536, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) tab[number[i]] = first_idx + uint32_t(i) + 1;
      });
      count += n;
      return n;
    }
    parallel_chunks(n, 65536, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; ++i) {
        const uint32_t v = first_idx + uint32_t(i) + 1;
        uint32_t* slot = tab + number[i];
        uint32_t cur = __atomic_load_n(slot, __ATOMIC_RELAXED);
        while ((cur == 0 || cur > v) &&
               !__atomic_compare_exchange_n(slot, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
        }
      }
    });
    std::atomic<size_t> first(n);
    parallel_chunks(n, 65536, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; ++i)
        if (tab[number[i]] != first_idx + uint32_t(i) + 1) {
          size_t cur = first.load();
          while (i < cur && !first.compare_exchange_weak(cur, i)) {
          }
          return;  // later positions of this chunk cannot be the first
        }
    });
    const size_t stop = first.load();
    for (size_t i = stop; i < n; ++i)  // claims of the positions that are not inserted
      if (tab[number[i]] == first_idx + uint32_t(i) + 1) tab[number[i]] = 0;
    count += stop;
    return stop;
  }
  void erase(uint32_t number) {
    if (number < dense.size() && dense[number]) {
      dense[number] = 0;
      --count;
      return;
    }
    if (!sparse.empty() && sparse.erase(number)) --count;
  }
  // forget the labels, keep the table (zeroed by all cores: a re-used instance does not value-initialise it again
  // entry by entry when the next batch sizes it)
  void clear() {
    uint32_t* tab = dense.data();
    parallel_chunks(dense.size(), size_t(1) << 20, [&](size_t b, size_t e) { memset(tab + b, 0, (e - b) * sizeof(uint32_t)); });
    sparse.clear();
    count = 0;
    max_label = 0;
  }
};

struct PlateKey {
  uint32_t n[4];  // sorted
  bool operator==(const PlateKey& o) const {
    return n[0] == o.n[0] && n[1] == o.n[1] && n[2] == o.n[2] && n[3] == o.n[3];
  }
};
struct PlateKeyHash {
  size_t operator()(const PlateKey& k) const {
    uint64_t a = (uint64_t(k.n[0]) << 32) | k.n[1], b = (uint64_t(k.n[2]) << 32) | k.n[3];
    a *= 0x9E3779B97F4A7C15ull;
    a ^= (b + 0x9E3779B97F4A7C15ull + (a << 6) + (a >> 2));
    return size_t(a);
  }
};

constexpr int kFamilies = 3;
constexpr int kNodesPerElem[kFamilies] = {2, 2, 4};
constexpr int kPairsPerElem[kFamilies] = {4, 4, 16};
constexpr int kPropsPerElem[kFamilies] = {3, 11, 4};  // truss: E,A,A2; beam: 8 props + axis[3]; plate: 4
// per-element record written by the prep kernels, laid out like its shared-memory slot in the assembly kernel
// (kTrussSlotDoubles / kBeamSlotDoubles / 16 + 4 for a plate: geometry + material), so a run of consecutive
// elements is one contiguous TMA bulk copy: truss 4 used of 6, beam 16 used of 18, plate 20
constexpr int kTrussSlotDoubles = 6;       // 4 used
constexpr int kBeamSlotDoubles = 18;       // 16 used
constexpr int kPlateRawDoubles = 20;       // geometry record 16 + material 4
constexpr int kRecDoubles[kFamilies] = {kTrussSlotDoubles, kBeamSlotDoubles, kPlateRawDoubles};

struct FamilyHost {
  std::vector<uint32_t> number;       // user label
  std::vector<uint32_t> conn[4];      // node indices (0-based), SoA
  std::vector<uint32_t> conn_number[4];  // node numbers as given (for messages)
  std::vector<double> props[11];      // SoA
  // Offset of every element's first contribution in global insertion order (= accumulation order of the reference).
  // Inside one accepted batch it is affine in the element index, so the host keeps one run per batch and the device
  // array is filled by a kernel (api.cu upload_pending) instead of being written, copied and uploaded per element.
  struct CbaseRun {
    size_t start, count;  // elements [start, start + count) of the family
    int64_t base;         // cbase of element `start`; + pairs per element for each following one
  };
  std::vector<CbaseRun> cbase_runs;
  int64_t cbase_at(size_t i, int pairs) const {
    size_t lo = 0, hi = cbase_runs.size();
    while (hi - lo > 1) {
      size_t mid = (lo + hi) / 2;
      if (cbase_runs[mid].start <= i) lo = mid;
      else hi = mid;
    }
    return cbase_runs[lo].base + int64_t(i - cbase_runs[lo].start) * pairs;
  }
  // the element whose first contribution sits at `cb` (runs ascend in base as they do in start)
  size_t index_of_cbase(int64_t cb, int pairs) const {
    size_t lo = 0, hi = cbase_runs.size();
    while (hi - lo > 1) {
      size_t mid = (lo + hi) / 2;
      if (cbase_runs[mid].base <= cb) lo = mid;
      else hi = mid;
    }
    return cbase_runs[lo].start + size_t((cb - cbase_runs[lo].base) / pairs);
  }
  void cbase_append(size_t start, size_t count, int64_t base, int pairs) {
    if (!count) return;
    if (!cbase_runs.empty()) {
      CbaseRun& b = cbase_runs.back();
      if (b.start + b.count == start && b.base + int64_t(b.count) * pairs == base) {
        b.count += count;
        return;
      }
    }
    cbase_runs.push_back({start, count, base});
  }
  void cbase_truncate(size_t keep) {
    while (!cbase_runs.empty() && cbase_runs.back().start >= keep) cbase_runs.pop_back();
    if (!cbase_runs.empty()) cbase_runs.back().count = std::min(cbase_runs.back().count, keep - cbase_runs.back().start);
  }
  NumberMap by_number;
  size_t size() const { return number.size(); }
  // FEM::reset: forget the elements, keep the storage (a re-used instance re-fills the same pages instead of
  // faulting in half a gigabyte of fresh ones)
  void clear() {
    number.clear();
    for (auto& v : conn) v.clear();
    for (auto& v : conn_number) v.clear();
    for (auto& v : props) v.clear();
    cbase_runs.clear();
    by_number.clear();
  }
};

struct FamilyDev {
  DevBuf<uint32_t> conn[4];
  DevBuf<double> props[11];
  DevBuf<int64_t> cbase;
  DevBuf<double> rec;    // kRecDoubles per element
  DevBuf<int32_t> err;   // per-element validation code
  size_t uploaded = 0;   // elements already on the device
  size_t validated = 0;  // elements already checked by the prep kernel
};

// block metadata consumed by the assembly kernel, in thread order (slab-major)
struct BlockMeta {
  uint32_t seg0;     // slab-relative offset of the block's first entry in dof rows 0..2
  uint32_t seg3;     // same for dof rows 3..5, 0xFFFFFFFF for a 3x3 (truss-only) block
  uint32_t strides;  // row length of rows 0..2 (low 16 bits) | rows 3..5 (high 16 bits)
  uint32_t count;    // contributions of this block (consecutive in contrib[], thread order)
};

// One thread's share of a slab: a run of consecutive thread-ordered blocks (and therefore of
// consecutive contributions). Stored as a dense [n_slabs][asm_threads] table so a thread can fetch
// its item without first reading the slab descriptor.
struct WorkItem {
  uint32_t blk_begin;  // first block (thread order)
  uint32_t blk_count;  // low 16 bits: blocks, 0 = idle thread. Bit 24: the item is one chunk of a block
                       // split over several lanes, bits 16..17 its chunk index; the chunk covers
                       // c_count contributions (execution order) from c_begin - first contribution of the block
  uint32_t c_begin;    // first contribution
  uint32_t c_count;
};

struct SlabDesc {        // 48 bytes = three 16-byte loads
  int64_t val_base;     // first CSR value of the slab
  uint32_t val_count;   // values in the slab
  uint32_t flags;       // bit 0: too large for the staged kernel -> assemble_unstaged_kernel;
                        // bits 8..9: merge rounds of the slab (max chunk index of a split block);
                        // bits 16..31: plates in the element list (staged slabs)
  uint32_t blk_begin;   // first block (thread order)
  uint32_t blk_count;
  uint32_t c_begin;     // the slab's contribution entries (consecutive, lane-major)
  uint32_t c_count;
  uint32_t el_begin;    // the slab's distinct elements in elist_compact[], sorted by family:
  uint32_t el_count;    // [trusses][beams][plates][placeholders of remote contributions]
  uint32_t n_truss;
  uint32_t n_beam;
};

// Contribution entry of a staged slab (32 bits), in the owning lane's execution order: family-major
// (placeholders, plates, beams, trusses) over the lane's blocks, insertion order inside a group.
//   31..30 family   29..26 local node pair   25 group end: flush the accumulator into the block
//   24 the block already holds an earlier group's sum (read-modify-write); on a chunk of a split
//      block: this lane is a sender
//   23..22 merge round: the block is split over several adjacent lanes of one warp. Chunk j >= 1 (bit 24
//          set) hands its partial sum to the lane of chunk 0 in round j (warp shuffle); chunk 0 (bit 24
//          clear) carries the number of rounds it receives and stores the block after the last one
//   21..11 block index inside the slab   10..0 record offset inside the family's record area (16 B units)
// Unstaged slabs keep family<<30 | pair<<26 | slot in the slab's element list, block-major.
constexpr uint32_t kEntEnd = 1u << 25, kEntRmw = 1u << 24;
constexpr int kEntDeferShift = 22, kEntBlkShift = 11;
constexpr uint32_t kEntBlkMask = 0x7FFu, kEntRecMask = 0x7FFu, kEntDeferMask = 3u;
#ifndef FEMGPU_MAX_CHUNKS
#define FEMGPU_MAX_CHUNKS 4
#endif
constexpr int kMaxChunks = FEMGPU_MAX_CHUNKS;  // a block is split over at most this many lanes

struct DistState {
  bool enabled = false;
  int rank = 0, world = 1;
  void* comm = nullptr;  // ncclComm_t
  uint32_t own_begin = 0, own_end = 0;
  bool ownership_set = false;
  // ghost exchange plan (built by the symbolic pass); all counts are node-pair blocks
  std::vector<int64_t> send_blocks, recv_blocks;  // per peer
  std::vector<int64_t> send_off, recv_off;        // per peer: prefix offsets (blocks)
  std::vector<int64_t> send_first_block;          // per peer: first ghost block in blk_key order
  DevBuf<uint64_t> remote_keys;   // received ghost block keys, bit 63 = 6x6, grouped by source rank
  DevBuf<double> send_buf, recv_buf;  // 36 doubles per block
  DevBuf<uint32_t> recv_dst_block;    // per received block: index of the owner's block
  DevBuf<uint8_t> recv_full;          // per received block: 1 = 6x6
  uint64_t last_sent = 0, last_recv = 0;
  std::vector<int64_t> count_matrix;  // [world][world] ghost blocks sender -> owner (the symbolic pass gathers it)
  // Peer-to-peer exchange (dist.cu): every rank exposes one cudaMalloc'ed "window" through CUDA IPC,
  //   [arrived[world] | consumed[world]] (64-byte slots) [ring slot 0] [ring slot 1], a slot = this rank's whole
  // receive area (36 doubles per block, grouped by source rank). The pack kernel of a sender stores its ghost
  // blocks straight into the owner's window over NVLink and then raises arrived[sender] there; the owner's apply
  // kernel waits for that flag (in its own memory), adds the blocks and raises consumed[owner] in the sender's
  // window, which frees the ring slot for the sender's pass after next. No NCCL call in a numeric pass.
  static constexpr int kRing = 2;
  bool p2p = false;
  unsigned char* win = nullptr;
  std::vector<unsigned char*> peer_win;  // peers' windows mapped into this process (nullptr: no traffic with that rank)
  size_t win_slot_bytes = 0, win_bytes = 0;
  std::vector<int64_t> peer_recv_off;    // first block of my run in peer r's receive area
  std::vector<size_t> peer_slot_bytes;   // ring-slot size of peer r's window (its receive area, rounded like mine)
  uint64_t epoch = 0;                    // numeric passes since the plan was built (the same on every rank)
  uint32_t* done_count = nullptr;        // device: per-peer "last CTA" counters of the pack / apply kernels
  volatile uint32_t* h_err = nullptr;    // mapped pinned host word: a kernel timed out waiting for a peer
  uint32_t* d_err = nullptr;
};

struct Handle {
  // ---- properties (structs/props.rs) ----
  double rel_tol = 0, abs_tol = 0;
  uint32_t nodes_number = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  static constexpr int kEvRing = 64;  // event quadruples of the last kEvRing numeric passes
  cudaEvent_t ev[kEvRing][4] = {};
  uint64_t n_numeric = 0;  // numeric passes issued
  mutable Status last;
  size_t dev_bytes = 0;
  uint64_t launches = 0;
  float last_ms[4] = {0, 0, 0, 0};

  // ---- nodes ----
  // Node window (multi-GPU, femgpu_dist_set_node_window): the handle is given the nodes with the global insertion
  // indices [node_index_base, node_index_base + n_nodes()) only — its own rows plus the halo its elements touch. Host
  // and device coordinate arrays are local to the window; element connectivity and everything the symbolic pass
  // builds keep GLOBAL node indices (kernels read coordinates through x_global() etc.).
  uint32_t node_index_base = 0;
  std::vector<uint32_t> node_number;
  std::vector<double> nx, ny, nz;
  NumberMap node_by_number;
  ShardedIndex node_by_xyz;  // hash of the coordinate bit patterns -> node index
  DevBuf<double> d_x, d_y, d_z;
  size_t nodes_uploaded = 0;

  // ---- elements ----
  FamilyHost fh[kFamilies];
  FamilyDev fd[kFamilies];
  ShardedIndex nodeset_seen[kFamilies];  // per family: hash of the sorted node-index set -> element
  int64_t n_contrib = 0;  // running contribution counter (global insertion order)
  // insertion journal for prefix rollback: (family, count) runs in global order
  std::vector<std::pair<int, size_t>> journal;

  // scratch of the batched adds (node indices of the batch, hashes), kept between calls
  std::vector<uint32_t> add_idx[4];
  std::vector<uint64_t> add_hash;
  std::vector<ShardedIndex::Item> add_items;  // scratch of ShardedIndex::insert_batch (the batch partitioned by shard)

  // ---- symbolic products ----
  bool symbolic_valid = false;
  bool values_valid = false;  // `values` holds a finished numeric pass of the CURRENT model (set by femgpu_numeric,
                              // cleared by every add_*, reset and symbolic rebuild)
  int64_t n_rows = 0, nnz = 0;
  uint32_t n_blocks = 0, n_slabs = 0;
  int key_bits = 1;
  DevBuf<uint64_t> blk_key;        // sorted unique (a << key_bits | b)
  DevBuf<uint8_t> blk_full;        // 1 = 6x6, 0 = 3x3
  DevBuf<uint32_t> blk_cptr;       // [n_blocks+1] contributions, thread order
  DevBuf<uint32_t> contrib;        // family<<30 | pair<<26 | element, thread order
  DevBuf<BlockMeta> blk_meta;      // thread order
  DevBuf<uint32_t> blk_order;      // thread position -> sorted block id
  DevBuf<WorkItem> items;          // [n_slabs][asm_threads] balanced per-thread work lists
  DevBuf<uint32_t> items_c;        // the same table as the staged kernel reads it: (first entry - the slab's
                                   // first entry) | entries << 16
  DevBuf<uint32_t> elist;          // [n_slabs][kElistStride]: family<<26 | element, 0xFFFFFFFF = empty;
                                   // (compact list when a slab overflows the table: unstaged path)
  DevBuf<uint32_t> elist_compact;
  // shared-memory regions of the staged assembly kernel (maxima over the staged slabs, bytes)
  uint32_t smem_img = 0, smem_form = 0, smem_rawp = 0, smem_stage = 0;
  uint32_t n_unstaged = 0;         // slabs handled by assemble_unstaged_kernel
  int sm_count = 0;
  uint32_t slab_quota = 72;        // node-pair blocks per slab the symbolic pass aimed for
  int asm_threads = 32;            // 32 or 64, fixed by the symbolic pass
  bool asm_split = false;          // some slab splits a block over several lanes (kernel variant with merge rounds)
  bool asm_bulk = false;           // element records are staged run-wise with TMA bulk copies (consecutive numbering)
  uint32_t asm_smem_set = 0;       // dynamic shared memory assemble_kernel is currently configured for
  int asm_ctas_per_sm = 0;         // persistent CTAs per SM at that size (occupancy query)
  DevBuf<uint32_t> node_blk_ptr;   // [n_nodes_total+1]
  DevBuf<int64_t> node_base;       // [n_nodes_total+1] first value of the node's rows
  DevBuf<uint32_t> node_len;       // [2*n_nodes_total] len03, len35
  DevBuf<uint32_t> blk_off;        // [2*n_blocks] off03, off35 within the node's rows
  DevBuf<SlabDesc> slabs;
  DevBuf<int64_t> row_ptr;         // [n_rows+1]
  DevBuf<int32_t> col_idx;         // [nnz]
  DevBuf<double> values;           // [nnz]
  // the matrix compacted to its entries != 0.0 (femgpu_get_nonzero_csr), valid until the next numeric pass
  bool nz_valid = false;
  uint64_t nz_pass = 0;            // n_numeric the compaction was made for
  int64_t nz_count = 0;
  DevBuf<int64_t> nz_row_ptr;
  DevBuf<int32_t> nz_col;
  DevBuf<double> nz_val;
  DevBuf<double> trig_table;       // element_math.cuh trig_table_init: acos(0), acos(-1), their sin / cos (prep.cu)
  DevBuf<uint8_t> scratch;         // CUB temp storage and sort double-buffers
  DevBuf<int32_t> d_flag;          // small device scalars

  DistState dist;
  struct DistScratch {
    DevBuf<int64_t> i64;
  } dist_scratch;

  // ---- boundary conditions and the separated matrix (separate.cu) ----
  struct BoundaryConditions {
    std::vector<uint8_t> constrained;         // imposed_constraints, [6 * nodes_number] (fem.rs:24,45)
    std::vector<double> displacement, force;  // displacements_vector / forces_vector (fem.rs:18-19);
                                              // `force` holds the concentrated loads only
    // uniformly distributed line / surface loads in call order (family, element index, dof, value): their
    // nodal equivalents are evaluated on the device and added to the forces vector at the next flush
    std::vector<int32_t> load_family, load_dof;
    std::vector<uint32_t> load_elem;
    std::vector<double> load_value;
    bool uploaded = false;
  } bc;
  struct Separated {
    bool valid = false;
    int64_t n_aa = 0, n_bb = 0;
    int64_t nnz[4] = {0, 0, 0, 0};  // aa, ab, ba, bb
    float last_ms = 0.f;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    DevBuf<uint8_t> d_constrained;
    DevBuf<double> d_disp, d_force, rhs;
    DevBuf<uint32_t> cls_pos;         // per DOF: class (0 inactive, 1 a, 2 b) << 30 | local index
    DevBuf<int64_t> aa_idx, bb_idx;   // k_aa_indexes / k_bb_indexes
    DevBuf<int64_t> row_ptr[4];
    DevBuf<int32_t> col[4];
    DevBuf<double> val[4];
    DevBuf<int32_t> tmp[8];           // scratch: class flags, their scans, per-row counts of the 4 quadrants
    DevBuf<unsigned long long> onepass;  // one-pass separation: tile states of the four running counts, ticket, flags
    bool last_onepass = false;        // which path the last separation took
    // direct separation: k_aa_skyline and K_aa in the compacted column form (a, maxa) of the skyline solver
    bool sky_valid = false;
    int64_t sky_total = 0;
    DevBuf<int32_t> sky;
    DevBuf<int64_t> maxa;
    DevBuf<double> sky_a;
  } sep;

  // ---- global analysis (solve.cu): u_a, reactions, composed vectors, element results (results.cu) ----
  struct Solution {
    bool ua_valid = false, composed = false, disp_valid = false;
    int64_t iterations = 0;
    double residual = 0.0;
    float last_ms = 0.f;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    DevBuf<double> u_a, r_r;          // [n_aa], [n_bb]
    DevBuf<double> r, z, p, ap, minv; // PCG work vectors; minv: 1/diag, or 6 doubles per row (block Jacobi)
    DevBuf<uint32_t> blk;             // block Jacobi: first row of the row's block | size << 28
    DevBuf<double> partial, scal;     // reduction partials, device scalars
    DevBuf<double> disp, force;       // [6 * nodes_number] displacements / forces after compose
    DevBuf<double> res[kFamilies];    // element results
  } sol;

  // Range plan of the numeric pass (symbolic.cu build_range_plan): the slabs are cut into n_ranges consecutive
  // ranges, and every family's elements are listed by the first range that needs their record (prep_order, with
  // range_elem_end[f][r] = end of range r's share of the list). The record kernels of range r + 1 run on a low-priority
  // stream while the assembly kernel works on range r, so only the first range's records sit on the critical path.
  static constexpr int kMaxRanges = 16;
  int n_ranges = 1;
  uint32_t range_slab_end[kMaxRanges] = {};
  uint32_t range_elem_end[kFamilies][kMaxRanges] = {};
  DevBuf<uint32_t> prep_order[kFamilies];
  cudaStream_t prep_stream = nullptr;
  cudaStream_t xchg_stream = nullptr;   // ghost-first multi-GPU pass: the pack kernels run here, under the assembly
  cudaEvent_t ghost_ev = nullptr, pack_ev = nullptr;
  cudaEvent_t range_ev[kMaxRanges] = {};
  cudaEvent_t range_t0[kEvRing][kMaxRanges] = {}, range_t1[kEvRing][kMaxRanges] = {};  // per-range assembly launch times
  int range_count[kEvRing] = {};  // ranges of the pass in that ring slot (0: one launch between ev[1] and ev[2])

  // side streams of the numeric pass: the element-record kernels of the three families are independent and
  // each is bound by gather latency at a third of the SM's warp slots, so they run side by side (prep.cu)
  cudaStream_t side_stream[2] = {nullptr, nullptr};
  cudaEvent_t fork_ev = nullptr, join_ev[2] = {nullptr, nullptr};

  // Pinned bounce buffers of the bulk host->device path (api.cu h2d_staged): the host staging vectors
  // are pageable, so large uploads are copied chunk-wise into pinned memory by several host threads
  // while the previous chunk is on its way over PCIe.
  // host staging vectors registered with CUDA (api.cu: pin_take / make_room); on from the first femgpu_reset
  struct PinnedRegion {
    const void* base;
    size_t bytes;
  };
  std::vector<PinnedRegion> pinned;
  bool pin_host = false;
  static constexpr size_t kPinChunk = size_t(16) << 20;
  void* pin_buf[2] = {nullptr, nullptr};
  cudaEvent_t pin_ev[2] = {nullptr, nullptr};
  bool pin_busy[2] = {false, false};
  int pin_next = 0;

  Handle() {
    auto tie = [&](auto& b) {
      b.tally = &dev_bytes;
      b.stream = &stream;
    };
    tie(d_x); tie(d_y); tie(d_z);
    for (auto& f : fd) {
      for (auto& c : f.conn) tie(c);
      for (auto& p : f.props) tie(p);
      tie(f.cbase); tie(f.rec); tie(f.err);
    }
    tie(blk_key); tie(blk_full); tie(blk_cptr); tie(contrib); tie(blk_meta); tie(blk_order); tie(items); tie(items_c); tie(elist); tie(elist_compact);
    tie(node_blk_ptr); tie(node_base); tie(node_len); tie(blk_off); tie(slabs); tie(row_ptr);
    tie(col_idx); tie(values); tie(scratch); tie(d_flag); tie(nz_row_ptr); tie(nz_col); tie(nz_val); tie(trig_table);
    for (auto& o : prep_order) tie(o);
    tie(dist.send_buf); tie(dist.recv_buf); tie(dist.recv_dst_block); tie(dist.recv_full);
    tie(dist.remote_keys); tie(dist_scratch.i64);
    tie(sep.d_constrained); tie(sep.d_disp); tie(sep.d_force); tie(sep.rhs); tie(sep.cls_pos); tie(sep.aa_idx);
    tie(sep.bb_idx);
    for (int q = 0; q < 4; ++q) {
      tie(sep.row_ptr[q]); tie(sep.col[q]); tie(sep.val[q]);
    }
    for (auto& t : sep.tmp) tie(t);
    tie(sep.onepass);
    tie(sep.sky); tie(sep.maxa); tie(sep.sky_a);
    tie(sol.u_a); tie(sol.r_r); tie(sol.r); tie(sol.z); tie(sol.p); tie(sol.ap); tie(sol.minv); tie(sol.blk);
    tie(sol.partial); tie(sol.scal); tie(sol.disp); tie(sol.force);
    for (auto& r : sol.res) tie(r);
  }

  int32_t fail(int32_t code, const std::string& text) const {
    last.code = code;
    last.text = text;
    return code;
  }
  size_t n_nodes() const { return node_number.size(); }
  // device coordinate arrays addressed by GLOBAL node index (only the window may be dereferenced)
  const double* x_global() const { return d_x.p - node_index_base; }
  const double* y_global() const { return d_y.p - node_index_base; }
  const double* z_global() const { return d_z.p - node_index_base; }
};

// ---- kernels' host entry points (one per translation unit) -------------------------------------
int32_t upload_pending(Handle* h);                       // api.cu
int32_t run_prep(Handle* h, bool validate_only);         // prep.cu
int32_t run_prep_range(Handle* h, int range, cudaStream_t st);  // prep.cu: records of the elements first needed by `range`
int32_t first_error(Handle* h, int* family, size_t* index, int* code);  // prep.cu
int32_t run_symbolic(Handle* h);                         // symbolic.cu
int32_t run_assembly(Handle* h, uint32_t slab_begin, uint32_t slab_end);  // numeric.cu: staged slabs of [begin, end)
int32_t run_assembly_unstaged(Handle* h);                // numeric.cu: the oversized slabs
int32_t element_matrix(Handle* h, int family, size_t index, double* out_host);   // numeric.cu
int32_t element_rotation(Handle* h, int family, size_t index, double* out_host); // prep.cu
int32_t element_slots(Handle* h, int family, size_t index, int64_t* out_host);   // symbolic.cu
int32_t nonzero_coo(Handle* h, int64_t* count, int64_t* rows, int64_t* cols, double* vals);  // symbolic.cu
int32_t nonzero_csr(Handle* h, int64_t* count, int64_t* row_ptr, int32_t* cols, double* vals);  // symbolic.cu
int32_t run_load_kernel(Handle* h, uint32_t n, const int32_t* d_family, const uint32_t* d_elem, const int32_t* d_dof,
                        const double* d_value, uint32_t* d_key, double* d_val);  // prep.cu
int32_t forces_flush(Handle* h);                         // separate.cu: bc -> sep.d_constrained / d_disp / d_force
int32_t run_separate(Handle* h, bool direct);            // separate.cu
int32_t run_skyline(Handle* h);                          // separate.cu
void sep_release(Handle* h);                             // separate.cu
void bc_clear(Handle* h);                                // separate.cu
int32_t run_element_results(Handle* h, int family, const double* d_u, double* d_out);  // results.cu
void sol_release(Handle* h);                             // solve.cu
void sol_invalidate(Handle* h);                          // solve.cu
int32_t dist_numeric_exchange(Handle* h);                // dist.cu
bool dist_ghost_first(Handle* h, uint32_t* first_ghost_slab);  // dist.cu: can the ghost slabs be assembled and sent first?
int32_t dist_begin_pass(Handle* h);                      // dist.cu: ghost-first pass: next epoch
int32_t dist_pack(Handle* h, cudaStream_t st);           // dist.cu: my ghost blocks -> the owners' windows
int32_t dist_apply(Handle* h);                           // dist.cu: received partials += into my rows
int32_t dist_setup_p2p(Handle* h);                       // dist.cu: collective, called when the exchange plan is final
int32_t dist_check(Handle* h);                           // dist.cu: after a stream sync — did an exchange time out?
void dist_destroy(Handle* h);                            // dist.cu
int32_t dist_allgather_i64(Handle* h, const int64_t* send, int64_t* recv, size_t n);  // dist.cu
int32_t dist_exchange_8(Handle* h, const void* send, const int64_t* send_offs, const int64_t* send_counts,
                        void* recv, const int64_t* recv_offs, const int64_t* recv_counts);  // dist.cu

// Threads per assembly CTA = lanes that share one slab: 32 or 64, chosen by the symbolic pass
// (Handle::asm_threads). One warp has the lowest per-slab overhead and wins on single-family plate /
// truss meshes; two warps double the resident warps per SM at the same shared memory and balance
// mixed and beam meshes better (measured, see profiles/).
constexpr int kAsmThreadsMax = 64;
constexpr int kSlabQuota = 72;             // node-pair blocks a slab aims for (8 plate-grid nodes)
// per-slab capacity of the staged kernel's shared-memory regions; a slab exceeding any of them (a
// node with hundreds of neighbours) goes to the unstaged kernel
constexpr int kCapImgBytes = 40 * 1024;    // slab image (CSR values of the slab)
constexpr int kCapFormBytes = 24 * 1024;   // shared forms of the slab's plates
constexpr int kCapStageBytes = 20 * 1024;  // block metadata + contribution entries + truss/beam records
constexpr int kCapBlocks = 2048;           // blocks per staged slab (11-bit index in the entries)
// Record slots in the CTA's shared-memory record area, in doubles. Odd multiples of 16 bytes so that
// lanes reading the same field of different elements spread over the banks.
constexpr int kPlateSlotDoubles = 66;      // 64 used: the plate's shared form (element_math.cuh)
constexpr int kElistStride = 64;           // element slots per slab in the dense elist table; a slab
                                           // touching more elements takes the unstaged path
constexpr int kRecStride = 20;             // doubles per staged element record (plate: 16 + 4)
// relative cost of one contribution, used to balance the per-thread work lists
#ifndef FEMGPU_COST_T
#define FEMGPU_COST_T 1
#endif
#ifndef FEMGPU_COST_B
#define FEMGPU_COST_B 8
#endif
#ifndef FEMGPU_COST_P
#define FEMGPU_COST_P 16
#endif
constexpr uint32_t kCostTruss = FEMGPU_COST_T, kCostBeam = FEMGPU_COST_B, kCostPlate = FEMGPU_COST_P;

inline uint32_t div_up(uint64_t a, uint64_t b) { return uint32_t((a + b - 1) / b); }

}  // namespace femgpu

struct femgpu_handle : femgpu::Handle {};
