// One-time symbolic pass (integer work only, bit-exact by construction):
//   global DOF numbering  row = 6*node_index + dof            (structs/node.rs:8, fem.rs:37)
//   node-pair block list  sort/unique of (row node, col node) keys of every element's local pairs
//                         — the start_positions tables of methods_for_truss_data_handle.rs:83-109,
//                         methods_for_beam_data_handle.rs:96-123, methods_for_plate_data_handle.rs:113-132
//   CSR row_ptr / col_idx on the structural pattern (3x3 per truss-only pair, 6x6 otherwise)
//   gather map            per block, the (family, local pair, element) contributions in global
//                         insertion order (= the order the reference accumulates in)
//   slabs                 contiguous node ranges whose CSR values one CTA stages in shared memory,
//                         blocks inside a slab ordered by contribution count so the threads of a
//                         warp loop the same number of times.
// Sorting/scanning uses CUB device primitives; everything else is hand-written kernels.
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace femgpu {

namespace {

// stream the scratch allocations of the running pass are ordered on (set by the entry points below)
thread_local cudaStream_t t_tmp_stream = nullptr;

struct CastToI64 {
  __host__ __device__ int64_t operator()(int32_t v) const { return int64_t(v); }
};

struct Tmp {  // RAII scratch allocation from the stream-ordered pool
  void* p = nullptr;
  cudaStream_t s = nullptr;
  cudaError_t alloc(size_t bytes) {
    release();
    s = t_tmp_stream;
    return s ? pool_malloc(&p, bytes ? bytes : 16, s) : cudaMalloc(&p, bytes ? bytes : 16);
  }
  void release() {
    if (p) {
      if (s) cudaFreeAsync(p, s);
      else cudaFree(p);
    }
    p = nullptr;
  }
  ~Tmp() { release(); }
  template <typename T>
  T* as() {
    return static_cast<T*>(p);
  }
};

template <int kNodes>
__global__ void gen_contrib_kernel(uint32_t n_elem, int family, const uint32_t* __restrict__ c0,
                                   const uint32_t* __restrict__ c1, const uint32_t* __restrict__ c2,
                                   const uint32_t* __restrict__ c3,
                                   const int64_t* __restrict__ cbase, int key_bits,
                                   uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  constexpr int kPairs = kNodes * kNodes;
  uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t e = uint32_t(t / kPairs);
  int pair = int(t % kPairs);
  if (e >= n_elem) return;
  int la = pair / kNodes, lb = pair % kNodes;
  const uint32_t* cs[4] = {c0, c1, c2, c3};
  uint32_t a = cs[la][e], b = cs[lb][e];
  int64_t pos = cbase[e] + pair;
  keys[pos] = (uint64_t(a) << key_bits) | b;
  vals[pos] = (uint32_t(family) << 30) | (uint32_t(pair) << 26) | e;
}

__global__ void block_kind_kernel(uint32_t n_blocks, const uint32_t* __restrict__ cptr,
                                  const uint32_t* __restrict__ contrib, uint8_t* __restrict__ full) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_blocks) return;
  uint8_t f = 0;
  for (uint32_t c = cptr[i]; c < cptr[i + 1]; ++c) {
    uint32_t fam = contrib[c] >> 30;
    // family 3 = placeholder for a block another rank contributes to; bit 26 says whether it is 6x6
    if (fam == FEMGPU_BEAM || fam == FEMGPU_PLATE || (fam == 3u && ((contrib[c] >> 26) & 1u))) {
      f = 1;
      break;
    }
  }
  full[i] = f;
}

__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t* __restrict__ a, uint32_t n,
                                                    uint64_t v) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* __restrict__ a, uint32_t n,
                                                    uint32_t v) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// node_blk_ptr[a] = first block whose row node is >= a, for a in [0, n_nodes]
__global__ void node_ptr_kernel(uint32_t n_nodes, uint32_t n_blocks, int key_bits,
                                const uint64_t* __restrict__ blk_key,
                                uint32_t* __restrict__ node_blk_ptr) {
  uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a > n_nodes) return;
  node_blk_ptr[a] = (a == n_nodes) ? n_blocks : lower_bound_u64(blk_key, n_blocks, uint64_t(a) << key_bits);
}

// per node: offsets of its blocks inside dof rows 0..2 / 3..5, the two row lengths, value count
__global__ void node_layout_kernel(uint32_t n_nodes, const uint32_t* __restrict__ node_blk_ptr,
                                   const uint8_t* __restrict__ full, uint32_t* __restrict__ blk_off,
                                   uint32_t* __restrict__ node_len, int64_t* __restrict__ node_size,
                                   int32_t* __restrict__ overflow) {
  uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_nodes) return;
  uint32_t l03 = 0, l35 = 0;
  for (uint32_t b = node_blk_ptr[a]; b < node_blk_ptr[a + 1]; ++b) {
    blk_off[2 * b] = l03;
    blk_off[2 * b + 1] = full[b] ? l35 : 0xFFFFFFFFu;
    l03 += full[b] ? 6 : 3;
    l35 += full[b] ? 6 : 0;
  }
  node_len[2 * a] = l03;
  node_len[2 * a + 1] = l35;
  node_size[a] = 3 * int64_t(l03) + 3 * int64_t(l35);
  if (l03 >= 65536u) *overflow = 1;
}

__global__ void row_ptr_kernel(uint32_t n_nodes, const int64_t* __restrict__ node_base,
                               const uint32_t* __restrict__ node_len, int64_t* __restrict__ row_ptr) {
  uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a > n_nodes) return;
  if (a == n_nodes) {
    row_ptr[6 * size_t(a)] = node_base[a];
    return;
  }
  int64_t base = node_base[a];
  uint32_t l03 = node_len[2 * a], l35 = node_len[2 * a + 1];
#pragma unroll
  for (int i = 0; i < 3; ++i) row_ptr[6 * size_t(a) + i] = base + int64_t(i) * l03;
#pragma unroll
  for (int i = 0; i < 3; ++i) row_ptr[6 * size_t(a) + 3 + i] = base + 3 * int64_t(l03) + int64_t(i) * l35;
}

// one warp per node: lanes stride over (block, dof row, dof col) of the node's blocks
__global__ void col_idx_kernel(uint32_t n_nodes, int key_bits, const uint64_t* __restrict__ blk_key,
                               const uint32_t* __restrict__ node_blk_ptr,
                               const uint8_t* __restrict__ full, const uint32_t* __restrict__ blk_off,
                               const uint32_t* __restrict__ node_len,
                               const int64_t* __restrict__ node_base, int32_t* __restrict__ col_idx) {
  uint32_t warp = uint32_t((uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n_nodes) return;
  uint32_t a = warp;
  uint32_t b0 = node_blk_ptr[a], b1 = node_blk_ptr[a + 1];
  if (b0 == b1) return;
  uint32_t l03 = node_len[2 * a], l35 = node_len[2 * a + 1];
  int64_t base = node_base[a];
  const uint64_t mask = (uint64_t(1) << key_bits) - 1;
  uint32_t total = (b1 - b0) * 36;
  for (uint32_t t = lane; t < total; t += 32) {
    uint32_t blk = b0 + t / 36, i = (t % 36) / 6, j = t % 6;
    bool f = full[blk] != 0;
    if (!f && (i >= 3 || j >= 3)) continue;
    uint32_t col_node = uint32_t(blk_key[blk] & mask);
    int64_t pos = (i < 3) ? base + int64_t(i) * l03 + blk_off[2 * blk] + j
                          : base + 3 * int64_t(l03) + int64_t(i - 3) * l35 + blk_off[2 * blk + 1] + j;
    col_idx[pos] = int32_t(6 * col_node + j);
  }
}

__global__ void slab_kernel(uint32_t n_slabs, uint32_t quota, uint32_t n_nodes,
                            const uint32_t* __restrict__ node_blk_ptr,
                            const int64_t* __restrict__ node_base,
                            SlabDesc* __restrict__ slabs, int32_t* __restrict__ overflow) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_slabs) return;
  uint32_t first = lower_bound_u32(node_blk_ptr, n_nodes, k * quota);
  uint32_t next = (k + 1 == n_slabs) ? n_nodes : lower_bound_u32(node_blk_ptr, n_nodes, (k + 1) * quota);
  SlabDesc d;
  d.val_base = node_base[first];
  int64_t cnt = node_base[next] - d.val_base;
  if (cnt >= (int64_t(1) << 32)) *overflow = 2;
  d.val_count = uint32_t(cnt);
  d.blk_begin = node_blk_ptr[first];
  d.blk_count = node_blk_ptr[next] - d.blk_begin;
  d.flags = 0;
  d.c_begin = d.c_count = 0;
  d.el_begin = d.el_count = 0;
  d.n_truss = d.n_beam = 0;
  slabs[k] = d;
}

// relative cost of one contribution per family (truss, beam, plate), used to balance the per-lane work lists; the
// compile-time defaults can be overridden per process (FEMGPU_COST_T / _B / _P, tuning knob)
__device__ uint32_t g_cost[3] = {kCostTruss, kCostBeam, kCostPlate};

__device__ __forceinline__ uint32_t block_cost(const uint32_t* __restrict__ cptr,
                                               const uint32_t* __restrict__ contrib, uint32_t i) {
  uint32_t cost = 0;
  for (uint32_t c = cptr[i]; c < cptr[i + 1]; ++c) {
    uint32_t f = contrib[c] >> 30;
    cost += f < 3u ? g_cost[f] : 0u;
  }
  return min(cost, 65535u);
}

// sort key for the in-slab thread order: slab id, then descending cost of the block
__global__ void order_key_kernel(uint32_t n_blocks, uint32_t quota, int key_bits,
                                 const uint64_t* __restrict__ blk_key,
                                 const uint32_t* __restrict__ node_blk_ptr,
                                 const uint32_t* __restrict__ cptr,
                                 const uint32_t* __restrict__ contrib, uint64_t* __restrict__ okey,
                                 uint32_t* __restrict__ oval) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_blocks) return;
  uint32_t a = uint32_t(blk_key[i] >> key_bits);
  uint32_t slab = node_blk_ptr[a] / quota;
  okey[i] = (uint64_t(slab) << 16) | (65535u - block_cost(cptr, contrib, i));
  oval[i] = i;
}

__global__ void ordered_count_kernel(uint32_t n_blocks, const uint32_t* __restrict__ order,
                                     const uint32_t* __restrict__ cptr,
                                     const uint32_t* __restrict__ contrib, uint32_t* __restrict__ cnt,
                                     uint32_t* __restrict__ cost) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_blocks) return;
  uint32_t i = order[p];
  cnt[p] = cptr[i + 1] - cptr[i];
  cost[p] = block_cost(cptr, contrib, i);
}

// Work items: consecutive thread-ordered blocks of a slab are grouped greedily (blocks arrive in
// descending cost) so that every thread of the CTA gets about the same cost W. A block heavier
// than W (the diagonal block of a plate-grid node: 4 plate contributions, twice the per-thread
// share) is split into up to kMaxChunks chunks of consecutive contributions (execution order), one
// thread each, in adjacent lanes of one warp: after the contribution loop the lane of chunk 0 collects the
// partial sums of chunks 1, 2, .. with warp shuffles, in that order ("merge round" j), and stores the
// block. W starts at total / threads and is raised until the slab fits.
// Only blocks at least 25 % heavier than W are split (a split costs a merge round: 72 shuffles).
// One thread per slab; items are written to the dense [slab][thread] table.
__device__ __forceinline__ uint32_t family_cost(uint32_t f) {
  return f < 3u ? g_cost[f] : 0u;
}
// How the contributions of a block split over kk lanes are dealt to its chunks: every family
// (execution order: placeholders, plates, beams, trusses) is spread evenly, chunk j taking a
// consecutive run of the family's contributions — n / kk each, the n % kk left over going one by
// one to the chunks that are lightest so far (cost, then count, then index). The lanes of a warp
// run their entries family by family, so a warp's time is the sum over the families of the LARGEST
// per-lane count: a chunk holding "the second half" of a mixed block (2 plates + 2 beams + 2
// trusses) costs its warp six loop trips, two chunks of (2 plates + beam + truss) cost it four.
// cnt[g][j] = contributions of family group g in chunk j. Every chunk is non-empty when kk <= total.
__device__ __forceinline__ void chunk_family_counts(const uint32_t n[4], uint32_t kk, uint32_t cnt[4][kMaxChunks]) {
  const uint32_t cost[4] = {0u, g_cost[FEMGPU_PLATE], g_cost[FEMGPU_BEAM], g_cost[FEMGPU_TRUSS]};
  uint32_t load[kMaxChunks], items[kMaxChunks];
  for (uint32_t j = 0; j < uint32_t(kMaxChunks); ++j) load[j] = items[j] = 0;
  for (int g = 0; g < 4; ++g) {
    const uint32_t base = n[g] / kk, extra = n[g] % kk;
    uint32_t got = 0;  // bit j: chunk j already has one of this family's left-over contributions
    for (uint32_t j = 0; j < uint32_t(kMaxChunks); ++j) cnt[g][j] = j < kk ? base : 0u;
    for (uint32_t e = 0; e < extra; ++e) {
      uint32_t best = kk;
      for (uint32_t j = 0; j < kk; ++j) {
        if (got & (1u << j)) continue;
        if (best == kk || load[j] < load[best] || (load[j] == load[best] && items[j] < items[best])) best = j;
      }
      got |= 1u << best;
      cnt[g][best]++;
    }
    for (uint32_t j = 0; j < kk; ++j) {
      load[j] += cnt[g][j] * cost[g];
      items[j] += cnt[g][j];
    }
  }
}
__device__ __forceinline__ int family_group(uint32_t f) {
  return f == 3u ? 0 : (f == uint32_t(FEMGPU_PLATE) ? 1 : (f == uint32_t(FEMGPU_BEAM) ? 2 : 3));
}

__global__ void work_item_kernel(uint32_t n_slabs, uint32_t threads, SlabDesc* __restrict__ slabs,
                                 const uint32_t* __restrict__ cost, const uint32_t* __restrict__ cptr_ord,
                                 const uint32_t* __restrict__ contrib, WorkItem* __restrict__ items,
                                 int32_t* __restrict__ flags, const BlockMeta* __restrict__ meta, int spread_banks) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_slabs) return;
  const SlabDesc d = slabs[k];
  const uint32_t b0 = d.blk_begin, b1 = d.blk_begin + d.blk_count;
  const bool staged = !(d.flags & 1u);
  uint64_t total = 0;
  uint32_t mx = 1;
  for (uint32_t b = b0; b < b1; ++b) {
    total += cost[b];
    mx = max(mx, cost[b]);
  }
  uint64_t W = max(uint64_t(1), (total + threads - 1) / threads);
  if (!staged) W = max(W, uint64_t(mx));  // the unstaged kernel takes whole blocks only
  for (;;) {
    uint32_t n = 0;
    uint64_t cur = 0;
    bool open = false, ok = true;
    for (uint32_t b = b0; b < b1; ++b) {
      const uint32_t c = cost[b], cnt = cptr_ord[b + 1] - cptr_ord[b];
      if (staged && 4 * uint64_t(c) > 5 * W && cnt > 1) {
        const uint64_t kk = min(uint64_t(cnt), (c + W - 1) / W);
        if (kk > uint64_t(kMaxChunks)) ok = false;
        if ((n & 31u) + uint32_t(kk) > 32u) n = (n + 31u) & ~31u;  // chunks of a block stay inside one warp
        n += uint32_t(kk);
        open = false;
      } else {
        if (!open || cur + c > W) {
          ++n;
          cur = 0;
          open = true;
        }
        cur += c;
      }
    }
    if (ok && n <= threads) break;
    W += max(uint64_t(1), W / 8);
  }
  WorkItem* out = items + size_t(k) * threads;
  uint32_t n = 0, max_round = 0;
  uint64_t cur = 0;
  bool open = false;
  uint32_t start = b0;
  auto close_run = [&](uint32_t b_end) {
    if (open && b_end > start) {
      WorkItem w;
      w.blk_begin = start;
      w.blk_count = b_end - start;
      w.c_begin = cptr_ord[start];
      w.c_count = cptr_ord[b_end] - cptr_ord[start];
      out[n++] = w;
    }
    open = false;
  };
  for (uint32_t b = b0; b < b1; ++b) {
    const uint32_t c = cost[b], c0 = cptr_ord[b], cnt = cptr_ord[b + 1] - c0;
    if (staged && 4 * uint64_t(c) > 5 * W && cnt > 1) {
      close_run(b);
      const uint32_t kk = uint32_t(min(uint64_t(cnt), (c + W - 1) / W));
      // the chunks are merged with warp shuffles (lane of chunk 0 <- lane + j): keep them in one warp
      if ((n & 31u) + kk > 32u)
        while (n & 31u) out[n++] = WorkItem{0u, 0u, 0u, 0u};
      uint32_t nf[4] = {0, 0, 0, 0};  // placeholders, plates, beams, trusses
      for (uint32_t i = 0; i < cnt; ++i) nf[family_group(contrib[c0 + i] >> 30)]++;
      uint32_t fc[4][kMaxChunks];
      chunk_family_counts(nf, kk, fc);
      uint32_t s_prev = 0;
      for (uint32_t j = 0; j < kk; ++j) {
        const uint32_t take = fc[0][j] + fc[1][j] + fc[2][j] + fc[3][j];
        WorkItem w;
        w.blk_begin = b;
        w.blk_count = 1u | (j << 16) | ((kk - 1u) << 20) | (1u << 24);
        w.c_begin = c0 + s_prev;  // where the chunk's entries go; WHICH contributions: chunk_family_counts
        w.c_count = take;
        out[n++] = w;
        max_round = max(max_round, j);
        s_prev += take;
      }
    } else {
      if (!open || cur + c > W) {
        close_run(b);
        start = b;
        cur = 0;
        open = true;
      }
      cur += c;
    }
  }
  close_run(b1);
  for (; n < threads; ++n) out[n] = WorkItem{0u, 0u, 0u, 0u};
  // Lane order inside a warp. A block is flushed into the slab image with 16-byte stores, which the
  // shared-memory pipe serves a quarter-warp (8 lanes) at a time: conflict-free when the eight blocks start in
  // eight different 16-byte bank groups. Which group a block starts in is fixed by its place in the CSR rows
  // ((seg0 / 2) mod 8; the rows of a block shift every lane alike), and blocks of equal cost sit in
  // neighbouring lanes whatever their group: the diagonal blocks of eight grid nodes occupy four groups, two
  // each. The warp's time only depends on WHICH items it holds, not on their lanes, so the items are dealt to
  // the four quarter-warps greedily — the items that flush most often first, each to the quarter where it adds
  // the fewest wavefronts. The chunks of a split block stay in adjacent lanes (they are merged with shuffles).
  if (spread_banks && staged) {
    for (uint32_t w0 = 0; w0 < threads; w0 += 32) {
      WorkItem tmp[32];
      // a unit = one lane, or the adjacent lanes of a split block; its flush events = (loop trip, bank group)
      unsigned char u_first[32], u_len[32], u_nev[32], u_ev[32][8], order[32];
      uint32_t n_units = 0;
      for (uint32_t l = 0; l < 32; ++l) tmp[l] = out[w0 + l];
      for (uint32_t l = 0; l < 32;) {
        const WorkItem w = tmp[l];
        const uint32_t n_blk = w.blk_count & 0xFFFFu;
        const bool chunk = (w.blk_count & (1u << 24)) != 0;
        const uint32_t len = chunk ? ((w.blk_count >> 20) & 3u) + 1u : 1u;
        uint32_t n_ev = 0;
        if (chunk) {  // stored after the merge rounds, together with the other split blocks of the warp
          u_ev[n_units][n_ev++] = (unsigned char)((15u << 3) | ((meta[w.blk_begin].seg0 >> 1) & 7u));
        } else {
          for (uint32_t p = w.blk_begin; p < w.blk_begin + n_blk && n_ev < 8u; ++p)
            u_ev[n_units][n_ev++] =
                (unsigned char)((min(14u, cptr_ord[p + 1] - cptr_ord[w.blk_begin]) << 3) | ((meta[p].seg0 >> 1) & 7u));
        }
        u_first[n_units] = (unsigned char)l;
        u_len[n_units] = (unsigned char)len;
        u_nev[n_units] = (unsigned char)n_ev;
        order[n_units] = (unsigned char)n_units;
        ++n_units;
        l += len;
      }
      // units that flush most often are placed first (stable insertion sort)
      for (uint32_t i = 1; i < n_units; ++i) {
        const unsigned char u = order[i];
        uint32_t j = i;
        while (j > 0 && u_nev[order[j - 1]] < u_nev[u]) {
          order[j] = order[j - 1];
          --j;
        }
        order[j] = u;
      }
      unsigned char seen[4][16][8], peak[4][16], q_used[4] = {0, 0, 0, 0}, slot_of[32];
      for (int q = 0; q < 4; ++q)
        for (int t = 0; t < 16; ++t) {
          peak[q][t] = 0;
          for (int g = 0; g < 8; ++g) seen[q][t][g] = 0;
        }
      bool ok = true;
      for (uint32_t i = 0; i < n_units && ok; ++i) {
        const uint32_t u = order[i], len = u_len[u];
        int best = -1;
        uint32_t best_cost = 0xFFFFFFFFu, best_tie = 0xFFFFFFFFu;
        for (int q = 0; q < 4; ++q) {
          if (q_used[q] + len > 8u) continue;
          // wavefronts this unit adds to the quarter: an STS.128 of a quarter-warp costs as many wavefronts as
          // its most crowded bank group holds lanes
          uint32_t c = 0, tie = 0;
          for (uint32_t e = 0; e < u_nev[u]; ++e) {
            const uint32_t t = u_ev[u][e] >> 3, g = u_ev[u][e] & 7u, m = seen[q][t][g] + 1u;
            c += m > peak[q][t] ? 1u : 0u;
            tie += m;
          }
          if (c < best_cost || (c == best_cost && tie < best_tie)) {
            best = q;
            best_cost = c;
            best_tie = tie;
          }
        }
        if (best < 0) {
          ok = false;
          break;
        }
        for (uint32_t e = 0; e < u_nev[u]; ++e) {
          const uint32_t t = u_ev[u][e] >> 3, g = u_ev[u][e] & 7u;
          const unsigned char m = ++seen[best][t][g];
          if (m > peak[best][t]) peak[best][t] = m;
        }
        for (uint32_t j = 0; j < len; ++j) slot_of[u_first[u] + j] = (unsigned char)(8u * best + q_used[best] + j);
        q_used[best] = (unsigned char)(q_used[best] + len);
      }
      if (!ok) continue;  // a split block did not fit any quarter: keep the order of this warp
      for (uint32_t l = 0; l < 32; ++l) out[w0 + l] = WorkItem{0u, 0u, 0u, 0u};
      for (uint32_t l = 0; l < 32; ++l) out[w0 + slot_of[l]] = tmp[l];
    }
  }
  if (max_round) {
    slabs[k].flags = d.flags | (max_round << 8);
    atomicMax(flags + 13, 1);  // some staged slab splits a block: the kernel variant with merge rounds is needed
  }
}

// the work items as the staged kernel stages them: 4 bytes per lane
__global__ void compact_items_kernel(uint32_t n_slabs, uint32_t threads, const SlabDesc* __restrict__ slabs,
                                     const WorkItem* __restrict__ items, uint32_t* __restrict__ out) {
  uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t k = uint32_t(t / threads);
  if (k >= n_slabs) return;
  const WorkItem w = items[t];
  const SlabDesc d = slabs[k];
  const bool live = (w.blk_count & 0xFFFFu) != 0 && !(d.flags & 1u);  // staged slabs: < 2^16 entries (stage cap)
  out[t] = live ? ((w.c_begin - d.c_begin) & 0xFFFFu) | (w.c_count << 16) : 0u;
}

// ---- per-slab element lists -----------------------------------------------------------------
// key of a contribution: (slab << 28) | family << 26 | element; sorting + unique gives, per slab,
// the distinct elements whose records the CTA stages in shared memory.
__global__ void slab_elem_key_kernel(uint32_t n_blocks, uint32_t quota, int key_bits,
                                     const uint32_t* __restrict__ order,
                                     const uint64_t* __restrict__ blk_key,
                                     const uint32_t* __restrict__ node_blk_ptr,
                                     const uint32_t* __restrict__ cptr_ord,
                                     const uint32_t* __restrict__ contrib_ord,
                                     uint64_t* __restrict__ keys, uint32_t* __restrict__ payload) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_blocks) return;
  uint32_t a = uint32_t(blk_key[order[p]] >> key_bits);
  uint64_t slab = node_blk_ptr[a] / quota;
  for (uint32_t c = cptr_ord[p]; c < cptr_ord[p + 1]; ++c) {
    uint32_t code = contrib_ord[c];
    uint32_t fe = ((code >> 30) << 26) | (code & 0x03FFFFFFu);
    keys[c] = (slab << 28) | fe;
    payload[c] = c;
  }
}

__global__ void head_flag_kernel(uint32_t n, const uint64_t* __restrict__ keys, uint32_t* __restrict__ flag) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// uidx = inclusive scan of the head flags; element u = uidx - 1
__global__ void elist_kernel(uint32_t n, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ flag,
                             const uint32_t* __restrict__ uidx, uint64_t* __restrict__ ukey,
                             uint32_t* __restrict__ elist) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  uint32_t u = uidx[i] - 1;
  ukey[u] = keys[i];
  elist[u] = uint32_t(keys[i] & 0x0FFFFFFFu);
}

// bytes of a slab's double-buffered stage: block metadata, contribution entries (fetched from the
// 16-byte aligned address below the first one), truss and beam records
__host__ __device__ inline uint32_t slab_stage_bytes(uint32_t blk_count, uint32_t c_begin, uint32_t c_count,
                                                     uint32_t nt, uint32_t nb) {
  uint32_t ent = (((c_begin & 3u) + c_count + 1u) * 4u + 15u) & ~15u;  // +1: the loop peeks one entry ahead
  return blk_count * 16u + ent + nt * uint32_t(kTrussSlotDoubles * 8) + nb * uint32_t(kBeamSlotDoubles * 8);
}

__global__ void slab_elist_kernel(uint32_t n_slabs, uint32_t n_unique, const uint64_t* __restrict__ ukey,
                                  const uint32_t* __restrict__ cptr_ord, SlabDesc* __restrict__ slabs,
                                  int32_t* __restrict__ flags) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_slabs) return;
  // the list is sorted by family: [trusses][beams][plates][remote placeholders]
  uint32_t b = lower_bound_u64(ukey, n_unique, uint64_t(k) << 28);
  uint32_t b1 = lower_bound_u64(ukey, n_unique, (uint64_t(k) << 28) | (uint64_t(FEMGPU_BEAM) << 26));
  uint32_t b2 = lower_bound_u64(ukey, n_unique, (uint64_t(k) << 28) | (uint64_t(FEMGPU_PLATE) << 26));
  uint32_t b3 = lower_bound_u64(ukey, n_unique, (uint64_t(k) << 28) | (uint64_t(3) << 26));
  uint32_t e = lower_bound_u64(ukey, n_unique, uint64_t(k + 1) << 28);
  SlabDesc d = slabs[k];
  d.el_begin = b;
  d.el_count = e - b;
  d.n_truss = b1 - b;
  d.n_beam = b2 - b1;
  const uint32_t np = b3 - b2;
  d.c_begin = cptr_ord[d.blk_begin];
  d.c_count = cptr_ord[d.blk_begin + d.blk_count] - d.c_begin;
  const uint64_t img = (uint64_t(d.val_count) * 8 + 15) & ~uint64_t(15);
  const uint32_t form = np * uint32_t(kPlateSlotDoubles * 8), rawp = np * 160u;
  const uint32_t stage = slab_stage_bytes(d.blk_count, d.c_begin, d.c_count, d.n_truss, d.n_beam);
  if (img > uint64_t(kCapImgBytes) || form > uint32_t(kCapFormBytes) || stage > uint32_t(kCapStageBytes) ||
      d.el_count > uint32_t(kElistStride) || d.blk_count > uint32_t(kCapBlocks))
    d.flags |= 1u;
  if (d.flags & 1u) {
    atomicAdd(flags + 12, 1);
  } else {  // integer maxima: order independent
    d.flags |= np << 16;
    atomicMax(flags + 8, int32_t(img));
    atomicMax(flags + 9, int32_t(form));
    atomicMax(flags + 10, int32_t(rawp));
    atomicMax(flags + 11, int32_t(stage));
  }
  slabs[k] = d;
}

// dense [slab][kElistStride] copy of the compact lists: addressable from the slab id alone
__global__ void elist_table_kernel(uint32_t n_slabs, const SlabDesc* __restrict__ slabs,
                                   const uint32_t* __restrict__ compact, uint32_t* __restrict__ table) {
  uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t k = uint32_t(t / kElistStride), e = uint32_t(t % kElistStride);
  if (k >= n_slabs) return;
  const SlabDesc d = slabs[k];
  table[t] = (e < d.el_count && !(d.flags & 1u)) ? compact[d.el_begin + e] : 0xFFFFFFFFu;
}

// How long are the runs of consecutively numbered elements in the slabs' (sorted) element lists? One thread per
// table entry counts the record bytes in 16-byte units (flags[14]) and the run heads (flags[15]) exactly as the
// bulk staging of the assembly kernel delimits them (a run also ends at a warp's 32 slots). Integer sums: order
// independent.
__global__ void run_stat_kernel(uint64_t n, const uint32_t* __restrict__ table, int32_t* __restrict__ flags) {
  uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  bool copy = false, head = false;
  if (t < n) {
    const uint32_t fe = table[t], slot = uint32_t(t % kElistStride);
    copy = fe != 0xFFFFFFFFu && (fe >> 26) < 3u;
    if (copy) {
      const uint32_t prev = (slot & 31u) ? table[t - 1] : 0xFFFFFFFFu;
      head = (slot & 31u) == 0u || fe != prev + 1u || (prev >> 26) != (fe >> 26);
    }
  }
  uint32_t units = 0;
  if (copy) {
    const uint32_t family = table[t] >> 26;
    units = uint32_t(family == FEMGPU_PLATE ? kPlateRawDoubles : (family == FEMGPU_BEAM ? kBeamSlotDoubles : kTrussSlotDoubles)) / 2u;
  }
  units = __reduce_add_sync(0xFFFFFFFFu, units);
  const uint32_t nh = __popc(__ballot_sync(0xFFFFFFFFu, head));
  if ((threadIdx.x & 31u) == 0u && units) {
    atomicAdd(flags + 14, int32_t(units >> 3));  // 128-byte units: 2^31 of them = 256 GB of records
    atomicAdd(flags + 15, int32_t(nh));
  }
}

// contrib codes become family<<30 | pair<<26 | where the element's record sits: its offset inside
// the family's record area in 16-byte units (staged slabs), or its slot in the slab's element list
__global__ void relabel_kernel(uint32_t n, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ payload,
                               const uint32_t* __restrict__ uidx, const SlabDesc* __restrict__ slabs,
                               uint32_t* __restrict__ contrib_ord) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t slab = uint32_t(keys[i] >> 28);
  const SlabDesc d = slabs[slab];
  uint32_t local = (uidx[i] - 1) - d.el_begin;
  uint32_t c = payload[i];
  const uint32_t code = contrib_ord[c], family = code >> 30;
  if (!(d.flags & 1u)) {
    if (family == FEMGPU_TRUSS) local = local * uint32_t(kTrussSlotDoubles / 2);
    else if (family == FEMGPU_BEAM) local = (local - d.n_truss) * uint32_t(kBeamSlotDoubles / 2);
    else if (family == FEMGPU_PLATE) local = (local - d.n_truss - d.n_beam) * uint32_t(kPlateSlotDoubles / 2);
    else local = 0;
  }
  contrib_ord[c] = (code & 0xFC000000u) | local;
}

// Staged slabs: rewrite each lane's contributions from block-major order into the lane's execution
// order — family-major (placeholders, plates, beams, trusses) over the lane's blocks, insertion order
// inside a (block, family) group — and attach the group-end / read-modify-write / merge-round
// flags and the block index. One thread per work item; entries stay inside the item's own range.
__global__ void item_program_kernel(uint32_t n_slabs, uint32_t threads, const SlabDesc* __restrict__ slabs,
                                    const WorkItem* __restrict__ items, const uint32_t* __restrict__ cptr_ord,
                                    const uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
  uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t k = uint32_t(t / threads);
  if (k >= n_slabs) return;
  const WorkItem w = items[t];
  const uint32_t n_blk = w.blk_count & 0xFFFFu;
  if (n_blk == 0) return;
  const SlabDesc d = slabs[k];
  if (d.flags & 1u) {
    for (uint32_t c = w.c_begin; c < w.c_begin + w.c_count; ++c) dst[c] = src[c];
    return;
  }
  uint32_t out = w.c_begin;
  const uint32_t order[4] = {3u, uint32_t(FEMGPU_PLATE), uint32_t(FEMGPU_BEAM), uint32_t(FEMGPU_TRUSS)};
  if ((w.blk_count & (1u << 24)) || n_blk == 1) {
    // one chunk of a split block — or a whole block that is the item's only one (= the single chunk
    // of an unsplit block): its share of every family (chunk_family_counts), accumulated as a single
    // group (no flush between the families)
    const uint32_t p = w.blk_begin, c0 = cptr_ord[p], c1 = cptr_ord[p + 1];
    const bool is_chunk = (w.blk_count & (1u << 24)) != 0;
    const uint32_t round = is_chunk ? (w.blk_count >> 16) & 3u : 0u;
    const uint32_t more = is_chunk ? (w.blk_count >> 20) & 3u : 0u;  // chunks after chunk 0
    const uint32_t blk = p - d.blk_begin;
    uint32_t nf[4] = {0, 0, 0, 0};
    for (uint32_t c = c0; c < c1; ++c) nf[family_group(src[c] >> 30)]++;
    uint32_t fc[4][kMaxChunks];
    chunk_family_counts(nf, more + 1u, fc);
    uint32_t left = w.c_count;
    for (int f = 0; f < 4; ++f) {
      uint32_t lo = 0;
      for (uint32_t j = 0; j < round; ++j) lo += fc[f][j];
      const uint32_t hi = lo + fc[f][round];
      uint32_t pos = 0;
      for (uint32_t c = c0; c < c1; ++c) {
        const uint32_t code = src[c];
        if ((code >> 30) != order[f]) continue;
        if (pos >= lo && pos < hi) {
          uint32_t e = (code & 0xFC000000u) | (blk << kEntBlkShift) | (code & kEntRecMask);
          // chunk j >= 1: sender of merge round j; chunk 0: holds its flush and receives `more` rounds
          if (--left == 0) e |= kEntEnd | ((round ? round : more) << kEntDeferShift) | (round ? kEntRmw : 0u);
          dst[out++] = e;
        }
        ++pos;
      }
    }
    return;
  }
  for (int f = 0; f < 4; ++f) {
    const uint32_t fam = order[f];
    for (uint32_t p = w.blk_begin; p < w.blk_begin + n_blk; ++p) {
      const uint32_t c0 = cptr_ord[p], c1 = cptr_ord[p + 1];
      uint32_t in_group = 0, earlier = 0;
      for (uint32_t c = c0; c < c1; ++c) {
        const uint32_t cf = src[c] >> 30;
        if (cf == fam) ++in_group;
        else {
          int pos = 0;  // does family cf come before fam in the execution order?
          for (int g = 0; g < 4; ++g) if (order[g] == cf) pos = g;
          if (pos < f) ++earlier;
        }
      }
      if (!in_group) continue;
      const uint32_t blk = p - d.blk_begin;
      uint32_t left = in_group;
      for (uint32_t c = c0; c < c1; ++c) {
        const uint32_t code = src[c];
        if ((code >> 30) != fam) continue;
        uint32_t e = (code & 0xFC000000u) | (blk << kEntBlkShift) | (code & kEntRecMask);
        if (--left == 0) e |= kEntEnd | (earlier ? kEntRmw : 0u);
        dst[out++] = e;
      }
    }
  }
}

__global__ void ordered_meta_kernel(uint32_t n_blocks, uint32_t quota, int key_bits,
                                    const uint32_t* __restrict__ order,
                                    const uint64_t* __restrict__ blk_key,
                                    const uint32_t* __restrict__ node_blk_ptr,
                                    const uint32_t* __restrict__ blk_off,
                                    const uint32_t* __restrict__ node_len,
                                    const int64_t* __restrict__ node_base,
                                    const SlabDesc* __restrict__ slabs,
                                    const uint32_t* __restrict__ cptr_sorted,
                                    const uint32_t* __restrict__ contrib_sorted,
                                    const uint32_t* __restrict__ cptr_ord,
                                    uint32_t* __restrict__ contrib_ord, BlockMeta* __restrict__ meta) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_blocks) return;
  uint32_t i = order[p];
  uint32_t a = uint32_t(blk_key[i] >> key_bits);
  uint32_t slab = node_blk_ptr[a] / quota;
  uint32_t rel = uint32_t(node_base[a] - slabs[slab].val_base);
  uint32_t l03 = node_len[2 * a], l35 = node_len[2 * a + 1];
  uint32_t o03 = blk_off[2 * i], o35 = blk_off[2 * i + 1];
  BlockMeta m;
  m.seg0 = rel + o03;
  m.seg3 = (o35 == 0xFFFFFFFFu) ? 0xFFFFFFFFu : rel + 3 * l03 + o35;
  m.strides = l03 | (l35 << 16);
  m.count = cptr_ord[p + 1] - cptr_ord[p];
  meta[p] = m;
  uint32_t src = cptr_sorted[i], n = cptr_sorted[i + 1] - src, dst = cptr_ord[p];
  for (uint32_t c = 0; c < n; ++c) contrib_ord[dst + c] = contrib_sorted[src + c];
}

// ---- range plan (see Handle::n_ranges) ----
// first range whose slabs list the element; one thread per slab walks its compact element list
__global__ void first_range_kernel(uint32_t n_slabs, uint32_t slabs_per_range, const SlabDesc* __restrict__ slabs,
                                   const uint32_t* __restrict__ elist_compact, uint32_t* __restrict__ first_t,
                                   uint32_t* __restrict__ first_b, uint32_t* __restrict__ first_p) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_slabs) return;
  const SlabDesc d = slabs[k];
  const uint32_t r = k / slabs_per_range;
  for (uint32_t i = 0; i < d.el_count; ++i) {
    const uint32_t fe = elist_compact[d.el_begin + i], family = fe >> 26, e = fe & 0x03FFFFFFu;
    if (family == FEMGPU_TRUSS) atomicMin(first_t + e, r);       // integer min: order independent
    else if (family == FEMGPU_BEAM) atomicMin(first_b + e, r);
    else if (family == FEMGPU_PLATE) atomicMin(first_p + e, r);
  }
}
__global__ void iota_u32_kernel(uint32_t n, uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}
// ends[r] = number of sorted keys <= r (keys of unreferenced elements were clamped to the last range)
__global__ void range_ends_kernel(uint32_t n, const uint32_t* __restrict__ sorted_keys, uint32_t n_ranges,
                                  uint32_t* __restrict__ ends) {
  uint32_t r = threadIdx.x;
  if (r >= n_ranges) return;
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (min(sorted_keys[mid], n_ranges - 1u) <= r) lo = mid + 1;
    else hi = mid;
  }
  ends[r] = lo;
}

__global__ void element_slots_kernel(int n_nodes_elem, int dof, const uint32_t* __restrict__ nodes,
                                     uint32_t n_blocks, int key_bits,
                                     const uint64_t* __restrict__ blk_key,
                                     const uint32_t* __restrict__ blk_off,
                                     const uint32_t* __restrict__ node_len,
                                     const int64_t* __restrict__ node_base, int64_t* __restrict__ out) {
  int n = n_nodes_elem * dof;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * n) return;
  int row = t / n, col = t % n;
  int la = row / dof, i = row % dof, lb = col / dof, j = col % dof;
  uint32_t a = nodes[la], b = nodes[lb];
  uint64_t key = (uint64_t(a) << key_bits) | b;
  uint32_t blk = lower_bound_u64(blk_key, n_blocks, key);
  int64_t slot = -1;
  if (blk < n_blocks && blk_key[blk] == key) {
    uint32_t l03 = node_len[2 * a], l35 = node_len[2 * a + 1];
    uint32_t o03 = blk_off[2 * blk], o35 = blk_off[2 * blk + 1];
    bool full = o35 != 0xFFFFFFFFu;
    if (i < 3 && (j < 3 || full)) slot = node_base[a] + int64_t(i) * l03 + o03 + j;
    if (i >= 3 && full) slot = node_base[a] + 3 * int64_t(l03) + int64_t(i - 3) * l35 + o35 + j;
  }
  out[t] = slot;
}

__global__ void nz_flag_kernel(int64_t nnz, const double* __restrict__ values, int64_t* __restrict__ flag) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < nnz) flag[i] = values[i] != 0.0 ? 1 : 0;
}

__global__ void nz_write_kernel(int64_t nnz, int64_t n_rows, const double* __restrict__ values,
                                const int64_t* __restrict__ pos, const int64_t* __restrict__ row_ptr,
                                const int32_t* __restrict__ col_idx, int64_t* __restrict__ rows,
                                int64_t* __restrict__ cols, double* __restrict__ vals) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  double v = values[i];
  if (v == 0.0) return;
  // row = last r with row_ptr[r] <= i
  int64_t lo = 0, hi = n_rows;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (row_ptr[mid] <= i) lo = mid; else hi = mid;
  }
  int64_t p = pos[i];
  rows[p] = lo;
  cols[p] = col_idx[i];
  vals[p] = v;
}

// ---- compaction to the entries != 0.0, CSR in place of the COO above (femgpu_get_nonzero_csr) ----
// Rows of a structural model are short (18-54 stored entries): eight lanes share a row.
constexpr int kNzLanes = 8;
__global__ void __launch_bounds__(256)
nz_row_count_kernel(int64_t row_begin, int64_t row_end, const int64_t* __restrict__ row_ptr,
                    const double* __restrict__ values, int32_t* __restrict__ cnt /* [n_rows + 1], zeroed */) {
  const int64_t row = row_begin + (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) / kNzLanes;
  const uint32_t sub = threadIdx.x & (kNzLanes - 1);
  int c = 0;
  if (row < row_end) {
    const int64_t b = row_ptr[row], e = row_ptr[row + 1];
    for (int64_t i = b + sub; i < e; i += kNzLanes) c += values[i] != 0.0 ? 1 : 0;
  }
#pragma unroll
  for (int d = kNzLanes / 2; d > 0; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
  if (row < row_end && sub == 0) cnt[row] = c;
}
__global__ void __launch_bounds__(256)
nz_row_fill_kernel(int64_t row_begin, int64_t row_end, const int64_t* __restrict__ row_ptr,
                   const int32_t* __restrict__ col_idx, const double* __restrict__ values,
                   const int64_t* __restrict__ nz_ptr, int32_t* __restrict__ out_col, double* __restrict__ out_val) {
  const int64_t row = row_begin + (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) / kNzLanes;
  const uint32_t lane = threadIdx.x & 31u, sub = lane & (kNzLanes - 1), shift = lane & ~uint32_t(kNzLanes - 1);
  const bool live = row < row_end;
  const int64_t b = live ? row_ptr[row] : 0, e = live ? row_ptr[row + 1] : 0;
  int64_t w = live ? nz_ptr[row] : 0;
  // the four row groups of a warp advance together (ballots are warp-wide): as many steps as the longest row needs
  int64_t steps = (e - b + kNzLanes - 1) / kNzLanes;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) steps = max(steps, __shfl_xor_sync(0xFFFFFFFFu, steps, d));
  for (int64_t k = 0; k < steps; ++k) {
    const int64_t i = b + k * kNzLanes + sub;
    const double v = i < e ? values[i] : 0.0;
    const bool keep = v != 0.0;
    const uint32_t mask = (__ballot_sync(0xFFFFFFFFu, keep) >> shift) & ((1u << kNzLanes) - 1u);
    if (keep) {
      const int64_t p = w + __popc(mask & ((1u << sub) - 1u));
      out_col[p] = col_idx[i];
      out_val[p] = v;
    }
    w += __popc(mask);
  }
}

// placeholder contributions for blocks other ranks will add to (multi-GPU): they only create the
// slot; the assembly kernel skips family 3
__global__ void remote_contrib_kernel(uint32_t n, const uint64_t* __restrict__ remote_keys, int64_t base,
                                      uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = remote_keys[i];
  keys[base + i] = k & ~(uint64_t(1) << 63);
  vals[base + i] = (3u << 30) | (uint32_t(k >> 63) << 26);
}

__global__ void ghost_key_kernel(uint32_t n, const uint64_t* __restrict__ blk_key, const uint8_t* __restrict__ full,
                                 uint64_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = blk_key[i] | (uint64_t(full[i]) << 63);
}

// bounds[2r], bounds[2r+1] = first / last+1 block whose row node lies in rank r's node range
__global__ void rank_bounds_kernel(int world, const int64_t* __restrict__ ranges, int key_bits,
                                   const uint64_t* __restrict__ blk_key, uint32_t n_blocks,
                                   int64_t* __restrict__ bounds) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * world) return;
  bounds[t] = lower_bound_u64(blk_key, n_blocks, uint64_t(ranges[t]) << key_bits);
}

__global__ void ghost_dst_kernel(uint32_t n, const uint64_t* __restrict__ remote_keys,
                                 const uint64_t* __restrict__ blk_key, uint32_t n_blocks,
                                 uint32_t* __restrict__ dst, uint8_t* __restrict__ full, int32_t* __restrict__ bad) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = remote_keys[i] & ~(uint64_t(1) << 63);
  uint32_t b = lower_bound_u64(blk_key, n_blocks, k);
  if (b >= n_blocks || blk_key[b] != k) *bad = 1;
  dst[i] = b;
  full[i] = uint8_t(remote_keys[i] >> 63);
}

struct BlockBuild {
  Tmp keys_a, keys_b, vals_a, vals_b, uniq, counts, nruns, cptr_sorted;
  uint32_t nblk = 0;
  int64_t total = 0;  // local contributions + remote placeholders
  void release() {
    keys_a.release(); keys_b.release(); vals_a.release(); vals_b.release();
    uniq.release(); counts.release(); nruns.release(); cptr_sorted.release();
    nblk = 0;
    total = 0;
  }
};

}  // namespace

#define SYM_CHECK(expr) FEMGPU_CUDA_CHECK(h, (expr))

// steps 1-3a: contributions in global insertion order (+ remote placeholders), stable sort by
// (row node, col node), unique blocks with contribution ranges, block kinds
static int32_t build_blocks(Handle* h, int64_t n_extra, BlockBuild& bb) {
  cudaStream_t s = h->stream;
  const int kb = h->key_bits;
  const int64_t NC = h->n_contrib;
  const int64_t T = NC + n_extra;
  bb.total = T;
  SYM_CHECK(bb.keys_a.alloc(size_t(T) * 8));
  SYM_CHECK(bb.keys_b.alloc(size_t(T) * 8));
  SYM_CHECK(bb.vals_a.alloc(size_t(T) * 4));
  SYM_CHECK(bb.vals_b.alloc(size_t(T) * 4));
  for (int f = 0; f < kFamilies; ++f) {
    FamilyDev& fd = h->fd[f];
    size_t n = h->fh[f].size();
    if (!n) continue;
    uint64_t threads = uint64_t(n) * kPairsPerElem[f];
    uint32_t grid = div_up(threads, 256);
    if (f == FEMGPU_PLATE)
      gen_contrib_kernel<4><<<grid, 256, 0, s>>>(uint32_t(n), f, fd.conn[0].p, fd.conn[1].p, fd.conn[2].p,
                                                 fd.conn[3].p, fd.cbase.p, kb, bb.keys_a.as<uint64_t>(),
                                                 bb.vals_a.as<uint32_t>());
    else
      gen_contrib_kernel<2><<<grid, 256, 0, s>>>(uint32_t(n), f, fd.conn[0].p, fd.conn[1].p, nullptr, nullptr,
                                                 fd.cbase.p, kb, bb.keys_a.as<uint64_t>(), bb.vals_a.as<uint32_t>());
    h->launches++;
  }
  if (n_extra) {
    remote_contrib_kernel<<<div_up(n_extra, 256), 256, 0, s>>>(uint32_t(n_extra), h->dist.remote_keys.p, NC,
                                                               bb.keys_a.as<uint64_t>(), bb.vals_a.as<uint32_t>());
    h->launches++;
  }
  SYM_CHECK(cudaGetLastError());
  {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, bb.keys_a.as<uint64_t>(), bb.keys_b.as<uint64_t>(),
                                    bb.vals_a.as<uint32_t>(), bb.vals_b.as<uint32_t>(), int(T), 0, 2 * kb, s);
    Tmp t;
    SYM_CHECK(t.alloc(tb));
    SYM_CHECK(cub::DeviceRadixSort::SortPairs(t.p, tb, bb.keys_a.as<uint64_t>(), bb.keys_b.as<uint64_t>(),
                                              bb.vals_a.as<uint32_t>(), bb.vals_b.as<uint32_t>(), int(T), 0,
                                              2 * kb, s));
    SYM_CHECK(cudaStreamSynchronize(s));
  }
  SYM_CHECK(bb.uniq.alloc(size_t(T) * 8));
  SYM_CHECK(bb.counts.alloc((size_t(T) + 1) * 4));
  SYM_CHECK(bb.nruns.alloc(16));
  {
    size_t tb = 0;
    cub::DeviceRunLengthEncode::Encode(nullptr, tb, bb.keys_b.as<uint64_t>(), bb.uniq.as<uint64_t>(),
                                       bb.counts.as<uint32_t>(), bb.nruns.as<uint32_t>(), int(T), s);
    Tmp t;
    SYM_CHECK(t.alloc(tb));
    SYM_CHECK(cub::DeviceRunLengthEncode::Encode(t.p, tb, bb.keys_b.as<uint64_t>(), bb.uniq.as<uint64_t>(),
                                                 bb.counts.as<uint32_t>(), bb.nruns.as<uint32_t>(), int(T), s));
    SYM_CHECK(cudaStreamSynchronize(s));
  }
  uint32_t nblk = 0;
  SYM_CHECK(cudaMemcpy(&nblk, bb.nruns.p, 4, cudaMemcpyDeviceToHost));
  bb.nblk = nblk;
  h->n_blocks = nblk;
  SYM_CHECK(h->blk_key.reserve(nblk));
  SYM_CHECK(cudaMemcpyAsync(h->blk_key.p, bb.uniq.p, size_t(nblk) * 8, cudaMemcpyDeviceToDevice, s));
  SYM_CHECK(bb.cptr_sorted.alloc((size_t(nblk) + 1) * 4));
  {
    // exclusive scan over nblk+1 items (the extra item makes the last entry the total)
    SYM_CHECK(cudaMemsetAsync(bb.counts.as<uint32_t>() + nblk, 0, 4, s));
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, bb.counts.as<uint32_t>(), bb.cptr_sorted.as<uint32_t>(), int(nblk + 1), s);
    Tmp t;
    SYM_CHECK(t.alloc(tb));
    SYM_CHECK(cub::DeviceScan::ExclusiveSum(t.p, tb, bb.counts.as<uint32_t>(), bb.cptr_sorted.as<uint32_t>(),
                                            int(nblk + 1), s));
    SYM_CHECK(cudaStreamSynchronize(s));
  }
  SYM_CHECK(h->blk_full.reserve(nblk));
  block_kind_kernel<<<div_up(nblk, 256), 256, 0, s>>>(nblk, bb.cptr_sorted.as<uint32_t>(), bb.vals_b.as<uint32_t>(),
                                                      h->blk_full.p);
  h->launches++;
  SYM_CHECK(cudaGetLastError());
  return 0;
}

// Multi-GPU, symbolic half (collective): send the keys of the blocks in rows this rank does not own
// to the owning ranks; what arrives becomes placeholder contributions of the second build.
static int32_t dist_collect_ghost_keys(Handle* h, const BlockBuild& bb, int64_t* n_extra,
                                       std::vector<int64_t>& ranges) {
  DistState& D = h->dist;
  cudaStream_t s = h->stream;
  const int W = D.world;
  if (!D.ownership_set) return h->fail(FEMGPU_ERR_USAGE, "femgpu_dist_set_ownership was not called");
  // everyone's node ranges
  int64_t mine[2] = {int64_t(D.own_begin), int64_t(D.own_end)};
  ranges.assign(size_t(2) * W, 0);
  int32_t st = dist_allgather_i64(h, mine, ranges.data(), 2);
  if (st) return st;
  // per destination rank: the contiguous run of my (sorted) blocks whose row node it owns
  std::vector<int64_t> bounds(size_t(2) * W, 0);
  Tmp d_ranges, d_bounds, sendkeys;
  if (bb.nblk) {
    SYM_CHECK(d_ranges.alloc(size_t(2) * W * 8));
    SYM_CHECK(d_bounds.alloc(size_t(2) * W * 8));
    SYM_CHECK(cudaMemcpyAsync(d_ranges.p, ranges.data(), size_t(2) * W * 8, cudaMemcpyHostToDevice, s));
    rank_bounds_kernel<<<1, 2 * W, 0, s>>>(W, d_ranges.as<int64_t>(), h->key_bits, h->blk_key.p, bb.nblk,
                                           d_bounds.as<int64_t>());
    SYM_CHECK(cudaMemcpyAsync(bounds.data(), d_bounds.p, size_t(2) * W * 8, cudaMemcpyDeviceToHost, s));
    SYM_CHECK(cudaStreamSynchronize(s));
    SYM_CHECK(sendkeys.alloc(size_t(bb.nblk) * 8));
    ghost_key_kernel<<<div_up(bb.nblk, 256), 256, 0, s>>>(bb.nblk, h->blk_key.p, h->blk_full.p, sendkeys.as<uint64_t>());
    h->launches += 2;
  }
  std::vector<int64_t> send_counts(W, 0), send_offs(W, 0);
  for (int r = 0; r < W; ++r) {
    if (r == D.rank) continue;
    send_offs[r] = bounds[2 * r];
    send_counts[r] = bounds[2 * r + 1] - bounds[2 * r];
  }
  // count matrix: row = sender
  std::vector<int64_t> matrix(size_t(W) * W, 0);
  if ((st = dist_allgather_i64(h, send_counts.data(), matrix.data(), W))) return st;
  std::vector<int64_t> recv_counts(W, 0), recv_offs(W + 1, 0);
  for (int r = 0; r < W; ++r) recv_counts[r] = (r == D.rank) ? 0 : matrix[size_t(r) * W + D.rank];
  for (int r = 0; r < W; ++r) recv_offs[r + 1] = recv_offs[r] + recv_counts[r];
  const int64_t n_recv = recv_offs[W];
  SYM_CHECK(D.remote_keys.reserve(size_t(n_recv) + 1));
  if ((st = dist_exchange_8(h, sendkeys.p, send_offs.data(), send_counts.data(), D.remote_keys.p, recv_offs.data(),
                            recv_counts.data())))
    return st;
  SYM_CHECK(cudaStreamSynchronize(s));
  D.count_matrix = matrix;
  D.send_blocks = send_counts;
  D.recv_blocks = recv_counts;
  D.recv_off = recv_offs;
  D.send_off.assign(W + 1, 0);
  for (int r = 0; r < W; ++r) D.send_off[r + 1] = D.send_off[r] + send_counts[r];
  *n_extra = n_recv;
  return 0;
}

// Multi-GPU, after the final pattern exists: where my ghost blocks sit, where received ones land
static int32_t dist_finalize_plan(Handle* h, const std::vector<int64_t>& ranges) {
  DistState& D = h->dist;
  cudaStream_t s = h->stream;
  const int W = D.world;
  std::vector<int64_t> bounds(size_t(2) * W, 0);
  if (h->n_blocks) {
    Tmp d_ranges, d_bounds;
    SYM_CHECK(d_ranges.alloc(size_t(2) * W * 8));
    SYM_CHECK(d_bounds.alloc(size_t(2) * W * 8));
    SYM_CHECK(cudaMemcpyAsync(d_ranges.p, ranges.data(), size_t(2) * W * 8, cudaMemcpyHostToDevice, s));
    rank_bounds_kernel<<<1, 2 * W, 0, s>>>(W, d_ranges.as<int64_t>(), h->key_bits, h->blk_key.p, h->n_blocks,
                                           d_bounds.as<int64_t>());
    SYM_CHECK(cudaMemcpyAsync(bounds.data(), d_bounds.p, size_t(2) * W * 8, cudaMemcpyDeviceToHost, s));
    SYM_CHECK(cudaStreamSynchronize(s));
    h->launches++;
  }
  D.send_first_block.assign(W, 0);
  for (int r = 0; r < W; ++r) {
    if (r == D.rank) continue;
    D.send_first_block[r] = bounds[2 * r];
    if (bounds[2 * r + 1] - bounds[2 * r] != D.send_blocks[r])
      return h->fail(FEMGPU_ERR_USAGE, "internal: ghost block count changed between the two symbolic builds");
  }
  SYM_CHECK(D.send_buf.reserve(size_t(D.send_off[W]) * 36 + 1));
  SYM_CHECK(D.recv_buf.reserve(size_t(D.recv_off[W]) * 36 + 1));
  const int64_t n_recv = D.recv_off[W];
  SYM_CHECK(D.recv_dst_block.reserve(size_t(n_recv) + 1));
  SYM_CHECK(D.recv_full.reserve(size_t(n_recv) + 1));
  if (n_recv) {
    SYM_CHECK(cudaMemsetAsync(h->d_flag.p + 4, 0, 4, s));
    ghost_dst_kernel<<<div_up(n_recv, 256), 256, 0, s>>>(uint32_t(n_recv), D.remote_keys.p, h->blk_key.p, h->n_blocks,
                                                         D.recv_dst_block.p, D.recv_full.p, h->d_flag.p + 4);
    h->launches++;
    int32_t bad = 0;
    SYM_CHECK(cudaMemcpy(&bad, h->d_flag.p + 4, 4, cudaMemcpyDeviceToHost));
    if (bad) return h->fail(FEMGPU_ERR_USAGE, "internal: a received ghost block has no slot in the owner's pattern");
  }
  return dist_setup_p2p(h);
}

// Cut the slabs into consecutive ranges and list every family's elements by the first range that needs them.
static int32_t build_range_plan(Handle* h) {
  cudaStream_t s = h->stream;
  const uint32_t n_slabs = h->n_slabs;
  // Off by default (one range = the element records of all elements first, then one assembly launch). Measured on
  // B200 (profiles/README.md, round 2): the record kernels running in the SM resources the assembly kernel leaves
  // free (one 256-thread block per SM next to 4 x 54 KB assembly CTAs) are ~25 % slower than the assembly of the same
  // range, so the pipeline is bound by the records again and every extra launch adds its ramp: M 3.93 ms with one
  // range, 4.01 / 4.02 / 4.04 / 4.12 ms with 2 / 3 / 4 / 8. FEMGPU_NUMERIC_RANGES=n turns it on.
  int want = 1;
  if (const char* q = getenv("FEMGPU_NUMERIC_RANGES")) want = atoi(q);
  want = std::max(1, std::min(want, Handle::kMaxRanges));
  if (uint32_t(want) > n_slabs) want = int(std::max<uint32_t>(1u, n_slabs));
  h->n_ranges = want;
  const uint32_t per = div_up(n_slabs, uint32_t(want));
  for (int r = 0; r < want; ++r) h->range_slab_end[r] = std::min<uint64_t>(n_slabs, uint64_t(per) * (r + 1));
  h->range_slab_end[want - 1] = n_slabs;
  if (want == 1) {
    for (int f = 0; f < kFamilies; ++f) h->range_elem_end[f][0] = uint32_t(h->fh[f].size());
    return 0;
  }
  Tmp first[kFamilies];
  for (int f = 0; f < kFamilies; ++f) {
    const size_t n = h->fh[f].size();
    SYM_CHECK(first[f].alloc((n + 1) * 4));
    SYM_CHECK(cudaMemsetAsync(first[f].p, 0xFF, (n + 1) * 4, s));
  }
  first_range_kernel<<<div_up(n_slabs, 128), 128, 0, s>>>(n_slabs, per, h->slabs.p, h->elist_compact.p,
                                                           first[0].as<uint32_t>(), first[1].as<uint32_t>(),
                                                           first[2].as<uint32_t>());
  h->launches++;
  SYM_CHECK(cudaGetLastError());
  Tmp ends;
  SYM_CHECK(ends.alloc(size_t(kFamilies) * Handle::kMaxRanges * 4));
  for (int f = 0; f < kFamilies; ++f) {
    const uint32_t n = uint32_t(h->fh[f].size());
    if (n == 0) {
      for (int r = 0; r < want; ++r) h->range_elem_end[f][r] = 0;
      continue;
    }
    SYM_CHECK(h->prep_order[f].reserve(n));
    Tmp keys_out, idx_in;
    SYM_CHECK(keys_out.alloc(size_t(n) * 4));
    SYM_CHECK(idx_in.alloc(size_t(n) * 4));
    iota_u32_kernel<<<div_up(n, 256), 256, 0, s>>>(n, idx_in.as<uint32_t>());
    h->launches++;
    // stable sort by first range: inside a range the elements keep their index order (coalesced property loads);
    // 32 key bits so that the 0xFFFFFFFF of an unreferenced element sorts last
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, first[f].as<uint32_t>(), keys_out.as<uint32_t>(), idx_in.as<uint32_t>(),
                                    h->prep_order[f].p, int(n), 0, 32, s);
    Tmp t;
    SYM_CHECK(t.alloc(tb));
    SYM_CHECK(cub::DeviceRadixSort::SortPairs(t.p, tb, first[f].as<uint32_t>(), keys_out.as<uint32_t>(),
                                              idx_in.as<uint32_t>(), h->prep_order[f].p, int(n), 0, 32, s));
    range_ends_kernel<<<1, 32, 0, s>>>(n, keys_out.as<uint32_t>(), uint32_t(want),
                                       ends.as<uint32_t>() + f * Handle::kMaxRanges);
    h->launches++;
    SYM_CHECK(cudaMemcpyAsync(h->range_elem_end[f], ends.as<uint32_t>() + f * Handle::kMaxRanges, size_t(want) * 4,
                              cudaMemcpyDeviceToHost, s));
    SYM_CHECK(cudaStreamSynchronize(s));
    h->range_elem_end[f][want - 1] = n;
  }
  // nothing to overlap when the first range already needs (almost) every record: one range then
  uint64_t first_share = 0, total = 0;
  for (int f = 0; f < kFamilies; ++f) {
    first_share += h->range_elem_end[f][0];
    total += h->fh[f].size();
  }
  if (!getenv("FEMGPU_NUMERIC_RANGES") && first_share * 10 > total * 6) {
    h->n_ranges = 1;
    h->range_slab_end[0] = n_slabs;
    for (int f = 0; f < kFamilies; ++f) h->range_elem_end[f][0] = uint32_t(h->fh[f].size());
  }
  if (getenv("FEMGPU_ASM_INFO")) {
    fprintf(stderr, "[femgpu ranges] %d ranges of ~%u slabs; first range needs %.1f %% of the element records\n", h->n_ranges,
            per, total ? 100.0 * double(first_share) / double(total) : 0.0);
  }
  return 0;
}

int32_t run_symbolic(Handle* h) {
  SYM_CHECK(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  t_tmp_stream = s;
  const bool timing = getenv("FEMGPU_SYM_TIMING") != nullptr;  // per-stage wall clock to stderr
  auto t_last = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(s);
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[femgpu symbolic] %-28s %8.2f ms\n", what,
            std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  {
    uint32_t cost[3] = {kCostTruss, kCostBeam, kCostPlate};
    const char* names[3] = {"FEMGPU_COST_T", "FEMGPU_COST_B", "FEMGPU_COST_P"};
    for (int f = 0; f < 3; ++f)
      if (const char* q = getenv(names[f])) cost[f] = uint32_t(std::max(1, atoi(q)));
    SYM_CHECK(cudaMemcpyToSymbolAsync(g_cost, cost, sizeof cost, 0, cudaMemcpyHostToDevice, s));
  }
  const uint32_t N = h->nodes_number;
  h->n_rows = 6 * int64_t(N);
  const int64_t NC = h->n_contrib;
  int kb = 1;
  while ((uint64_t(1) << kb) < uint64_t(N)) ++kb;
  h->key_bits = kb;

  SYM_CHECK(h->node_blk_ptr.reserve(size_t(N) + 1));
  SYM_CHECK(h->node_base.reserve(size_t(N) + 1));
  SYM_CHECK(h->node_len.reserve(2 * size_t(N) + 2));
  SYM_CHECK(h->row_ptr.reserve(size_t(h->n_rows) + 1));
  SYM_CHECK(h->d_flag.reserve(16));
  SYM_CHECK(cudaMemsetAsync(h->d_flag.p, 0, 64, s));

  mark("reserve node arrays");
  BlockBuild bb;
  int64_t n_extra = 0;
  std::vector<int64_t> ranges;
  if (h->dist.enabled) {
    // first build sees local elements only; its ghost-row block keys go to the owning ranks
    if (NC) {
      int32_t st = build_blocks(h, 0, bb);
      if (st) return st;
    }
    int32_t st = dist_collect_ghost_keys(h, bb, &n_extra, ranges);
    if (st) return st;
    bb.release();
  }
  const int64_t NCt = NC + n_extra;
  if (NCt >= (int64_t(1) << 31)) return h->fail(FEMGPU_ERR_LIMIT, "more than 2^31 node-pair contributions on one device");

  if (NCt == 0) {
    SYM_CHECK(cudaMemsetAsync(h->node_blk_ptr.p, 0, (size_t(N) + 1) * 4, s));
    SYM_CHECK(cudaMemsetAsync(h->node_base.p, 0, (size_t(N) + 1) * 8, s));
    SYM_CHECK(cudaMemsetAsync(h->node_len.p, 0, (2 * size_t(N) + 2) * 4, s));
    SYM_CHECK(cudaMemsetAsync(h->row_ptr.p, 0, (size_t(h->n_rows) + 1) * 8, s));
    SYM_CHECK(cudaStreamSynchronize(s));
    h->n_blocks = h->n_slabs = 0;
    h->nnz = 0;
    h->n_ranges = 1;
    h->range_slab_end[0] = 0;
    for (int f = 0; f < kFamilies; ++f) h->range_elem_end[f][0] = uint32_t(h->fh[f].size());
    if (h->dist.enabled) return dist_finalize_plan(h, ranges);
    return 0;
  }

  // ---- 1-3a. sorted contributions, unique blocks, block kinds
  {
    int32_t st = build_blocks(h, n_extra, bb);
    if (st) return st;
  }
  mark("sort + unique blocks");
  const uint32_t nblk = bb.nblk;
  Tmp& keys_a = bb.keys_a;
  Tmp& vals_a = bb.vals_a;
  Tmp& uniq = bb.uniq;
  Tmp& cptr_sorted = bb.cptr_sorted;
  uint32_t* contrib_sorted = bb.vals_b.as<uint32_t>();

  // ---- 3b. node ranges, row layout
  SYM_CHECK(h->blk_off.reserve(2 * size_t(nblk)));
  node_ptr_kernel<<<div_up(size_t(N) + 1, 256), 256, 0, s>>>(N, nblk, kb, h->blk_key.p, h->node_blk_ptr.p);
  Tmp node_size;
  SYM_CHECK(node_size.alloc((size_t(N) + 1) * 8));
  SYM_CHECK(cudaMemsetAsync(node_size.p, 0, (size_t(N) + 1) * 8, s));
  node_layout_kernel<<<div_up(N, 256), 256, 0, s>>>(N, h->node_blk_ptr.p, h->blk_full.p, h->blk_off.p,
                                                    h->node_len.p, node_size.as<int64_t>(), h->d_flag.p);
  h->launches += 2;
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, node_size.as<int64_t>(), h->node_base.p, int(N + 1), s);
    Tmp t;
    SYM_CHECK(t.alloc(tb));
    SYM_CHECK(cub::DeviceScan::ExclusiveSum(t.p, tb, node_size.as<int64_t>(), h->node_base.p, int(N + 1), s));
    SYM_CHECK(cudaStreamSynchronize(s));
  }
  int64_t nnz = 0;
  SYM_CHECK(cudaMemcpy(&nnz, h->node_base.p + N, 8, cudaMemcpyDeviceToHost));
  h->nnz = nnz;
  int32_t overflow = 0;
  SYM_CHECK(cudaMemcpy(&overflow, h->d_flag.p, 4, cudaMemcpyDeviceToHost));
  if (overflow) return h->fail(FEMGPU_ERR_LIMIT, "a node has 65536 or more entries per row (>= 10923 neighbours)");

  SYM_CHECK(h->col_idx.reserve(size_t(nnz)));
  SYM_CHECK(h->values.reserve(size_t(nnz)));
  mark("node layout + alloc col_idx/values");
  row_ptr_kernel<<<div_up(size_t(N) + 1, 256), 256, 0, s>>>(N, h->node_base.p, h->node_len.p, h->row_ptr.p);
  col_idx_kernel<<<div_up(uint64_t(N) * 32, 256), 256, 0, s>>>(N, kb, h->blk_key.p, h->node_blk_ptr.p, h->blk_full.p,
                                                              h->blk_off.p, h->node_len.p, h->node_base.p,
                                                              h->col_idx.p);
  h->launches += 2;

  mark("row layout, row_ptr, col_idx");
  // ---- 4. slabs and the in-slab thread order
  uint32_t quota = kSlabQuota;
  if (const char* q = getenv("FEMGPU_SLAB_QUOTA")) quota = std::max(1, atoi(q));  // tuning knob
  uint32_t n_slabs = div_up(nblk, quota);
  h->n_slabs = n_slabs;
  h->slab_quota = quota;
  SYM_CHECK(h->slabs.reserve(n_slabs));
  slab_kernel<<<div_up(n_slabs, 256), 256, 0, s>>>(n_slabs, quota, N, h->node_blk_ptr.p, h->node_base.p,
                                                   h->slabs.p, h->d_flag.p);
  h->launches++;
  Tmp okey_a, okey_b, oval_a;
  SYM_CHECK(okey_a.alloc(size_t(nblk) * 8));
  SYM_CHECK(okey_b.alloc(size_t(nblk) * 8));
  SYM_CHECK(oval_a.alloc(size_t(nblk) * 4));
  SYM_CHECK(h->blk_order.reserve(nblk));
  order_key_kernel<<<div_up(nblk, 256), 256, 0, s>>>(nblk, quota, kb, h->blk_key.p, h->node_blk_ptr.p,
                                                     cptr_sorted.as<uint32_t>(), contrib_sorted,
                                                     okey_a.as<uint64_t>(), oval_a.as<uint32_t>());
  h->launches++;
  {
    int sbits = 16;
    while ((uint64_t(1) << (sbits - 16)) < uint64_t(n_slabs) + 1) ++sbits;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, okey_a.as<uint64_t>(), okey_b.as<uint64_t>(),
                                    oval_a.as<uint32_t>(), h->blk_order.p, int(nblk), 0, sbits, s);
    Tmp t;
    SYM_CHECK(t.alloc(tb));
    SYM_CHECK(cub::DeviceRadixSort::SortPairs(t.p, tb, okey_a.as<uint64_t>(), okey_b.as<uint64_t>(),
                                              oval_a.as<uint32_t>(), h->blk_order.p, int(nblk), 0, sbits, s));
    SYM_CHECK(cudaStreamSynchronize(s));
  }
  Tmp ocnt, ocost;
  SYM_CHECK(ocnt.alloc((size_t(nblk) + 1) * 4));
  SYM_CHECK(ocost.alloc((size_t(nblk) + 1) * 4));
  SYM_CHECK(cudaMemsetAsync(ocnt.as<uint32_t>() + nblk, 0, 4, s));
  ordered_count_kernel<<<div_up(nblk, 256), 256, 0, s>>>(nblk, h->blk_order.p, cptr_sorted.as<uint32_t>(),
                                                         contrib_sorted, ocnt.as<uint32_t>(),
                                                         ocost.as<uint32_t>());
  h->launches++;
  SYM_CHECK(h->blk_cptr.reserve(size_t(nblk) + 1));
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, ocnt.as<uint32_t>(), h->blk_cptr.p, int(nblk + 1), s);
    Tmp t;
    SYM_CHECK(t.alloc(tb));
    SYM_CHECK(cub::DeviceScan::ExclusiveSum(t.p, tb, ocnt.as<uint32_t>(), h->blk_cptr.p, int(nblk + 1), s));
    SYM_CHECK(cudaStreamSynchronize(s));
  }
  SYM_CHECK(h->contrib.reserve(size_t(NCt) + 8));  // +8: entries are fetched in 16-byte chunks, one entry ahead
  SYM_CHECK(h->blk_meta.reserve(size_t(nblk) + 1));  // +1: the kernel loads one block ahead
  ordered_meta_kernel<<<div_up(nblk, 256), 256, 0, s>>>(nblk, quota, kb, h->blk_order.p, h->blk_key.p,
                                                        h->node_blk_ptr.p, h->blk_off.p, h->node_len.p,
                                                        h->node_base.p, h->slabs.p, cptr_sorted.as<uint32_t>(),
                                                        contrib_sorted, h->blk_cptr.p, h->contrib.p,
                                                        h->blk_meta.p);
  h->launches++;
  SYM_CHECK(cudaGetLastError());
  SYM_CHECK(cudaStreamSynchronize(s));
  SYM_CHECK(cudaMemsetAsync(h->contrib.p + NCt, 0, 32, s));
  SYM_CHECK(cudaMemsetAsync(h->blk_meta.p + nblk, 0, sizeof(BlockMeta), s));
  mark("slabs, thread order, meta");
  // per-slab element lists; contrib codes are relabelled to slab-local element slots
  {
    if (uint64_t(n_slabs) >= (uint64_t(1) << 36)) return h->fail(FEMGPU_ERR_LIMIT, "too many slabs");
    uint64_t* ekey_a = keys_a.as<uint64_t>();  // the contribution sort buffers are free again
    uint64_t* ekey_b = uniq.as<uint64_t>();
    uint32_t* epay_a = vals_a.as<uint32_t>();
    Tmp epay_b, eflag, euidx;
    SYM_CHECK(epay_b.alloc(size_t(NCt) * 4));
    SYM_CHECK(eflag.alloc(size_t(NCt) * 4));
    SYM_CHECK(euidx.alloc(size_t(NCt) * 4));
    slab_elem_key_kernel<<<div_up(nblk, 256), 256, 0, s>>>(nblk, quota, kb, h->blk_order.p, h->blk_key.p,
                                                           h->node_blk_ptr.p, h->blk_cptr.p, h->contrib.p, ekey_a,
                                                           epay_a);
    int ebits = 28;
    while ((uint64_t(1) << (ebits - 28)) < uint64_t(n_slabs) + 1) ++ebits;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, ekey_a, ekey_b, epay_a, epay_b.as<uint32_t>(), int(NCt), 0, ebits, s);
    {
      Tmp t;
      SYM_CHECK(t.alloc(tb));
      SYM_CHECK(cub::DeviceRadixSort::SortPairs(t.p, tb, ekey_a, ekey_b, epay_a, epay_b.as<uint32_t>(), int(NCt), 0,
                                                ebits, s));
      SYM_CHECK(cudaStreamSynchronize(s));
    }
    head_flag_kernel<<<div_up(NCt, 256), 256, 0, s>>>(uint32_t(NCt), ekey_b, eflag.as<uint32_t>());
    tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, eflag.as<uint32_t>(), euidx.as<uint32_t>(), int(NCt), s);
    {
      Tmp t;
      SYM_CHECK(t.alloc(tb));
      SYM_CHECK(cub::DeviceScan::InclusiveSum(t.p, tb, eflag.as<uint32_t>(), euidx.as<uint32_t>(), int(NCt), s));
      SYM_CHECK(cudaStreamSynchronize(s));
    }
    uint32_t n_unique = 0;
    SYM_CHECK(cudaMemcpy(&n_unique, euidx.as<uint32_t>() + (NCt - 1), 4, cudaMemcpyDeviceToHost));
    SYM_CHECK(h->elist_compact.reserve(n_unique));
    SYM_CHECK(h->elist.reserve(size_t(n_slabs) * kElistStride));
    uint64_t* ukey = ekey_a;  // reuse: ekey_a is dead after the sort
    elist_kernel<<<div_up(NCt, 256), 256, 0, s>>>(uint32_t(NCt), ekey_b, eflag.as<uint32_t>(), euidx.as<uint32_t>(),
                                                 ukey, h->elist_compact.p);
    slab_elist_kernel<<<div_up(n_slabs, 256), 256, 0, s>>>(n_slabs, n_unique, ukey, h->blk_cptr.p, h->slabs.p,
                                                           h->d_flag.p);
    relabel_kernel<<<div_up(NCt, 256), 256, 0, s>>>(uint32_t(NCt), ekey_b, epay_b.as<uint32_t>(), euidx.as<uint32_t>(),
                                                   h->slabs.p, h->contrib.p);
    elist_table_kernel<<<div_up(uint64_t(n_slabs) * kElistStride, 256), 256, 0, s>>>(n_slabs, h->slabs.p,
                                                                                  h->elist_compact.p, h->elist.p);
    run_stat_kernel<<<div_up(uint64_t(n_slabs) * kElistStride, 256), 256, 0, s>>>(uint64_t(n_slabs) * kElistStride,
                                                                               h->elist.p, h->d_flag.p);
    h->launches += 7;
    SYM_CHECK(cudaGetLastError());
    SYM_CHECK(cudaStreamSynchronize(s));
  }
  mark("element lists, relabel");
  // balanced per-thread work lists (needs blk_cptr, filled above)
  {
    // lanes per slab: two warps for beam and mixed-family meshes, one for plate-only / truss-only
    const bool has_t = h->fh[FEMGPU_TRUSS].size() != 0, has_b = h->fh[FEMGPU_BEAM].size() != 0,
               has_p = h->fh[FEMGPU_PLATE].size() != 0;
    h->asm_threads = (has_b || (has_t && has_p)) ? 64 : 32;
    if (const char* q = getenv("FEMGPU_ASM_THREADS")) h->asm_threads = (atoi(q) == 64) ? 64 : 32;  // tuning knob
  }
  const uint32_t threads = uint32_t(h->asm_threads);
  SYM_CHECK(h->items.reserve(size_t(n_slabs) * threads));
  int spread_banks = 0;
  if (const char* q = getenv("FEMGPU_SPREAD_BANKS")) spread_banks = atoi(q);  // tuning knob
  work_item_kernel<<<div_up(n_slabs, 128), 128, 0, s>>>(n_slabs, threads, h->slabs.p, ocost.as<uint32_t>(),
                                                         h->blk_cptr.p, h->contrib.p, h->items.p, h->d_flag.p,
                                                         h->blk_meta.p, spread_banks);
  SYM_CHECK(h->items_c.reserve(size_t(n_slabs) * threads));
  compact_items_kernel<<<div_up(uint64_t(n_slabs) * threads, 256), 256, 0, s>>>(n_slabs, threads, h->slabs.p,
                                                                                h->items.p, h->items_c.p);
  h->launches += 2;
  SYM_CHECK(cudaGetLastError());
  SYM_CHECK(cudaStreamSynchronize(s));
  mark("work items");
  {
    Tmp prog;
    SYM_CHECK(prog.alloc((size_t(NCt) + 1) * 4));
    item_program_kernel<<<div_up(uint64_t(n_slabs) * threads, 256), 256, 0, s>>>(
        n_slabs, threads, h->slabs.p, h->items.p, h->blk_cptr.p, h->contrib.p, prog.as<uint32_t>());
    h->launches++;
    SYM_CHECK(cudaGetLastError());
    SYM_CHECK(cudaMemcpyAsync(h->contrib.p, prog.p, size_t(NCt) * 4, cudaMemcpyDeviceToDevice, s));
    SYM_CHECK(cudaStreamSynchronize(s));
  }
  mark("item programs");
  int32_t flags[16] = {0};
  SYM_CHECK(cudaMemcpy(flags, h->d_flag.p, 64, cudaMemcpyDeviceToHost));
  if (flags[0]) return h->fail(FEMGPU_ERR_LIMIT, "a slab holds 2^32 or more values");
  h->smem_img = uint32_t(flags[8]);
  h->smem_form = uint32_t(flags[9]);
  h->smem_rawp = uint32_t(flags[10]);
  h->smem_stage = uint32_t(flags[11]);
  h->n_unstaged = uint32_t(flags[12]);
  h->asm_split = flags[13] != 0;
  // run-wise bulk staging of the element records pays off when a run moves 512 bytes or more on average (mixed
  // grid 0.9 KB: -1.7 %, beam frame 1.5 KB: -6 %; truss lattice 0.4 KB: +8 %, where LDGSTS stays)
  h->asm_bulk = int64_t(flags[14]) >= 4 * int64_t(flags[15]) && flags[15] > 0;
  if (const char* q = getenv("FEMGPU_ASM_BULK")) h->asm_bulk = atoi(q) != 0;  // tuning knob

  {
    int32_t st = build_range_plan(h);
    if (st) return st;
  }
  mark("range plan");
  if (h->dist.enabled) {
    int32_t st = dist_finalize_plan(h, ranges);
    if (st) return st;
  }
  return 0;
}

int32_t element_slots(Handle* h, int family, size_t index, int64_t* out_host) {
  SYM_CHECK(cudaSetDevice(h->device));
  t_tmp_stream = h->stream;
  const int nn = kNodesPerElem[family], dof = family == FEMGPU_TRUSS ? 3 : 6;
  const int n = nn * dof;
  uint32_t nodes[4];
  for (int c = 0; c < nn; ++c) nodes[c] = h->fh[family].conn[c][index];
  Tmp d_nodes, d_out;
  SYM_CHECK(d_nodes.alloc(16));
  SYM_CHECK(d_out.alloc(size_t(n) * n * 8));
  SYM_CHECK(cudaMemcpyAsync(d_nodes.p, nodes, 16, cudaMemcpyHostToDevice, h->stream));
  element_slots_kernel<<<div_up(n * n, 128), 128, 0, h->stream>>>(nn, dof, d_nodes.as<uint32_t>(), h->n_blocks,
                                                                  h->key_bits, h->blk_key.p, h->blk_off.p,
                                                                  h->node_len.p, h->node_base.p, d_out.as<int64_t>());
  h->launches++;
  SYM_CHECK(cudaGetLastError());
  SYM_CHECK(cudaMemcpyAsync(out_host, d_out.p, size_t(n) * n * 8, cudaMemcpyDeviceToHost, h->stream));
  SYM_CHECK(cudaStreamSynchronize(h->stream));
  return 0;
}

// The assembled matrix compacted to its entries != 0.0, as CSR (same rows, columns ascending): computed on the device
// once per numeric pass and kept in HBM; `count` only -> no copy. Multi-GPU: the rows this rank owns (the others empty).
int32_t nonzero_csr(Handle* h, int64_t* count, int64_t* row_ptr, int32_t* cols, double* vals) {
  SYM_CHECK(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  t_tmp_stream = s;
  const int64_t n_rows = h->n_rows;
  if (!h->nz_valid || h->nz_pass != h->n_numeric) {
    int64_t rb = 0, re = n_rows;
    if (h->dist.enabled && h->dist.ownership_set) {
      rb = 6 * int64_t(h->dist.own_begin);
      re = 6 * int64_t(h->dist.own_end);
    }
    Tmp cnt;
    SYM_CHECK(cnt.alloc((size_t(n_rows) + 1) * 4));
    SYM_CHECK(cudaMemsetAsync(cnt.p, 0, (size_t(n_rows) + 1) * 4, s));
    SYM_CHECK(h->nz_row_ptr.reserve(size_t(n_rows) + 1));
    const uint32_t grid = div_up(uint64_t(re - rb) * kNzLanes, 256);
    if (re > rb && h->nnz) {
      nz_row_count_kernel<<<grid, 256, 0, s>>>(rb, re, h->row_ptr.p, h->values.p, cnt.as<int32_t>());
      h->launches++;
    }
    {
      cub::TransformInputIterator<int64_t, CastToI64, const int32_t*> it(cnt.as<int32_t>(), CastToI64());
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, it, h->nz_row_ptr.p, int(n_rows + 1), s);
      Tmp t;
      SYM_CHECK(t.alloc(tb));
      SYM_CHECK(cub::DeviceScan::ExclusiveSum(t.p, tb, it, h->nz_row_ptr.p, int(n_rows + 1), s));
    }
    int64_t nz = 0;
    SYM_CHECK(cudaMemcpyAsync(&nz, h->nz_row_ptr.p + n_rows, 8, cudaMemcpyDeviceToHost, s));
    SYM_CHECK(cudaStreamSynchronize(s));
    SYM_CHECK(h->nz_col.reserve(size_t(nz) + 1));
    SYM_CHECK(h->nz_val.reserve(size_t(nz) + 1));
    if (nz) {
      nz_row_fill_kernel<<<grid, 256, 0, s>>>(rb, re, h->row_ptr.p, h->col_idx.p, h->values.p, h->nz_row_ptr.p,
                                              h->nz_col.p, h->nz_val.p);
      h->launches++;
      SYM_CHECK(cudaGetLastError());
    }
    h->nz_count = nz;
    h->nz_pass = h->n_numeric;
    h->nz_valid = true;
  }
  if (count) *count = h->nz_count;
  if (row_ptr) SYM_CHECK(cudaMemcpyAsync(row_ptr, h->nz_row_ptr.p, (size_t(n_rows) + 1) * 8, cudaMemcpyDeviceToHost, s));
  if (cols && h->nz_count)
    SYM_CHECK(cudaMemcpyAsync(cols, h->nz_col.p, size_t(h->nz_count) * 4, cudaMemcpyDeviceToHost, s));
  if (vals && h->nz_count)
    SYM_CHECK(cudaMemcpyAsync(vals, h->nz_val.p, size_t(h->nz_count) * 8, cudaMemcpyDeviceToHost, s));
  SYM_CHECK(cudaStreamSynchronize(s));
  return 0;
}

int32_t nonzero_coo(Handle* h, int64_t* count, int64_t* rows, int64_t* cols, double* vals) {
  SYM_CHECK(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  t_tmp_stream = s;
  int64_t nnz = h->nnz;
  if (nnz == 0) {
    if (count) *count = 0;
    return 0;
  }
  if (nnz >= (int64_t(1) << 31)) return h->fail(FEMGPU_ERR_LIMIT, "nonzero compaction supports < 2^31 entries");
  Tmp flag, pos;
  SYM_CHECK(flag.alloc((size_t(nnz) + 1) * 8));
  SYM_CHECK(pos.alloc((size_t(nnz) + 1) * 8));
  SYM_CHECK(cudaMemsetAsync(flag.as<int64_t>() + nnz, 0, 8, s));
  nz_flag_kernel<<<div_up(nnz, 256), 256, 0, s>>>(nnz, h->values.p, flag.as<int64_t>());
  h->launches++;
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, flag.as<int64_t>(), pos.as<int64_t>(), int(nnz + 1), s);
    Tmp t;
    SYM_CHECK(t.alloc(tb));
    SYM_CHECK(cub::DeviceScan::ExclusiveSum(t.p, tb, flag.as<int64_t>(), pos.as<int64_t>(), int(nnz + 1), s));
    SYM_CHECK(cudaStreamSynchronize(s));
  }
  int64_t nz = 0;
  SYM_CHECK(cudaMemcpy(&nz, pos.as<int64_t>() + nnz, 8, cudaMemcpyDeviceToHost));
  if (count) *count = nz;
  if (!rows || !cols || !vals || nz == 0) return 0;
  Tmp d_rows, d_cols, d_vals;
  SYM_CHECK(d_rows.alloc(size_t(nz) * 8));
  SYM_CHECK(d_cols.alloc(size_t(nz) * 8));
  SYM_CHECK(d_vals.alloc(size_t(nz) * 8));
  nz_write_kernel<<<div_up(nnz, 256), 256, 0, s>>>(nnz, h->n_rows, h->values.p, pos.as<int64_t>(), h->row_ptr.p,
                                                   h->col_idx.p, d_rows.as<int64_t>(), d_cols.as<int64_t>(),
                                                   d_vals.as<double>());
  h->launches++;
  SYM_CHECK(cudaGetLastError());
  SYM_CHECK(cudaMemcpyAsync(rows, d_rows.p, size_t(nz) * 8, cudaMemcpyDeviceToHost, s));
  SYM_CHECK(cudaMemcpyAsync(cols, d_cols.p, size_t(nz) * 8, cudaMemcpyDeviceToHost, s));
  SYM_CHECK(cudaMemcpyAsync(vals, d_vals.p, size_t(nz) * 8, cudaMemcpyDeviceToHost, s));
  SYM_CHECK(cudaStreamSynchronize(s));
  return 0;
}

}  // namespace femgpu
