// Boundary conditions and the sparse separation of the assembled matrix — the first consumer of K
// in the reference:
//   FEM::add_displacement / add_concentrated_load        methods_for_bc_data_handle.rs:31-56,175-203
//   FEM::separate_stiffness_matrix_sparse_iterative      methods_for_separate_stiffness_matrix.rs:217-320
//   compose_r_a_vector / compose_u_b_vector / find_b_sparse
//                                methods_for_separate_stiffness_matrix.rs:322-343, methods_for_global_analysis.rs:28-48
//
// The reference walks the position-keyed map of K three times on one core. Here the CSR values stay
// in HBM and the separation is integer stream compaction:
//   1. classify every DOF from its diagonal entry and the constraint flags
//      (zero diagonal -> inactive, constrained -> "b", else "a"), first error = smallest index;
//   2. exclusive scans give the local numbering of the a- and b-sets (ascending global index, like
//      the reference's push order);
//   3. one warp per row counts its non-zero entries per quadrant, scans give four CSR row pointers;
//   4. the same walk writes (local column, value) in row order — deterministic and column-sorted.
// Entries equal to 0.0 are dropped (the reference skips them, :277-279) and entries touching an
// inactive DOF are ignored (:296-299), so the four matrices hold exactly the reference's triplets,
// as CSR (its consumer builds CsrMatrix::from_coo out of them, methods_for_global_analysis.rs:205).
// Bound: HBM — col_idx and values are read twice, the compacted copies written once.
#include <cub/cub.cuh>

#include <functional>

#include "common.cuh"

namespace femgpu {

namespace {

constexpr uint32_t kClsNone = 0u, kClsA = 1u, kClsB = 2u;

struct CastI64 {
  __host__ __device__ int64_t operator()(const int32_t& v) const { return int64_t(v); }
};

// diag(K)[i]: the row's columns ascend, so the diagonal is found by bisection; an absent entry is 0
__global__ void classify_kernel(int64_t n_rows, const int64_t* __restrict__ row_ptr,
                                const int32_t* __restrict__ col_idx, const double* __restrict__ values,
                                const uint8_t* __restrict__ constrained, const double* __restrict__ force,
                                int32_t* __restrict__ is_a, int32_t* __restrict__ is_b,
                                unsigned long long* __restrict__ first_bad) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  int64_t lo = row_ptr[i], hi = row_ptr[i + 1];
  double d = 0.0;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    int32_t c = col_idx[mid];
    if (c == int32_t(i)) {
      d = values[mid];
      break;
    }
    if (c < int32_t(i)) lo = mid + 1;
    else hi = mid;
  }
  const bool fixed = constrained[i] != 0;
  int32_t a = 0, b = 0;
  if (d == 0.0) {
    // check_excluded_index of the direct separation (methods_for_separate_stiffness_matrix.rs:36-61) also rejects
    // a load on a DOF without stiffness; the sparse separation only looks at the constraints (force == nullptr)
    if (fixed || (force && force[i] != 0.0)) atomicMin(first_bad, (unsigned long long)i);  // integer min: order independent
  } else if (fixed) {
    b = 1;
  } else {
    a = 1;
  }
  is_a[i] = a;
  is_b[i] = b;
}

// cls_pos[i] = class << 30 | local index; index lists in ascending global order
__global__ void number_kernel(int64_t n_rows, const int32_t* __restrict__ is_a, const int32_t* __restrict__ is_b,
                              const int32_t* __restrict__ pos_a, const int32_t* __restrict__ pos_b,
                              uint32_t* __restrict__ cls_pos, int64_t* __restrict__ aa_idx,
                              int64_t* __restrict__ bb_idx) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  uint32_t v = 0;
  if (is_a[i]) {
    v = (kClsA << 30) | uint32_t(pos_a[i]);
    aa_idx[pos_a[i]] = i;
  } else if (is_b[i]) {
    v = (kClsB << 30) | uint32_t(pos_b[i]);
    bb_idx[pos_b[i]] = i;
  }
  cls_pos[i] = v;
}

// One warp per global row. kFill = false: count the row's non-zero entries whose column is in the
// a-set / b-set. kFill = true: write them, in row order, at the positions the scanned counts give.
// Quadrants: 0 = aa, 1 = ab, 2 = ba, 3 = bb.
struct SepOut {
  int64_t* row_ptr[4];
  int32_t* col[4];
  double* val[4];
  int32_t* cnt[4];
};

// one row, any length: chunks of 32 entries, ballots give the in-row positions
template <bool kFill>
__device__ __forceinline__ void quadrant_row(uint32_t lane, uint32_t rc, int64_t begin, int64_t end,
                                             const int32_t* __restrict__ col_idx, const double* __restrict__ values,
                                             const uint32_t* __restrict__ cls_pos, const SepOut& out) {
  const uint32_t rcls = rc >> 30, rloc = rc & 0x3FFFFFFFu;
  if (rcls == kClsNone) return;
  const int qa = rcls == kClsA ? 0 : 2, qb = qa + 1;
  int64_t wa = 0, wb = 0;
  if (kFill) {
    wa = out.row_ptr[qa][rloc];
    wb = out.row_ptr[qb][rloc];
  }
  uint32_t na = 0, nb = 0;
  for (int64_t p = begin + lane; p - lane < end; p += 32) {
    bool in_a = false, in_b = false;
    double v = 0.0;
    uint32_t cloc = 0;
    if (p < end) {
      v = values[p];
      if (v != 0.0) {
        const uint32_t cc = cls_pos[col_idx[p]];
        in_a = (cc >> 30) == kClsA;
        in_b = (cc >> 30) == kClsB;
        cloc = cc & 0x3FFFFFFFu;
      }
    }
    const uint32_t ma = __ballot_sync(0xFFFFFFFFu, in_a), mb = __ballot_sync(0xFFFFFFFFu, in_b);
    if (kFill) {
      const uint32_t below = (1u << lane) - 1u;
      if (in_a) {
        const int64_t w = wa + na + __popc(ma & below);
        out.col[qa][w] = int32_t(cloc);
        out.val[qa][w] = v;
      } else if (in_b) {
        const int64_t w = wb + nb + __popc(mb & below);
        out.col[qb][w] = int32_t(cloc);
        out.val[qb][w] = v;
      }
    }
    na += __popc(ma);
    nb += __popc(mb);
  }
  if (!kFill && lane == 0) {
    out.cnt[qa][rloc] = int32_t(na);
    out.cnt[qb][rloc] = int32_t(nb);
  }
}

// One warp per kRowsPerWarp consecutive rows. Rows of a structural model are short (54 entries on a
// plate grid) and each needs a chain of dependent loads (row_ptr -> values/col_idx -> class of the
// column), so a warp that walks one row at a time is latency-bound. When its rows have at most 64
// entries each — two chunks of 32 — the warp issues the loads of all of them back to back before
// the first use; longer rows take the generic loop.
#ifndef FEMGPU_SEP_ROWS
#define FEMGPU_SEP_ROWS 4
#endif
constexpr int kRowsPerWarp = FEMGPU_SEP_ROWS;

template <bool kFill>
__global__ void __launch_bounds__(256)
quadrant_kernel(int64_t n_rows, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                const double* __restrict__ values, const uint32_t* __restrict__ cls_pos, SepOut out) {
  const int64_t row0 = ((int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * kRowsPerWarp;
  const uint32_t lane = threadIdx.x & 31u;
  if (row0 >= n_rows) return;
  // lanes 0..4 fetch the five row pointers, lanes 0..3 the rows' classes
  const int64_t my_row = row0 + lane;
  const int64_t ptr_l = (lane <= uint32_t(kRowsPerWarp) && my_row <= n_rows) ? row_ptr[my_row] : 0;
  const uint32_t cls_l = (lane < uint32_t(kRowsPerWarp) && my_row < n_rows) ? cls_pos[my_row] : 0u;
  int64_t begin[kRowsPerWarp], end[kRowsPerWarp];
  uint32_t rc[kRowsPerWarp];
  bool all_short = true;
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) {
    begin[r] = __shfl_sync(0xFFFFFFFFu, ptr_l, r);
    end[r] = __shfl_sync(0xFFFFFFFFu, ptr_l, r + 1);
    rc[r] = __shfl_sync(0xFFFFFFFFu, cls_l, r);
    if (row0 + r >= n_rows) {
      rc[r] = 0u;
      end[r] = begin[r];
    }
    all_short = all_short && (end[r] - begin[r] <= 64);
  }
  if (!all_short) {
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) quadrant_row<kFill>(lane, rc[r], begin[r], end[r], col_idx, values, cls_pos, out);
    return;
  }
  double v[kRowsPerWarp][2];
  int32_t c[kRowsPerWarp][2];
  uint32_t cc[kRowsPerWarp][2];
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int64_t p = begin[r] + lane + 32 * k;
      const bool on = (rc[r] >> 30) != kClsNone && p < end[r];
      v[r][k] = on ? values[p] : 0.0;
      c[r][k] = on ? col_idx[p] : 0;
    }
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
    for (int k = 0; k < 2; ++k) cc[r][k] = (v[r][k] != 0.0) ? cls_pos[c[r][k]] : 0u;
  int64_t wa[kRowsPerWarp], wb[kRowsPerWarp];
  if (kFill) {
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const uint32_t rcls = rc[r] >> 30, rloc = rc[r] & 0x3FFFFFFFu;
      const int qa = rcls == kClsA ? 0 : 2;
      wa[r] = rcls != kClsNone ? out.row_ptr[qa][rloc] : 0;
      wb[r] = rcls != kClsNone ? out.row_ptr[qa + 1][rloc] : 0;
    }
  }
  const uint32_t below = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) {
    const uint32_t rcls = rc[r] >> 30, rloc = rc[r] & 0x3FFFFFFFu;
    if (rcls == kClsNone) continue;  // warp-uniform
    const int qa = rcls == kClsA ? 0 : 2, qb = qa + 1;
    uint32_t na = 0, nb = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const bool in_a = (cc[r][k] >> 30) == kClsA, in_b = (cc[r][k] >> 30) == kClsB;
      const uint32_t ma = __ballot_sync(0xFFFFFFFFu, in_a), mb = __ballot_sync(0xFFFFFFFFu, in_b);
      if (kFill) {
        if (in_a) {
          const int64_t w = wa[r] + na + __popc(ma & below);
          out.col[qa][w] = int32_t(cc[r][k] & 0x3FFFFFFFu);
          out.val[qa][w] = v[r][k];
        } else if (in_b) {
          const int64_t w = wb[r] + nb + __popc(mb & below);
          out.col[qb][w] = int32_t(cc[r][k] & 0x3FFFFFFFu);
          out.val[qb][w] = v[r][k];
        }
      }
      na += __popc(ma);
      nb += __popc(mb);
    }
    if (!kFill && lane == 0) {
      out.cnt[qa][rloc] = int32_t(na);
      out.cnt[qb][rloc] = int32_t(nb);
    }
  }
}

// ---- one-pass separation ------------------------------------------------------------------------------------------
// count + fill above read col_idx and values from HBM twice: the counts of ALL rows are scanned before the first
// entry can be placed. Here a CTA takes a tile of 128 consecutive rows (~80 KB of K), counts it, joins four chained
// scans over the tiles with the tile's counts (aa, ab over its a-rows; ba, bb over its b-rows) — it publishes its own
// counts at once, then adds up its predecessors' words until it meets one that already holds a running total
// (decoupled look-back) — and walks the tile a second time to place the entries, while those 80 KB are still in the
// L2. So K comes out of HBM once. Tiles take their number from a ticket counter: a tile's predecessors are always
// running or done. The outputs are sized by upper bounds (run_separate); a spin that does not end or a bound that
// does not hold raises a flag and the caller falls back to the two passes.
// (A tile of 32 rows kept in registers between the two steps was measured first: 28.5 ms on config M against 12.7 —
// 750 000 tiles at three CTAs per SM, each waiting on its look-back with its loads already spent.)
#ifndef FEMGPU_SEP_TILE_ITERS
#define FEMGPU_SEP_TILE_ITERS 4
#endif
constexpr int kTileIters = FEMGPU_SEP_TILE_ITERS;             // rounds per tile (build knob: 1, 2 or 4)
constexpr int kTileRows = 8 * kRowsPerWarp * kTileIters;       // 256 threads: 8 warps x 4 rows x 4 rounds
constexpr int kRowsPerLane = kTileRows / 32;                   // rows per lane in the scan step
constexpr unsigned long long kTileMask = (1ull << 62) - 1ull;  // word = state << 62 | count; state 1: own count, 2: running total
constexpr uint32_t kSpinLimit = 1u << 21;

struct OnePass {
  unsigned long long* state[4];  // [n_tiles] per quadrant
  unsigned int* ticket;
  int* error;
  int64_t cap[4];
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// counts of one row of any length (the loop of quadrant_row<false>)
__device__ __forceinline__ void count_long_row(uint32_t lane, int64_t begin, int64_t end, const int32_t* __restrict__ col_idx,
                                               const double* __restrict__ values, const uint32_t* __restrict__ cls_pos,
                                               uint32_t& na, uint32_t& nb) {
  na = nb = 0;
  for (int64_t p = begin + lane; p - lane < end; p += 32) {
    bool in_a = false, in_b = false;
    if (p < end && values[p] != 0.0) {
      const uint32_t cc = cls_pos[col_idx[p]];
      in_a = (cc >> 30) == kClsA;
      in_b = (cc >> 30) == kClsB;
    }
    na += __popc(__ballot_sync(0xFFFFFFFFu, in_a));
    nb += __popc(__ballot_sync(0xFFFFFFFFu, in_b));
  }
}

__device__ __forceinline__ void fill_long_row(uint32_t lane, int qa, int64_t wa, int64_t wb, int64_t begin, int64_t end,
                                              const int32_t* __restrict__ col_idx, const double* __restrict__ values,
                                              const uint32_t* __restrict__ cls_pos, const SepOut& out) {
  const int qb = qa + 1;
  const uint32_t below = (1u << lane) - 1u;
  uint32_t na = 0, nb = 0;
  for (int64_t p = begin + lane; p - lane < end; p += 32) {
    bool in_a = false, in_b = false;
    double v = 0.0;
    uint32_t cloc = 0;
    if (p < end) {
      v = values[p];
      if (v != 0.0) {
        const uint32_t cc = cls_pos[col_idx[p]];
        in_a = (cc >> 30) == kClsA;
        in_b = (cc >> 30) == kClsB;
        cloc = cc & 0x3FFFFFFFu;
      }
    }
    const uint32_t ma = __ballot_sync(0xFFFFFFFFu, in_a), mb = __ballot_sync(0xFFFFFFFFu, in_b);
    if (in_a) {
      const int64_t w = wa + na + __popc(ma & below);
      out.col[qa][w] = int32_t(cloc);
      out.val[qa][w] = v;
    } else if (in_b) {
      const int64_t w = wb + nb + __popc(mb & below);
      out.col[qb][w] = int32_t(cloc);
      out.val[qb][w] = v;
    }
    na += __popc(ma);
    nb += __popc(mb);
  }
}

// kRowsPerWarp consecutive rows of one warp: pointers, classes and — when none is longer than 64 entries — the
// entries with their columns' classes in registers (the loads of all rows issued back to back)
struct WarpRows {
  int64_t begin[kRowsPerWarp], end[kRowsPerWarp];
  uint32_t rc[kRowsPerWarp];
  double v[kRowsPerWarp][2];
  uint32_t cc[kRowsPerWarp][2];
  bool all_short;
};

__device__ __forceinline__ void load_warp_rows(WarpRows& w, int64_t row0, uint32_t lane, int64_t n_rows,
                                               const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                               const double* __restrict__ values, const uint32_t* __restrict__ cls_pos) {
  const int64_t my_row = row0 + lane;
  const int64_t ptr_l = (lane <= uint32_t(kRowsPerWarp) && my_row <= n_rows) ? row_ptr[my_row] : 0;
  const uint32_t cls_l = (lane < uint32_t(kRowsPerWarp) && my_row < n_rows) ? cls_pos[my_row] : 0u;
  w.all_short = true;
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) {
    w.begin[r] = __shfl_sync(0xFFFFFFFFu, ptr_l, r);
    w.end[r] = __shfl_sync(0xFFFFFFFFu, ptr_l, r + 1);
    w.rc[r] = __shfl_sync(0xFFFFFFFFu, cls_l, r);
    if (row0 + r >= n_rows) {
      w.rc[r] = 0u;
      w.end[r] = w.begin[r];
    }
    w.all_short = w.all_short && (w.end[r] - w.begin[r] <= 64);
  }
  if (!w.all_short) return;
  int32_t c[kRowsPerWarp][2];
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int64_t p = w.begin[r] + lane + 32 * k;
      const bool on = (w.rc[r] >> 30) != kClsNone && p < w.end[r];
      w.v[r][k] = on ? values[p] : 0.0;
      c[r][k] = on ? col_idx[p] : 0;
    }
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
    for (int k = 0; k < 2; ++k) w.cc[r][k] = (w.v[r][k] != 0.0) ? cls_pos[c[r][k]] : 0u;
}

__global__ void __launch_bounds__(256, 4)
quadrant_onepass_kernel(int64_t n_rows, uint32_t n_tiles, const int64_t* __restrict__ row_ptr,
                        const int32_t* __restrict__ col_idx, const double* __restrict__ values,
                        const uint32_t* __restrict__ cls_pos, const __grid_constant__ SepOut out,
                        const __grid_constant__ OnePass op) {
  // (__grid_constant__: out / op are indexed by the quadrant at run time; without it every thread copied both structs
  // to its local-memory stack frame first — 248 B x 48 M threads = 12 GB of stores per launch, ncu: 16.7 GB written
  // to DRAM against 6.3 GB of output)
  static_assert(kTileRows % 32 == 0 && kRowsPerLane >= 1, "whole rows per lane in the scan step");
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_n[2][kTileRows], s_rc[kTileRows], s_ex[2][kTileRows];
  __shared__ unsigned long long s_base[4];
  if (threadIdx.x == 0) s_tile = atomicAdd(op.ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  if (tile >= n_tiles) return;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const int64_t tile_row0 = int64_t(tile) * kTileRows;

  // ---- 1. count: in round `it` the eight warps take 32 consecutive rows
#pragma unroll 1
  for (int it = 0; it < kTileIters; ++it) {
    const uint32_t t0 = (uint32_t(it) * 8 + warp) * kRowsPerWarp;
    WarpRows w;
    load_warp_rows(w, tile_row0 + t0, lane, n_rows, row_ptr, col_idx, values, cls_pos);
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      uint32_t na = 0, nb = 0;
      if (w.all_short) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          na += __popc(__ballot_sync(0xFFFFFFFFu, (w.cc[r][k] >> 30) == kClsA));
          nb += __popc(__ballot_sync(0xFFFFFFFFu, (w.cc[r][k] >> 30) == kClsB));
        }
      } else if ((w.rc[r] >> 30) != kClsNone) {
        count_long_row(lane, w.begin[r], w.end[r], col_idx, values, cls_pos, na, nb);
      }
      if (lane == 0) {
        s_n[0][t0 + r] = na;
        s_n[1][t0 + r] = nb;
        s_rc[t0 + r] = w.rc[r];
      }
    }
  }
  __syncthreads();

  // ---- 2. warps 0..3: quadrant q = warp. In-tile offsets (four rows per lane), then the tile's place among the tiles
  if (warp < 4) {
    const int q = int(warp);
    const uint32_t want = q < 2 ? kClsA : kClsB;
    uint32_t cnt[kRowsPerLane], lane_sum = 0;
    bool mine[kRowsPerLane];
#pragma unroll
    for (int j = 0; j < kRowsPerLane; ++j) {
      mine[j] = (s_rc[kRowsPerLane * lane + j] >> 30) == want;
      cnt[j] = mine[j] ? s_n[q & 1][kRowsPerLane * lane + j] : 0u;
      lane_sum += cnt[j];
    }
    uint32_t incl = lane_sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= uint32_t(d)) incl += t;
    }
    uint32_t run = incl - lane_sum;
#pragma unroll
    for (int j = 0; j < kRowsPerLane; ++j) {
      if (mine[j]) s_ex[q & 1][kRowsPerLane * lane + j] = run;
      run += cnt[j];
    }
    const unsigned long long agg = __shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned long long* state = op.state[q];
    unsigned long long before = 0;
    if (tile == 0) {
      if (lane == 0) st_relaxed_u64(state, (2ull << 62) | agg);
    } else {
      if (lane == 0) st_relaxed_u64(state + tile, (1ull << 62) | agg);
      int64_t look = int64_t(tile) - 1;
      uint32_t spins = 0;
      for (;;) {
        const int64_t idx = look - lane;  // lane 0: the nearest predecessor
        unsigned long long w = idx >= 0 ? ld_relaxed_u64(state + idx) : (2ull << 62);
        while (__any_sync(0xFFFFFFFFu, (w >> 62) == 0)) {
          if ((w >> 62) == 0) w = ld_relaxed_u64(state + idx);
          if (++spins > kSpinLimit) break;
        }
        if (spins > kSpinLimit) {  // warp-uniform (every lane counts the same rounds)
          if (lane == 0) atomicExch(op.error, 1);
          break;
        }
        const uint32_t done = __ballot_sync(0xFFFFFFFFu, (w >> 62) == 2);
        const uint32_t upto = done ? uint32_t(__ffs(done) - 1) : 31u;
        unsigned long long part = lane <= upto ? (w & kTileMask) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, d);
        before += part;
        if (done) break;
        look -= 32;
      }
      if (lane == 0) st_relaxed_u64(state + tile, (2ull << 62) | ((before + agg) & kTileMask));
    }
    if (lane == 0) {
      s_base[q] = before;
      if (int64_t(before + agg) > op.cap[q]) atomicExch(op.error, 2);
    }
  }
  __syncthreads();

  // ---- 3. fill: the same walk again (the tile is in the L2), row pointers and entries at their final places
  const uint32_t below = (1u << lane) - 1u;
#pragma unroll 1
  for (int it = 0; it < kTileIters; ++it) {
    const uint32_t t0 = (uint32_t(it) * 8 + warp) * kRowsPerWarp;
    WarpRows w;
    load_warp_rows(w, tile_row0 + t0, lane, n_rows, row_ptr, col_idx, values, cls_pos);
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const uint32_t rcls = w.rc[r] >> 30, rloc = w.rc[r] & 0x3FFFFFFFu;
      if (rcls == kClsNone) continue;  // warp-uniform
      const int qa = rcls == kClsA ? 0 : 2, qb = qa + 1;
      const int64_t wa = int64_t(s_base[qa]) + s_ex[0][t0 + r], wb = int64_t(s_base[qb]) + s_ex[1][t0 + r];
      if (lane == 0) {
        out.row_ptr[qa][rloc] = wa;
        out.row_ptr[qb][rloc] = wb;
      }
      if (wa + s_n[0][t0 + r] > op.cap[qa] || wb + s_n[1][t0 + r] > op.cap[qb]) continue;  // flagged in step 2: nothing goes out of bounds
      if (!w.all_short) {
        fill_long_row(lane, qa, wa, wb, w.begin[r], w.end[r], col_idx, values, cls_pos, out);
        continue;
      }
      uint32_t ka = 0, kb = 0;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const bool in_a = (w.cc[r][k] >> 30) == kClsA, in_b = (w.cc[r][k] >> 30) == kClsB;
        const uint32_t ma = __ballot_sync(0xFFFFFFFFu, in_a), mb = __ballot_sync(0xFFFFFFFFu, in_b);
        if (in_a) {
          const int64_t pos = wa + ka + __popc(ma & below);
          out.col[qa][pos] = int32_t(w.cc[r][k] & 0x3FFFFFFFu);
          out.val[qa][pos] = w.v[r][k];
        } else if (in_b) {
          const int64_t pos = wb + kb + __popc(mb & below);
          out.col[qb][pos] = int32_t(w.cc[r][k] & 0x3FFFFFFFu);
          out.val[qb][pos] = w.v[r][k];
        }
        ka += __popc(ma);
        kb += __popc(mb);
      }
    }
  }
}

// closes the four row-pointer arrays with the totals of the last tile and hands the totals to the host
__global__ void onepass_finish_kernel(uint32_t n_tiles, OnePass op, SepOut out, int64_t n_aa, int64_t n_bb,
                                      unsigned long long* __restrict__ totals) {
  const int q = threadIdx.x;
  if (q >= 4) return;
  const unsigned long long total = n_tiles ? (op.state[q][n_tiles - 1] & kTileMask) : 0ull;
  out.row_ptr[q][q < 2 ? n_aa : n_bb] = int64_t(total);
  totals[q] = total;
}

// structural entries of the constrained rows: with the symmetric block pattern, no quadrant but aa holds more
__global__ void b_rows_entries_kernel(int64_t n_rows, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ is_b,
                                      unsigned long long* __restrict__ sum) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  unsigned long long len = (i < n_rows && is_b[i]) ? (unsigned long long)(row_ptr[i + 1] - row_ptr[i]) : 0ull;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) len += __shfl_xor_sync(0xFFFFFFFFu, len, d);
  if ((threadIdx.x & 31u) == 0 && len) atomicAdd(sum, len);
}

// b = R_a - K_ab u_b, one thread per a-row, summed in row order (find_b_sparse)
__global__ void rhs_kernel(int64_t n_aa, const int64_t* __restrict__ aa_idx, const int64_t* __restrict__ bb_idx,
                           const double* __restrict__ force, const double* __restrict__ disp,
                           const int64_t* __restrict__ ab_ptr, const int32_t* __restrict__ ab_col,
                           const double* __restrict__ ab_val, double* __restrict__ b) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_aa) return;
  double acc = force[aa_idx[i]];
  for (int64_t p = ab_ptr[i]; p < ab_ptr[i + 1]; ++p) acc = acc - ab_val[p] * disp[bb_idx[ab_col[p]]];
  b[i] = acc;
}

// dense row-major copy of one CSR quadrant (SeparatedStiffnessMatrix's k_aa_matrix ... of the reference): one thread per row
__global__ void densify_kernel(int64_t rows, int64_t cols, const int64_t* __restrict__ row_ptr,
                               const int32_t* __restrict__ col, const double* __restrict__ val, double* __restrict__ out) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (int64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) out[i * cols + col[p]] = val[p];
}

__global__ void iota_kernel(uint32_t n, uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

// contributions sorted by global DOF index (stable: insertion order inside a run); the thread at the
// head of a run adds the run to the forces vector one value after the other, like the reference's
// sequence of `+=` (methods_for_bc_data_handle.rs:79-101,126-172)
__global__ void run_sum_kernel(uint32_t m, const uint32_t* __restrict__ key, const uint32_t* __restrict__ pos,
                               const double* __restrict__ val, double* __restrict__ force) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= m) return;
  const uint32_t k = key[p];
  if (k == 0xFFFFFFFFu || (p > 0 && key[p - 1] == k)) return;
  double acc = force[k];
  for (uint32_t q = p; q < m && key[q] == k; ++q) acc += val[pos[q]];
  force[k] = acc;
}

const char* dof_name(int d) {
  static const char* names[6] = {"X", "Y", "Z", "ThX", "ThY", "ThZ"};  // DOFParameter's {:?}
  return names[d];
}

void ensure_bc(Handle* h) {
  const size_t n = size_t(h->nodes_number) * 6;
  if (h->bc.constrained.size() != n) {
    h->bc.constrained.assign(n, 0);
    h->bc.displacement.assign(n, 0.0);
    h->bc.force.assign(n, 0.0);
  }
}

template <typename T>
cudaError_t scan_i32_to_i64(const int32_t* in, int64_t* out, int64_t n, cudaStream_t s) {
  cub::TransformInputIterator<int64_t, CastI64, const int32_t*> it(in, CastI64());
  size_t tb = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tb, it, out, n, s);
  if (e != cudaSuccess) return e;
  void* t = nullptr;
  if ((e = pool_malloc(&t, tb ? tb : 16, s)) != cudaSuccess) return e;
  e = cub::DeviceScan::ExclusiveSum(t, tb, it, out, n, s);
  cudaFreeAsync(t, s);
  return e;
}

cudaError_t scan_i32(const int32_t* in, int32_t* out, int64_t n, cudaStream_t s) {
  size_t tb = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, s);
  if (e != cudaSuccess) return e;
  void* t = nullptr;
  if ((e = pool_malloc(&t, tb ? tb : 16, s)) != cudaSuccess) return e;
  e = cub::DeviceScan::ExclusiveSum(t, tb, in, out, n, s);
  cudaFreeAsync(t, s);
  return e;
}

}  // namespace

void bc_clear(Handle* h) {
  h->bc.constrained.clear();
  h->bc.displacement.clear();
  h->bc.force.clear();
  h->bc.load_family.clear();
  h->bc.load_dof.clear();
  h->bc.load_elem.clear();
  h->bc.load_value.clear();
  h->bc.uploaded = false;
  h->sep.valid = false;
  sol_invalidate(h);
}

void sep_release(Handle* h) {
  Handle::Separated& S = h->sep;
  S.d_constrained.release(); S.d_disp.release(); S.d_force.release(); S.cls_pos.release();
  S.aa_idx.release(); S.bb_idx.release(); S.rhs.release(); S.sky.release(); S.maxa.release(); S.sky_a.release();
  S.sky_valid = false;
  for (auto& t : S.tmp) t.release();
  S.onepass.release();
  for (int q = 0; q < 4; ++q) {
    S.row_ptr[q].release(); S.col[q].release(); S.val[q].release();
  }
  for (auto& e : S.ev) {
    if (e) cudaEventDestroy(e);
    e = nullptr;
  }
  S.valid = false;
  h->bc.uploaded = false;
}

// n x FEM::add_displacement / add_concentrated_load, prefix semantics like the element batches
int32_t bc_add(Handle* h, bool displacement, size_t n, const uint32_t* node_number, const int32_t* dof,
               const double* value) {
  if (n == 0) return 0;
  if (!node_number || !dof || !value) return h->fail(FEMGPU_ERR_USAGE, "null boundary-condition array");
  ensure_bc(h);
  for (size_t k = 0; k < n; ++k) {
    if (dof[k] < 0 || dof[k] > 5) return h->fail(FEMGPU_ERR_USAGE, "dof parameter must be 0..5 (X, Y, Z, ThX, ThY, ThZ)");
    uint32_t idx;
    if (!h->node_by_number.find(node_number[k], &idx))
      // check_node_exist, methods_for_node_data_handle.rs:80-86
      return h->fail(FEMGPU_E_NODE_NOT_EXIST, "Node with number " + std::to_string(node_number[k]) + " does not exist!");
    const size_t i = size_t(idx) * 6 + size_t(dof[k]);
    if (displacement) {
      if (h->bc.constrained[i])  // methods_for_bc_data_handle.rs:190-194
        return h->fail(FEMGPU_E_DISPLACEMENT_EXISTS, std::string("Displacement ") + dof_name(dof[k]) +
                                                         " already applied to node " + std::to_string(node_number[k]) + "!");
      h->bc.constrained[i] = 1;
      h->bc.displacement[i] = value[k];
    } else if (h->bc.load_family.empty()) {
      h->bc.force[i] += value[k];  // methods_for_bc_data_handle.rs:47-53
    } else {
      // distributed loads were recorded before this call: keep the reference's `+=` sequence per DOF by
      // queueing the concentrated load behind them in the same ordered list (kind 3: elem = node index)
      h->bc.load_family.push_back(3);
      h->bc.load_elem.push_back(idx);
      h->bc.load_dof.push_back(dof[k]);
      h->bc.load_value.push_back(value[k]);
    }
    h->bc.uploaded = false;
    h->sep.valid = false;
    sol_invalidate(h);
  }
  return 0;
}

// n x FEM::add_uniformly_distributed_line_load / add_uniformly_distributed_surface_load
// (methods_for_bc_data_handle.rs:58-102, :104-173): recorded on the host, evaluated at the next flush
int32_t load_add(Handle* h, int family, size_t n, const uint32_t* number, const int32_t* dof, const double* value) {
  if (n == 0) return 0;
  if (!number || !dof || !value) return h->fail(FEMGPU_ERR_USAGE, "null load array");
  for (size_t k = 0; k < n; ++k) {
    if (dof[k] < 0 || dof[k] > 5) return h->fail(FEMGPU_ERR_USAGE, "dof parameter must be 0..5 (X, Y, Z, ThX, ThY, ThZ)");
    uint32_t idx;
    if (!h->fh[family].by_number.find(number[k], &idx))
      // check_beam_element_exist / check_plate_element_exist
      return h->fail(FEMGPU_E_ELEMENT_NOT_EXIST, std::string(family == FEMGPU_BEAM ? "Beam" : "Plate") +
                                                     " element with number " + std::to_string(number[k]) + " does not exist!");
    h->bc.load_family.push_back(family);
    h->bc.load_elem.push_back(idx);
    h->bc.load_dof.push_back(dof[k]);
    h->bc.load_value.push_back(value[k]);
    h->bc.uploaded = false;
    h->sep.valid = false;
    sol_invalidate(h);
  }
  return 0;
}

// host boundary-condition state -> device: constraint flags, displacements, and the forces vector =
// concentrated loads + the nodal equivalents of the distributed loads (evaluated here, on the device)
int32_t forces_flush(Handle* h) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  Handle::Separated& S = h->sep;
  const int64_t n = 6 * int64_t(h->nodes_number);
  ensure_bc(h);
  FEMGPU_CUDA_CHECK(h, S.d_constrained.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, S.d_disp.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, S.d_force.reserve(size_t(n) + 1));
  if (h->bc.uploaded || n == 0) return 0;
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(S.d_constrained.p, h->bc.constrained.data(), size_t(n), cudaMemcpyHostToDevice, s));
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(S.d_disp.p, h->bc.displacement.data(), size_t(n) * 8, cudaMemcpyHostToDevice, s));
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(S.d_force.p, h->bc.force.data(), size_t(n) * 8, cudaMemcpyHostToDevice, s));
  const size_t nl = h->bc.load_family.size();
  if (nl) {
    if (4 * nl >= (size_t(1) << 32) || uint64_t(n) >= 0xFFFFFFFFull)
      return h->fail(FEMGPU_ERR_LIMIT, "too many distributed loads / degrees of freedom for 32-bit load keys");
    int32_t st = upload_pending(h);  // the loaded elements' connectivity and the nodes must be on the device
    if (st) return st;
    const uint32_t m = uint32_t(4 * nl);
    DevBuf<int32_t> d_fam, d_dof;
    DevBuf<uint32_t> d_elem, key_a, key_b, pos_a, pos_b;
    DevBuf<double> d_value, val;
    DevBuf<uint8_t> tmp;
    auto tie = [&](auto& b) { b.stream = &h->stream; };
    tie(d_fam); tie(d_dof); tie(d_elem); tie(key_a); tie(key_b); tie(pos_a); tie(pos_b); tie(d_value); tie(val); tie(tmp);
    struct Guard {
      std::function<void()> f;
      ~Guard() { f(); }
    } guard{[&] {
      d_fam.release(); d_dof.release(); d_elem.release(); key_a.release(); key_b.release(); pos_a.release();
      pos_b.release(); d_value.release(); val.release(); tmp.release();
    }};
    FEMGPU_CUDA_CHECK(h, d_fam.reserve(nl));
    FEMGPU_CUDA_CHECK(h, d_dof.reserve(nl));
    FEMGPU_CUDA_CHECK(h, d_elem.reserve(nl));
    FEMGPU_CUDA_CHECK(h, d_value.reserve(nl));
    FEMGPU_CUDA_CHECK(h, key_a.reserve(m));
    FEMGPU_CUDA_CHECK(h, key_b.reserve(m));
    FEMGPU_CUDA_CHECK(h, pos_a.reserve(m));
    FEMGPU_CUDA_CHECK(h, pos_b.reserve(m));
    FEMGPU_CUDA_CHECK(h, val.reserve(m));
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(d_fam.p, h->bc.load_family.data(), nl * 4, cudaMemcpyHostToDevice, s));
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(d_dof.p, h->bc.load_dof.data(), nl * 4, cudaMemcpyHostToDevice, s));
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(d_elem.p, h->bc.load_elem.data(), nl * 4, cudaMemcpyHostToDevice, s));
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(d_value.p, h->bc.load_value.data(), nl * 8, cudaMemcpyHostToDevice, s));
    if ((st = run_load_kernel(h, uint32_t(nl), d_fam.p, d_elem.p, d_dof.p, d_value.p, key_a.p, val.p))) return st;
    iota_kernel<<<div_up(m, 256), 256, 0, s>>>(m, pos_a.p);
    int bits = 1;
    while (bits < 32 && (uint64_t(1) << bits) <= uint64_t(n)) ++bits;
    bits = 32;  // the padding key 0xFFFFFFFF must sort last
    size_t tb = 0;
    FEMGPU_CUDA_CHECK(h, cub::DeviceRadixSort::SortPairs(nullptr, tb, key_a.p, key_b.p, pos_a.p, pos_b.p, int(m), 0, bits, s));
    FEMGPU_CUDA_CHECK(h, tmp.reserve(tb + 16));
    FEMGPU_CUDA_CHECK(h, cub::DeviceRadixSort::SortPairs(tmp.p, tb, key_a.p, key_b.p, pos_a.p, pos_b.p, int(m), 0, bits, s));
    run_sum_kernel<<<div_up(m, 256), 256, 0, s>>>(m, key_b.p, pos_b.p, val.p, S.d_force.p);
    h->launches += 3;
    FEMGPU_CUDA_CHECK(h, cudaGetLastError());
    FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(s));
  }
  h->bc.uploaded = true;
  return 0;
}

int32_t run_separate(Handle* h, bool direct) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  Handle::Separated& S = h->sep;
  const int64_t n = h->n_rows;
  ensure_bc(h);
  if (!S.ev[0]) {
    FEMGPU_CUDA_CHECK(h, cudaEventCreate(&S.ev[0]));
    FEMGPU_CUDA_CHECK(h, cudaEventCreate(&S.ev[1]));
  }
  cudaEvent_t e0 = S.ev[0], e1 = S.ev[1];
  FEMGPU_CUDA_CHECK(h, cudaEventRecord(e0, s));
  {
    int32_t st = forces_flush(h);
    if (st) return st;
  }
  S.valid = false;
  S.sky_valid = false;
  sol_invalidate(h);
  S.n_aa = S.n_bb = 0;
  for (auto& z : S.nnz) z = 0;

  // ---- 1. classes, 2. local numbering (scratch arrays stay with the handle for the next call)
  DevBuf<int32_t>&is_a = S.tmp[0], &is_b = S.tmp[1], &pos_a = S.tmp[2], &pos_b = S.tmp[3];
  FEMGPU_CUDA_CHECK(h, is_a.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, is_b.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, pos_a.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, pos_b.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, h->d_flag.reserve(16));
  unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(h->d_flag.p);
  unsigned long long none = ~0ull;
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(d_bad, &none, 8, cudaMemcpyHostToDevice, s));
  FEMGPU_CUDA_CHECK(h, cudaMemsetAsync(is_a.p + n, 0, 4, s));
  FEMGPU_CUDA_CHECK(h, cudaMemsetAsync(is_b.p + n, 0, 4, s));
  if (n) {
    classify_kernel<<<div_up(n, 256), 256, 0, s>>>(n, h->row_ptr.p, h->col_idx.p, h->values.p, S.d_constrained.p,
                                                  direct ? S.d_force.p : nullptr, is_a.p, is_b.p, d_bad);
    h->launches++;
  }
  FEMGPU_CUDA_CHECK(h, cudaMemsetAsync(d_bad + 1, 0, 8, s));
  if (n) {
    b_rows_entries_kernel<<<div_up(n, 256), 256, 0, s>>>(n, h->row_ptr.p, is_b.p, d_bad + 1);
    h->launches++;
  }
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  FEMGPU_CUDA_CHECK(h, scan_i32(is_a.p, pos_a.p, n + 1, s));
  FEMGPU_CUDA_CHECK(h, scan_i32(is_b.p, pos_b.p, n + 1, s));
  unsigned long long bad = 0, b_entries = 0;
  int32_t n_aa = 0, n_bb = 0;
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(&b_entries, d_bad + 1, 8, cudaMemcpyDeviceToHost, s));
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, s));
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(&n_aa, pos_a.p + n, 4, cudaMemcpyDeviceToHost, s));
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(&n_bb, pos_b.p + n, 4, cudaMemcpyDeviceToHost, s));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(s));
  if (bad != ~0ull) {
    // methods_for_separate_stiffness_matrix.rs:233-243
    const size_t node = size_t(bad / 6);
    const size_t local = node - h->node_index_base;
    const uint32_t number = node >= h->node_index_base && local < h->n_nodes() ? h->node_number[local] : 0u;
    if (direct && !h->bc.constrained[size_t(bad)])  // :49-58
      return h->fail(FEMGPU_E_NO_STIFFNESS_FOR_LOAD, std::string("There are no stiffness to withstand load ") +
                                                         dof_name(int(bad % 6)) + " applied to node " +
                                                         std::to_string(number) + "!");
    return h->fail(FEMGPU_E_NO_STIFFNESS_FOR_DISPLACEMENT,
                   std::string("There are no stiffness to withstand displacement ") + dof_name(int(bad % 6)) +
                       " applied to node " + std::to_string(number) + "!");
  }
  if (n_bb == 0)  // :257-259 ("No restraints"), :87-89 ("There are no restraints applied!")
    return h->fail(FEMGPU_E_NO_RESTRAINTS, direct ? "There are no restraints applied!" : "No restraints");
  if (n >= (int64_t(1) << 30)) return h->fail(FEMGPU_ERR_LIMIT, "separation supports < 2^30 degrees of freedom");

  FEMGPU_CUDA_CHECK(h, S.cls_pos.reserve(size_t(n)));
  FEMGPU_CUDA_CHECK(h, S.aa_idx.reserve(size_t(n_aa) + 1));
  FEMGPU_CUDA_CHECK(h, S.bb_idx.reserve(size_t(n_bb) + 1));
  number_kernel<<<div_up(n, 256), 256, 0, s>>>(n, is_a.p, is_b.p, pos_a.p, pos_b.p, S.cls_pos.p, S.aa_idx.p,
                                              S.bb_idx.p);
  h->launches++;

  const int64_t rows_q[4] = {n_aa, n_aa, n_bb, n_bb};
  SepOut out{};
  // ---- 3 + 4 in one pass over K (quadrant_onepass_kernel), on request: FEMGPU_SEP_ONE_PASS=1. Same result, K out of
  // HBM once instead of twice — and slower than count + fill on every configuration measured (16.4 ms against 12.7 on
  // config M, profiles/README.md round 2): the second walk over a tile pays the same dependent load chains out of the
  // L2, and a CTA waits on two barriers and its look-back in between. Kept with its tests as the starting point for
  // a tile staged in shared memory by bulk copies. The four outputs are sized by upper bounds: K_aa by the entries
  // != 0.0 of K when the compacted copy of this pass exists (femgpu_get_nonzero_csr), else by the structural entries;
  // the other three by the structural entries of the constrained rows. Not enough memory for that, a bound that
  // does not hold or a scan that does not finish -> the two passes below.
  S.last_onepass = false;
  const char* one_pass_env = getenv("FEMGPU_SEP_ONE_PASS");
  if (one_pass_env && atoi(one_pass_env) == 1 && n > 0) {
    const uint32_t n_tiles = uint32_t(div_up(n, kTileRows));
    int64_t all = h->nnz;
    if (h->nz_valid && h->nz_pass == h->n_numeric) all = std::min(all, h->nz_count);
    const int64_t side = std::min<int64_t>(all, int64_t(b_entries));
    OnePass op{};
    const int64_t cap[4] = {all, side, side, side};
    bool room = S.onepass.reserve(4 * size_t(n_tiles) + 16) == cudaSuccess;
    for (int q = 0; q < 4 && room; ++q) {
      room = S.row_ptr[q].reserve(size_t(rows_q[q]) + 1) == cudaSuccess && S.col[q].reserve(size_t(cap[q]) + 1) == cudaSuccess &&
             S.val[q].reserve(size_t(cap[q]) + 1) == cudaSuccess;
      op.cap[q] = cap[q];
    }
    if (!room) cudaGetLastError();
    if (room) {
      FEMGPU_CUDA_CHECK(h, cudaMemsetAsync(S.onepass.p, 0, (4 * size_t(n_tiles) + 16) * 8, s));
      unsigned long long* tail = S.onepass.p + 4 * size_t(n_tiles);  // [0] ticket, [1] error, [2..5] totals
      for (int q = 0; q < 4; ++q) {
        op.state[q] = S.onepass.p + size_t(q) * n_tiles;
        out.row_ptr[q] = S.row_ptr[q].p;
        out.col[q] = S.col[q].p;
        out.val[q] = S.val[q].p;
      }
      op.ticket = reinterpret_cast<unsigned int*>(tail);
      op.error = reinterpret_cast<int*>(tail + 1);
      quadrant_onepass_kernel<<<n_tiles, 256, 0, s>>>(n, n_tiles, h->row_ptr.p, h->col_idx.p, h->values.p, S.cls_pos.p, out, op);
      onepass_finish_kernel<<<1, 32, 0, s>>>(n_tiles, op, out, n_aa, n_bb, tail + 2);
      h->launches += 2;
      FEMGPU_CUDA_CHECK(h, cudaGetLastError());
      unsigned long long back[5] = {0, 0, 0, 0, 0};
      FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(back, tail + 1, sizeof back, cudaMemcpyDeviceToHost, s));
      FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(s));
      if (int(back[0] & 0xFFFFFFFFull) == 0) {
        for (int q = 0; q < 4; ++q) S.nnz[q] = int64_t(back[1 + q]);
        S.last_onepass = true;
      } else if (getenv("FEMGPU_ASM_INFO")) {
        fprintf(stderr, "[femgpu separate] one-pass kernel gave up (flag %d): two passes\n", int(back[0] & 0xFFFFFFFFull));
      }
    }
  }
  if (S.last_onepass && !direct && S.nnz[0] == 0)
    return h->fail(FEMGPU_E_KAA_EMPTY, "Sparse separation: K_aa is empty (structure has no free stiffness?)");

  if (!S.last_onepass) {
  // ---- 3. counts per row and quadrant -> row pointers
  DevBuf<int32_t>* cnt = S.tmp + 4;
  for (int q = 0; q < 4; ++q) {
    FEMGPU_CUDA_CHECK(h, cnt[q].reserve(size_t(rows_q[q]) + 1));
    FEMGPU_CUDA_CHECK(h, cudaMemsetAsync(cnt[q].p, 0, (size_t(rows_q[q]) + 1) * 4, s));
    FEMGPU_CUDA_CHECK(h, S.row_ptr[q].reserve(size_t(rows_q[q]) + 1));
    out.cnt[q] = cnt[q].p;
    out.row_ptr[q] = S.row_ptr[q].p;
  }
  const uint32_t warp_grid = div_up(uint64_t(div_up(n, kRowsPerWarp)) * 32, 256);
  quadrant_kernel<false><<<warp_grid, 256, 0, s>>>(n, h->row_ptr.p, h->col_idx.p, h->values.p, S.cls_pos.p, out);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  for (int q = 0; q < 4; ++q) {
    FEMGPU_CUDA_CHECK(h, scan_i32_to_i64<int>(cnt[q].p, S.row_ptr[q].p, rows_q[q] + 1, s));
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(&S.nnz[q], S.row_ptr[q].p + rows_q[q], 8, cudaMemcpyDeviceToHost, s));
  }
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(s));
  if (!direct && S.nnz[0] == 0)  // :303-307 (the direct variant, :63-215, has no such check)
    return h->fail(FEMGPU_E_KAA_EMPTY, "Sparse separation: K_aa is empty (structure has no free stiffness?)");

  // ---- 4. fill
  for (int q = 0; q < 4; ++q) {
    FEMGPU_CUDA_CHECK(h, S.col[q].reserve(size_t(S.nnz[q]) + 1));
    FEMGPU_CUDA_CHECK(h, S.val[q].reserve(size_t(S.nnz[q]) + 1));
    out.col[q] = S.col[q].p;
    out.val[q] = S.val[q].p;
  }
  quadrant_kernel<true><<<warp_grid, 256, 0, s>>>(n, h->row_ptr.p, h->col_idx.p, h->values.p, S.cls_pos.p, out);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  }

  // ---- b = R_a - K_ab u_b
  FEMGPU_CUDA_CHECK(h, S.rhs.reserve(size_t(n_aa) + 1));
  rhs_kernel<<<div_up(n_aa, 256), 256, 0, s>>>(n_aa, S.aa_idx.p, S.bb_idx.p, S.d_force.p, S.d_disp.p, S.row_ptr[1].p,
                                              S.col[1].p, S.val[1].p, S.rhs.p);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  FEMGPU_CUDA_CHECK(h, cudaEventRecord(e1, s));
  FEMGPU_CUDA_CHECK(h, cudaEventSynchronize(e1));
  FEMGPU_CUDA_CHECK(h, cudaEventElapsedTime(&S.last_ms, e0, e1));
  S.n_aa = n_aa;
  S.n_bb = n_bb;
  S.valid = true;
  return 0;
}

// ---- direct separation: skyline of K_aa ---------------------------------------------------------------
// FEM::separate_stiffness_matrix_direct (methods_for_separate_stiffness_matrix.rs:63-215) walks a DENSE copy of
// K to build k_aa_skyline[j] = the largest j - i with K_aa[i, j] != 0 (i < j) and dense quadrants; its only
// consumer, find_ua_vector_direct (methods_for_global_analysis.rs:161-187), turns K_aa into the compacted
// column form (a, maxa) of convert_k_aa_into_compacted_form (:50-80): column j holds K_aa[j, j], K_aa[j-1, j],
// ..., K_aa[j - skyline[j], j] from maxa[j]. Here both come straight from the CSR quadrant, no dense detour.
namespace {

// one thread per K_aa row i: every stored entry (i, j), j > i, is != 0 by construction
__global__ void skyline_height_kernel(int64_t n_aa, const int64_t* __restrict__ rp, const int32_t* __restrict__ ci,
                                      int32_t* __restrict__ sky) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_aa) return;
  for (int64_t p = rp[i]; p < rp[i + 1]; ++p) {
    const int32_t j = ci[p];
    if (j > int32_t(i)) atomicMax(sky + j, j - int32_t(i));  // integer max: order independent
  }
}

__global__ void skyline_len_kernel(int64_t n_aa, const int32_t* __restrict__ sky, int32_t* __restrict__ len) {
  int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j <= n_aa) len[j] = j < n_aa ? sky[j] + 1 : 0;
}

__global__ void skyline_fill_kernel(int64_t n_aa, const int64_t* __restrict__ rp, const int32_t* __restrict__ ci,
                                    const double* __restrict__ va, const int64_t* __restrict__ maxa,
                                    double* __restrict__ a) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_aa) return;
  for (int64_t p = rp[i]; p < rp[i + 1]; ++p) {
    const int32_t j = ci[p];
    if (j >= int32_t(i)) a[maxa[j] + (j - int32_t(i))] = va[p];
  }
}

}  // namespace

int32_t run_skyline(Handle* h) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  Handle::Separated& S = h->sep;
  const int64_t n = S.n_aa;
  S.sky_valid = false;
  FEMGPU_CUDA_CHECK(h, S.sky.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, S.maxa.reserve(size_t(n) + 2));
  DevBuf<int32_t>& len = S.tmp[0];  // the class flags of run_separate are no longer needed
  FEMGPU_CUDA_CHECK(h, len.reserve(size_t(n) + 2));
  FEMGPU_CUDA_CHECK(h, cudaMemsetAsync(S.sky.p, 0, (size_t(n) + 1) * 4, s));
  skyline_height_kernel<<<div_up(n, 256), 256, 0, s>>>(n, S.row_ptr[0].p, S.col[0].p, S.sky.p);
  skyline_len_kernel<<<div_up(n + 1, 256), 256, 0, s>>>(n, S.sky.p, len.p);
  FEMGPU_CUDA_CHECK(h, scan_i32_to_i64<int>(len.p, S.maxa.p, n + 1, s));
  int64_t total = 0;
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(&total, S.maxa.p + n, 8, cudaMemcpyDeviceToHost, s));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(s));
  size_t free_b = 0, total_b = 0;
  FEMGPU_CUDA_CHECK(h, cudaMemGetInfo(&free_b, &total_b));
  if (uint64_t(total) * 8 > uint64_t(free_b) + S.sky_a.cap * 8)
    return h->fail(FEMGPU_ERR_LIMIT, "the skyline of K_aa holds " + std::to_string(total) +
                                         " values: too large for the device (use the sparse separation)");
  FEMGPU_CUDA_CHECK(h, S.sky_a.reserve(size_t(total) + 1));
  FEMGPU_CUDA_CHECK(h, cudaMemsetAsync(S.sky_a.p, 0, size_t(total) * 8, s));
  skyline_fill_kernel<<<div_up(n, 256), 256, 0, s>>>(n, S.row_ptr[0].p, S.col[0].p, S.val[0].p, S.maxa.p, S.sky_a.p);
  h->launches += 3;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  S.sky_total = total;
  S.sky_valid = true;
  return 0;
}

}  // namespace femgpu

using femgpu::Handle;

extern "C" {

static int32_t sep_ready(femgpu_t* h) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0)
    return h->fail(FEMGPU_ERR_NO_DEVICE, "this handle was created without a CUDA device (staging only); "
                                         "femgpu has no CPU fallback");
  if (!h->sep.valid) return h->fail(FEMGPU_ERR_USAGE, "no separated matrix: call femgpu_separate_sparse first");
  return 0;
}

int32_t femgpu_add_displacement(femgpu_t* h, size_t n, const uint32_t* node_number, const int32_t* dof,
                                const double* value) {
  if (!h) return FEMGPU_ERR_USAGE;
  return femgpu::bc_add(h, true, n, node_number, dof, value);
}

int32_t femgpu_add_concentrated_load(femgpu_t* h, size_t n, const uint32_t* node_number, const int32_t* dof,
                                     const double* value) {
  if (!h) return FEMGPU_ERR_USAGE;
  return femgpu::bc_add(h, false, n, node_number, dof, value);
}

// the reference only accepts loads on elements that passed *::create: settle pending validation first
static int32_t settle_validation(femgpu_t* h) {
  return h->device >= 0 ? femgpu_validate(h, nullptr, nullptr, nullptr) : 0;
}

int32_t femgpu_add_line_load(femgpu_t* h, size_t n, const uint32_t* beam_number, const int32_t* dof,
                             const double* value) {
  if (!h) return FEMGPU_ERR_USAGE;
  int32_t st = settle_validation(h);
  if (st) return st;
  return femgpu::load_add(h, FEMGPU_BEAM, n, beam_number, dof, value);
}

int32_t femgpu_add_surface_load(femgpu_t* h, size_t n, const uint32_t* plate_number, const int32_t* dof,
                                const double* value) {
  if (!h) return FEMGPU_ERR_USAGE;
  int32_t st = settle_validation(h);
  if (st) return st;
  return femgpu::load_add(h, FEMGPU_PLATE, n, plate_number, dof, value);
}

int32_t femgpu_get_forces(femgpu_t* h, double* forces, const double** forces_device) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0)
    return h->fail(FEMGPU_ERR_NO_DEVICE, "this handle was created without a CUDA device (staging only); "
                                         "femgpu has no CPU fallback");
  int32_t st = femgpu::forces_flush(h);
  if (st) return st;
  const size_t n = size_t(h->nodes_number) * 6;
  if (forces && n) {
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(forces, h->sep.d_force.p, n * 8, cudaMemcpyDeviceToHost, h->stream));
    FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  }
  if (forces_device) *forces_device = h->sep.d_force.p;
  return 0;
}

int32_t femgpu_separate_sparse(femgpu_t* h, int64_t* n_aa, int64_t* n_bb, int64_t nnz[4]) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0)
    return h->fail(FEMGPU_ERR_NO_DEVICE, "this handle was created without a CUDA device (staging only); "
                                         "femgpu has no CPU fallback");
  if (!h->symbolic_valid || !h->values_valid)
    return h->fail(FEMGPU_ERR_USAGE, "femgpu_separate_sparse needs an assembled matrix (femgpu_assemble)");
  if (h->dist.enabled)
    return h->fail(FEMGPU_ERR_USAGE, "femgpu_separate_sparse is single-GPU for now (the rows of a multi-GPU "
                                     "assembly stay partitioned)");
  int32_t st = femgpu::run_separate(h, false);
  if (st) return st;
  if (n_aa) *n_aa = h->sep.n_aa;
  if (n_bb) *n_bb = h->sep.n_bb;
  if (nnz)
    for (int q = 0; q < 4; ++q) nnz[q] = h->sep.nnz[q];
  return 0;
}

int32_t femgpu_separate_direct(femgpu_t* h, int64_t* n_aa, int64_t* n_bb, int64_t* skyline_values) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0)
    return h->fail(FEMGPU_ERR_NO_DEVICE, "this handle was created without a CUDA device (staging only); "
                                         "femgpu has no CPU fallback");
  if (!h->symbolic_valid || !h->values_valid)
    return h->fail(FEMGPU_ERR_USAGE, "femgpu_separate_direct needs an assembled matrix (femgpu_assemble)");
  if (h->dist.enabled)
    return h->fail(FEMGPU_ERR_USAGE, "femgpu_separate_direct is single-GPU (the rows of a multi-GPU assembly stay partitioned)");
  int32_t st = femgpu::run_separate(h, true);
  if (st) return st;
  if ((st = femgpu::run_skyline(h))) return st;
  if (n_aa) *n_aa = h->sep.n_aa;
  if (n_bb) *n_bb = h->sep.n_bb;
  if (skyline_values) *skyline_values = h->sep.sky_total;
  return 0;
}

int32_t femgpu_get_skyline(femgpu_t* h, int64_t* k_aa_skyline, double* a, int64_t* maxa) {
  int32_t st = sep_ready(h);
  if (st) return st;
  if (!h->sep.sky_valid) return h->fail(FEMGPU_ERR_USAGE, "no skyline: call femgpu_separate_direct first");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  const size_t n = size_t(h->sep.n_aa);
  if (k_aa_skyline && n) {  // widened on the host: the reference's Vec<usize>
    std::vector<int32_t> tmp(n);
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(tmp.data(), h->sep.sky.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < n; ++i) k_aa_skyline[i] = tmp[i];
  }
  if (a && h->sep.sky_total)
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(a, h->sep.sky_a.p, size_t(h->sep.sky_total) * 8, cudaMemcpyDeviceToHost, h->stream));
  if (maxa) FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(maxa, h->sep.maxa.p, (n + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int32_t femgpu_get_separated_indexes(femgpu_t* h, int64_t* k_aa_indexes, int64_t* k_bb_indexes) {
  int32_t st = sep_ready(h);
  if (st) return st;
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  if (k_aa_indexes && h->sep.n_aa)
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(k_aa_indexes, h->sep.aa_idx.p, size_t(h->sep.n_aa) * 8, cudaMemcpyDeviceToHost, h->stream));
  if (k_bb_indexes && h->sep.n_bb)
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(k_bb_indexes, h->sep.bb_idx.p, size_t(h->sep.n_bb) * 8, cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int32_t femgpu_get_separated_csr(femgpu_t* h, int32_t which, int64_t* row_ptr, int32_t* col_idx, double* values) {
  int32_t st = sep_ready(h);
  if (st) return st;
  if (which < 0 || which > 3) return h->fail(FEMGPU_ERR_USAGE, "quadrant must be 0..3 (aa, ab, ba, bb)");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  const int64_t rows = which < 2 ? h->sep.n_aa : h->sep.n_bb, nnz = h->sep.nnz[which];
  if (row_ptr)
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(row_ptr, h->sep.row_ptr[which].p, size_t(rows + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
  if (col_idx && nnz)
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(col_idx, h->sep.col[which].p, size_t(nnz) * 4, cudaMemcpyDeviceToHost, h->stream));
  if (values && nnz)
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(values, h->sep.val[which].p, size_t(nnz) * 8, cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int32_t femgpu_get_separated_dense(femgpu_t* h, int32_t which, double* out) {
  int32_t st = sep_ready(h);
  if (st) return st;
  if (which < 0 || which > 3) return h->fail(FEMGPU_ERR_USAGE, "quadrant must be 0..3 (aa, ab, ba, bb)");
  if (!out) return h->fail(FEMGPU_ERR_USAGE, "null output");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  const int64_t rows = which < 2 ? h->sep.n_aa : h->sep.n_bb, cols = (which == 0 || which == 2) ? h->sep.n_aa : h->sep.n_bb;
  if (rows == 0 || cols == 0) return 0;
  if (rows > (int64_t(1) << 28) / cols)   // 2 GB of doubles: the dense form is for the model sizes the reference's own
    return h->fail(FEMGPU_ERR_LIMIT,      // dense separation could handle
                   "dense quadrant of " + std::to_string(rows) + " x " + std::to_string(cols) +
                       " entries is too large; use femgpu_get_separated_csr / femgpu_get_skyline");
  femgpu::DevBuf<double> dense;
  dense.stream = &h->stream;
  FEMGPU_CUDA_CHECK(h, dense.reserve(size_t(rows) * size_t(cols)));
  cudaError_t e = cudaMemsetAsync(dense.p, 0, size_t(rows) * size_t(cols) * 8, h->stream);
  if (e == cudaSuccess && h->sep.nnz[which]) {
    femgpu::densify_kernel<<<femgpu::div_up(rows, 128), 128, 0, h->stream>>>(rows, cols, h->sep.row_ptr[which].p, h->sep.col[which].p,
                                                                     h->sep.val[which].p, dense.p);
    h->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dense.p, size_t(rows) * size_t(cols) * 8, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  dense.release();
  FEMGPU_CUDA_CHECK(h, e);
  return 0;
}

int32_t femgpu_get_separated_csr_device(femgpu_t* h, int32_t which, const int64_t** row_ptr, const int32_t** col_idx,
                                        const double** values) {
  int32_t st = sep_ready(h);
  if (st) return st;
  if (which < 0 || which > 3) return h->fail(FEMGPU_ERR_USAGE, "quadrant must be 0..3 (aa, ab, ba, bb)");
  if (row_ptr) *row_ptr = h->sep.row_ptr[which].p;
  if (col_idx) *col_idx = h->sep.col[which].p;
  if (values) *values = h->sep.val[which].p;
  return 0;
}

int32_t femgpu_separated_rhs(femgpu_t* h, double* b, const double** b_device) {
  int32_t st = sep_ready(h);
  if (st) return st;
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  if (b && h->sep.n_aa) {
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(b, h->sep.rhs.p, size_t(h->sep.n_aa) * 8, cudaMemcpyDeviceToHost, h->stream));
    FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  }
  if (b_device) *b_device = h->sep.rhs.p;
  return 0;
}

int32_t femgpu_last_separate_path(femgpu_t* h, int32_t* one_pass) {
  int32_t st = sep_ready(h);
  if (st) return st;
  if (one_pass) *one_pass = h->sep.last_onepass ? 1 : 0;
  return 0;
}

int32_t femgpu_last_separate_ms(femgpu_t* h, float* ms) {
  int32_t st = sep_ready(h);
  if (st) return st;
  if (ms) *ms = h->sep.last_ms;
  return 0;
}

}  // extern "C"
