// Global analysis on the device, downstream of the separated matrix in HBM:
//   femgpu_solve_pcg        find_ua_vector_iterative_pcg_jacobi_sparse / ..._pcg_block_jacobi_sparse
//                           (methods_for_global_analysis.rs:189-275; block starts as :121-147)
//   femgpu_global_analysis  find_r_r_vector_sparse (:334-360, find_r_r_sparse :100-137) +
//                           compose_global_analysis_result (:362-385)
//   femgpu_get_global_result  extract_global_analysis_result (:387-...)
//   femgpu_element_results  extract_elements_analysis_result (methods_for_element_analysis.rs:27-58),
//                           kernels in results.cu
//
// The reference delegates the PCG arithmetic to the un-vendored crate iterative_solvers_smpl 0.1.5
// (pcg_jacobi_csr / pcg_block_jacobi_csr): what is restated here is the textbook preconditioned conjugate
// gradient from x0 = 0 with the stopping test ||r||_2 <= max(rel_tol ||b||_2, abs_tol), the iteration count
// being the number of search directions used. PARITY UNPINNED beyond the reference's own test
// (`iterations == 1`, u = 0.0015 on the two-node truss, tests/fem/test_fem.rs:83-225).
//
// Everything is deterministic: SpMV rows are summed by eight lanes in a fixed tree, dot products are
// reduced CTA-wise into a partial array of fixed length and then by one CTA in a fixed order; no atomics.
// All scalars (alpha, beta, r.z) stay on the device; the host reads ||r||^2 once per iteration for the
// stopping test while the next direction update is already queued.
//
// Each iteration is HBM-bound: 12 B per stored entry of K_aa (value + column) + ~10 vector passes of 8 B.
#include "common.cuh"

namespace femgpu {

namespace {

constexpr int kVecThreads = 256;
#ifndef FEMGPU_SPMV_LANES
#define FEMGPU_SPMV_LANES 4
#endif
constexpr int kLanesPerRow = FEMGPU_SPMV_LANES;   // rows of a structural K_aa hold ~20-54 entries
constexpr int kMaxPartials = 2048;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  // fixed-shape tree over the CTA: warp shuffles, then the warp leaders
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = l < int(blockDim.x >> 5) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
  }
  return v;  // valid in thread 0
}

// y = A x (CSR, kLanesPerRow lanes per row), partial[blockIdx] = sum over the CTA's rows of x_i * y_i
__global__ void __launch_bounds__(kVecThreads)
spmv_dot_kernel(int64_t n, const int64_t* __restrict__ rp, const int32_t* __restrict__ ci,
                const double* __restrict__ va, const double* __restrict__ x, double* __restrict__ y,
                double* __restrict__ partial) {
  __shared__ double sh[32];
  const int sub = threadIdx.x % kLanesPerRow;
  const int64_t rows_per_cta = kVecThreads / kLanesPerRow;
  double dot = 0.0;
  const int64_t step = int64_t(gridDim.x) * rows_per_cta;
  int64_t i = int64_t(blockIdx.x) * rows_per_cta + threadIdx.x / kLanesPerRow;
  // the row pointers of the NEXT row are requested before this row's entries are summed: one dependent
  // round trip (row_ptr -> entries -> x) less on the critical path of a thread
  int64_t b = i < n ? rp[i] : 0, e = i < n ? rp[i + 1] : 0;
  for (int64_t base = int64_t(blockIdx.x) * rows_per_cta; base < n; base += step) {
    const int64_t i_next = i + step;
    const int64_t b_next = i_next < n ? rp[i_next] : 0, e_next = i_next < n ? rp[i_next + 1] : 0;
    double acc = 0.0;
    if (i < n)
      for (int64_t p = b + sub; p < e; p += kLanesPerRow) acc += va[p] * __ldg(x + ci[p]);
#pragma unroll
    for (int o = kLanesPerRow / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o, kLanesPerRow);
    if (sub == 0 && i < n) {
      y[i] = acc;
      dot += __ldg(x + i) * acc;
    }
    i = i_next;
    b = b_next;
    e = e_next;
  }
  const double t = block_sum(dot, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// scal layout: 0 rz, 1 pAp, 2 alpha, 3 beta, 4 rr, 5 bb
// op 0: pAp = sum(pa); alpha = rz / pAp
// op 1: rr = sum(pa); rz_new = sum(pb); beta = rz_new / rz; rz = rz_new
// op 2 (initialisation): bb = rr = sum(pa); rz = sum(pb)
__global__ void __launch_bounds__(1024)
scalar_kernel(int op, int m, const double* __restrict__ pa, const double* __restrict__ pb, double* __restrict__ scal) {
  __shared__ double sh[32];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    a += pa[i];
    if (op) b += pb[i];
  }
  a = block_sum(a, sh);
  __syncthreads();
  if (op) b = block_sum(b, sh);
  if (threadIdx.x == 0) {
    if (op == 0) {
      scal[1] = a;
      scal[2] = scal[0] / a;
    } else if (op == 1) {
      scal[4] = a;
      scal[3] = b / scal[0];
      scal[0] = b;
    } else {
      scal[5] = a;
      scal[4] = a;
      scal[0] = b;
    }
  }
}

// z_i = (M^-1 r)_i. Jacobi: minv[i] = 1 / K_ii. Block Jacobi: row i of the inverse of its node's diagonal
// block (minv[6 i + j], blk[i] = first row of the block | size << 28).
template <bool kBlock>
__device__ __forceinline__ double precond(int64_t i, const double* __restrict__ r, double ri,
                                          const double* __restrict__ minv, const uint32_t* __restrict__ blk) {
  if (!kBlock) return minv[i] * ri;
  const uint32_t w = blk[i], start = w & 0x0FFFFFFFu, size = w >> 28;
  double acc = 0.0;
  for (uint32_t j = 0; j < size; ++j) acc += minv[6 * i + j] * r[start + j];
  return acc;
}

// initialisation: x = 0, r = b, z = M^-1 r, p = z; partials of b.b and r.z
template <bool kBlock>
__global__ void __launch_bounds__(kVecThreads)
init_kernel(int64_t n, const double* __restrict__ b, const double* __restrict__ minv, const uint32_t* __restrict__ blk,
            double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, double* __restrict__ p,
            double* __restrict__ pa, double* __restrict__ pb) {
  __shared__ double sh[32];
  double s_bb = 0.0, s_rz = 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double bi = b[i];
    const double zi = precond<kBlock>(i, b, bi, minv, blk);
    x[i] = 0.0;
    r[i] = bi;
    z[i] = zi;
    p[i] = zi;
    s_bb += bi * bi;
    s_rz += bi * zi;
  }
  const double t0 = block_sum(s_bb, sh);
  __syncthreads();
  const double t1 = block_sum(s_rz, sh);
  if (threadIdx.x == 0) {
    pa[blockIdx.x] = t0;
    pb[blockIdx.x] = t1;
  }
}

// x += alpha p; r -= alpha Ap (Jacobi: also z = M^-1 r and the partials of r.r, r.z)
template <bool kBlock>
__global__ void __launch_bounds__(kVecThreads)
update_kernel(int64_t n, const double* __restrict__ scal, const double* __restrict__ p, const double* __restrict__ ap,
              const double* __restrict__ minv, double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
              double* __restrict__ pa, double* __restrict__ pb) {
  __shared__ double sh[32];
  const double alpha = scal[2];
  double s_rr = 0.0, s_rz = 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    x[i] = x[i] + alpha * p[i];
    const double ri = r[i] - alpha * ap[i];
    r[i] = ri;
    if (!kBlock) {
      const double zi = minv[i] * ri;
      z[i] = zi;
      s_rr += ri * ri;
      s_rz += ri * zi;
    }
  }
  if (!kBlock) {
    const double t0 = block_sum(s_rr, sh);
    __syncthreads();
    const double t1 = block_sum(s_rz, sh);
    if (threadIdx.x == 0) {
      pa[blockIdx.x] = t0;
      pb[blockIdx.x] = t1;
    }
  }
}

// block Jacobi: z = M^-1 r needs the whole updated r of a block, hence its own pass
__global__ void __launch_bounds__(kVecThreads)
block_precond_kernel(int64_t n, const double* __restrict__ r, const double* __restrict__ minv,
                     const uint32_t* __restrict__ blk, double* __restrict__ z, double* __restrict__ pa,
                     double* __restrict__ pb) {
  __shared__ double sh[32];
  double s_rr = 0.0, s_rz = 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double ri = r[i];
    const double zi = precond<true>(i, r, ri, minv, blk);
    z[i] = zi;
    s_rr += ri * ri;
    s_rz += ri * zi;
  }
  const double t0 = block_sum(s_rr, sh);
  __syncthreads();
  const double t1 = block_sum(s_rz, sh);
  if (threadIdx.x == 0) {
    pa[blockIdx.x] = t0;
    pb[blockIdx.x] = t1;
  }
}

// p = z + beta p
__global__ void __launch_bounds__(kVecThreads)
direction_kernel(int64_t n, const double* __restrict__ scal, const double* __restrict__ z, double* __restrict__ p) {
  const double beta = scal[3];
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    p[i] = z[i] + beta * p[i];
}

// Jacobi: 1 / K_ii (columns of a row ascend: bisection). A zero or missing diagonal cannot happen for an
// a-row (the separation made it active because K_ii != 0), reported all the same.
__global__ void jacobi_setup_kernel(int64_t n, const int64_t* __restrict__ rp, const int32_t* __restrict__ ci,
                                    const double* __restrict__ va, double* __restrict__ minv, int32_t* __restrict__ bad) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t lo = rp[i], hi = rp[i + 1];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (ci[mid] < int32_t(i)) lo = mid + 1;
    else hi = mid;
  }
  double d = 0.0;
  if (lo < rp[i + 1] && ci[lo] == int32_t(i)) d = va[lo];
  if (d == 0.0) *bad = 1;
  minv[i] = 1.0 / d;
}

// Block Jacobi: the rows of one node are consecutive in K_aa (build_block_starts_from_k_aa_indexes,
// methods_for_global_analysis.rs:121-147). The first row of a block gathers the dense diagonal block
// (<= 6 x 6) from the CSR rows and inverts it by Gauss-Jordan elimination (symmetric positive definite:
// no pivoting); every row keeps its row of the inverse.
__global__ void block_setup_kernel(int64_t n, const int64_t* __restrict__ aa_idx, const int64_t* __restrict__ rp,
                                   const int32_t* __restrict__ ci, const double* __restrict__ va,
                                   double* __restrict__ minv, uint32_t* __restrict__ blk, int32_t* __restrict__ bad) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t node = aa_idx[i] / 6;
  if (i > 0 && aa_idx[i - 1] / 6 == node) return;  // not the first row of its block
  int size = 1;
  while (size < 6 && i + size < n && aa_idx[i + size] / 6 == node) ++size;
  double a[6][6], inv[6][6];
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) {
      a[r][c] = 0.0;
      inv[r][c] = r == c ? 1.0 : 0.0;
    }
  for (int r = 0; r < size; ++r)
    for (int64_t p = rp[i + r]; p < rp[i + r + 1]; ++p) {
      const int64_t c = int64_t(ci[p]) - i;
      if (c >= 0 && c < size) a[r][c] = va[p];
    }
  for (int k = 0; k < size; ++k) {
    const double piv = a[k][k];
    if (piv == 0.0) *bad = 1;
    const double ip = 1.0 / piv;
    for (int c = 0; c < size; ++c) {
      a[k][c] *= ip;
      inv[k][c] *= ip;
    }
    for (int r = 0; r < size; ++r) {
      if (r == k) continue;
      const double f = a[r][k];
      for (int c = 0; c < size; ++c) {
        a[r][c] -= f * a[k][c];
        inv[r][c] -= f * inv[k][c];
      }
    }
  }
  for (int r = 0; r < size; ++r) {
    blk[i + r] = uint32_t(i) | (uint32_t(size) << 28);
    for (int c = 0; c < 6; ++c) minv[6 * (i + r) + c] = c < size ? inv[r][c] : 0.0;
  }
}

// find_r_r_sparse (methods_for_global_analysis.rs:100-137): r_r = K_ba u_a + K_bb u_b - R_b, one thread per
// b-row, each product summed from zero in (row, column) order (the reference's triplet order is hash-map
// order, i.e. unspecified)
__global__ void reactions_kernel(int64_t n_bb, const int64_t* __restrict__ bb_idx, const int64_t* __restrict__ ba_ptr,
                                 const int32_t* __restrict__ ba_col, const double* __restrict__ ba_val,
                                 const int64_t* __restrict__ bb_ptr, const int32_t* __restrict__ bb_col,
                                 const double* __restrict__ bb_val, const double* __restrict__ u_a,
                                 const double* __restrict__ disp, const double* __restrict__ force,
                                 double* __restrict__ r_r) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_bb) return;
  double y_ba = 0.0, y_bb = 0.0;
  // product rounded, then added: the reference's `y[i] = y[i] + a_ij * u` is not contracted into an FMA
  for (int64_t p = ba_ptr[i]; p < ba_ptr[i + 1]; ++p) y_ba = __dadd_rn(y_ba, __dmul_rn(ba_val[p], u_a[ba_col[p]]));
  for (int64_t p = bb_ptr[i]; p < bb_ptr[i + 1]; ++p)
    y_bb = __dadd_rn(y_bb, __dmul_rn(bb_val[p], disp[bb_idx[bb_col[p]]]));
  r_r[i] = y_ba + y_bb - force[bb_idx[i]];
}

// compose_global_analysis_result (:362-385): displacements[k_aa_indexes[i]] = u_a[i], forces[k_bb_indexes[i]] = r_r[i]
__global__ void scatter_kernel(int64_t n, const int64_t* __restrict__ idx, const double* __restrict__ src,
                               double* __restrict__ dst) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = src[i];
}

// ---- direct solve: active-column (skyline) LDL^T, FEM::find_ua_vector_direct --------------------------
// methods_for_global_analysis.rs:161-187 hands (a, maxa) to the un-vendored crate colsol 1.0.1
// (`factorization`, `find_unknown`), i.e. Bathe's COLSOL: column by column, every entry of a column is reduced by
// the dot product of the rows above it with the matching column of L (top to bottom), then the column is divided
// by the pivots and the diagonal reduced; forward reduction, division by D and back-substitution follow. The
// columns — and the entries inside a column — depend on each other, so ONE warp walks them in COLSOL's order and
// only the dot products are spread over its lanes (fixed shuffle tree: deterministic, but not the sequential
// summation order of the scalar code — parity unpinned beyond the reference's 1 x 1 test). This is the path for
// the model sizes the reference's dense separation could handle; large models take the PCG.
// status: 0 ok, n + 1 = "stiffness matrix not positive definite" at column n.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

__global__ void __launch_bounds__(32)
colsol_kernel(int64_t nn, const int64_t* __restrict__ maxa, double* a, double* v, int64_t* __restrict__ status) {
  const int lane = threadIdx.x;
  // factorisation (COLSOL, KKK = 1)
  for (int64_t n = 0; n < nn; ++n) {
    const int64_t kn = maxa[n], kl = kn + 1, ku = maxa[n + 1] - 1, kh = ku - kl;
    if (kh > 0) {
      int64_t k = n - kh, klt = ku;
      for (int64_t ic = 1; ic <= kh; ++ic, ++k) {
        --klt;
        const int64_t ki = maxa[k], nd = maxa[k + 1] - ki - 1;
        if (nd > 0) {
          const int64_t kk = ic < nd ? ic : nd;
          double c = 0.0;
          for (int64_t l = 1 + lane; l <= kk; l += 32) c += a[ki + l] * a[klt + l];
          c = warp_sum(c);
          if (lane == 0) a[klt] -= c;
          __syncwarp();
        }
      }
    }
    if (kh >= 0) {
      double b = 0.0;
      for (int64_t kk = kl + lane; kk <= ku; kk += 32) {
        const int64_t k = n - 1 - (kk - kl);
        const double c = a[kk] / a[maxa[k]];
        b += c * a[kk];
        a[kk] = c;
      }
      b = warp_sum(b);
      if (lane == 0) a[kn] -= b;
      __syncwarp();
    }
    if (!(a[kn] > 0.0)) {
      if (lane == 0) *status = n + 1;
      return;
    }
  }
  // forward reduction of the right-hand side (KKK = 2)
  for (int64_t n = 0; n < nn; ++n) {
    const int64_t kl = maxa[n] + 1, ku = maxa[n + 1] - 1;
    if (ku - kl >= 0) {
      double c = 0.0;
      for (int64_t kk = kl + lane; kk <= ku; kk += 32) c += a[kk] * v[n - 1 - (kk - kl)];
      c = warp_sum(c);
      if (lane == 0) v[n] -= c;
      __syncwarp();
    }
  }
  for (int64_t n = lane; n < nn; n += 32) v[n] = v[n] / a[maxa[n]];
  __syncwarp();
  // back-substitution
  for (int64_t n = nn - 1; n >= 1; --n) {
    const int64_t kl = maxa[n] + 1, ku = maxa[n + 1] - 1;
    const double vn = v[n];
    for (int64_t kk = kl + lane; kk <= ku; kk += 32) v[n - 1 - (kk - kl)] -= a[kk] * vn;
    __syncwarp();
  }
  if (lane == 0) *status = 0;
}

int grid_for(int64_t n, int per_cta, int sm_count) {
  const int64_t want = (n + per_cta - 1) / per_cta;
  const int64_t cap = std::min<int64_t>(kMaxPartials, int64_t(sm_count > 0 ? sm_count : 148) * 8);
  return int(std::max<int64_t>(1, std::min(want, cap)));
}

int32_t need_sep(Handle* h) {
  if (h->device < 0)
    return h->fail(FEMGPU_ERR_NO_DEVICE, "this handle was created without a CUDA device (staging only); "
                                         "femgpu has no CPU fallback");
  if (!h->sep.valid) return h->fail(FEMGPU_ERR_USAGE, "no separated matrix: call femgpu_separate_sparse first");
  return 0;
}

}  // namespace

void sol_release(Handle* h) {
  Handle::Solution& Z = h->sol;
  Z.u_a.release(); Z.r_r.release(); Z.r.release(); Z.z.release(); Z.p.release(); Z.ap.release(); Z.minv.release();
  Z.blk.release(); Z.partial.release(); Z.scal.release(); Z.disp.release(); Z.force.release();
  for (auto& r : Z.res) r.release();
  for (auto& e : Z.ev) {
    if (e) cudaEventDestroy(e);
    e = nullptr;
  }
  Z.ua_valid = Z.composed = Z.disp_valid = false;
}

void sol_invalidate(Handle* h) { h->sol.ua_valid = h->sol.composed = h->sol.disp_valid = false; }

int32_t run_pcg(Handle* h, int preconditioner, int64_t max_iter, int64_t* iterations) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  Handle::Separated& S = h->sep;
  Handle::Solution& Z = h->sol;
  const int64_t n = S.n_aa;
  if (h->sm_count == 0)
    FEMGPU_CUDA_CHECK(h, cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
  if (!Z.ev[0]) {
    FEMGPU_CUDA_CHECK(h, cudaEventCreate(&Z.ev[0]));
    FEMGPU_CUDA_CHECK(h, cudaEventCreate(&Z.ev[1]));
  }
  FEMGPU_CUDA_CHECK(h, cudaEventRecord(Z.ev[0], s));
  Z.ua_valid = Z.composed = false;
  const bool block = preconditioner == 1;
  if (block && n >= (int64_t(1) << 28))  // Z.blk packs (first row of the block | size << 28) into 32 bits
    return h->fail(FEMGPU_ERR_LIMIT, "block-Jacobi PCG supports fewer than 2^28 free degrees of freedom");
  FEMGPU_CUDA_CHECK(h, Z.u_a.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, Z.r.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, Z.z.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, Z.p.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, Z.ap.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, Z.minv.reserve((block ? 6 : 1) * size_t(n) + 1));
  if (block) FEMGPU_CUDA_CHECK(h, Z.blk.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, Z.partial.reserve(2 * kMaxPartials));
  FEMGPU_CUDA_CHECK(h, Z.scal.reserve(8));
  FEMGPU_CUDA_CHECK(h, h->d_flag.reserve(16));
  double* pa = Z.partial.p;
  double* pb = Z.partial.p + kMaxPartials;
  const int64_t* rp = S.row_ptr[0].p;
  const int32_t* ci = S.col[0].p;
  const double* va = S.val[0].p;
  int32_t* d_bad = h->d_flag.p;
  FEMGPU_CUDA_CHECK(h, cudaMemsetAsync(d_bad, 0, 4, s));
  FEMGPU_CUDA_CHECK(h, cudaMemsetAsync(Z.scal.p, 0, 64, s));
  const int g_vec = grid_for(n, kVecThreads, h->sm_count);
  const int g_spmv = grid_for(n, kVecThreads / kLanesPerRow, h->sm_count);
  const uint32_t g_rows = div_up(n, 256);

  if (block)
    block_setup_kernel<<<g_rows, 256, 0, s>>>(n, S.aa_idx.p, rp, ci, va, Z.minv.p, Z.blk.p, d_bad);
  else
    jacobi_setup_kernel<<<g_rows, 256, 0, s>>>(n, rp, ci, va, Z.minv.p, d_bad);
  if (block)
    init_kernel<true><<<g_vec, kVecThreads, 0, s>>>(n, S.rhs.p, Z.minv.p, Z.blk.p, Z.u_a.p, Z.r.p, Z.z.p, Z.p.p, pa, pb);
  else
    init_kernel<false><<<g_vec, kVecThreads, 0, s>>>(n, S.rhs.p, Z.minv.p, nullptr, Z.u_a.p, Z.r.p, Z.z.p, Z.p.p, pa, pb);
  scalar_kernel<<<1, 1024, 0, s>>>(2, g_vec, pa, pb, Z.scal.p);
  h->launches += 3;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  double host[8] = {0};
  int32_t bad = 0;
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(host, Z.scal.p, 64, cudaMemcpyDeviceToHost, s));
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(s));
  if (bad) return h->fail(FEMGPU_E_SOLVER, "PCG: zero pivot in the preconditioner (K_aa diagonal)");
  const double bb = host[5];
  const double tol = std::max(h->rel_tol * std::sqrt(bb), h->abs_tol);
  int64_t it = 0;
  bool converged = std::sqrt(host[4]) <= tol;  // b = 0: u_a = 0
  while (!converged && it < max_iter) {
    spmv_dot_kernel<<<g_spmv, kVecThreads, 0, s>>>(n, rp, ci, va, Z.p.p, Z.ap.p, pa);
    scalar_kernel<<<1, 1024, 0, s>>>(0, g_spmv, pa, pb, Z.scal.p);
    if (block) {
      update_kernel<true><<<g_vec, kVecThreads, 0, s>>>(n, Z.scal.p, Z.p.p, Z.ap.p, Z.minv.p, Z.u_a.p, Z.r.p, Z.z.p, pa, pb);
      block_precond_kernel<<<g_vec, kVecThreads, 0, s>>>(n, Z.r.p, Z.minv.p, Z.blk.p, Z.z.p, pa, pb);
      h->launches++;
    } else {
      update_kernel<false><<<g_vec, kVecThreads, 0, s>>>(n, Z.scal.p, Z.p.p, Z.ap.p, Z.minv.p, Z.u_a.p, Z.r.p, Z.z.p, pa, pb);
    }
    scalar_kernel<<<1, 1024, 0, s>>>(1, g_vec, pa, pb, Z.scal.p);
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(host, Z.scal.p, 64, cudaMemcpyDeviceToHost, s));
    direction_kernel<<<g_vec, kVecThreads, 0, s>>>(n, Z.scal.p, Z.z.p, Z.p.p);
    h->launches += 5;
    FEMGPU_CUDA_CHECK(h, cudaGetLastError());
    FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(s));
    ++it;
    if (!(host[1] > 0.0) || !std::isfinite(host[4]))
      return h->fail(FEMGPU_E_SOLVER, "PCG: breakdown (p.Ap <= 0 or non-finite residual): K_aa is not positive definite");
    converged = std::sqrt(host[4]) <= tol;
  }
  FEMGPU_CUDA_CHECK(h, cudaEventRecord(Z.ev[1], s));
  FEMGPU_CUDA_CHECK(h, cudaEventSynchronize(Z.ev[1]));
  FEMGPU_CUDA_CHECK(h, cudaEventElapsedTime(&Z.last_ms, Z.ev[0], Z.ev[1]));
  Z.iterations = it;
  Z.residual = std::sqrt(host[4]);
  if (iterations) *iterations = it;
  if (!converged)
    return h->fail(FEMGPU_E_SOLVER, "PCG did not converge in " + std::to_string(max_iter) + " iterations (residual " +
                                        std::to_string(Z.residual) + ")");
  Z.ua_valid = true;
  return 0;
}

// r_r = K_ba u_a + K_bb u_b - R_b, then displacements / forces vectors of the model
int32_t run_global_analysis(Handle* h) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  Handle::Separated& S = h->sep;
  Handle::Solution& Z = h->sol;
  const int64_t n = 6 * int64_t(h->nodes_number);
  FEMGPU_CUDA_CHECK(h, Z.r_r.reserve(size_t(S.n_bb) + 1));
  FEMGPU_CUDA_CHECK(h, Z.disp.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, Z.force.reserve(size_t(n) + 1));
  reactions_kernel<<<div_up(S.n_bb, 128), 128, 0, s>>>(S.n_bb, S.bb_idx.p, S.row_ptr[2].p, S.col[2].p, S.val[2].p,
                                                       S.row_ptr[3].p, S.col[3].p, S.val[3].p, Z.u_a.p, S.d_disp.p,
                                                       S.d_force.p, Z.r_r.p);
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(Z.disp.p, S.d_disp.p, size_t(n) * 8, cudaMemcpyDeviceToDevice, s));
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(Z.force.p, S.d_force.p, size_t(n) * 8, cudaMemcpyDeviceToDevice, s));
  scatter_kernel<<<div_up(S.n_aa, 256), 256, 0, s>>>(S.n_aa, S.aa_idx.p, Z.u_a.p, Z.disp.p);
  scatter_kernel<<<div_up(S.n_bb, 256), 256, 0, s>>>(S.n_bb, S.bb_idx.p, Z.r_r.p, Z.force.p);
  h->launches += 3;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  Z.composed = true;
  Z.disp_valid = true;
  return 0;
}

// K_aa u_a = b by the skyline LDL^T on copies of (a, b): u_a replaces any earlier solution
int32_t run_colsol(Handle* h) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  Handle::Separated& S = h->sep;
  Handle::Solution& Z = h->sol;
  const int64_t n = S.n_aa;
  Z.ua_valid = Z.composed = false;
  if (!Z.ev[0]) {
    FEMGPU_CUDA_CHECK(h, cudaEventCreate(&Z.ev[0]));
    FEMGPU_CUDA_CHECK(h, cudaEventCreate(&Z.ev[1]));
  }
  FEMGPU_CUDA_CHECK(h, cudaEventRecord(Z.ev[0], s));
  FEMGPU_CUDA_CHECK(h, Z.u_a.reserve(size_t(n) + 1));
  FEMGPU_CUDA_CHECK(h, Z.ap.reserve(size_t(S.sky_total) + 1));  // the factor overwrites a copy of `a`
  FEMGPU_CUDA_CHECK(h, h->d_flag.reserve(16));
  int64_t* d_status = reinterpret_cast<int64_t*>(h->d_flag.p);
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(Z.ap.p, S.sky_a.p, size_t(S.sky_total) * 8, cudaMemcpyDeviceToDevice, s));
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(Z.u_a.p, S.rhs.p, size_t(n) * 8, cudaMemcpyDeviceToDevice, s));
  colsol_kernel<<<1, 32, 0, s>>>(n, S.maxa.p, Z.ap.p, Z.u_a.p, d_status);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  int64_t status = -1;
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(&status, d_status, 8, cudaMemcpyDeviceToHost, s));
  FEMGPU_CUDA_CHECK(h, cudaEventRecord(Z.ev[1], s));
  FEMGPU_CUDA_CHECK(h, cudaEventSynchronize(Z.ev[1]));
  FEMGPU_CUDA_CHECK(h, cudaEventElapsedTime(&Z.last_ms, Z.ev[0], Z.ev[1]));
  Z.iterations = 0;
  Z.residual = 0.0;
  if (status != 0)
    return h->fail(FEMGPU_E_SOLVER, "Direct solve: stiffness matrix is not positive definite (pivot of equation " +
                                        std::to_string(status) + " is not positive)");
  Z.ua_valid = true;
  return 0;
}

}  // namespace femgpu

using femgpu::Handle;

extern "C" {

int32_t femgpu_solve_pcg(femgpu_t* h, int32_t preconditioner, int64_t max_iter, int64_t* iterations) {
  if (!h) return FEMGPU_ERR_USAGE;
  int32_t st = femgpu::need_sep(h);
  if (st) return st;
  if (preconditioner != 0 && preconditioner != 1)
    return h->fail(FEMGPU_ERR_USAGE, "preconditioner must be 0 (Jacobi) or 1 (block Jacobi)");
  if (max_iter < 0) return h->fail(FEMGPU_ERR_USAGE, "max_iter must be >= 0");
  return femgpu::run_pcg(h, preconditioner, max_iter, iterations);
}

int32_t femgpu_solve_direct(femgpu_t* h) {
  if (!h) return FEMGPU_ERR_USAGE;
  int32_t st = femgpu::need_sep(h);
  if (st) return st;
  if (!h->sep.sky_valid) return h->fail(FEMGPU_ERR_USAGE, "no skyline: call femgpu_separate_direct first");
  return femgpu::run_colsol(h);
}

int32_t femgpu_set_ua(femgpu_t* h, const double* u_a) {
  if (!h) return FEMGPU_ERR_USAGE;
  int32_t st = femgpu::need_sep(h);
  if (st) return st;
  if (!u_a) return h->fail(FEMGPU_ERR_USAGE, "null u_a");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  FEMGPU_CUDA_CHECK(h, h->sol.u_a.reserve(size_t(h->sep.n_aa) + 1));
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(h->sol.u_a.p, u_a, size_t(h->sep.n_aa) * 8, cudaMemcpyHostToDevice, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  h->sol.ua_valid = true;
  h->sol.composed = false;
  return 0;
}

int32_t femgpu_get_ua(femgpu_t* h, double* u_a, const double** u_a_device) {
  if (!h) return FEMGPU_ERR_USAGE;
  int32_t st = femgpu::need_sep(h);
  if (st) return st;
  if (!h->sol.ua_valid) return h->fail(FEMGPU_ERR_USAGE, "no u_a: call femgpu_solve_pcg or femgpu_set_ua first");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  if (u_a && h->sep.n_aa) {
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(u_a, h->sol.u_a.p, size_t(h->sep.n_aa) * 8, cudaMemcpyDeviceToHost, h->stream));
    FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  }
  if (u_a_device) *u_a_device = h->sol.u_a.p;
  return 0;
}

int32_t femgpu_solve_info(femgpu_t* h, int64_t* iterations, double* residual, float* ms) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (iterations) *iterations = h->sol.iterations;
  if (residual) *residual = h->sol.residual;
  if (ms) *ms = h->sol.last_ms;
  return 0;
}

int32_t femgpu_global_analysis(femgpu_t* h) {
  if (!h) return FEMGPU_ERR_USAGE;
  int32_t st = femgpu::need_sep(h);
  if (st) return st;
  if (!h->sol.ua_valid) return h->fail(FEMGPU_ERR_USAGE, "no u_a: call femgpu_solve_pcg or femgpu_set_ua first");
  return femgpu::run_global_analysis(h);
}

int32_t femgpu_get_reactions(femgpu_t* h, double* r_r, const double** r_r_device) {
  if (!h) return FEMGPU_ERR_USAGE;
  int32_t st = femgpu::need_sep(h);
  if (st) return st;
  if (!h->sol.composed) return h->fail(FEMGPU_ERR_USAGE, "no reactions: call femgpu_global_analysis first");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  if (r_r && h->sep.n_bb) {
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(r_r, h->sol.r_r.p, size_t(h->sep.n_bb) * 8, cudaMemcpyDeviceToHost, h->stream));
    FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  }
  if (r_r_device) *r_r_device = h->sol.r_r.p;
  return 0;
}

int32_t femgpu_get_global_result(femgpu_t* h, double* displacements, double* forces) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0)
    return h->fail(FEMGPU_ERR_NO_DEVICE, "this handle was created without a CUDA device (staging only); "
                                         "femgpu has no CPU fallback");
  if (!h->sol.composed) return h->fail(FEMGPU_ERR_USAGE, "no global analysis result: call femgpu_global_analysis first");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  const size_t n = size_t(h->nodes_number) * 6;
  if (displacements && n)
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(displacements, h->sol.disp.p, n * 8, cudaMemcpyDeviceToHost, h->stream));
  if (forces && n)
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(forces, h->sol.force.p, n * 8, cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int32_t femgpu_set_displacements(femgpu_t* h, const double* displacements) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0)
    return h->fail(FEMGPU_ERR_NO_DEVICE, "this handle was created without a CUDA device (staging only); "
                                         "femgpu has no CPU fallback");
  if (!displacements) return h->fail(FEMGPU_ERR_USAGE, "null displacements");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  const size_t n = size_t(h->nodes_number) * 6;
  FEMGPU_CUDA_CHECK(h, h->sol.disp.reserve(n + 1));
  if (n) FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(h->sol.disp.p, displacements, n * 8, cudaMemcpyHostToDevice, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  h->sol.disp_valid = true;
  return 0;
}

int32_t femgpu_element_results(femgpu_t* h, int32_t family, double* out, const double** out_device) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0)
    return h->fail(FEMGPU_ERR_NO_DEVICE, "this handle was created without a CUDA device (staging only); "
                                         "femgpu has no CPU fallback");
  if (family < 0 || family >= femgpu::kFamilies) return h->fail(FEMGPU_ERR_USAGE, "family must be 0 (truss), 1 (beam) or 2 (plate)");
  if (!h->sol.disp_valid)
    return h->fail(FEMGPU_ERR_USAGE, "no displacements: call femgpu_global_analysis or femgpu_set_displacements first");
  // the reference only holds elements that passed *::create: settle pending validation (uploads everything)
  int32_t st = femgpu_validate(h, nullptr, nullptr, nullptr);
  if (st) return st;
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  static const int kComp[femgpu::kFamilies] = {1, 10, 8};
  const size_t n = h->fh[family].size(), count = n * size_t(kComp[family]);
  FEMGPU_CUDA_CHECK(h, h->sol.res[family].reserve(count + 2));
  st = femgpu::run_element_results(h, family, h->sol.disp.p, h->sol.res[family].p);
  if (st) return st;
  if (out && count) FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(out, h->sol.res[family].p, count * 8, cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  if (out_device) *out_device = h->sol.res[family].p;
  return 0;
}

}  // extern "C"
