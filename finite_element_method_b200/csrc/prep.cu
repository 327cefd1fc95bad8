// Element-record kernels: one thread per element computes everything that needs transcendental
// functions or the abs_tol clipping (rotation matrices, principal inertia, Jacobian scalars) and
// the element-level validity checks of Truss/Beam/Plate::create, and writes a compact record the
// assembly kernel reads (truss 32 B, beam 128 B, plate 128 B + 32 B material; stored at the stride of the
// kernel's shared-memory slots, kRecDoubles).
//
// Compiled with -fmad=false: this file follows the reference's operation order, and Rust does not
// contract a*b+c.
//
// Loads are struct-of-arrays and coalesced (thread i reads element i of every property array);
// node coordinates are gathered through the read-only path. Records are written as 16-byte
// vectors.
#include "common.cuh"
#include "element_math.cuh"

namespace femgpu {

namespace {

constexpr int kPrepThreads = 256;
// The record kernels are latency-bound (dependent FP64 chains through acos / sincos / sqrt / divisions): resident warps
// are what hides that latency, so the beam and plate kernels are held to 64 registers (4 x 256 threads per SM).
#ifndef FEMGPU_PREP_BLOCKS
#define FEMGPU_PREP_BLOCKS 4
#endif
constexpr int kPrepBlocksPerSm = FEMGPU_PREP_BLOCKS;

__device__ __forceinline__ void load_xyz(const double* __restrict__ x, const double* __restrict__ y,
                                         const double* __restrict__ z, uint32_t i, double p[3]) {
  p[0] = __ldg(x + i);
  p[1] = __ldg(y + i);
  p[2] = __ldg(z + i);
}

// the library's own acos(0), acos(-1) and their sin / cos, computed once per handle (element_math.cuh trig_table_init)
__global__ void trig_table_kernel(double* __restrict__ trig, double opaque_zero) { trig_table_init(trig, opaque_zero); }

// Record stores. A thread holds its element's record (kRec doubles) in registers; written directly, one 16-byte store
// per thread lands in 32 different 128-byte lines per warp instruction (record stride 48 / 144 / 160 B) and the LSU
// queue throttles (ncu: lg_throttle 6.6 stalls per issue in the plate kernel). When the warp's elements are consecutive
// (no order list) the records go through a per-warp shared-memory tile laid out exactly like the global range —
// element-major, unpadded — and leave as fully coalesced 16-byte stores (512 contiguous bytes per instruction).
template <int kRec>
__device__ __forceinline__ void write_records(double* __restrict__ rec, uint32_t e, bool live, bool consecutive,
                                              uint32_t warp_first, uint32_t warp_count, const double (&v)[kRec],
                                              double* __restrict__ tile /* this warp's 32 * kRec doubles */) {
  static_assert(kRec % 2 == 0, "records are written as 16-byte pairs");
  const uint32_t lane = threadIdx.x & 31u;
  if (!consecutive) {
    if (live) {
      double2* out = reinterpret_cast<double2*>(rec + size_t(e) * kRec);
#pragma unroll
      for (int i = 0; i < kRec / 2; ++i) out[i] = make_double2(v[2 * i], v[2 * i + 1]);
    }
    return;
  }
  double2* t2 = reinterpret_cast<double2*>(tile);
  if (live) {
#pragma unroll
    for (int i = 0; i < kRec / 2; ++i) t2[lane * (kRec / 2) + i] = make_double2(v[2 * i], v[2 * i + 1]);
  }
  __syncwarp();
  double2* out = reinterpret_cast<double2*>(rec + size_t(warp_first) * kRec);
  const uint32_t chunks = warp_count * uint32_t(kRec / 2);
#pragma unroll
  for (int j = 0; j < kRec / 2; ++j) {
    const uint32_t c = uint32_t(j) * 32u + lane;
    if (c < chunks) out[c] = t2[c];
  }
  __syncwarp();
}

template <bool kWriteErr>
__global__ void __launch_bounds__(kPrepThreads)
truss_prep_kernel(uint32_t from, uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ n1,
                  const uint32_t* __restrict__ n2, const double* __restrict__ E,
                  const double* __restrict__ A, const double* __restrict__ A2,
                  const double* __restrict__ x, const double* __restrict__ y,
                  const double* __restrict__ z, double abs_tol, double* __restrict__ rec,
                  int32_t* __restrict__ err, const double* __restrict__ trig) {
  __shared__ __align__(16) double tile[kPrepThreads * kTrussSlotDoubles];
  const uint32_t pos = from + blockIdx.x * blockDim.x + threadIdx.x, warp_first = pos - (threadIdx.x & 31u);
  const bool live = pos < n;
  const uint32_t e = live && order ? order[pos] : pos;  // positions [from, n) of the range plan's list (Handle::prep_order)
  double v[kTrussSlotDoubles] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  int code = 0;
  if (live) {
    double p1[3], p2[3], q[9], k00 = 0.0;
    load_xyz(x, y, z, n1[e], p1);
    load_xyz(x, y, z, n2[e], p2);
    code = truss_record(p1, p2, E[e], A[e], A2[e], abs_tol, q, &k00, trig);
    if (!code) {
      v[0] = q[0];
      v[1] = q[1];
      v[2] = q[2];
      v[3] = k00;
    }
  }
  write_records<kTrussSlotDoubles>(rec, e, live, order == nullptr, warp_first, min(32u, n > warp_first ? n - warp_first : 0u), v,
                                   tile + (threadIdx.x >> 5) * 32 * kTrussSlotDoubles);
  if (kWriteErr && live) err[e] = code;
}

template <bool kWriteErr>
__global__ void __launch_bounds__(kPrepThreads, kPrepBlocksPerSm)
beam_prep_kernel(uint32_t from, uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ n1,
                 const uint32_t* __restrict__ n2, const double* __restrict__ E,
                 const double* __restrict__ nu, const double* __restrict__ A,
                 const double* __restrict__ I11, const double* __restrict__ I22,
                 const double* __restrict__ I12, const double* __restrict__ It,
                 const double* __restrict__ ks, const double* __restrict__ ax,
                 const double* __restrict__ ay, const double* __restrict__ az,
                 const double* __restrict__ x, const double* __restrict__ y,
                 const double* __restrict__ z, double rel_tol, double abs_tol,
                 double* __restrict__ rec, int32_t* __restrict__ err, const double* __restrict__ trig) {
  __shared__ __align__(16) double tile[kPrepThreads * kBeamSlotDoubles];
  const uint32_t pos = from + blockIdx.x * blockDim.x + threadIdx.x, warp_first = pos - (threadIdx.x & 31u);
  const bool live = pos < n;
  const uint32_t e = live && order ? order[pos] : pos;
  double v[kBeamSlotDoubles];
#pragma unroll
  for (int i = 0; i < kBeamSlotDoubles; ++i) v[i] = 0.0;
  int code = 0;
  if (live) {
    double p1[3], p2[3], r[16];
    load_xyz(x, y, z, n1[e], p1);
    load_xyz(x, y, z, n2[e], p2);
    double axis[3] = {ax[e], ay[e], az[e]};
    code = beam_record(p1, p2, E[e], nu[e], A[e], I11[e], I22[e], I12[e], It[e], ks[e], axis, rel_tol, abs_tol, r, trig);
    if (!code) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = r[i];
    }
  }
  write_records<kBeamSlotDoubles>(rec, e, live, order == nullptr, warp_first, min(32u, n > warp_first ? n - warp_first : 0u), v,
                                  tile + (threadIdx.x >> 5) * 32 * kBeamSlotDoubles);
  if (kWriteErr && live) err[e] = code;
}

template <bool kWriteErr>
__global__ void __launch_bounds__(kPrepThreads, kPrepBlocksPerSm)
plate_prep_kernel(uint32_t from, uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ n1,
                  const uint32_t* __restrict__ n2, const uint32_t* __restrict__ n3,
                  const uint32_t* __restrict__ n4, const double* __restrict__ E,
                  const double* __restrict__ nu, const double* __restrict__ t,
                  const double* __restrict__ ks, const double* __restrict__ x,
                  const double* __restrict__ y, const double* __restrict__ z, double abs_tol,
                  double* __restrict__ rec, int32_t* __restrict__ err, const double* __restrict__ trig) {
  __shared__ __align__(16) double tile[kPrepThreads * kPlateRawDoubles];
  const uint32_t pos = from + blockIdx.x * blockDim.x + threadIdx.x, warp_first = pos - (threadIdx.x & 31u);
  const bool live = pos < n;
  const uint32_t e = live && order ? order[pos] : pos;
  double v[kPlateRawDoubles];
#pragma unroll
  for (int i = 0; i < kPlateRawDoubles; ++i) v[i] = 0.0;
  int code = 0;
  if (live) {
    double p1[3], p2[3], p3[3], p4[3], r[16], m[4];
    load_xyz(x, y, z, n1[e], p1);
    load_xyz(x, y, z, n2[e], p2);
    load_xyz(x, y, z, n3[e], p3);
    load_xyz(x, y, z, n4[e], p4);
    code = plate_record<kWriteErr>(p1, p2, p3, p4, E[e], nu[e], t[e], ks[e], abs_tol, r, m, trig);
    if (!code) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = r[i];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[16 + i] = m[i];
    }
  }
  write_records<kPlateRawDoubles>(rec, e, live, order == nullptr, warp_first, min(32u, n > warp_first ? n - warp_first : 0u), v,
                                  tile + (threadIdx.x >> 5) * 32 * kPlateRawDoubles);
  if (kWriteErr && live) err[e] = code;
}

// Uniformly distributed loads -> nodal loads, one thread per load:
//   Beam::convert_uniformly_distributed_line_load_to_nodal_loads    structs/beam.rs:775-797
//       over the beam's integration points [(r = 0, alpha = 2)] (:729): f_a = (h_a(r) q)(det J alpha)
//   Plate::convert_uniformly_distributed_surface_load_to_nodal_loads structs/plate.rs:1145-1185
//       over the four Gauss points of :1066-1091: f_a += (h_a(r, s) q)(det J(r, s) alpha_r alpha_s)
// Each load writes four (global DOF index, value) contributions (a beam's last two are padding with
// the key 0xFFFFFFFF; kind 3 = a concentrated load recorded after distributed ones: one contribution);
// the caller sorts them by key and sums every run in insertion order.
__global__ void __launch_bounds__(kPrepThreads)
load_kernel(uint32_t n, const int32_t* __restrict__ family, const uint32_t* __restrict__ elem,
            const int32_t* __restrict__ dof, const double* __restrict__ value,
            const uint32_t* __restrict__ b1, const uint32_t* __restrict__ b2,
            const uint32_t* __restrict__ q1, const uint32_t* __restrict__ q2,
            const uint32_t* __restrict__ q3, const uint32_t* __restrict__ q4,
            const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
            double abs_tol, uint32_t* __restrict__ key, double* __restrict__ val) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t e = elem[k];
  const double q = value[k];
  uint32_t node[4] = {0u, 0u, 0u, 0u};
  double f[4] = {0.0, 0.0, 0.0, 0.0};
  int nn;
  if (family[k] == 3) {  // a concentrated load queued behind distributed ones (separate.cu bc_add): elem = node index
    nn = 1;
    node[0] = e;
    f[0] = q;
  } else if (family[k] == FEMGPU_BEAM) {
    nn = 2;
    node[0] = b1[e];
    node[1] = b2[e];
    double p1[3], p2[3];
    load_xyz(x, y, z, node[0], p1);
    load_xyz(x, y, z, node[1], p2);
    const double v[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    const double det = bar_jacobian(norm3(v)), r = 0.0, alpha = 2.0;
    f[0] = 0.0 + ((0.5 * (1.0 - r)) * q) * (det * alpha);
    f[1] = 0.0 + ((0.5 * (1.0 + r)) * q) * (det * alpha);
  } else {
    nn = 4;
    node[0] = q1[e];
    node[1] = q2[e];
    node[2] = q3[e];
    node[3] = q4[e];
    double p1[3], p2[3], p3[3], p4[3], rec[16], mat[4];
    load_xyz(x, y, z, node[0], p1);
    load_xyz(x, y, z, node[1], p2);
    load_xyz(x, y, z, node[2], p3);
    load_xyz(x, y, z, node[3], p4);
    plate_record<false>(p1, p2, p3, p4, 1.0, 0.5, 1.0, 1.0, abs_tol, rec, mat);
    const double x1 = rec[9], y1 = rec[10], x2 = rec[11], y2 = rec[12], x4 = rec[13], y4 = rec[14];
    const double g = 0.57735027779281512;  // sqrt((double)(1.0f / 3.0f)), plate.rs:1066-1091
#pragma unroll
    for (int ip = 0; ip < 4; ++ip) {
      const double r = (ip == 0 || ip == 3) ? g : -g;
      const double s = (ip < 2) ? g : -g;
      // J of quadrilateral_4n_element_functions.rs:252-446 with node 3 at the local origin
      const double x_r = 0.25 * ((x1 - x2) * (1.0 + s) + x4 * (1.0 - s));
      const double y_r = 0.25 * ((y1 - y2) * (1.0 + s) + y4 * (1.0 - s));
      const double x_s = 0.25 * ((x1 - x4) * (1.0 + r) + x2 * (1.0 - r));
      const double y_s = 0.25 * ((y1 - y4) * (1.0 + r) + y2 * (1.0 - r));
      const double scale = (x_r * y_s - y_r * x_s) * 1.0 * 1.0;
      const double h[4] = {0.25 * (1.0 + r) * (1.0 + s), 0.25 * (1.0 - r) * (1.0 + s),
                           0.25 * (1.0 - r) * (1.0 - s), 0.25 * (1.0 + r) * (1.0 - s)};
#pragma unroll
      for (int a = 0; a < 4; ++a) f[a] = f[a] + (h[a] * q) * scale;
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    key[4 * size_t(k) + a] = a < nn ? 6u * node[a] + uint32_t(dof[k]) : 0xFFFFFFFFu;
    val[4 * size_t(k) + a] = f[a];
  }
}

// smallest insertion position (cbase) among failing elements; family in the low 2 bits
__global__ void first_error_kernel(uint32_t from, uint32_t n, int family,
                                   const int32_t* __restrict__ err,
                                   const int64_t* __restrict__ cbase,
                                   unsigned long long* __restrict__ out) {
  uint32_t e = from + blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (err[e] != 0) {
    unsigned long long key = ((unsigned long long)cbase[e] << 2) | (unsigned)family;
    atomicMin(out, key);  // integer min: order independent
  }
}

__global__ void rotation_kernel(int family, uint32_t e, const uint32_t* n1, const uint32_t* n2,
                                const uint32_t* n3, const uint32_t* n4, const double* const* props,
                                const double* x, const double* y, const double* z, double rel_tol,
                                double abs_tol, double* out) {
  double p1[3], p2[3], p3[3], p4[3];
  load_xyz(x, y, z, n1[e], p1);
  load_xyz(x, y, z, n2[e], p2);
  if (family == FEMGPU_TRUSS) {
    double q[9], k00;
    truss_record(p1, p2, 1.0, 1.0, NAN, abs_tol, q, &k00);
    for (int i = 0; i < 9; ++i) out[i] = q[i];
  } else if (family == FEMGPU_BEAM) {
    double r[16];
    double axis[3] = {props[8][e], props[9][e], props[10][e]};
    beam_record(p1, p2, props[0][e], props[1][e], props[2][e], props[3][e], props[4][e],
                props[5][e], props[6][e], props[7][e], axis, rel_tol, abs_tol, r);
    for (int i = 0; i < 9; ++i) out[i] = r[i];
  } else {
    load_xyz(x, y, z, n3[e], p3);
    load_xyz(x, y, z, n4[e], p4);
    double q[9];
    plate_rotation(p2, p3, p4, abs_tol, q);
    for (int i = 0; i < 9; ++i) out[i] = q[i];
  }
}

}  // namespace

static int32_t ensure_trig_table(Handle* h) {
  if (h->trig_table.p) return 0;
  FEMGPU_CUDA_CHECK(h, h->trig_table.reserve(8));
  trig_table_kernel<<<1, 1, 0, h->stream>>>(h->trig_table.p, h->abs_tol * 0.0);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  return 0;
}

int32_t run_prep(Handle* h, bool validate_only) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  {
    int32_t st = ensure_trig_table(h);
    if (st) return st;
  }
  // buffers first (stream-ordered allocations on the handle's stream), then the kernels
  size_t from_of[kFamilies] = {0, 0, 0};
  int n_live = 0;
  for (int f = 0; f < kFamilies; ++f) {
    FamilyDev& fd = h->fd[f];
    const size_t n = h->fh[f].size();
    from_of[f] = validate_only ? fd.validated : 0;
    if (n == 0) continue;
    // records are needed for every element either way (the buffers may have been reallocated)
    const size_t before = fd.rec.cap;
    FEMGPU_CUDA_CHECK(h, fd.rec.reserve(n * size_t(kRecDoubles[f])));
    FEMGPU_CUDA_CHECK(h, fd.err.reserve(n));
    if (fd.rec.cap != before) from_of[f] = validate_only ? fd.validated : 0;
    if (from_of[f] < n) ++n_live;
  }
  // numeric pass with several families: fork the kernels onto side streams, join before returning
  const bool fork = !validate_only && n_live > 1;
  if (fork) {
    if (!h->fork_ev) {
      FEMGPU_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
      for (int b = 0; b < 2; ++b) {
        FEMGPU_CUDA_CHECK(h, cudaStreamCreateWithFlags(&h->side_stream[b], cudaStreamNonBlocking));
        FEMGPU_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->join_ev[b], cudaEventDisableTiming));
      }
    }
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->fork_ev, h->stream));
  }
  int lane = 0;  // 0 = the handle's stream, 1 / 2 = side streams
  for (int f = kFamilies - 1; f >= 0; --f) {  // plates and beams (the long ones) first
    FamilyDev& fd = h->fd[f];
    const size_t n = h->fh[f].size(), from = from_of[f];
    if (n == 0 || from >= n) continue;
    cudaStream_t st = h->stream;
    if (fork && lane > 0) {
      st = h->side_stream[lane - 1];
      FEMGPU_CUDA_CHECK(h, cudaStreamWaitEvent(st, h->fork_ev, 0));
    }
    uint32_t grid = div_up(n - from, kPrepThreads);
    const double* x = h->x_global();
    const double* y = h->y_global();
    const double* z = h->z_global();
    auto P = [&](int k) { return (const double*)fd.props[k].p; };
    if (f == FEMGPU_TRUSS) {
      if (validate_only)
        truss_prep_kernel<true><<<grid, kPrepThreads, 0, st>>>(
            uint32_t(from), uint32_t(n), nullptr, fd.conn[0].p, fd.conn[1].p, P(0), P(1), P(2), x, y, z,
            h->abs_tol, fd.rec.p, fd.err.p, h->trig_table.p);
      else
        truss_prep_kernel<false><<<grid, kPrepThreads, 0, st>>>(
            uint32_t(from), uint32_t(n), nullptr, fd.conn[0].p, fd.conn[1].p, P(0), P(1), P(2), x, y, z,
            h->abs_tol, fd.rec.p, fd.err.p, h->trig_table.p);
    } else if (f == FEMGPU_BEAM) {
      if (validate_only)
        beam_prep_kernel<true><<<grid, kPrepThreads, 0, st>>>(
            uint32_t(from), uint32_t(n), nullptr, fd.conn[0].p, fd.conn[1].p, P(0), P(1), P(2), P(3), P(4),
            P(5), P(6), P(7), P(8), P(9), P(10), x, y, z, h->rel_tol, h->abs_tol, fd.rec.p, fd.err.p, h->trig_table.p);
      else
        beam_prep_kernel<false><<<grid, kPrepThreads, 0, st>>>(
            uint32_t(from), uint32_t(n), nullptr, fd.conn[0].p, fd.conn[1].p, P(0), P(1), P(2), P(3), P(4),
            P(5), P(6), P(7), P(8), P(9), P(10), x, y, z, h->rel_tol, h->abs_tol, fd.rec.p, fd.err.p, h->trig_table.p);
    } else {
      if (validate_only)
        plate_prep_kernel<true><<<grid, kPrepThreads, 0, st>>>(
            uint32_t(from), uint32_t(n), nullptr, fd.conn[0].p, fd.conn[1].p, fd.conn[2].p, fd.conn[3].p,
            P(0), P(1), P(2), P(3), x, y, z, h->abs_tol, fd.rec.p, fd.err.p, h->trig_table.p);
      else
        plate_prep_kernel<false><<<grid, kPrepThreads, 0, st>>>(
            uint32_t(from), uint32_t(n), nullptr, fd.conn[0].p, fd.conn[1].p, fd.conn[2].p, fd.conn[3].p,
            P(0), P(1), P(2), P(3), x, y, z, h->abs_tol, fd.rec.p, fd.err.p, h->trig_table.p);
    }
    h->launches++;
    FEMGPU_CUDA_CHECK(h, cudaGetLastError());
    if (fork && lane > 0) {
      FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->join_ev[lane - 1], st));
      FEMGPU_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, h->join_ev[lane - 1], 0));
    }
    ++lane;
  }
  return 0;
}

// Numeric pass with a range plan: the records of the elements `range` is the first to need, the three families one
// after the other on stream `st` (the low-priority stream of api.cu femgpu_numeric; the buffers exist: the symbolic
// call validated, i.e. ran the record kernels over, every element).
int32_t run_prep_range(Handle* h, int range, cudaStream_t st) {
  {
    int32_t e = ensure_trig_table(h);  // (already there: the symbolic call ran the record kernels over every element)
    if (e) return e;
  }
  const double* x = h->x_global();
  const double* y = h->y_global();
  const double* z = h->z_global();
  for (int f = kFamilies - 1; f >= 0; --f) {
    FamilyDev& fd = h->fd[f];
    const uint32_t from = range ? h->range_elem_end[f][range - 1] : 0u, to = h->range_elem_end[f][range];
    if (to <= from) continue;
    FEMGPU_CUDA_CHECK(h, fd.rec.reserve(h->fh[f].size() * size_t(kRecDoubles[f])));
    FEMGPU_CUDA_CHECK(h, fd.err.reserve(h->fh[f].size()));
    const uint32_t grid = div_up(to - from, kPrepThreads);
    const uint32_t* order = h->prep_order[f].p;
    auto P = [&](int k) { return (const double*)fd.props[k].p; };
    if (f == FEMGPU_TRUSS)
      truss_prep_kernel<false><<<grid, kPrepThreads, 0, st>>>(from, to, order, fd.conn[0].p, fd.conn[1].p, P(0), P(1), P(2),
                                                             x, y, z, h->abs_tol, fd.rec.p, fd.err.p, h->trig_table.p);
    else if (f == FEMGPU_BEAM)
      beam_prep_kernel<false><<<grid, kPrepThreads, 0, st>>>(from, to, order, fd.conn[0].p, fd.conn[1].p, P(0), P(1), P(2),
                                                            P(3), P(4), P(5), P(6), P(7), P(8), P(9), P(10), x, y, z,
                                                            h->rel_tol, h->abs_tol, fd.rec.p, fd.err.p, h->trig_table.p);
    else
      plate_prep_kernel<false><<<grid, kPrepThreads, 0, st>>>(from, to, order, fd.conn[0].p, fd.conn[1].p, fd.conn[2].p,
                                                             fd.conn[3].p, P(0), P(1), P(2), P(3), x, y, z, h->abs_tol,
                                                             fd.rec.p, fd.err.p, h->trig_table.p);
    h->launches++;
    FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  }
  return 0;
}

int32_t run_load_kernel(Handle* h, uint32_t n, const int32_t* d_family, const uint32_t* d_elem,
                        const int32_t* d_dof, const double* d_value, uint32_t* d_key, double* d_val) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  if (n == 0) return 0;
  const FamilyDev& fb = h->fd[FEMGPU_BEAM];
  const FamilyDev& fp = h->fd[FEMGPU_PLATE];
  load_kernel<<<div_up(n, kPrepThreads), kPrepThreads, 0, h->stream>>>(
      n, d_family, d_elem, d_dof, d_value, fb.conn[0].p, fb.conn[1].p, fp.conn[0].p, fp.conn[1].p, fp.conn[2].p,
      fp.conn[3].p, h->x_global(), h->y_global(), h->z_global(), h->abs_tol, d_key, d_val);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  return 0;
}

int32_t first_error(Handle* h, int* family, size_t* index, int* code) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  FEMGPU_CUDA_CHECK(h, h->d_flag.reserve(16));
  unsigned long long* d_key = reinterpret_cast<unsigned long long*>(h->d_flag.p);
  unsigned long long init = ~0ull;
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(d_key, &init, 8, cudaMemcpyHostToDevice, h->stream));
  for (int f = 0; f < kFamilies; ++f) {
    FamilyDev& fd = h->fd[f];
    size_t n = h->fh[f].size(), from = fd.validated;
    if (from >= n) continue;
    first_error_kernel<<<div_up(n - from, 256), 256, 0, h->stream>>>(uint32_t(from), uint32_t(n), f,
                                                                      fd.err.p, fd.cbase.p, d_key);
    h->launches++;
  }
  unsigned long long key = 0;
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(&key, d_key, 8, cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  if (key == ~0ull) {
    *family = -1;
    return 0;
  }
  int f = int(key & 3);
  int64_t cb = int64_t(key >> 2);
  size_t idx = h->fh[f].index_of_cbase(cb, kPairsPerElem[f]);
  int32_t ec = 0;
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(&ec, h->fd[f].err.p + idx, 4, cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  *family = f;
  *index = idx;
  *code = ec;
  return 0;
}

int32_t element_rotation(Handle* h, int family, size_t index, double* out_host) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  FamilyDev& fd = h->fd[family];
  FEMGPU_CUDA_CHECK(h, h->scratch.reserve(4096));
  const double* hp[11];
  for (int k = 0; k < 11; ++k) hp[k] = fd.props[k].p;
  const double** d_props = reinterpret_cast<const double**>(h->scratch.p);
  double* d_out = reinterpret_cast<double*>(h->scratch.p + 1024);
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(d_props, hp, sizeof hp, cudaMemcpyHostToDevice, h->stream));
  rotation_kernel<<<1, 1, 0, h->stream>>>(family, uint32_t(index), fd.conn[0].p, fd.conn[1].p,
                                          fd.conn[2].p, fd.conn[3].p, d_props, h->x_global(), h->y_global(),
                                          h->z_global(), h->rel_tol, h->abs_tol, d_out);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(out_host, d_out, 72, cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

}  // namespace femgpu
