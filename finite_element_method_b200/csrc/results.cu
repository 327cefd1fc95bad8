// Element result recovery: element forces / moments from the nodal
// displacements, one thread per element — the stiffness kernels run "backwards" (gather instead of
// scatter). Reference: FEM::extract_elements_analysis_result (methods_for_element_analysis.rs:27-58) ->
//   Truss::extract_element_analysis_result   structs/truss.rs:281-333    1 value  (ForceR)
//   Beam::extract_element_analysis_result    structs/beam.rs:803-993     10 values
//   Plate::extract_element_analysis_result   structs/plate.rs:1196-1409  8 values
// Every kernel gathers the element's slice of the global displacement vector, rotates it with the
// element's rotation matrix (the same record builders the stiffness path uses), applies the summed
// strain-displacement rows in the reference's column order and scales by the section constants.
//
// Compiled with -fmad=false like prep.cu: the file follows the reference's operation order.
// Loads are struct-of-arrays and coalesced; results are written element-major (the reference returns
// one list per element), as 16-byte vectors where the row size allows.
#include "common.cuh"
#include "element_math.cuh"

namespace femgpu {

namespace {

constexpr int kResThreads = 256;

__device__ __forceinline__ void load_xyz(const double* __restrict__ x, const double* __restrict__ y,
                                         const double* __restrict__ z, uint32_t i, double p[3]) {
  p[0] = __ldg(x + i);
  p[1] = __ldg(y + i);
  p[2] = __ldg(z + i);
}

// o = Q d, inner index ascending from a zero accumulator (the assumed arithmetic of extended_matrix's
// Matrix::multiply)
__device__ __forceinline__ void rot3(const double* __restrict__ q, const double* __restrict__ d, double* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double acc = 0.0;
    acc += q[3 * i + 0] * d[0];
    acc += q[3 * i + 1] * d[1];
    acc += q[3 * i + 2] * d[2];
    o[i] = acc;
  }
}

// structs/truss.rs:281-333
__global__ void __launch_bounds__(kResThreads)
truss_result_kernel(uint32_t n, const uint32_t* __restrict__ n1, const uint32_t* __restrict__ n2,
                    const double* __restrict__ E, const double* __restrict__ A, const double* __restrict__ A2,
                    const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                    double abs_tol, const double* __restrict__ u, double* __restrict__ out) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint32_t a = n1[e], b = n2[e];
  double p1[3], p2[3], q[9], k00;
  load_xyz(x, y, z, a, p1);
  load_xyz(x, y, z, b, p2);
  const double area = A[e], area_2 = A2[e], young = E[e];
  truss_record(p1, p2, young, area, area_2, abs_tol, q, &k00);
  double ug[6], ul[6];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    ug[i] = __ldg(u + size_t(a) * 6 + i);
    ug[3 + i] = __ldg(u + size_t(b) * 6 + i);
  }
  rot3(q, ug, ul);
  rot3(q, ug + 3, ul + 3);
  // one integration point (r = 0, alpha = 2), truss.rs:244
  const double r = 0.0;
  const double v[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  const double inv_j = 1.0 / bar_jacobian(norm3(v));
  const double b0 = (0.5 * 0.0 - 0.5 * 1.0) * inv_j, b3 = (0.5 * 0.0 + 0.5 * 1.0) * inv_j;  // truss.rs:95-118
  const bool has2 = !isnan(area_2);
  const double area_sum = 0.0 + (has2 ? (area_2 - area) / 2.0 * r + area - (area_2 - area) / 2.0 * -1.0 : area);
  double strain = 0.0;
  strain += b0 * ul[0];
  strain += b3 * ul[3];
  out[e] = strain * (young * area_sum / 1.0);
}

// structs/beam.rs:803-993
__global__ void __launch_bounds__(kResThreads)
beam_result_kernel(uint32_t n, const uint32_t* __restrict__ n1, const uint32_t* __restrict__ n2,
                   const double* __restrict__ E, const double* __restrict__ nu, const double* __restrict__ A,
                   const double* __restrict__ I11, const double* __restrict__ I22, const double* __restrict__ I12,
                   const double* __restrict__ It, const double* __restrict__ ks, const double* __restrict__ ax,
                   const double* __restrict__ ay, const double* __restrict__ az, const double* __restrict__ x,
                   const double* __restrict__ y, const double* __restrict__ z, double rel_tol, double abs_tol,
                   const double* __restrict__ u, double* __restrict__ out) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint32_t a = n1[e], b = n2[e];
  double p1[3], p2[3], rec[16];
  load_xyz(x, y, z, a, p1);
  load_xyz(x, y, z, b, p2);
  const double young = E[e], poisson = nu[e], area = A[e], it = It[e], shear_factor = ks[e];
  const double axis[3] = {ax[e], ay[e], az[e]};
  beam_record(p1, p2, young, poisson, area, I11[e], I22[e], I12[e], it, shear_factor, axis, rel_tol, abs_tol, rec);
  double i11_p, i22_p, angle;
  beam_principal_inertia(I11[e], I22[e], I12[e], rel_tol, &i11_p, &i22_p, &angle);
  double ug[12], ul[12];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    ug[i] = __ldg(u + size_t(a) * 6 + i);
    ug[6 + i] = __ldg(u + size_t(b) * 6 + i);
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) rot3(rec, ug + 3 * g, ul + 3 * g);
  // summed strain-displacement rows over the single integration point (r = 0), beam.rs:260-493:
  // derivative columns w / w + 6, and for v (w) the coupling with thw (thv): lhs - rhs = -h at columns 5 / 11 (4 / 10)
  const double v[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  const double len = norm3(v);
  const double inv_j = 1.0 / bar_jacobian(len);
  const double d1 = (0.5 * 0.0 - 0.5 * 1.0) * inv_j, d2 = (0.5 * 0.0 + 0.5 * 1.0) * inv_j;
  const double h1 = 0.0 - 0.5 * (1.0 - 0.0), h2 = 0.0 - 0.5 * (1.0 + 0.0);
  double s[6];
#pragma unroll
  for (int w = 0; w < 6; ++w) {
    double acc = 0.0;
    if (w == 1) {  // columns 1, 5, 7, 11
      acc += d1 * ul[1];
      acc += h1 * ul[5];
      acc += d2 * ul[7];
      acc += h2 * ul[11];
    } else if (w == 2) {  // columns 2, 4, 8, 10
      acc += d1 * ul[2];
      acc += h1 * ul[4];
      acc += d2 * ul[8];
      acc += h2 * ul[10];
    } else {
      acc += d1 * ul[w];
      acc += d2 * ul[w + 6];
    }
    s[w] = acc;
  }
  const double n_ip = 1.0;
  const double shear_modulus = young / (2.0 * (1.0 + poisson));
  const double force_r = s[0] * (young * area / n_ip);
  const double force_s = s[1] * (shear_modulus * area * shear_factor / n_ip);
  const double force_t = s[2] * (shear_modulus * area * shear_factor / n_ip);
  const double moment_r = s[3] * (shear_modulus * it / n_ip);
  const double moment_s = s[4] * (young * i22_p / n_ip);
  const double moment_t = s[5] * (young * i11_p / n_ip);
  double2* o = reinterpret_cast<double2*>(out + size_t(e) * 10);
  o[0] = make_double2(force_r, force_s);
  o[1] = make_double2(force_t, moment_r);
  o[2] = make_double2(moment_s + len * force_t / 2.0, moment_s);
  o[3] = make_double2(moment_s - len * force_t / 2.0, moment_t + len * force_s / 2.0);
  o[4] = make_double2(moment_t, moment_t - len * force_s / 2.0);
}

// structs/plate.rs:1196-1409. The strain-displacement matrices are summed over the four NODES
// (r, s = +-1), in the order (1,1), (-1,1), (-1,-1), (1,-1).
__global__ void __launch_bounds__(kResThreads)
plate_result_kernel(uint32_t n, const uint32_t* __restrict__ n1, const uint32_t* __restrict__ n2,
                    const uint32_t* __restrict__ n3, const uint32_t* __restrict__ n4,
                    const double* __restrict__ E, const double* __restrict__ nu, const double* __restrict__ th,
                    const double* __restrict__ ks, const double* __restrict__ x, const double* __restrict__ y,
                    const double* __restrict__ z, double abs_tol, const double* __restrict__ u,
                    double* __restrict__ out) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint32_t nd[4] = {n1[e], n2[e], n3[e], n4[e]};
  double p[4][3], rec[16], mat[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) load_xyz(x, y, z, nd[a], p[a]);
  const double young = E[e], poisson = nu[e], t = th[e], shear_factor = ks[e];
  plate_record<false>(p[0], p[1], p[2], p[3], young, poisson, t, shear_factor, abs_tol, rec, mat);
  double ul[24];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double ug[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) ug[i] = __ldg(u + size_t(nd[a]) * 6 + i);
    rot3(rec, ug, ul + 6 * a);
    rot3(rec, ug + 3, ul + 6 * a + 3);
  }
  const double x1 = rec[9], y1 = rec[10], x2 = rec[11], y2 = rec[12], x3 = 0.0, y3 = 0.0, x4 = rec[13], y4 = rec[14];
  // sums over the four node points of dh_a/dx, dh_a/dy (membrane, bending) and of the shear rows
  double dx[4] = {0.0, 0.0, 0.0, 0.0}, dy[4] = {0.0, 0.0, 0.0, 0.0};
  double ndx[4] = {0.0, 0.0, 0.0, 0.0}, ndy[4] = {0.0, 0.0, 0.0, 0.0};  // sums of -1 * dh/dx, -1 * dh/dy (bending rows)
  double sh0[12], sh1[12];  // per node: columns 2 (w), 3 (thx), 4 (thy)
#pragma unroll
  for (int i = 0; i < 12; ++i) sh0[i] = sh1[i] = 0.0;
  const double a_x = x1 - x2 - x3 + x4, b_x = x1 - x2 + x3 - x4, c_x = x1 + x2 - x3 - x4;
  const double a_y = y1 - y2 - y3 + y4, b_y = y1 - y2 + y3 - y4, c_y = y1 + y2 - y3 - y4;
#pragma unroll
  for (int pt = 0; pt < 4; ++pt) {
    const double r = (pt == 0 || pt == 3) ? 1.0 : -1.0;
    const double s = (pt < 2) ? 1.0 : -1.0;
    // Jacobian with node 3 at the local origin (quadrilateral_4n_element_functions.rs:252-446)
    const double j0 = 0.25 * ((x1 - x2) * (1.0 + s) + (x4 - x3) * (1.0 - s));
    const double j1 = 0.25 * ((y1 - y2) * (1.0 + s) + (y4 - y3) * (1.0 - s));
    const double j2 = 0.25 * ((x1 - x4) * (1.0 + r) + (x2 - x3) * (1.0 - r));
    const double j3 = 0.25 * ((y1 - y4) * (1.0 + r) + (y2 - y3) * (1.0 - r));
    const double det = j0 * j3 - j1 * j2;
    const double i0 = j3 / det, i1 = -1.0 * j1 / det, i2 = -1.0 * j2 / det, i3 = j0 / det;  // :448-475
    const double dr[4] = {0.25 * (1.0 + s), -0.25 * (1.0 + s), -0.25 * (1.0 - s), 0.25 * (1.0 - s)};   // :505-583
    const double ds[4] = {0.25 * (1.0 + r), 0.25 * (1.0 - r), -0.25 * (1.0 - r), -0.25 * (1.0 + r)};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double gx = 0.0, gy = 0.0;  // :613-653
      gx += i0 * dr[a];
      gx += i1 * ds[a];
      gy += i2 * dr[a];
      gy += i3 * ds[a];
      dx[a] = dx[a] + gx;
      dy[a] = dy[a] + gy;
      ndx[a] = ndx[a] + -1.0 * gx;
      ndy[a] = ndy[a] + -1.0 * gy;
    }
    // plate.rs:392-511
    const double ex = c_x + r * b_x, ey = c_y + r * b_y, fx = a_x + s * b_x, fy = a_y + s * b_y;
    const double grz = sqrt(ex * ex + ey * ey) / (8.0 * det);
    const double gsz = sqrt(fx * fx + fy * fy) / (8.0 * det);
    const double row0[12] = {(1.0 + s) / 2.0 * grz,        (1.0 + s) * -1.0 * (y1 - y2) / 4.0 * grz, (1.0 + s) * (x1 - x2) / 4.0 * grz,
                             -1.0 * (1.0 + s) / 2.0 * grz, (1.0 + s) * -1.0 * (y1 - y2) / 4.0 * grz, (1.0 + s) * (x1 - x2) / 4.0 * grz,
                             -1.0 * (1.0 - s) / 2.0 * grz, (1.0 - s) * -1.0 * (y4 - y3) / 4.0 * grz, (1.0 - s) * (x4 - x3) / 4.0 * grz,
                             (1.0 - s) / 2.0 * grz,        (1.0 - s) * -1.0 * (y4 - y3) / 4.0 * grz, (1.0 - s) * (x4 - x3) / 4.0 * grz};
    const double row1[12] = {(1.0 + r) / 2.0 * gsz,        (1.0 + r) * -1.0 * (y1 - y4) / 4.0 * gsz, (1.0 + r) * (x1 - x4) / 4.0 * gsz,
                             (1.0 - r) / 2.0 * gsz,        (1.0 - r) * -1.0 * (y2 - y3) / 4.0 * gsz, (1.0 - r) * (x2 - x3) / 4.0 * gsz,
                             -1.0 * (1.0 - r) / 2.0 * gsz, (1.0 - r) * -1.0 * (y2 - y3) / 4.0 * gsz, (1.0 - r) * (x2 - x3) / 4.0 * gsz,
                             -1.0 * (1.0 + r) / 2.0 * gsz, (1.0 + r) * -1.0 * (y1 - y4) / 4.0 * gsz, (1.0 + r) * (x1 - x4) / 4.0 * gsz};
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      sh0[i] = sh0[i] + row0[i];
      sh1[i] = sh1[i] + row1[i];
    }
  }
  // strains = B_sum * u_local, columns ascending
  double em[3] = {0.0, 0.0, 0.0}, eb[3] = {0.0, 0.0, 0.0}, es[2] = {0.0, 0.0};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double* w = ul + 6 * a;
    em[0] += dx[a] * w[0];   // plate.rs:160-274
    em[1] += dy[a] * w[1];
    em[2] += dy[a] * w[0];
    em[2] += dx[a] * w[1];
    eb[0] += ndx[a] * w[4];  // plate.rs:276-390
    eb[1] += dy[a] * w[3];
    eb[2] += dx[a] * w[3];
    eb[2] += ndy[a] * w[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      es[0] += sh0[3 * a + c] * w[2 + c];
      es[1] += sh1[3 * a + c] * w[2 + c];
    }
  }
  const double n_nodes = 4.0;
  const double c22 = (1.0 - poisson) / 2.0;
  double2* o = reinterpret_cast<double2*>(out + size_t(e) * 8);
  {  // membrane forces, plate.rs:1256-1292
    const double cm = young / (1.0 - poisson * poisson), sc = t / n_nodes;
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    f0 += (1.0 * cm) * em[0]; f0 += (poisson * cm) * em[1]; f0 += (0.0 * cm) * em[2];
    f1 += (poisson * cm) * em[0]; f1 += (1.0 * cm) * em[1]; f1 += (0.0 * cm) * em[2];
    f2 += (0.0 * cm) * em[0]; f2 += (0.0 * cm) * em[1]; f2 += (c22 * cm) * em[2];
    o[0] = make_double2(f0 * sc, f1 * sc);
    const double mrs = f2 * sc;
    // bending moments, plate.rs:1294-1332: BendingMomentR = row 1, BendingMomentS = row 0 (:1384-1391)
    const double cb = young * t / (2.0 * (1.0 - poisson * poisson)), sb = t * t / 24.0;
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    g0 += (1.0 * cb) * eb[0]; g0 += (poisson * cb) * eb[1]; g0 += (0.0 * cb) * eb[2];
    g1 += (poisson * cb) * eb[0]; g1 += (1.0 * cb) * eb[1]; g1 += (0.0 * cb) * eb[2];
    g2 += (0.0 * cb) * eb[0]; g2 += (0.0 * cb) * eb[1]; g2 += (c22 * cb) * eb[2];
    o[1] = make_double2(mrs, g1 * sb);
    o[2] = make_double2(g0 * sb, g2 * sb);
    // shear forces, plate.rs:1334-1366
    const double cs = young / (2.0 * (1.0 + poisson)), ss = t * shear_factor / n_nodes;
    double q0 = 0.0, q1 = 0.0;
    q0 += (1.0 * cs) * es[0]; q0 += (0.0 * cs) * es[1];
    q1 += (0.0 * cs) * es[0]; q1 += (1.0 * cs) * es[1];
    o[3] = make_double2(q0 * ss, q1 * ss);
  }
}

}  // namespace

// family results into d_out ([n] / [n][10] / [n][8] doubles) for the displacement vector d_u (6 per node)
int32_t run_element_results(Handle* h, int family, const double* d_u, double* d_out) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  const FamilyDev& fd = h->fd[family];
  const uint32_t n = uint32_t(h->fh[family].size());
  if (n == 0) return 0;
  const uint32_t grid = div_up(n, kResThreads);
  auto P = [&](int k) { return (const double*)fd.props[k].p; };
  const double *x = h->x_global(), *y = h->y_global(), *z = h->z_global();
  if (family == FEMGPU_TRUSS)
    truss_result_kernel<<<grid, kResThreads, 0, h->stream>>>(n, fd.conn[0].p, fd.conn[1].p, P(0), P(1), P(2), x, y, z,
                                                              h->abs_tol, d_u, d_out);
  else if (family == FEMGPU_BEAM)
    beam_result_kernel<<<grid, kResThreads, 0, h->stream>>>(n, fd.conn[0].p, fd.conn[1].p, P(0), P(1), P(2), P(3), P(4),
                                                             P(5), P(6), P(7), P(8), P(9), P(10), x, y, z, h->rel_tol,
                                                             h->abs_tol, d_u, d_out);
  else
    plate_result_kernel<<<grid, kResThreads, 0, h->stream>>>(n, fd.conn[0].p, fd.conn[1].p, fd.conn[2].p, fd.conn[3].p,
                                                              P(0), P(1), P(2), P(3), x, y, z, h->abs_tol, d_u, d_out);
  h->launches++;
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  return 0;
}

}  // namespace femgpu
