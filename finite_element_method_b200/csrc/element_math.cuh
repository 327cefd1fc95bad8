// Device-side element mathematics for the three families.
//
// Two layers:
//   (1) "record" builders — run once per element per numeric pass (prep kernels). They follow the
//       reference operation by operation because this is where the transcendental functions and
//       the abs_tol clipping live (rotation matrices, principal inertia). prep.cu is compiled with
//       -fmad=false so a*b+c is not contracted here, like the Rust original.
//   (2) node-pair block evaluators — the hot code. Given an element record and a local node pair
//       (la, lb) they add that element's 6x6 (3x3 for trusses) global-frame block
//       T^T k[la,lb] T into 36 accumulators held in registers. These are algebraically the
//       reference's B^T C B quadrature (same Gauss rule, same f32-rounded abscissa), evaluated
//       block-wise with structural zeros skipped; FMA contraction is allowed here.
//
// Reference files (relative to /root/reference/src/fem/) are cited per function.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace femgpu {

// element-level validation codes, numerically equal to FEMGPU_E_* in include/femgpu.h
enum : int {
  EV_OK = 0,
  EV_YOUNG = 20,
  EV_POISSON = 21,
  EV_AREA = 22,
  EV_AREA2 = 23,
  EV_I11 = 24,
  EV_I22 = 25,
  EV_IT = 26,
  EV_SHEAR_FACTOR = 27,
  EV_PARALLEL_AXIS = 28,
  EV_THICKNESS = 29,
  EV_ON_LINE = 30,
  EV_NOT_ON_PLANE = 31,
  EV_NOT_CONVEX = 32
};

// math_functions.rs:3-12
__device__ __forceinline__ double clip_tol(double v, double abs_tol) {
  return fabs(v) < abs_tol ? 0.0 : v;
}

// Transcendental functions at their exactly representable special values. acos(1) = 0, atan(+-0) = +-0,
// sin(+-0) = +-0 and cos(+-0) = 1 hold exactly, so skipping the library call there returns the same bits; it matters
// because structural models are full of such inputs (members along +x, flat plates with the normal along +z, sections
// given in their principal axes) and a double-precision acos / sincos costs a few hundred instructions.
// `trig` (may be null) is the table of trig_table_init(): the library's own results at two more arguments that
// axis-aligned structures produce all the time — acos(0) (a member along y or z against the x axis) and acos(-1) — and
// the sin / cos of those two angles, computed once per handle on the device with the same functions, so a hit returns
// the same bits. (A per-CTA copy in shared memory was measured first: the CTA waits ~500 dependent instructions for
// its thread 0, which cost more than the table saved.)
//   trig = { acos(0), sin(acos(0)), cos(acos(0)), acos(-1), sin(acos(-1)), cos(acos(-1)) }
__device__ __forceinline__ void trig_table_init(double* trig, double opaque_zero) {
  // opaque_zero is 0.0 computed from a kernel argument, so nothing here is folded at compile time (by a host libm)
  const double a1 = acos(opaque_zero), a2 = acos(opaque_zero - 1.0);
  trig[0] = a1;
  sincos(a1, &trig[1], &trig[2]);
  trig[3] = a2;
  sincos(a2, &trig[4], &trig[5]);
}
__device__ __forceinline__ double acos_x(double c, const double* trig = nullptr) {
  if (c == 1.0) return 0.0;
  if (trig) {
    if (c == 0.0) return trig[0];
    if (c == -1.0) return trig[3];
  }
  return acos(c);
}
__device__ __forceinline__ double atan_x(double t) { return t == 0.0 ? t : atan(t); }
__device__ __forceinline__ double sin_x(double a) { return a == 0.0 ? a : sin(a); }
__device__ __forceinline__ void sincos_x(double a, double* s, double* c, const double* trig = nullptr) {
  if (a == 0.0) {
    *s = a;
    *c = 1.0;
  } else if (trig && a == trig[0]) {
    *s = trig[1];
    *c = trig[2];
  } else if (trig && a == trig[3]) {
    *s = trig[4];
    *c = trig[5];
  } else {
    sincos(a, s, c);
  }
}

__device__ __forceinline__ double norm3(const double a[3]) {
  double acc = 0.0;
  acc += a[0] * a[0];
  acc += a[1] * a[1];
  acc += a[2] * a[2];
  return sqrt(acc);
}
__device__ __forceinline__ double dot3(const double a[3], const double b[3]) {
  double acc = 0.0;
  acc += a[0] * b[0];
  acc += a[1] * b[1];
  acc += a[2] * b[2];
  return acc;
}
__device__ __forceinline__ void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// extended_matrix Vector3::rotation_matrix_to_align_with_vector as called from truss.rs:78-84,
// beam.rs:177-183 and quadrilateral_4n_element_functions.rs:158-162: Rodrigues rotation about
// a x b by acos(a.b/(|a||b|)); cos, sin and every entry clipped by abs_tol; zero axis when the
// vectors are (anti)parallel.
__device__ inline void rotation_align(const double a[3], const double b[3], double abs_tol,
                                      double q[9], const double* trig = nullptr) {
  double na = norm3(a), nb = norm3(b);
  double cosv = dot3(a, b) / (na * nb);
  cosv = cosv > 1.0 ? 1.0 : (cosv < -1.0 ? -1.0 : cosv);
  double angle = acos_x(cosv, trig);
  double ax[3];
  cross3(a, b, ax);
  double n = norm3(ax);
  double x = 0.0, y = 0.0, z = 0.0;
  if (n != 0.0) {
    x = ax[0] / n;
    y = ax[1] / n;
    z = ax[2] / n;
  }
  double sn, cs;
  sincos_x(angle, &sn, &cs, trig);  // one argument reduction for both (the values are those of sin() and cos())
  double c = clip_tol(cs, abs_tol);
  double s = clip_tol(sn, abs_tol);
  double t = 1.0 - c;
  q[0] = clip_tol(t * x * x + c, abs_tol);
  q[1] = clip_tol(t * x * y - z * s, abs_tol);
  q[2] = clip_tol(t * x * z + y * s, abs_tol);
  q[3] = clip_tol(t * x * y + z * s, abs_tol);
  q[4] = clip_tol(t * y * y + c, abs_tol);
  q[5] = clip_tol(t * y * z - x * s, abs_tol);
  q[6] = clip_tol(t * x * z - y * s, abs_tol);
  q[7] = clip_tol(t * y * z + x * s, abs_tol);
  q[8] = clip_tol(t * z * z + c, abs_tol);
}

// bar_2n_element_functions.rs:35-59 at r: J = dx_dr(-L/2, L/2, r); the power/derivative helpers of
// math_functions.rs:14-28 reduce to (a*0) for n = 0 and (a*1) for n = 1.
__device__ __forceinline__ double bar_jacobian(double len) {
  double x_1 = -1.0 * len / 2.0;
  double x_2 = len / 2.0;
  return (x_1 * 0.5) * 0.0 - (x_1 * 0.5) * 1.0 + (x_2 * 0.5) * 0.0 + (x_2 * 0.5) * 1.0;
}

// ------------------------------------------------------------------------------------------
// Truss: record = {q11, q12, q13, k00}; k00 = EA/L as the reference rounds it.
// structs/truss.rs:44-64 (checks), :66-93 (rotation), :95-155 (B, area, k at the single IP r=0,
// alpha=2 of :244).
// ------------------------------------------------------------------------------------------
__device__ inline int truss_record(const double p1[3], const double p2[3], double young_modulus,
                                   double area, double area_2 /* NaN = None */, double abs_tol,
                                   double q[9], double* k00, const double* trig = nullptr) {
  if (young_modulus <= 0.0) return EV_YOUNG;
  if (area <= 0.0) return EV_AREA;
  bool has2 = !isnan(area_2);
  if (has2 && area_2 <= 0.0) return EV_AREA2;
  double v[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double len = norm3(v);
  double dir[3] = {len, 0.0, 0.0};
  rotation_align(v, dir, abs_tol, q, trig);
  const double r = 0.0, alpha = 2.0;
  double jac = bar_jacobian(len);
  double inv_j = 1.0 / jac;
  double dh1 = 0.5 * 0.0 - 0.5 * 1.0;  // dh1_dr, bar_2n_element_functions.rs:99-105
  double b1 = dh1 * inv_j;
  double a_r = has2 ? (area_2 - area) / 2.0 * r + area - (area_2 - area) / 2.0 * -1.0 : area;
  double c_at_r = a_r * young_modulus;
  *k00 = (b1 * b1) * (c_at_r * jac * alpha);
  return EV_OK;
}

// block (la, lb) of (R^T k) R for a truss: only k[3la][3lb] = +-k00 is non-zero, so the block is
// (q1i * k) * q1j — the product order the dense reference loops produce
// (methods_for_truss_data_handle.rs:75-81).
__device__ __forceinline__ void truss_block(double q0, double q1, double q2, double k00, int la,
                                            int lb, double acc[36]) {
  double kab = (la == lb) ? k00 : -k00;
  double qv[3] = {q0, q1, q2};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double t = qv[i] * kab;
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[6 * i + j] += t * qv[j];
  }
}

// ------------------------------------------------------------------------------------------
// Beam: record (16 doubles) = r[9], d2, s_u, s_v, s_thu, s_thv, s_thw, identity flag (r == I)
//   d2 = dh2/dr / J (d1 = -d2), s_* = c_* * J * alpha            structs/beam.rs:495-563
// ------------------------------------------------------------------------------------------

// structs/beam.rs:119-160
__device__ inline void beam_principal_inertia(double i11, double i22, double i12, double rel_tol,
                                              double* i11_p, double* i22_p, double* angle_out) {
  const double PI_F32 = (double)3.14159265358979323846f;  // V::from(std::f32::consts::PI)
  double angle;
  if (i11 != i22) {
    angle = atan_x(2.0 * i12 / (i22 - i11)) / 2.0;
  } else {
    double i11_mod, i22_mod;
    if (i22 < i11) {
      i11_mod = i11;
      i22_mod = (fabs(i22) - fabs(i22) * rel_tol) * i22 / fabs(i22);
    } else {
      i11_mod = (fabs(i11) - fabs(i11) * rel_tol) * i11 / fabs(i11);
      i22_mod = i22;
    }
    angle = atan_x(2.0 * i12 / (i22_mod - i11_mod)) / 2.0;
  }
  double ca, sa, s2 = sin_x(2.0 * angle);
  sincos_x(angle, &sa, &ca);
  double p11 = i11 * (ca * ca) + i22 * (sa * sa) - i12 * s2;
  double p22 = i11 * (sa * sa) + i22 * (ca * ca) + i12 * s2;
  int i = 1;
  while (p11 < p22 && i <= 64) {
    angle = (atan(2.0 * i12 / (i22 - i11)) + PI_F32 * (double)(float)i) / 2.0;
    sincos(angle, &sa, &ca);
    s2 = sin(2.0 * angle);
    p11 = i11 * (ca * ca) + i22 * (sa * sa) - i12 * s2;
    p22 = i11 * (sa * sa) + i22 * (ca * ca) + i12 * s2;
    i += 1;
  }
  *i11_p = p11;
  *i22_p = p22;
  *angle_out = angle;
}

__device__ inline int beam_record(const double p1[3], const double p2[3], double young_modulus,
                                  double poisson_ratio, double area, double i11, double i22,
                                  double i12, double it, double shear_factor,
                                  const double axis1[3], double rel_tol, double abs_tol,
                                  double rec[16], const double* trig = nullptr) {
  // structs/beam.rs:63-117
  if (young_modulus <= 0.0) return EV_YOUNG;
  if (poisson_ratio <= 0.0) return EV_POISSON;
  if (area <= 0.0) return EV_AREA;
  if (i11 <= 0.0) return EV_I11;
  if (i22 <= 0.0) return EV_I22;
  if (it <= 0.0) return EV_IT;
  if (shear_factor <= 0.0) return EV_SHEAR_FACTOR;
  double v[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  // projection_perpendicular_to_vector: a - v * (a.v / v.v)
  double f = dot3(axis1, v) / dot3(v, v);
  double proj[3] = {axis1[0] - v[0] * f, axis1[1] - v[1] * f, axis1[2] - v[2] * f};
  if (norm3(proj) == 0.0) return EV_PARALLEL_AXIS;

  double i11_p, i22_p, angle;
  beam_principal_inertia(i11, i22, i12, rel_tol, &i11_p, &i22_p, &angle);

  // structs/beam.rs:162-258
  double len = norm3(v);
  double dir[3] = {len, 0.0, 0.0};
  double qi[9];
  rotation_align(v, dir, abs_tol, qi, trig);
  double tp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double acc = 0.0;
    acc += qi[3 * i + 0] * proj[0];
    acc += qi[3 * i + 1] * proj[1];
    acc += qi[3 * i + 2] * proj[2];
    tp[i] = acc;
  }
  const double ez[3] = {0.0, 0.0, 1.0};
  double cosv = dot3(ez, tp) / (norm3(ez) * norm3(tp));
  cosv = cosv > 1.0 ? 1.0 : (cosv < -1.0 ? -1.0 : cosv);
  double total_angle = angle + acos_x(cosv, trig);
  double c_x = clip_tol(v[0] / len, abs_tol);
  double c_y = clip_tol(v[1] / len, abs_tol);
  double c_z = clip_tol(v[2] / len, abs_tol);
  double c_xz = clip_tol(sqrt(c_x * c_x + c_z * c_z), abs_tol);
  double sn, cs;
  sincos_x(total_angle, &sn, &cs, trig);
  double c = clip_tol(cs, abs_tol);
  double s = clip_tol(sn, abs_tol);
  bool nz = c_xz != 0.0;
  rec[0] = nz ? c_x : 0.0;
  rec[1] = c_y;
  rec[2] = nz ? c_z : 0.0;
  rec[3] = nz ? (-1.0 * c_x * c_y * c - c_z * s) / c_xz : -1.0 * c_y * c;
  rec[4] = nz ? c_xz * c : 0.0;
  rec[5] = nz ? (-1.0 * c_y * c_z * c + c_x * s) / c_xz : s;
  rec[6] = nz ? (c_x * c_y * s - c_z * c) / c_xz : c_y * s;
  rec[7] = nz ? -1.0 * c_xz * s : 0.0;
  rec[8] = nz ? (c_y * c_z * s + c_x * c) / c_xz : c;

  // structs/beam.rs:495-563 at the single IP (r = 0, alpha = 2) of :729
  const double alpha = 2.0;
  double jac = bar_jacobian(len);
  double inv_j = 1.0 / jac;
  double dh2 = 0.5 * 0.0 + 0.5 * 1.0;
  double shear_modulus = young_modulus / (2.0 * (1.0 + poisson_ratio));
  rec[9] = dh2 * inv_j;
  rec[10] = (area * young_modulus) * jac * alpha;
  rec[11] = (shear_modulus * area * shear_factor) * jac * alpha;
  rec[12] = (shear_modulus * it) * jac * alpha;
  rec[13] = (young_modulus * i22_p) * jac * alpha;
  rec[14] = (young_modulus * i11_p) * jac * alpha;
  // r == I exactly (a member along +x whose principal axes are the global y and z): beam_block then adds the ten
  // entries of the local block instead of evaluating the two 3x3 sandwiches (the products with 1.0 and 0.0 are exact,
  // so the general path would add the same numbers)
  bool ident = rec[0] == 1.0 && rec[4] == 1.0 && rec[8] == 1.0 && rec[1] == 0.0 && rec[2] == 0.0 && rec[3] == 0.0 &&
               rec[5] == 0.0 && rec[6] == 0.0 && rec[7] == 0.0;
  rec[15] = ident ? 1.0 : 0.0;
  return EV_OK;
}

// acc(3x3 at rows ro.., cols co..) += r^T diag(l0,l1,l2) r
__device__ __forceinline__ void sandwich_diag(const double* __restrict__ r, double l0, double l1,
                                              double l2, double* acc, int ro, int co) {
  const double l[3] = {l0, l1, l2};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double t = r[3 * k + i] * l[k];
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[6 * (ro + i) + co + j] += t * r[3 * k + j];
    }
}

// acc(3x3 at ro, co) += r^T X r where X has the two entries X[1][2] = x12 and X[2][1] = x21
__device__ __forceinline__ void sandwich_cross(const double* __restrict__ r, double x12, double x21,
                                               double* acc, int ro, int co) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double t1 = r[3 + i] * x12;  // row 1 of X
    double t2 = r[6 + i] * x21;  // row 2 of X
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[6 * (ro + i) + co + j] += t1 * r[6 + j] + t2 * r[3 + j];
  }
}

// block (la, lb) of (R^T k) R for a beam (methods_for_beam_data_handle.rs:88-95). The local block
// has the ten entries listed in beam.rs:260-563: diag(u,v,w), diag(thu,thv,thw), v<->thw, w<->thv.
__device__ __forceinline__ void beam_block(const double* __restrict__ rec, int la, int lb,
                                           double acc[36]) {
  const double* r = rec;
  double d2 = rec[9];
  double da = la ? d2 : -d2, db = lb ? d2 : -d2;
  const double hh = -0.5;  // 0 - h(r=0), beam.rs:326-341
  double dd = da * db;
  double s_u = rec[10], s_v = rec[11], s_thu = rec[12], s_thv = rec[13], s_thw = rec[14];
  double l00 = dd * s_u, l11 = dd * s_v, l22 = dd * s_v;
  double l15 = (da * hh) * s_v, l51 = (hh * db) * s_v;
  double l24 = (da * hh) * s_v, l42 = (hh * db) * s_v;
  double l33 = dd * s_thu;
  double l44 = (hh * hh) * s_v + dd * s_thv;
  double l55 = (hh * hh) * s_v + dd * s_thw;
  // every lane that is in the beam branch right now has an identity rotation (warp-uniform decision, so a frame with
  // members in all directions does not pay for both paths)
  if (__all_sync(__activemask(), rec[15] != 0.0)) {
    acc[0] += l00;
    acc[7] += l11;
    acc[14] += l22;
    acc[21] += l33;
    acc[28] += l44;
    acc[35] += l55;
    acc[11] += l15;  // k[v][thw]
    acc[16] += l24;  // k[w][thv]
    acc[26] += l42;  // k[thv][w]
    acc[31] += l51;  // k[thw][v]
    return;
  }
  sandwich_diag(r, l00, l11, l22, acc, 0, 0);
  sandwich_diag(r, l33, l44, l55, acc, 3, 3);
  // rows (u,v,w) x cols (thu,thv,thw): X[1][2] = k[v][thw], X[2][1] = k[w][thv]
  sandwich_cross(r, l15, l24, acc, 0, 3);
  // rows (thu,thv,thw) x cols (u,v,w): X[1][2] = k[thv][w], X[2][1] = k[thw][v]
  sandwich_cross(r, l42, l51, acc, 3, 0);
}

// ------------------------------------------------------------------------------------------
// Plate: record (16 doubles) = Q[9], x1, y1, x2, y2, x4, y4, identity-flag
//        material (4 doubles) = Cm, Cb, Cs, nu
// ------------------------------------------------------------------------------------------

// quadrilateral_4n_element_functions.rs:131-171
__device__ inline void plate_rotation(const double p2[3], const double p3[3], const double p4[3],
                                      double abs_tol, double q[9], const double* trig = nullptr) {
  double e34[3] = {p4[0] - p3[0], p4[1] - p3[1], p4[2] - p3[2]};
  double e32[3] = {p2[0] - p3[0], p2[1] - p3[1], p2[2] - p3[2]};
  double n[3];
  cross3(e34, e32, n);
  double len = norm3(n);
  double dir[3] = {0.0, 0.0, len};
  rotation_align(n, dir, abs_tol, q, trig);
}

__device__ __forceinline__ void mat3_vec(const double q[9], const double d[3], double o[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double acc = 0.0;
    acc += q[3 * i + 0] * d[0];
    acc += q[3 * i + 1] * d[1];
    acc += q[3 * i + 2] * d[2];
    o[i] = acc;
  }
}

// Graham-scan validity test of convex_hull_on_plane.rs:178-214 specialised to four points, with
// the reference's quick_sort (:14-60) and angle ordering (:81-121) kept as is. pts = (x, y) of
// nodes 1, 2, 3, 4 in the element plane. Returns the hull size.
__device__ inline double hull_angle_deg(double x, double y) {
  double dl = sqrt((0.0 - 1.0) * (0.0 - 1.0) + (0.0 - 0.0) * (0.0 - 0.0));
  double pl = sqrt((0.0 - x) * (0.0 - x) + (0.0 - y) * (0.0 - y));
  double sp = (0.0 - 1.0) * (0.0 - x) + (0.0 - 0.0) * (0.0 - y);
  return acos(sp / (dl * pl)) * (180.0 / 3.14159265358979323846264338327950288);
}

__device__ inline int hull_size4(double px[4], double py[4]) {
  double shift_x = px[0], min_y = py[0];
  int pos = 0;
  for (int i = 0; i < 4; ++i)
    if (py[i] < min_y) {
      shift_x = px[i];
      min_y = py[i];
      pos = i;
    }
  {
    double tx = px[0], ty = py[0];
    px[0] = px[pos];
    py[0] = py[pos];
    px[pos] = tx;
    py[pos] = ty;
  }
  double ang[4];
  for (int i = 0; i < 4; ++i) {
    px[i] -= shift_x;
    py[i] -= min_y;
    ang[i] = hull_angle_deg(px[i], py[i]);
  }
  // quick_sort(&mut data[1..]) with a[i] < a[j] <=> angle_i < angle_j (NaN compares false)
  struct Range {
    int lo, hi;
  } stack[4];
  int sp = 0;
  stack[sp++] = {0, 2};
  while (sp > 0) {
    Range rg = stack[--sp];
    if (rg.lo < rg.hi) {
      int pivot = rg.hi, store = rg.lo - 1, last = rg.hi;
      for (;;) {
        store += 1;
        while (ang[1 + store] < ang[1 + pivot]) store += 1;
        last -= 1;
        while (last >= 0 && ang[1 + last] > ang[1 + pivot]) last -= 1;
        if (store >= last) break;
        double t;
        t = ang[1 + store]; ang[1 + store] = ang[1 + last]; ang[1 + last] = t;
        t = px[1 + store]; px[1 + store] = px[1 + last]; px[1 + last] = t;
        t = py[1 + store]; py[1 + store] = py[1 + last]; py[1 + last] = t;
      }
      {
        double t;
        t = ang[1 + store]; ang[1 + store] = ang[1 + pivot]; ang[1 + pivot] = t;
        t = px[1 + store]; px[1 + store] = px[1 + pivot]; px[1 + pivot] = t;
        t = py[1 + store]; py[1 + store] = py[1 + pivot]; py[1 + pivot] = t;
      }
      stack[sp++] = {rg.lo, store - 1};
      stack[sp++] = {store + 1, rg.hi};
    }
  }
  int len = 4, i = 0;
  while (i + 2 < len) {
    double a2 = (px[i + 1] - px[i]) * (py[i + 2] - py[i]) - (py[i + 1] - py[i]) * (px[i + 2] - px[i]);
    if (a2 <= 0.0) {
      for (int k = i + 1; k + 1 < len; ++k) {
        px[k] = px[k + 1];
        py[k] = py[k + 1];
      }
      len -= 1;
    } else {
      i += 1;
    }
  }
  return len;
}

// kCheck = false builds the record only: the numeric pass runs on elements that already passed
// validation (femgpu_symbolic validates everything pending), so the validity-only work — collinearity,
// coplanarity and the Graham scan with its four acos — is not repeated per pass.
template <bool kCheck = true>
__device__ inline int plate_record(const double p1[3], const double p2[3], const double p3[3],
                                   const double p4[3], double young_modulus, double poisson_ratio,
                                   double thickness, double shear_factor, double abs_tol,
                                   double rec[16], double mat[4], const double* trig = nullptr) {
  // structs/plate.rs:58-158
  if (kCheck) {
    if (young_modulus <= 0.0) return EV_YOUNG;
    if (poisson_ratio <= 0.0) return EV_POISSON;
    if (thickness <= 0.0) return EV_THICKNESS;
    if (shear_factor <= 0.0) return EV_SHEAR_FACTOR;
  }
  if (kCheck) {  // quadrilateral_4n_element_functions.rs:14-86
    const double* P[4] = {p1, p2, p3, p4};
    const int pr[4][3] = {{0, 1, 3}, {1, 0, 2}, {2, 1, 3}, {3, 2, 0}};
    for (int k = 0; k < 4; ++k) {
      const double* o = P[pr[k][0]];
      const double* a = P[pr[k][1]];
      const double* b = P[pr[k][2]];
      double va[3] = {a[0] - o[0], a[1] - o[1], a[2] - o[2]};
      double vb[3] = {b[0] - o[0], b[1] - o[1], b[2] - o[2]};
      double cr[3];
      cross3(va, vb, cr);
      if (norm3(cr) == 0.0) return EV_ON_LINE;
    }
  }
  if (kCheck) {  // quadrilateral_4n_element_functions.rs:88-129
    double v32[3] = {p2[0] - p3[0], p2[1] - p3[1], p2[2] - p3[2]};
    double v34[3] = {p4[0] - p3[0], p4[1] - p3[1], p4[2] - p3[2]};
    double n[3];
    cross3(v32, v34, n);
    double d = -1.0 * (n[0] * p3[0] + n[1] * p3[1] + n[2] * p3[2]);
    if (clip_tol(n[0] * p1[0] + n[1] * p1[1] + n[2] * p1[2] + d, abs_tol) != 0.0)
      return EV_NOT_ON_PLANE;
  }
  double* q = rec;
  plate_rotation(p2, p3, p4, abs_tol, q, trig);
  double d1[3] = {p1[0] - p3[0], p1[1] - p3[1], p1[2] - p3[2]};
  double d2[3] = {p2[0] - p3[0], p2[1] - p3[1], p2[2] - p3[2]};
  double d4[3] = {p4[0] - p3[0], p4[1] - p3[1], p4[2] - p3[2]};
  double t1[3], t2[3], t4[3];
  mat3_vec(q, d1, t1);  // quadrilateral_4n_element_functions.rs:340-378
  mat3_vec(q, d2, t2);
  mat3_vec(q, d4, t4);
  if (kCheck) {  // quadrilateral_4n_element_functions.rs:173-250 + convex_hull_on_plane.rs
    double hx[4] = {t1[0], t2[0], 0.0, t4[0]};
    double hy[4] = {t1[1], t2[1], 0.0, t4[1]};
    if (hull_size4(hx, hy) != 4) return EV_NOT_CONVEX;
  }
  rec[9] = t1[0];
  rec[10] = t1[1];
  rec[11] = t2[0];
  rec[12] = t2[1];
  rec[13] = t4[0];
  rec[14] = t4[1];
  bool ident = q[0] == 1.0 && q[4] == 1.0 && q[8] == 1.0 && q[1] == 0.0 && q[2] == 0.0 &&
               q[3] == 0.0 && q[5] == 0.0 && q[6] == 0.0 && q[7] == 0.0;
  rec[15] = ident ? 1.0 : 0.0;
  // structs/plate.rs:532-543, :575-587, :618-636
  double one_m_nu2 = 1.0 - poisson_ratio * poisson_ratio;
  mat[0] = young_modulus * thickness / one_m_nu2;
  mat[1] = young_modulus * (thickness * thickness * thickness) / (12.0 * one_m_nu2);
  mat[2] = young_modulus * thickness * shear_factor / (2.0 * (1.0 + poisson_ratio));
  mat[3] = poisson_ratio;
  return EV_OK;
}

// acc(3x3 at ro, co) += Q^T X Q for a dense 3x3 X given row-major
__device__ __forceinline__ void sandwich_full(const double* __restrict__ q, const double x[9],
                                              double* acc, int ro, int co) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    // t = (Q^T X)[i][:]
    double t0 = q[i] * x[0] + q[3 + i] * x[3] + q[6 + i] * x[6];
    double t1 = q[i] * x[1] + q[3 + i] * x[4] + q[6 + i] * x[7];
    double t2 = q[i] * x[2] + q[3 + i] * x[5] + q[6 + i] * x[8];
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[6 * (ro + i) + co + j] += t0 * q[j] + t1 * q[3 + j] + t2 * q[6 + j];
  }
}

// Local block (la, lb) of the plate stiffness, 2x2 Gauss with the abscissa
// sqrt(f32(1/3)) of plate.rs:1066-1091, weights 1:
//   membrane  (plate.rs:160-274, :532-573)  on (u, v)
//   bending   (plate.rs:276-390, :575-616)  on (thx, thy)
//   shear     (plate.rs:392-511, :618-665)  on (w, thx, thy), assumed-strain edge-tied rows
//   drilling  (plate.rs:716-721)            +1 on (thz, thz) of diagonal blocks
// then T^T . T with T = diag(Q, Q) (plate.rs:726-1002, methods_for_plate_data_handle.rs:104-112).
//
// With n_a = J_adj * [dh_a/dr; dh_a/ds] (so dh_a/dx = n_a.x / det) the membrane/bending integrals
// reduce to four sums S = sum_ip n_a (x) n_b / det, and because the shear rows of a node are a
// fixed 3-vector times (1 +- s) * gamma_rz (or (1 +- r) * gamma_sz) with
// gamma_rz^2 * det = (x_s^2 + y_s^2) / (4 det), the shear integrals reduce to two scalars.
__device__ __forceinline__ void plate_block(const double* __restrict__ rec,
                                            const double* __restrict__ mat, int la, int lb,
                                            double acc[36]) {
  const double x1 = rec[9], y1 = rec[10], x2 = rec[11], y2 = rec[12], x4 = rec[13], y4 = rec[14];
  const double Cm = mat[0], Cb = mat[1], Cs = mat[2], nu = mat[3];
  const double g = 0.57735027779281512;  // sqrt((double)(1.0f / 3.0f))
  // natural-coordinate signs of nodes 1..4: (+,+), (-,+), (-,-), (+,-)
  const double xa = (la == 0 || la == 3) ? 1.0 : -1.0, ea = (la < 2) ? 1.0 : -1.0;
  const double xb = (lb == 0 || lb == 3) ? 1.0 : -1.0, eb = (lb < 2) ? 1.0 : -1.0;
  // edge differences (x3 = y3 = 0)
  const double ax12 = x1 - x2, ay12 = y1 - y2;  // edge 1-2 (s = +1)
  const double ax43 = x4, ay43 = y4;            // edge 4-3 (s = -1)
  const double ax14 = x1 - x4, ay14 = y1 - y4;  // edge 1-4 (r = +1)
  const double ax23 = x2, ay23 = y2;            // edge 2-3 (r = -1)

  double sxx = 0.0, sxy = 0.0, syx = 0.0, syy = 0.0, trz = 0.0, tsz = 0.0;
#pragma unroll
  for (int ip = 0; ip < 4; ++ip) {
    const double r = (ip == 0 || ip == 3) ? g : -g;
    const double s = (ip < 2) ? g : -g;
    // J = [[x_r, y_r], [x_s, y_s]]                 quadrilateral_4n_element_functions.rs:252-446
    double x_r = 0.25 * (ax12 * (1.0 + s) + ax43 * (1.0 - s));
    double y_r = 0.25 * (ay12 * (1.0 + s) + ay43 * (1.0 - s));
    double x_s = 0.25 * (ax14 * (1.0 + r) + ax23 * (1.0 - r));
    double y_s = 0.25 * (ay14 * (1.0 + r) + ay23 * (1.0 - r));
    double det = x_r * y_s - y_r * x_s;
    double rdet = 1.0 / det;
    // dh/dr, dh/ds of nodes a and b                quadrilateral_4n_element_functions.rs:505-583
    double dar = 0.25 * xa * (1.0 + ea * s), das = 0.25 * ea * (1.0 + xa * r);
    double dbr = 0.25 * xb * (1.0 + eb * s), dbs = 0.25 * eb * (1.0 + xb * r);
    // adj(J) * dh                                  quadrilateral_4n_element_functions.rs:613-653
    double nax = y_s * dar - y_r * das, nay = x_r * das - x_s * dar;
    double nbx = y_s * dbr - y_r * dbs, nby = x_r * dbs - x_s * dbr;
    double pax = nax * rdet, pay = nay * rdet;
    sxx += pax * nbx;
    sxy += pax * nby;
    syx += pay * nbx;
    syy += pay * nby;
    // shear: node factors (1 +- s) for gamma_rz rows, (1 +- r) for gamma_sz rows
    double fa = 1.0 + ea * s, fb = 1.0 + eb * s;
    double ga = 1.0 + xa * r, gb = 1.0 + xb * r;
    double q4 = 0.25 * rdet;
    trz += (fa * fb) * ((x_s * x_s + y_s * y_s) * q4);
    tsz += (ga * gb) * ((x_r * x_r + y_r * y_r) * q4);
  }

  const double gp = (1.0 - nu) * 0.5;
  // membrane on (u, v), bending on (thx, thy)
  double m00 = Cm * (sxx + gp * syy), m01 = Cm * (nu * sxy + gp * syx);
  double m10 = Cm * (nu * syx + gp * sxy), m11 = Cm * (syy + gp * sxx);
  double b33 = Cb * (syy + gp * sxx), b34 = -Cb * (nu * syx + gp * sxy);
  double b43 = -Cb * (nu * sxy + gp * syx), b44 = Cb * (sxx + gp * syy);
  // shear 3-vectors on (w, thx, thy): gamma_rz row uses the node's s-edge, gamma_sz its r-edge
  double erx = (la < 2) ? ax12 : ax43, ery = (la < 2) ? ay12 : ay43;
  double esx = (la == 0 || la == 3) ? ax14 : ax23, esy = (la == 0 || la == 3) ? ay14 : ay23;
  double arz[3] = {0.5 * xa, -0.25 * ery, 0.25 * erx};
  double asz[3] = {0.5 * ea, -0.25 * esy, 0.25 * esx};
  erx = (lb < 2) ? ax12 : ax43;
  ery = (lb < 2) ? ay12 : ay43;
  esx = (lb == 0 || lb == 3) ? ax14 : ax23;
  esy = (lb == 0 || lb == 3) ? ay14 : ay23;
  double brz[3] = {0.5 * xb, -0.25 * ery, 0.25 * erx};
  double bsz[3] = {0.5 * eb, -0.25 * esy, 0.25 * esx};
  double crz = Cs * trz, csz = Cs * tsz;
  double sh[9];
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int c = 0; c < 3; ++c) sh[3 * p + c] = (crz * arz[p]) * brz[c] + (csz * asz[p]) * bsz[c];
  double drill = (la == lb) ? 1.0 : 0.0;  // KROT6, plate.rs:25

  // local 6x6 (dof order u v w thx thy thz) as four 3x3 sub-blocks
  double uu[9] = {m00, m01, 0.0, m10, m11, 0.0, 0.0, 0.0, sh[0]};
  double ut[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, sh[1], sh[2], 0.0};
  double tu[9] = {0.0, 0.0, sh[3], 0.0, 0.0, sh[6], 0.0, 0.0, 0.0};
  double tt[9] = {b33 + sh[4], b34 + sh[5], 0.0, b43 + sh[7], b44 + sh[8], 0.0, 0.0, 0.0, drill};
  if (rec[15] != 0.0) {
    // Q == I exactly (flat plates in the global xy plane): (R^T k) R == k bit for bit
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        acc[6 * i + j] += uu[3 * i + j];
        acc[6 * i + 3 + j] += ut[3 * i + j];
        acc[6 * (3 + i) + j] += tu[3 * i + j];
        acc[6 * (3 + i) + 3 + j] += tt[3 * i + j];
      }
  } else {
    sandwich_full(rec, uu, acc, 0, 0);
    sandwich_full(rec, ut, acc, 0, 3);
    sandwich_full(rec, tu, acc, 3, 0);
    sandwich_full(rec, tt, acc, 3, 3);
  }
}

// ------------------------------------------------------------------------------------------
// Plate, shared form. Everything of a plate element that does not depend on the node pair
// (Jacobians, 1/det, adj(J)*dh of the four nodes at the four Gauss points, the assumed-strain
// shear sums, material constants) is evaluated ONCE per element per CTA into a 64-double record in
// shared memory (plate_shared_record); the node-pair evaluator (plate_block_shared) then only does
// the pair-specific part. Same quadrature and same f32-rounded abscissa as plate_block above.
//
// layout (doubles):
//    0..31  n[ip][node] = adj(J_ip) * [dh/dr; dh/ds]   as (x, y) pairs
//   32..35  1/det J_ip
//   36..38  Cs * sum_ip (1+ea s)(1+eb s) * gamma_rz^2 det  for (ea,eb) = (+,+), (-,-), (+,-)
//   39..41  same for gamma_sz with (xa,xb) and r
//   42..49  0.25 * edge vectors e12, e43, e14, e23 (x, y)
//   50..53  Cm, Cb, nu, (1-nu)/2
//   54..62  Q (row major)
//   63      1.0 when Q == I exactly
// ------------------------------------------------------------------------------------------
constexpr int kPlateSharedDoubles = 64;

// One Gauss point (r, s) of a plate with edge vectors e12, e43, e14, e23: n[node] = adj(J) dh
// (8 doubles) and 1/det into registers, returns the two shear weights gamma_rz^2 det and gamma_sz^2 det.
__device__ __forceinline__ void plate_gauss_point(const double e[8], double r, double s,
                                                  double n_out[8], double* rdet_out,
                                                  double* wr, double* ws) {
  // J = [[x_r, y_r], [x_s, y_s]]                 quadrilateral_4n_element_functions.rs:252-446
  const double x_r = 0.25 * (e[0] * (1.0 + s) + e[2] * (1.0 - s));
  const double y_r = 0.25 * (e[1] * (1.0 + s) + e[3] * (1.0 - s));
  const double x_s = 0.25 * (e[4] * (1.0 + r) + e[6] * (1.0 - r));
  const double y_s = 0.25 * (e[5] * (1.0 + r) + e[7] * (1.0 - r));
  const double det = x_r * y_s - y_r * x_s;
  const double rdet = 1.0 / det;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double xa = (a == 0 || a == 3) ? 1.0 : -1.0, ea = (a < 2) ? 1.0 : -1.0;
    // dh/dr, dh/ds                               quadrilateral_4n_element_functions.rs:505-583
    const double dar = 0.25 * xa * (1.0 + ea * s), das = 0.25 * ea * (1.0 + xa * r);
    // adj(J) * dh                                quadrilateral_4n_element_functions.rs:613-653
    n_out[2 * a] = y_s * dar - y_r * das;
    n_out[2 * a + 1] = x_r * das - x_s * dar;
  }
  *rdet_out = rdet;
  const double q4 = 0.25 * rdet;
  *wr = (x_s * x_s + y_s * y_s) * q4;  // gamma_rz^2 * det  (plate.rs:392-511)
  *ws = (x_r * x_r + y_r * y_r) * q4;  // gamma_sz^2 * det
}

// 16-byte store of two consecutive record fields (S is 16-byte aligned, i even): the shared-memory
// data pipe is the assembly kernel's busiest unit, and a 16-byte store costs the same wavefronts as
// an 8-byte one
__device__ __forceinline__ void store2(double* __restrict__ S, int i, double a, double b) {
  *reinterpret_cast<double2*>(S + i) = make_double2(a, b);
}

// raw = the prep kernel's record: Q[9], x1, y1, x2, y2, x4, y4, identity flag, Cm, Cb, Cs, nu
__device__ __forceinline__ void plate_shared_record(const double* __restrict__ raw,
                                                    double* __restrict__ S) {
  const double x1 = raw[9], y1 = raw[10], x2 = raw[11], y2 = raw[12], x4 = raw[13], y4 = raw[14];
  const double Cm = raw[16], Cb = raw[17], Cs = raw[18], nu = raw[19];
  const double g = 0.57735027779281512;  // sqrt((double)(1.0f / 3.0f)), plate.rs:1066-1091
  // edges 1-2 (s = +1), 4-3 (s = -1), 1-4 (r = +1), 2-3 (r = -1); x3 = y3 = 0
  const double e[8] = {x1 - x2, y1 - y2, x4, y4, x1 - x4, y1 - y4, x2, y2};
  double tr[3] = {0.0, 0.0, 0.0}, ts[3] = {0.0, 0.0, 0.0}, rd[4];
#pragma unroll
  for (int ip = 0; ip < 4; ++ip) {
    const double r = (ip == 0 || ip == 3) ? g : -g;
    const double s = (ip < 2) ? g : -g;
    double wr, ws, n[8];
    plate_gauss_point(e, r, s, n, &rd[ip], &wr, &ws);
#pragma unroll
    for (int a = 0; a < 4; ++a) store2(S, ip * 8 + 2 * a, n[2 * a], n[2 * a + 1]);
    tr[0] += ((1.0 + s) * (1.0 + s)) * wr;
    tr[1] += ((1.0 - s) * (1.0 - s)) * wr;
    tr[2] += ((1.0 + s) * (1.0 - s)) * wr;
    ts[0] += ((1.0 + r) * (1.0 + r)) * ws;
    ts[1] += ((1.0 - r) * (1.0 - r)) * ws;
    ts[2] += ((1.0 + r) * (1.0 - r)) * ws;
  }
  store2(S, 32, rd[0], rd[1]);
  store2(S, 34, rd[2], rd[3]);
  store2(S, 36, Cs * tr[0], Cs * tr[1]);
  store2(S, 38, Cs * tr[2], Cs * ts[0]);
  store2(S, 40, Cs * ts[1], Cs * ts[2]);
#pragma unroll
  for (int i = 0; i < 4; ++i) store2(S, 42 + 2 * i, 0.25 * e[2 * i], 0.25 * e[2 * i + 1]);
  store2(S, 50, Cm, Cb);
  store2(S, 52, nu, (1.0 - nu) * 0.5);
#pragma unroll
  for (int i = 0; i < 4; ++i) store2(S, 54 + 2 * i, raw[2 * i], raw[2 * i + 1]);
  store2(S, 62, raw[8], raw[15]);
}

// The same record built by TWO adjacent lanes (half = lane & 1): each takes two Gauss points
// (half 0: s = +g, half 1: s = -g), the shear sums are combined with one shuffle exchange, and the
// point-independent fields are shared out between the two. Must be called by both lanes of a pair.
__device__ __forceinline__ void plate_shared_record_half(const double* __restrict__ raw,
                                                         double* __restrict__ S, int half,
                                                         uint32_t pair_mask) {
  const double x1 = raw[9], y1 = raw[10], x2 = raw[11], y2 = raw[12], x4 = raw[13], y4 = raw[14];
  const double g = 0.57735027779281512;
  const double e[8] = {x1 - x2, y1 - y2, x4, y4, x1 - x4, y1 - y4, x2, y2};
  const double s = half ? -g : g;
  double tr[3] = {0.0, 0.0, 0.0}, ts[3] = {0.0, 0.0, 0.0}, rd[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    // Gauss points 0:(+,+) 1:(-,+) | 2:(-,-) 3:(+,-)
    const int ip = 2 * half + q;
    const double r = ((q == 0) != (half != 0)) ? g : -g;
    double wr, ws, n[8];
    plate_gauss_point(e, r, s, n, &rd[q], &wr, &ws);
#pragma unroll
    for (int a = 0; a < 4; ++a) store2(S, ip * 8 + 2 * a, n[2 * a], n[2 * a + 1]);
    tr[0] += ((1.0 + s) * (1.0 + s)) * wr;
    tr[1] += ((1.0 - s) * (1.0 - s)) * wr;
    tr[2] += ((1.0 + s) * (1.0 - s)) * wr;
    ts[0] += ((1.0 + r) * (1.0 + r)) * ws;
    ts[1] += ((1.0 - r) * (1.0 - r)) * ws;
    ts[2] += ((1.0 + r) * (1.0 - r)) * ws;
  }
  store2(S, 32 + 2 * half, rd[0], rd[1]);
  // half 0 finishes the gamma_rz sums, half 1 the gamma_sz sums (fixed order: half 0 + half 1)
  const double Cs = raw[18];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double mine = half ? ts[i] : tr[i], give = half ? tr[i] : ts[i];
    const double got = __shfl_xor_sync(pair_mask, give, 1);
    const double sum = half ? got + mine : mine + got;
    S[(half ? 39 : 36) + i] = Cs * sum;
  }
  if (half == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) store2(S, 42 + 2 * i, 0.25 * e[2 * i], 0.25 * e[2 * i + 1]);
    store2(S, 50, raw[16], raw[17]);
    store2(S, 52, raw[19], (1.0 - raw[19]) * 0.5);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) store2(S, 54 + 2 * i, raw[2 * i], raw[2 * i + 1]);
    store2(S, 62, raw[8], raw[15]);
  }
}

// Everything plate_block_shared derives from the local node pair (la, lb) alone, packed into 16 bytes:
// where the two nodes' entries sit inside the shared record (byte offsets < 512) and the signs of
// their natural coordinates. The staged kernel keeps the 16 possible entries in shared memory and
// fetches the entry of the NEXT contribution one loop trip ahead, so a contribution starts with its
// addresses in registers instead of a dependent chain of loads.
//   x: n[0][la] | n[0][lb] << 16                              (stride between Gauss points: 64 B)
//   y: node a's s-edge (1-2 or 4-3, for gamma_rz) | its r-edge (1-4 or 2-3, for gamma_sz) << 16
//   z: the same two for node b
//   w: shear sum of the (ea, eb) combination | of the (xa, xb) combination << 10
//      | bits 27..31: diagonal pair (KROT6, plate.rs:25), xi_a < 0, eta_a < 0, xi_b < 0, eta_b < 0
typedef uint4 PlatePair;
__host__ __device__ inline PlatePair make_plate_pair(int la, int lb) {
  // natural-coordinate signs of nodes 1..4: (+,+), (-,+), (-,-), (+,-)
  const uint32_t an = (la ^ (la >> 1)) & 1, bn = (lb ^ (lb >> 1)) & 1;  // 1 when xi = -1
  const uint32_t am = la >> 1, bm = lb >> 1;                            // 1 when eta = -1
  PlatePair t;
  t.x = uint32_t(la) * 16u | (uint32_t(lb) * 16u) << 16;
  t.y = (42u + 2u * am) * 8u | ((46u + 2u * an) * 8u) << 16;
  t.z = (42u + 2u * bm) * 8u | ((46u + 2u * bn) * 8u) << 16;
  t.w = (36u + ((am == bm) ? am : 2u)) * 8u | ((39u + ((an == bn) ? an : 2u)) * 8u) << 10 |
        (la == lb ? 1u << 27 : 0u) | an << 28 | am << 29 | bn << 30 | bm << 31;
  return t;
}
// +-0.5 with the sign taken from bit `bit` of w
__device__ __forceinline__ double plate_half_sign(uint32_t w, int bit) {
  return __hiloint2double(int(0x3FE00000u | ((w << (31 - bit)) & 0x80000000u)), 0);
}

// acc = keep * acc + block (la, lb) of (R^T k) R of the plate whose shared record is S; keep is 1
// (accumulate) or 0 (first contribution of a block: no separate zeroing of the accumulators).
// all_flat: every plate of the slab has Q == I exactly (flat plates in the global xy plane), then
// (R^T k) R == k and only the 14 structural entries of the local block are touched — the other 22
// accumulators are never written by a flat plate and must already be zero.
__device__ __forceinline__ void plate_block_shared(const double* __restrict__ S, const PlatePair pt,
                                                   double keep, bool all_flat, double acc[36]) {
  const uint4 o0 = make_uint4(pt.x & 0xFFFFu, pt.x >> 16, pt.y & 0xFFFFu, pt.y >> 16);
  const uint4 o1 = make_uint4(pt.z & 0xFFFFu, pt.z >> 16, pt.w & 0x3FFu, (pt.w >> 10) & 0x3FFu);
  const double2 sa = make_double2(plate_half_sign(pt.w, 28), plate_half_sign(pt.w, 29));
  const double2 sb = make_double2(plate_half_sign(pt.w, 30), plate_half_sign(pt.w, 31));
  const char* Sb = reinterpret_cast<const char*>(S);
  const double2 rd01 = *reinterpret_cast<const double2*>(S + 32);
  const double2 rd23 = *reinterpret_cast<const double2*>(S + 34);
  const double rd[4] = {rd01.x, rd01.y, rd23.x, rd23.y};
  double sxx = 0.0, sxy = 0.0, syx = 0.0, syy = 0.0;
#pragma unroll
  for (int ip = 0; ip < 4; ++ip) {
    const double2 na = *reinterpret_cast<const double2*>(Sb + o0.x + ip * 64);
    const double2 nb = *reinterpret_cast<const double2*>(Sb + o0.y + ip * 64);
    const double pax = na.x * rd[ip], pay = na.y * rd[ip];
    sxx += pax * nb.x;
    sxy += pax * nb.y;
    syx += pay * nb.x;
    syy += pay * nb.y;
  }
  const double2 c01 = *reinterpret_cast<const double2*>(S + 50);
  const double2 c23 = *reinterpret_cast<const double2*>(S + 52);
  const double Cm = c01.x, Cb = c01.y, nu = c23.x, gp = c23.y;
  const double t1 = sxx + gp * syy, t2 = nu * sxy + gp * syx;
  const double t3 = nu * syx + gp * sxy, t4 = syy + gp * sxx;
  const double m00 = Cm * t1, m01 = Cm * t2, m10 = Cm * t3, m11 = Cm * t4;
  const double b33 = Cb * t4, b34 = -(Cb * t3), b43 = -(Cb * t2), b44 = Cb * t1;
  // gamma_rz rows use the node's s-edge (1-2 or 4-3), gamma_sz rows its r-edge (1-4 or 2-3)
  const double2 era = *reinterpret_cast<const double2*>(Sb + o0.z);
  const double2 esa = *reinterpret_cast<const double2*>(Sb + o0.w);
  const double2 erb = *reinterpret_cast<const double2*>(Sb + o1.x);
  const double2 esb = *reinterpret_cast<const double2*>(Sb + o1.y);
  const double crz = *reinterpret_cast<const double*>(Sb + o1.z);
  const double csz = *reinterpret_cast<const double*>(Sb + o1.w);
  const double arz[3] = {crz * sa.x, crz * -era.y, crz * era.x};
  const double asz[3] = {csz * sa.y, csz * -esa.y, csz * esa.x};
  const double brz[3] = {sb.x, -erb.y, erb.x};
  const double bsz[3] = {sb.y, -esb.y, esb.x};
  double sh[9];
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int c = 0; c < 3; ++c) sh[3 * p + c] = arz[p] * brz[c] + asz[p] * bsz[c];
  const double drill = (pt.w & (1u << 27)) ? 1.0 : 0.0;
  if (all_flat) {
    acc[0] = fma(acc[0], keep, m00);
    acc[1] = fma(acc[1], keep, m01);
    acc[6] = fma(acc[6], keep, m10);
    acc[7] = fma(acc[7], keep, m11);
    acc[14] = fma(acc[14], keep, sh[0]);
    acc[15] = fma(acc[15], keep, sh[1]);
    acc[16] = fma(acc[16], keep, sh[2]);
    acc[20] = fma(acc[20], keep, sh[3]);
    acc[26] = fma(acc[26], keep, sh[6]);
    acc[21] = fma(acc[21], keep, b33 + sh[4]);
    acc[22] = fma(acc[22], keep, b34 + sh[5]);
    acc[27] = fma(acc[27], keep, b43 + sh[7]);
    acc[28] = fma(acc[28], keep, b44 + sh[8]);
    acc[35] = fma(acc[35], keep, drill);
  } else {
    // general orientation; for Q == I the sandwich reproduces the local block exactly (1*x + 0*y + 0*z)
#pragma unroll
    for (int i = 0; i < 36; ++i) acc[i] *= keep;
    const double* q = S + 54;
    const double uu[9] = {m00, m01, 0.0, m10, m11, 0.0, 0.0, 0.0, sh[0]};
    const double ut[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, sh[1], sh[2], 0.0};
    const double tu[9] = {0.0, 0.0, sh[3], 0.0, 0.0, sh[6], 0.0, 0.0, 0.0};
    const double tt[9] = {b33 + sh[4], b34 + sh[5], 0.0, b43 + sh[7], b44 + sh[8], 0.0, 0.0, 0.0, drill};
    sandwich_full(q, uu, acc, 0, 0);
    sandwich_full(q, ut, acc, 0, 3);
    sandwich_full(q, tu, acc, 3, 0);
    sandwich_full(q, tt, acc, 3, 3);
  }
}

}  // namespace femgpu
