// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement (C++17, templated on the float type V so the reference's own
// f32 known-answer test can be replayed) of the structural-FEM stiffness hot
// path of RomanShushakov/finite_element_method v0.9.12:
//   truss / beam / plate local stiffness -> R^T k R -> zero-skip block scatter
//   into the global matrix.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may build, load or call this. The product library
// (finite_element_method_b200/csrc) never includes or links it.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/src/fem/).
//
// PARITY STATUS
//   * Truss scalar EA/L and the solved displacement are pinned by the
//     reference's own f32 test (tests/fem/test_fem.rs:5-64): see
//     tests/test_oracle_golden.py.
//   * Everything that depends on the un-vendored crate `extended_matrix`
//     ("0.9.9", Cargo.toml:12; no lockfile, not on this machine) is restated
//     from its published behaviour and is PARITY UNPINNED beyond that test:
//       - Matrix::multiply            : triple loop, inner index ascending,
//                                       accumulator starts at 0
//       - Vector3::norm/dot/cross     : textbook, left-to-right sums
//       - rotation_matrix_to_align_with_vector : Rodrigues about a x b by
//                                       acos(a.b/(|a||b|)), cos/sin and each
//                                       entry clipped by abs_tol; zero axis
//                                       when a x b == 0
//       - projection_perpendicular_to_vector : a - b*(a.b/(b.b))
//       - 2x2 inverse / determinant   : closed form
//     Beam and plate element matrices are therefore "parity unpinned"
//     (the reference has no beam/plate test); analytic checks in tests/ back them.
//     `Variants` below switches the two restatement choices that could matter (the anti-parallel
//     branch of the Rodrigues rotation, the 2x2 inverse); tests/test_oracle_variants.py shows which
//     results do not depend on them (trusses always; plates and beams unless a normal / member is
//     EXACTLY anti-parallel to its target axis) and quantifies the ones that do.
//
// Build: g++ -O2 -ffp-contract=off (Rust never contracts a*b+c into an FMA).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

namespace oracle {

// --------------------------------------------------------------------------
// Variants of the two places where the un-vendored extended_matrix crate could plausibly differ from the
// restatement below (tests/test_oracle_variants.py runs every family through all of them to show which
// results depend on the unknown and which do not). Test-only switches; the defaults are the restatement.
//   antiparallel: what rotation_matrix_to_align_with_vector(a, b) returns when a x b == 0 and a.b < 0
//     0  the same guard as the parallel case: zero axis, c = -1  ->  Q = -I          (default)
//     1  rotation by pi about the y axis   diag(-1, 1, -1)
//     2  rotation by pi about the z axis   diag(-1, -1, 1)
//     3  identity (the guard returns before the formula)
//   inverse2: SquareMatrix::inverse / determinant of the 2x2 Jacobian (quadrilateral...rs:448-503)
//     0  closed form (adjugate / determinant)                                        (default)
//     1  LU (Doolittle, no pivoting): forward / back substitution per column, det = u11 u22
//     2  Gauss-Jordan with partial pivoting
// --------------------------------------------------------------------------
struct Variants {
  int antiparallel = 0;
  int inverse2 = 0;
  long antiparallel_hits = 0;  // how often the anti-parallel branch was taken since the last reset
};
inline Variants& variants() {
  static Variants v;
  return v;
}

// --------------------------------------------------------------------------
// math_functions.rs
// --------------------------------------------------------------------------

// math_functions.rs:3-12
template <typename V>
inline V compare_with_tolerance(V value, V abs_tol) {
  if (std::fabs(value) < abs_tol) return V(0.0f);
  return value;
}

// math_functions.rs:14-19   a * x^n by repeated multiplication (n<=0 -> a)
template <typename V>
inline V power_func_x(V a, V x, int n) {
  V acc = a;
  for (int i = 0; i < n; ++i) acc = acc * x;
  return acc;
}

// math_functions.rs:21-28   d/dx (a x^n) = (a*n) x^(n-1); n is rebuilt by adding 1.0 n times
template <typename V>
inline V derivative_x(V a, V x, int n) {
  V converted_n = V(0.0f);
  for (int i = 0; i < n; ++i) converted_n += V(1.0f);
  return power_func_x(a * converted_n, x, n - 1);
}

// --------------------------------------------------------------------------
// extended_matrix stand-ins (assumed behaviour, see header)
// --------------------------------------------------------------------------

template <typename V>
struct Vec3 {
  V c[3];
};

template <typename V>
inline V v3_norm(const Vec3<V>& a) {
  V acc = V(0.0f);
  for (int i = 0; i < 3; ++i) acc += a.c[i] * a.c[i];
  return std::sqrt(acc);
}

template <typename V>
inline V v3_dot(const Vec3<V>& a, const Vec3<V>& b) {
  V acc = V(0.0f);
  for (int i = 0; i < 3; ++i) acc += a.c[i] * b.c[i];
  return acc;
}

template <typename V>
inline Vec3<V> v3_cross(const Vec3<V>& a, const Vec3<V>& b) {
  return Vec3<V>{{a.c[1] * b.c[2] - a.c[2] * b.c[1], a.c[2] * b.c[0] - a.c[0] * b.c[2],
                  a.c[0] * b.c[1] - a.c[1] * b.c[0]}};
}

template <typename V>
inline V v3_cosine_angle_between(const Vec3<V>& a, const Vec3<V>& b) {
  return v3_dot(a, b) / (v3_norm(a) * v3_norm(b));
}

// a - b * (a.b / b.b)
template <typename V>
inline Vec3<V> v3_projection_perpendicular_to(const Vec3<V>& a, const Vec3<V>& b) {
  V f = v3_dot(a, b) / v3_dot(b, b);
  return Vec3<V>{{a.c[0] - b.c[0] * f, a.c[1] - b.c[1] * f, a.c[2] - b.c[2] * f}};
}

// Rodrigues rotation taking `a` onto `b` (call sites: truss.rs:78-84,
// beam.rs:177-183, quadrilateral_4n_element_functions.rs:158-162).
template <typename V>
inline void rotation_matrix_to_align_with_vector(const Vec3<V>& a, const Vec3<V>& b, V /*rel_tol*/,
                                                 V abs_tol, V q[9]) {
  V na = v3_norm(a), nb = v3_norm(b);
  V cosv = v3_dot(a, b) / (na * nb);
  // guard acos domain against 1+ulp (Rust acos would give NaN; a well-formed
  // input never exceeds 1 by more than rounding)
  if (cosv > V(1.0f)) cosv = V(1.0f);
  if (cosv < V(-1.0f)) cosv = V(-1.0f);
  V angle = std::acos(cosv);
  Vec3<V> axis = v3_cross(a, b);
  V n = v3_norm(axis);
  V x = V(0.0f), y = V(0.0f), z = V(0.0f);
  if (n != V(0.0f)) {
    x = axis.c[0] / n;
    y = axis.c[1] / n;
    z = axis.c[2] / n;
  }
  if (n == V(0.0f) && cosv < V(0.0f)) {  // anti-parallel: the one input class whose result is not pinned
#pragma omp atomic
    variants().antiparallel_hits++;
    const int mode = variants().antiparallel;
    if (mode != 0) {
      const V one = V(1.0f), m1 = V(-1.0f), z0 = V(0.0f);
      const V d[3][3] = {{m1, one, m1}, {m1, m1, one}, {one, one, one}};
      for (int i = 0; i < 9; ++i) q[i] = z0;
      q[0] = d[mode - 1][0];
      q[4] = d[mode - 1][1];
      q[8] = d[mode - 1][2];
      return;
    }
  }
  V c = compare_with_tolerance(std::cos(angle), abs_tol);
  V s = compare_with_tolerance(std::sin(angle), abs_tol);
  V t = V(1.0f) - c;
  q[0] = compare_with_tolerance(t * x * x + c, abs_tol);
  q[1] = compare_with_tolerance(t * x * y - z * s, abs_tol);
  q[2] = compare_with_tolerance(t * x * z + y * s, abs_tol);
  q[3] = compare_with_tolerance(t * x * y + z * s, abs_tol);
  q[4] = compare_with_tolerance(t * y * y + c, abs_tol);
  q[5] = compare_with_tolerance(t * y * z - x * s, abs_tol);
  q[6] = compare_with_tolerance(t * x * z - y * s, abs_tol);
  q[7] = compare_with_tolerance(t * y * z + x * s, abs_tol);
  q[8] = compare_with_tolerance(t * z * z + c, abs_tol);
}

// Small dense row-major matrix with the (assumed) extended_matrix arithmetic.
template <typename V>
struct Mat {
  int r = 0, c = 0;
  std::vector<V> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a(size_t(r_) * c_, V(0.0f)) {}
  Mat(int r_, int c_, std::initializer_list<V> v) : r(r_), c(c_), a(v) {}
  V& at(int i, int j) { return a[size_t(i) * c + j]; }
  const V& at(int i, int j) const { return a[size_t(i) * c + j]; }
  Mat transpose() const {
    Mat t(c, r);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < c; ++j) t.at(j, i) = at(i, j);
    return t;
  }
  Mat multiply(const Mat& o) const {
    Mat m(r, o.c);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < o.c; ++j) {
        V acc = V(0.0f);
        for (int k = 0; k < c; ++k) acc += at(i, k) * o.at(k, j);
        m.at(i, j) = acc;
      }
    return m;
  }
  Mat multiply_by_scalar(V s) const {
    Mat m(*this);
    for (auto& v : m.a) v = v * s;
    return m;
  }
  Mat add(const Mat& o) const {
    Mat m(*this);
    for (size_t i = 0; i < a.size(); ++i) m.a[i] = a[i] + o.a[i];
    return m;
  }
  Mat subtract(const Mat& o) const {
    Mat m(*this);
    for (size_t i = 0; i < a.size(); ++i) m.a[i] = a[i] - o.a[i];
    return m;
  }
};

// --------------------------------------------------------------------------
// bar_2n_element_functions.rs
// --------------------------------------------------------------------------

// bar_2n_element_functions.rs:8-33  (node_2 - node_1)
template <typename V>
inline Vec3<V> find_2n_element_vector(const V p1[3], const V p2[3]) {
  return Vec3<V>{{p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}};
}

// bar_2n_element_functions.rs:35-43
template <typename V>
inline V bar_dx_dr(V x_1, V x_2, V r) {
  return derivative_x(x_1 * V(0.5f), V(0.0f), 0) - derivative_x(x_1 * V(0.5f), r, 1) +
         derivative_x(x_2 * V(0.5f), V(0.0f), 0) + derivative_x(x_2 * V(0.5f), r, 1);
}

// bar_2n_element_functions.rs:45-59
template <typename V>
inline V bar_jacobian_at_r(const V p1[3], const V p2[3], V r) {
  V len = v3_norm(find_2n_element_vector(p1, p2));
  V x_1 = V(-1.0f) * len / V(2.0f);
  V x_2 = len / V(2.0f);
  return bar_dx_dr(x_1, x_2, r);
}
// :61-71
template <typename V>
inline V bar_inverse_jacobian_at_r(const V p1[3], const V p2[3], V r) {
  return V(1.0f) / bar_jacobian_at_r(p1, p2, r);
}
// :73-83
template <typename V>
inline V bar_determinant_of_jacobian_at_r(const V p1[3], const V p2[3], V r) {
  return bar_jacobian_at_r(p1, p2, r);
}
// :85-113
template <typename V>
inline V h1_r(V r) { return V(0.5f) * (V(1.0f) - r); }
template <typename V>
inline V h2_r(V r) { return V(0.5f) * (V(1.0f) + r); }
template <typename V>
inline V dh1_dr(V r) { return derivative_x(V(0.5f), V(0.0f), 0) - derivative_x(V(0.5f), r, 1); }
template <typename V>
inline V dh2_dr(V r) { return derivative_x(V(0.5f), V(0.0f), 0) + derivative_x(V(0.5f), r, 1); }

// --------------------------------------------------------------------------
// Error codes shared by the three element families (texts in oracle_capi.cpp)
// --------------------------------------------------------------------------
enum ElementError : int {
  OK = 0,
  E_YOUNG = 1,
  E_POISSON = 2,
  E_AREA = 3,
  E_AREA2 = 4,
  E_I11 = 5,
  E_I22 = 6,
  E_IT = 7,
  E_SHEAR_FACTOR = 8,
  E_PARALLEL_AXIS = 9,
  E_THICKNESS = 10,
  E_ON_LINE = 11,
  E_NOT_ON_PLANE = 12,
  E_NOT_CONVEX = 13,
};

// --------------------------------------------------------------------------
// truss.rs
// --------------------------------------------------------------------------
template <typename V>
struct TrussOut {
  V q[9];
  Mat<V> k_local;   // 6x6
  Mat<V> k_global;  // 6x6 = (R^T k) R
};

// truss.rs:44-64
template <typename V>
inline int check_truss_properties(V young_modulus, V area, bool has_area_2, V area_2) {
  if (young_modulus <= V(0.0f)) return E_YOUNG;
  if (area <= V(0.0f)) return E_AREA;
  if (has_area_2 && area_2 <= V(0.0f)) return E_AREA2;
  return OK;
}

// truss.rs:66-93
template <typename V>
inline void truss_find_rotation_matrix_elements(const V p1[3], const V p2[3], V rel_tol, V abs_tol,
                                                V q[9]) {
  Vec3<V> v = find_2n_element_vector(p1, p2);
  V len = v3_norm(v);
  Vec3<V> dir{{len, V(0.0f), V(0.0f)}};
  rotation_matrix_to_align_with_vector(v, dir, rel_tol, abs_tol, q);
}

// truss.rs:95-118
template <typename V>
inline Mat<V> truss_strain_displacement_matrix_at_r(const V p1[3], const V p2[3], V r) {
  V inv_j = bar_inverse_jacobian_at_r(p1, p2, r);
  Mat<V> b(1, 6, {dh1_dr(r), V(0.0f), V(0.0f), dh2_dr(r), V(0.0f), V(0.0f)});
  return b.multiply_by_scalar(inv_j);
}

// truss.rs:120-130
template <typename V>
inline V truss_area_at_r(V area, bool has_area_2, V area_2, V r) {
  if (has_area_2) {
    return (area_2 - area) / V(2.0f) * r + area - (area_2 - area) / V(2.0f) * V(-1.0f);
  }
  return area;
}

// truss.rs:132-155
template <typename V>
inline Mat<V> truss_local_stiffness_matrix_at_ip(const V p1[3], const V p2[3], V young_modulus,
                                                 V area, bool has_area_2, V area_2, V r, V alpha) {
  Mat<V> b = truss_strain_displacement_matrix_at_r(p1, p2, r);
  Mat<V> bt = b.transpose();
  V c_at_r = truss_area_at_r(area, has_area_2, area_2, r) * young_modulus;
  return bt.multiply(b).multiply_by_scalar(c_at_r * bar_determinant_of_jacobian_at_r(p1, p2, r) *
                                           alpha);
}

// truss.rs:191-214   R = diag(Q, Q)
template <typename V>
inline Mat<V> compose_rotation_matrix_3dof(const V q[9], int n_nodes) {
  Mat<V> rm(3 * n_nodes, 3 * n_nodes);
  for (int n = 0; n < n_nodes; ++n)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) rm.at(3 * n + i, 3 * n + j) = q[3 * i + j];
  return rm;
}

// Truss::create (truss.rs:230-255) + extract_local_stiffness_matrix (:266-279,
// :157-189) + (R^T k) R of add_truss (methods_for_truss_data_handle.rs:75-81)
template <typename V>
inline int truss_element(const V p1[3], const V p2[3], V young_modulus, V area, bool has_area_2,
                         V area_2, V rel_tol, V abs_tol, TrussOut<V>& out) {
  int err = check_truss_properties(young_modulus, area, has_area_2, area_2);
  if (err) return err;
  truss_find_rotation_matrix_elements(p1, p2, rel_tol, abs_tol, out.q);
  const V ips[1][2] = {{V(0.0f), V(2.0f)}};  // truss.rs:244
  Mat<V> k(6, 6);
  for (auto& ip : ips)
    k = k.add(truss_local_stiffness_matrix_at_ip(p1, p2, young_modulus, area, has_area_2, area_2,
                                                 ip[0], ip[1]));
  out.k_local = k;
  Mat<V> rm = compose_rotation_matrix_3dof(out.q, 2);
  out.k_global = rm.transpose().multiply(k).multiply(rm);
  return OK;
}

// --------------------------------------------------------------------------
// beam.rs
// --------------------------------------------------------------------------
template <typename V>
struct BeamOut {
  V q[9];
  V i11_p, i22_p, angle;
  Mat<V> k_local;   // 12x12
  Mat<V> k_global;  // 12x12
};

// beam.rs:63-117 (note the reference reports young_modulus in the Poisson message;
// only the code matters here)
template <typename V>
inline int check_beam_properties(V young_modulus, V poisson_ratio, V area, V i11, V i22, V it,
                                 V shear_factor, const V p1[3], const V p2[3], const V axis1[3]) {
  if (young_modulus <= V(0.0f)) return E_YOUNG;
  if (poisson_ratio <= V(0.0f)) return E_POISSON;
  if (area <= V(0.0f)) return E_AREA;
  if (i11 <= V(0.0f)) return E_I11;
  if (i22 <= V(0.0f)) return E_I22;
  if (it <= V(0.0f)) return E_IT;
  if (shear_factor <= V(0.0f)) return E_SHEAR_FACTOR;
  Vec3<V> v = find_2n_element_vector(p1, p2);
  Vec3<V> a1{{axis1[0], axis1[1], axis1[2]}};
  Vec3<V> proj = v3_projection_perpendicular_to(a1, v);
  if (v3_norm(proj) == V(0.0f)) return E_PARALLEL_AXIS;
  return OK;
}

// beam.rs:119-160
template <typename V>
inline void find_principal_moments_of_inertia(V i11, V i22, V i12, V rel_tol, V& i11_p, V& i22_p,
                                              V& angle) {
  const V PI_F32 = V(3.14159265358979323846f);  // V::from(std::f32::consts::PI), beam.rs:7,149
  if (i11 != i22) {
    angle = std::atan(V(2.0f) * i12 / (i22 - i11)) / V(2.0f);
  } else {
    V i11_mod, i22_mod;
    if (i22 < i11) {
      i11_mod = i11;
      i22_mod = (std::fabs(i22) - std::fabs(i22) * rel_tol) * i22 / std::fabs(i22);
    } else {
      i11_mod = (std::fabs(i11) - std::fabs(i11) * rel_tol) * i11 / std::fabs(i11);
      i22_mod = i22;
    }
    angle = std::atan(V(2.0f) * i12 / (i22_mod - i11_mod)) / V(2.0f);
  }
  auto sq = [](V x) { return x * x; };  // my_powi(2)
  i11_p = i11 * sq(std::cos(angle)) + i22 * sq(std::sin(angle)) - i12 * std::sin(V(2.0f) * angle);
  i22_p = i11 * sq(std::sin(angle)) + i22 * sq(std::cos(angle)) + i12 * std::sin(V(2.0f) * angle);
  int i = 1;
  while (i11_p < i22_p) {
    angle = (std::atan(V(2.0f) * i12 / (i22 - i11)) + PI_F32 * V(float(i))) / V(2.0f);
    i11_p = i11 * sq(std::cos(angle)) + i22 * sq(std::sin(angle)) - i12 * std::sin(V(2.0f) * angle);
    i22_p = i11 * sq(std::sin(angle)) + i22 * sq(std::cos(angle)) + i12 * std::sin(V(2.0f) * angle);
    i += 1;
    if (i > 64) break;  // the reference would spin forever on NaN input; the oracle does not
  }
}

// beam.rs:162-258
template <typename V>
inline void beam_find_rotation_matrix_elements(const V p1[3], const V p2[3], const V axis1[3],
                                               V angle, V rel_tol, V abs_tol, V r[9]) {
  Vec3<V> v = find_2n_element_vector(p1, p2);
  V len = v3_norm(v);
  Vec3<V> dir{{len, V(0.0f), V(0.0f)}};
  V qi[9];
  rotation_matrix_to_align_with_vector(v, dir, rel_tol, abs_tol, qi);
  Vec3<V> a1{{axis1[0], axis1[1], axis1[2]}};
  Vec3<V> proj = v3_projection_perpendicular_to(a1, v);
  // interim_rotation_matrix.multiply(&projection)
  Vec3<V> tp;
  for (int i = 0; i < 3; ++i) {
    V acc = V(0.0f);
    for (int k = 0; k < 3; ++k) acc += qi[3 * i + k] * proj.c[k];
    tp.c[i] = acc;
  }
  Vec3<V> ez{{V(0.0f), V(0.0f), V(1.0f)}};
  V cosv = v3_cosine_angle_between(ez, tp);
  if (cosv > V(1.0f)) cosv = V(1.0f);
  if (cosv < V(-1.0f)) cosv = V(-1.0f);
  V angle_t = std::acos(cosv);
  V total_angle = angle + angle_t;
  V x = v.c[0], y = v.c[1], z = v.c[2];
  V c_x = compare_with_tolerance(x / len, abs_tol);
  V c_y = compare_with_tolerance(y / len, abs_tol);
  V c_z = compare_with_tolerance(z / len, abs_tol);
  V c_xz = compare_with_tolerance(std::sqrt(c_x * c_x + c_z * c_z), abs_tol);
  V c = compare_with_tolerance(std::cos(total_angle), abs_tol);
  V s = compare_with_tolerance(std::sin(total_angle), abs_tol);
  const V zero = V(0.0f);
  bool nz = c_xz != zero;
  r[0] = nz ? c_x : zero;
  r[1] = c_y;
  r[2] = nz ? c_z : zero;
  r[3] = nz ? (V(-1.0f) * c_x * c_y * c - c_z * s) / c_xz : V(-1.0f) * c_y * c;
  r[4] = nz ? c_xz * c : zero;
  r[5] = nz ? (V(-1.0f) * c_y * c_z * c + c_x * s) / c_xz : s;
  r[6] = nz ? (c_x * c_y * s - c_z * c) / c_xz : c_y * s;
  r[7] = nz ? V(-1.0f) * c_xz * s : zero;
  r[8] = nz ? (c_y * c_z * s + c_x * c) / c_xz : c;
}

// beam.rs:260-493: six 1x12 strain-displacement rows at r
template <typename V>
inline Mat<V> beam_b_row(const V p1[3], const V p2[3], V r, int which) {
  V inv_j = bar_inverse_jacobian_at_r(p1, p2, r);
  Mat<V> lhs(1, 12);
  // u,v,w,thu,thv,thw -> derivative columns 0..5 (+6 for node 2)
  lhs.at(0, which) = dh1_dr(r);
  lhs.at(0, which + 6) = dh2_dr(r);
  lhs = lhs.multiply_by_scalar(inv_j);
  if (which == 1 || which == 2) {
    // v couples with thw (col 5/11), w couples with thv (col 4/10): lhs - rhs
    Mat<V> rhs(1, 12);
    int col = (which == 1) ? 5 : 4;
    rhs.at(0, col) = h1_r(r);
    rhs.at(0, col + 6) = h2_r(r);
    return lhs.subtract(rhs);
  }
  return lhs;
}

// beam.rs:495-563
template <typename V>
inline Mat<V> beam_local_stiffness_matrix_at_ip(const V p1[3], const V p2[3], V young_modulus,
                                                V poisson_ratio, V area, V i11_p, V i22_p, V it,
                                                V shear_factor, V r, V alpha) {
  V det = bar_determinant_of_jacobian_at_r(p1, p2, r);
  V shear_modulus = young_modulus / (V(2.0f) * (V(1.0f) + poisson_ratio));
  V cs[6] = {area * young_modulus,
             shear_modulus * area * shear_factor,
             shear_modulus * area * shear_factor,
             shear_modulus * it,
             young_modulus * i22_p,
             young_modulus * i11_p};
  Mat<V> k;
  for (int w = 0; w < 6; ++w) {
    Mat<V> b = beam_b_row(p1, p2, r, w);
    Mat<V> kw = b.transpose().multiply(b).multiply_by_scalar(cs[w] * det * alpha);
    k = (w == 0) ? kw : k.add(kw);
  }
  return k;
}

// beam.rs:607-667 / plate.rs:726-1002   R = diag(Q x (2*n_nodes)) for 6-dof nodes
template <typename V>
inline Mat<V> compose_rotation_matrix_6dof(const V q[9], int n_nodes) {
  Mat<V> rm(6 * n_nodes, 6 * n_nodes);
  for (int n = 0; n < 2 * n_nodes; ++n)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) rm.at(3 * n + i, 3 * n + j) = q[3 * i + j];
  return rm;
}

// Beam::create (beam.rs:687-745) + extract_local_stiffness_matrix (:756-773, :565-605)
// + (R^T k) R of add_beam (methods_for_beam_data_handle.rs:88-95)
template <typename V>
inline int beam_element(const V p1[3], const V p2[3], V young_modulus, V poisson_ratio, V area,
                        V i11, V i22, V i12, V it, V shear_factor, const V axis1[3], V rel_tol,
                        V abs_tol, BeamOut<V>& out) {
  int err = check_beam_properties(young_modulus, poisson_ratio, area, i11, i22, it, shear_factor,
                                  p1, p2, axis1);
  if (err) return err;
  find_principal_moments_of_inertia(i11, i22, i12, rel_tol, out.i11_p, out.i22_p, out.angle);
  beam_find_rotation_matrix_elements(p1, p2, axis1, out.angle, rel_tol, abs_tol, out.q);
  const V ips[1][2] = {{V(0.0f), V(2.0f)}};  // beam.rs:729
  Mat<V> k(12, 12);
  for (auto& ip : ips)
    k = k.add(beam_local_stiffness_matrix_at_ip(p1, p2, young_modulus, poisson_ratio, area,
                                                out.i11_p, out.i22_p, it, shear_factor, ip[0],
                                                ip[1]));
  out.k_local = k;
  Mat<V> rm = compose_rotation_matrix_6dof(out.q, 2);
  out.k_global = rm.transpose().multiply(k).multiply(rm);
  return OK;
}

// --------------------------------------------------------------------------
// convex_hull_on_plane.rs
// --------------------------------------------------------------------------
template <typename V>
struct HullPoint {
  uint32_t n;
  V x, y;
};

// Point::partial_cmp / eq (convex_hull_on_plane.rs:81-121): order by the angle (degrees) between
// +x and the vector origin->point; NaN (point at the origin) compares false both ways.
template <typename V>
inline V hull_angle_deg(const HullPoint<V>& p) {
  // directional (0,0)->(1,0) stored as point_1 - point_2 = (-1, 0); lhs = (0,0)->(p): (-x,-y)
  V dl = std::sqrt((V(0.0) - V(1.0)) * (V(0.0) - V(1.0)) + (V(0.0) - V(0.0)) * (V(0.0) - V(0.0)));
  V pl = std::sqrt((V(0.0) - p.x) * (V(0.0) - p.x) + (V(0.0) - p.y) * (V(0.0) - p.y));
  V sp = (V(0.0) - V(1.0)) * (V(0.0) - p.x) + (V(0.0) - V(0.0)) * (V(0.0) - p.y);
  V cosv = sp / (dl * pl);
  return std::acos(cosv) * (V(180.0) / V(3.14159265358979323846264338327950288));
}
template <typename V>
inline bool hull_lt(const HullPoint<V>& a, const HullPoint<V>& b) {
  return hull_angle_deg(a) < hull_angle_deg(b);
}
template <typename V>
inline bool hull_gt(const HullPoint<V>& a, const HullPoint<V>& b) {
  return hull_angle_deg(a) > hull_angle_deg(b);
}

// convex_hull_on_plane.rs:33-60
template <typename V>
inline long hull_partition(std::vector<HullPoint<V>>& arr, size_t base, long low, long high) {
  size_t pivot = size_t(high);
  long store_index = low - 1;
  long last_index = high;
  for (;;) {
    store_index += 1;
    while (hull_lt(arr[base + store_index], arr[base + pivot])) store_index += 1;
    last_index -= 1;
    while (last_index >= 0 && hull_gt(arr[base + last_index], arr[base + pivot])) last_index -= 1;
    if (store_index >= last_index) break;
    std::swap(arr[base + store_index], arr[base + last_index]);
  }
  std::swap(arr[base + store_index], arr[base + pivot]);
  return store_index;
}
// convex_hull_on_plane.rs:22-31
template <typename V>
inline void hull_quick_sort(std::vector<HullPoint<V>>& arr, size_t base, long low, long high) {
  if (low < high) {
    long p = hull_partition(arr, base, low, high);
    hull_quick_sort(arr, base, low, p - 1);
    hull_quick_sort(arr, base, p + 1, high);
  }
}

// convex_hull_on_plane.rs:178-214
template <typename V>
inline std::vector<HullPoint<V>> convex_hull_on_plane(const std::vector<HullPoint<V>>& data) {
  std::vector<HullPoint<V>> d = data;
  V shift_x = d[0].x, min_y = d[0].y;
  size_t min_y_position = 0;
  for (size_t i = 0; i < d.size(); ++i)
    if (d[i].y < min_y) {
      shift_x = d[i].x;
      min_y = d[i].y;
      min_y_position = i;
    }
  std::swap(d[0], d[min_y_position]);
  for (auto& p : d) {
    p.x -= shift_x;
    p.y -= min_y;
  }
  hull_quick_sort(d, 1, 0, long(d.size()) - 2);
  size_t i = 0;
  while (i + 2 < d.size()) {
    V area2 = (d[i + 1].x - d[i].x) * (d[i + 2].y - d[i].y) -
              (d[i + 1].y - d[i].y) * (d[i + 2].x - d[i].x);
    if (area2 <= V(0.0))
      d.erase(d.begin() + long(i) + 1);
    else
      i += 1;
  }
  for (auto& p : d) {
    p.x += shift_x;
    p.y += min_y;
  }
  return d;
}

// --------------------------------------------------------------------------
// quadrilateral_4n_element_functions.rs
// --------------------------------------------------------------------------

// :14-86
template <typename V>
inline bool is_points_of_quadrilateral_on_the_same_line(const V p1[3], const V p2[3], const V p3[3],
                                                        const V p4[3]) {
  auto d = [](const V* a, const V* b) { return Vec3<V>{{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; };
  Vec3<V> pairs[4][2] = {{d(p2, p1), d(p4, p1)},
                         {d(p1, p2), d(p3, p2)},
                         {d(p2, p3), d(p4, p3)},
                         {d(p3, p4), d(p1, p4)}};
  for (auto& pr : pairs)
    if (v3_norm(v3_cross(pr[0], pr[1])) == V(0.0f)) return true;
  return false;
}

// :88-129 (un-normalised normal)
template <typename V>
inline bool is_points_of_quadrilateral_on_the_same_plane(const V p1[3], const V p2[3],
                                                         const V p3[3], const V p4[3], V abs_tol) {
  Vec3<V> v32{{p2[0] - p3[0], p2[1] - p3[1], p2[2] - p3[2]}};
  Vec3<V> v34{{p4[0] - p3[0], p4[1] - p3[1], p4[2] - p3[2]}};
  Vec3<V> n = v3_cross(v32, v34);
  V a = n.c[0], b = n.c[1], c = n.c[2];
  V dd = V(-1.0f) * (a * p3[0] + b * p3[1] + c * p3[2]);
  return compare_with_tolerance(a * p1[0] + b * p1[1] + c * p1[2] + dd, abs_tol) == V(0.0f);
}

// :131-171
template <typename V>
inline void find_rotation_matrix_elements_of_quadrilateral(const V p2[3], const V p3[3],
                                                           const V p4[3], V rel_tol, V abs_tol,
                                                           V q[9]) {
  Vec3<V> e34{{p4[0] - p3[0], p4[1] - p3[1], p4[2] - p3[2]}};
  Vec3<V> e32{{p2[0] - p3[0], p2[1] - p3[1], p2[2] - p3[2]}};
  Vec3<V> n = v3_cross(e34, e32);
  V len = v3_norm(n);
  Vec3<V> dir{{V(0.0f), V(0.0f), len}};
  rotation_matrix_to_align_with_vector(n, dir, rel_tol, abs_tol, q);
}

template <typename V>
inline Vec3<V> mat3_mul_vec(const V q[9], const Vec3<V>& d) {
  Vec3<V> o;
  for (int i = 0; i < 3; ++i) {
    V acc = V(0.0f);
    for (int k = 0; k < 3; ++k) acc += q[3 * i + k] * d.c[k];
    o.c[i] = acc;
  }
  return o;
}

// :173-250
template <typename V>
inline size_t convex_hull_on_four_points_on_plane(const V p1[3], const V p2[3], const V p3[3],
                                                  const V p4[3], V rel_tol, V abs_tol) {
  V q[9];
  find_rotation_matrix_elements_of_quadrilateral(p2, p3, p4, rel_tol, abs_tol, q);
  auto d = [&](const V* a) { return Vec3<V>{{a[0] - p3[0], a[1] - p3[1], a[2] - p3[2]}}; };
  Vec3<V> t1 = mat3_mul_vec(q, d(p1)), t2 = mat3_mul_vec(q, d(p2)), t4 = mat3_mul_vec(q, d(p4));
  std::vector<HullPoint<V>> pts = {{1, t1.c[0], t1.c[1]},
                                   {2, t2.c[0], t2.c[1]},
                                   {3, V(0.0f), V(0.0f)},
                                   {4, t4.c[0], t4.c[1]}};
  return convex_hull_on_plane(pts).size();
}

// :252-338 (dx_dr and dy_dr share one form, as do dx_ds and dy_ds)
template <typename V>
inline V quad_dx_dr(V x_1, V x_2, V x_3, V x_4, V r_, V s) {
  const V q = V(0.25f), z = V(0.0f);
  return derivative_x(x_1 * q, z, 0) + derivative_x(x_1 * q * s, z, 0) + derivative_x(x_1 * q, r_, 1) +
         derivative_x(x_1 * q * s, r_, 1) + derivative_x(x_2 * q, z, 0) +
         derivative_x(x_2 * q * s, z, 0) - derivative_x(x_2 * q, r_, 1) -
         derivative_x(x_2 * q * s, r_, 1) + derivative_x(x_3 * q, z, 0) -
         derivative_x(x_3 * q * s, z, 0) - derivative_x(x_3 * q, r_, 1) +
         derivative_x(x_3 * q * s, r_, 1) + derivative_x(x_4 * q, z, 0) -
         derivative_x(x_4 * q * s, z, 0) + derivative_x(x_4 * q, r_, 1) -
         derivative_x(x_4 * q * s, r_, 1);
}
template <typename V>
inline V quad_dx_ds(V x_1, V x_2, V x_3, V x_4, V r, V s) {
  const V q = V(0.25f), z = V(0.0f);
  return derivative_x(x_1 * q, z, 0) + derivative_x(x_1 * q, s, 1) + derivative_x(x_1 * q * r, z, 0) +
         derivative_x(x_1 * q * r, s, 1) + derivative_x(x_2 * q, z, 0) + derivative_x(x_2 * q, s, 1) -
         derivative_x(x_2 * q * r, z, 0) - derivative_x(x_2 * q * r, s, 1) +
         derivative_x(x_3 * q, z, 0) - derivative_x(x_3 * q, s, 1) -
         derivative_x(x_3 * q * r, z, 0) + derivative_x(x_3 * q * r, s, 1) +
         derivative_x(x_4 * q, z, 0) - derivative_x(x_4 * q, s, 1) +
         derivative_x(x_4 * q * r, z, 0) - derivative_x(x_4 * q * r, s, 1);
}

template <typename V>
struct PlateGeom {
  V q[9];
  V x1, y1, x2, y2, x4, y4;  // local in-plane coordinates relative to node 3
};

// :340-378
template <typename V>
inline void extract_transformed_directions_of_nodes(const V p1[3], const V p2[3], const V p3[3],
                                                    const V p4[3], PlateGeom<V>& g) {
  Vec3<V> t1 = mat3_mul_vec(g.q, find_2n_element_vector(p3, p1));
  Vec3<V> t2 = mat3_mul_vec(g.q, find_2n_element_vector(p3, p2));
  Vec3<V> t4 = mat3_mul_vec(g.q, find_2n_element_vector(p3, p4));
  g.x1 = t1.c[0]; g.y1 = t1.c[1];
  g.x2 = t2.c[0]; g.y2 = t2.c[1];
  g.x4 = t4.c[0]; g.y4 = t4.c[1];
}

// :380-446  J = [[dx_dr, dy_dr],[dx_ds, dy_ds]]
template <typename V>
inline void quad_jacobian_at_r_s(const PlateGeom<V>& g, V r, V s, V j[4]) {
  const V z = V(0.0f);
  j[0] = quad_dx_dr(g.x1, g.x2, z, g.x4, r, s);
  j[1] = quad_dx_dr(g.y1, g.y2, z, g.y4, r, s);  // dy_dr has the identical form (:294-315)
  j[2] = quad_dx_ds(g.x1, g.x2, z, g.x4, r, s);
  j[3] = quad_dx_ds(g.y1, g.y2, z, g.y4, r, s);     // dy_ds identical form (:317-338)
}
// :477-503 (closed-form 2x2; extended_matrix's determinant(rel_tol) assumed equivalent)
template <typename V>
inline V quad_determinant_of_jacobian(const V j[4]) {
  const int mode = variants().inverse2;
  if (mode == 1) {  // LU without pivoting: det = u11 * u22
    const V l21 = j[2] / j[0];
    return j[0] * (j[3] - l21 * j[1]);
  }
  if (mode == 2) {  // partial pivoting: a row swap flips the sign
    const bool swap = std::fabs(j[2]) > std::fabs(j[0]);
    const V a = swap ? j[2] : j[0], b = swap ? j[3] : j[1], c = swap ? j[0] : j[2], d = swap ? j[1] : j[3];
    const V det = a * (d - (c / a) * b);
    return swap ? V(-1.0f) * det : det;
  }
  return j[0] * j[3] - j[1] * j[2];
}
// :448-475 (closed-form 2x2 inverse; variants: see the top of the file)
template <typename V>
inline void quad_inverse_jacobian(const V j[4], V inv[4]) {
  const int mode = variants().inverse2;
  if (mode == 1 || mode == 2) {
    // solve J x = e_k for k = 0, 1 by elimination (mode 2 picks the larger first-column entry as pivot)
    const bool swap = mode == 2 && std::fabs(j[2]) > std::fabs(j[0]);
    const V a = swap ? j[2] : j[0], b = swap ? j[3] : j[1], c = swap ? j[0] : j[2], d = swap ? j[1] : j[3];
    const V l = c / a, u22 = d - l * b;
    for (int k = 0; k < 2; ++k) {
      V r0 = (k == 0) ? V(1.0f) : V(0.0f), r1 = (k == 1) ? V(1.0f) : V(0.0f);
      if (swap) std::swap(r0, r1);
      const V y1 = r1 - l * r0;
      const V x1 = y1 / u22;
      const V x0 = (r0 - b * x1) / a;
      inv[k] = x0;
      inv[2 + k] = x1;
    }
    return;
  }
  V det = quad_determinant_of_jacobian(j);
  inv[0] = j[3] / det;
  inv[1] = V(-1.0f) * j[1] / det;
  inv[2] = V(-1.0f) * j[2] / det;
  inv[3] = j[0] / det;
}

// :505-583
template <typename V>
inline void quad_dh_dr_dh_ds(V r, V s, V dh[8]) {
  const V q = V(0.25f), z = V(0.0f);
  auto D = [](V a, V x, int n) { return derivative_x(a, x, n); };
  dh[0] = D(q, z, 0) + D(q * s, z, 0) + D(q, r, 1) + D(q * s, r, 1);
  dh[1] = D(q, z, 0) + D(q * s, z, 0) - D(q, r, 1) - D(q * s, r, 1);
  dh[2] = D(q, z, 0) - D(q * s, z, 0) - D(q, r, 1) + D(q * s, r, 1);
  dh[3] = D(q, z, 0) - D(q * s, z, 0) + D(q, r, 1) - D(q * s, r, 1);
  dh[4] = D(q, z, 0) + D(q, s, 1) + D(q * r, z, 0) + D(q * r, s, 1);
  dh[5] = D(q, z, 0) + D(q, s, 1) - D(q * r, z, 0) - D(q * r, s, 1);
  dh[6] = D(q, z, 0) - D(q, s, 1) - D(q * r, z, 0) + D(q * r, s, 1);
  dh[7] = D(q, z, 0) - D(q, s, 1) + D(q * r, z, 0) - D(q * r, s, 1);
}

// :613-653  [dh/dx; dh/dy] = J^-1 * [dh/dr; dh/ds]  (2x4)
template <typename V>
inline Mat<V> quad_dh_dx_dh_dy(const PlateGeom<V>& g, V r, V s) {
  V j[4], inv[4], dh[8];
  quad_jacobian_at_r_s(g, r, s, j);
  quad_inverse_jacobian(j, inv);
  quad_dh_dr_dh_ds(r, s, dh);
  Mat<V> mi(2, 2, {inv[0], inv[1], inv[2], inv[3]});
  Mat<V> md(2, 4, {dh[0], dh[1], dh[2], dh[3], dh[4], dh[5], dh[6], dh[7]});
  return mi.multiply(md);
}

// --------------------------------------------------------------------------
// plate.rs
// --------------------------------------------------------------------------
template <typename V>
struct PlateOut {
  V q[9];
  Mat<V> k_local;   // 24x24
  Mat<V> k_global;  // 24x24
};

// plate.rs:58-158
template <typename V>
inline int check_plate_properties(V young_modulus, V poisson_ratio, V thickness, V shear_factor,
                                  const V p1[3], const V p2[3], const V p3[3], const V p4[3],
                                  V rel_tol, V abs_tol) {
  if (young_modulus <= V(0.0f)) return E_YOUNG;
  if (poisson_ratio <= V(0.0f)) return E_POISSON;
  if (thickness <= V(0.0f)) return E_THICKNESS;
  if (shear_factor <= V(0.0f)) return E_SHEAR_FACTOR;
  if (is_points_of_quadrilateral_on_the_same_line(p1, p2, p3, p4)) return E_ON_LINE;
  if (!is_points_of_quadrilateral_on_the_same_plane(p1, p2, p3, p4, abs_tol)) return E_NOT_ON_PLANE;
  if (convex_hull_on_four_points_on_plane(p1, p2, p3, p4, rel_tol, abs_tol) != 4)
    return E_NOT_CONVEX;
  return OK;
}

// Raw row-major fills (shared with fem_oracle_fast.hpp); the Mat wrappers below are what the
// faithful path multiplies.
// plate.rs:160-274
template <typename V>
inline void plate_b_mem_raw(const V d[8] /* dh/dx[4], dh/dy[4] */, V* b /* 3x24, zeroed */) {
  for (int n = 0; n < 4; ++n) {
    b[0 * 24 + 6 * n + 0] = d[n];
    b[1 * 24 + 6 * n + 1] = d[4 + n];
    b[2 * 24 + 6 * n + 0] = d[4 + n];
    b[2 * 24 + 6 * n + 1] = d[n];
  }
}
// plate.rs:276-390
template <typename V>
inline void plate_b_bend_raw(const V d[8], V* b /* 3x24, zeroed */) {
  for (int n = 0; n < 4; ++n) {
    b[0 * 24 + 6 * n + 4] = V(-1.0f) * d[n];
    b[1 * 24 + 6 * n + 3] = d[4 + n];
    b[2 * 24 + 6 * n + 3] = d[n];
    b[2 * 24 + 6 * n + 4] = V(-1.0f) * d[4 + n];
  }
}
// plate.rs:392-511
template <typename V>
inline void plate_b_shear_raw(const PlateGeom<V>& g, V r, V s, V det, V* b /* 2x24, zeroed */) {
  V x_1 = g.x1, y_1 = g.y1, x_2 = g.x2, y_2 = g.y2, x_3 = V(0.0f), y_3 = V(0.0f), x_4 = g.x4,
    y_4 = g.y4;
  V a_x = x_1 - x_2 - x_3 + x_4;
  V b_x = x_1 - x_2 + x_3 - x_4;
  V c_x = x_1 + x_2 - x_3 - x_4;
  V a_y = y_1 - y_2 - y_3 + y_4;
  V b_y = y_1 - y_2 + y_3 - y_4;
  V c_y = y_1 + y_2 - y_3 - y_4;
  auto sq = [](V x) { return x * x; };
  V grz = std::sqrt(sq(c_x + r * b_x) + sq(c_y + r * b_y)) / (V(8.0f) * det);
  V gsz = std::sqrt(sq(a_x + s * b_x) + sq(a_y + s * b_y)) / (V(8.0f) * det);
  const V one = V(1.0f), two = V(2.0f), four = V(4.0f), m1 = V(-1.0f);
  V* r0 = b;
  V* r1 = b + 24;
  // row 0: gamma_rz
  r0[2] = (one + s) / two * grz;
  r0[3] = (one + s) * m1 * (y_1 - y_2) / four * grz;
  r0[4] = (one + s) * (x_1 - x_2) / four * grz;
  r0[8] = m1 * (one + s) / two * grz;
  r0[9] = (one + s) * m1 * (y_1 - y_2) / four * grz;
  r0[10] = (one + s) * (x_1 - x_2) / four * grz;
  r0[14] = m1 * (one - s) / two * grz;
  r0[15] = (one - s) * m1 * (y_4 - y_3) / four * grz;
  r0[16] = (one - s) * (x_4 - x_3) / four * grz;
  r0[20] = (one - s) / two * grz;
  r0[21] = (one - s) * m1 * (y_4 - y_3) / four * grz;
  r0[22] = (one - s) * (x_4 - x_3) / four * grz;
  // row 1: gamma_sz
  r1[2] = (one + r) / two * gsz;
  r1[3] = (one + r) * m1 * (y_1 - y_4) / four * gsz;
  r1[4] = (one + r) * (x_1 - x_4) / four * gsz;
  r1[8] = (one - r) / two * gsz;
  r1[9] = (one - r) * m1 * (y_2 - y_3) / four * gsz;
  r1[10] = (one - r) * (x_2 - x_3) / four * gsz;
  r1[14] = m1 * (one - r) / two * gsz;
  r1[15] = (one - r) * m1 * (y_2 - y_3) / four * gsz;
  r1[16] = (one - r) * (x_2 - x_3) / four * gsz;
  r1[20] = m1 * (one + r) / two * gsz;
  r1[21] = (one + r) * m1 * (y_1 - y_4) / four * gsz;
  r1[22] = (one + r) * (x_1 - x_4) / four * gsz;
}

// dh/dx, dh/dy as a flat [8] (row 0 = d/dx, row 1 = d/dy) and det J at (r, s)
template <typename V>
inline void quad_dh_dx_dh_dy_raw(const PlateGeom<V>& g, V r, V s, V d[8], V& det) {
  Mat<V> m = quad_dh_dx_dh_dy(g, r, s);
  for (int i = 0; i < 8; ++i) d[i] = m.a[i];
  V j[4];
  quad_jacobian_at_r_s(g, r, s, j);
  det = quad_determinant_of_jacobian(j);
}

template <typename V>
inline Mat<V> plate_b_mem(const PlateGeom<V>& g, V r, V s) {
  V d[8], det;
  quad_dh_dx_dh_dy_raw(g, r, s, d, det);
  Mat<V> b(3, 24);
  plate_b_mem_raw(d, b.a.data());
  return b;
}
template <typename V>
inline Mat<V> plate_b_bend(const PlateGeom<V>& g, V r, V s) {
  V d[8], det;
  quad_dh_dx_dh_dy_raw(g, r, s, d, det);
  Mat<V> b(3, 24);
  plate_b_bend_raw(d, b.a.data());
  return b;
}
template <typename V>
inline Mat<V> plate_b_shear(const PlateGeom<V>& g, V r, V s) {
  V j[4];
  quad_jacobian_at_r_s(g, r, s, j);
  Mat<V> b(2, 24);
  plate_b_shear_raw(g, r, s, quad_determinant_of_jacobian(j), b.a.data());
  return b;
}

// plate.rs:513-672
template <typename V>
inline Mat<V> plate_local_stiffness_matrix_at_ip(const PlateGeom<V>& g, V young_modulus,
                                                 V poisson_ratio, V thickness, V shear_factor, V r,
                                                 V s, V alpha) {
  const V one = V(1.0f), two = V(2.0f), zero = V(0.0f);
  V j[4];
  quad_jacobian_at_r_s(g, r, s, j);
  V det = quad_determinant_of_jacobian(j);

  V c_multiplier_mem = young_modulus * thickness / (one - poisson_ratio * poisson_ratio);
  Mat<V> c_mem = Mat<V>(3, 3, {one, poisson_ratio, zero, poisson_ratio, one, zero, zero, zero,
                               (one - poisson_ratio) / two})
                     .multiply_by_scalar(c_multiplier_mem);
  Mat<V> b_mem = plate_b_mem(g, r, s);
  Mat<V> k_mem = b_mem.transpose().multiply(c_mem).multiply(b_mem).multiply_by_scalar(det * alpha);

  V c_multiplier_bend = young_modulus * (thickness * thickness * thickness) /
                        (V(12.0f) * (one - poisson_ratio * poisson_ratio));
  Mat<V> c_bend = Mat<V>(3, 3, {one, poisson_ratio, zero, poisson_ratio, one, zero, zero, zero,
                                (one - poisson_ratio) / two})
                      .multiply_by_scalar(c_multiplier_bend);
  Mat<V> b_bend = plate_b_bend(g, r, s);
  Mat<V> k_bend =
      b_bend.transpose().multiply(c_bend).multiply(b_bend).multiply_by_scalar(det * alpha);

  V c_multiplier_shear =
      young_modulus * thickness * shear_factor / (two * (one + poisson_ratio));
  Mat<V> c_shear = Mat<V>(2, 2, {one, zero, zero, one}).multiply_by_scalar(c_multiplier_shear * alpha);
  Mat<V> b_shear = plate_b_shear(g, r, s);
  Mat<V> k_shear =
      b_shear.transpose().multiply(c_shear).multiply(b_shear).multiply_by_scalar(det * alpha);

  return k_mem.add(k_bend).add(k_shear);
}

// Plate::create (plate.rs:1021-1105) + extract_local_stiffness_matrix (:1124-1143, :674-724)
// + (R^T k) R of add_plate (methods_for_plate_data_handle.rs:104-112)
template <typename V>
inline int plate_element(const V p1[3], const V p2[3], const V p3[3], const V p4[3],
                         V young_modulus, V poisson_ratio, V thickness, V shear_factor, V rel_tol,
                         V abs_tol, PlateOut<V>& out) {
  int err = check_plate_properties(young_modulus, poisson_ratio, thickness, shear_factor, p1, p2,
                                   p3, p4, rel_tol, abs_tol);
  if (err) return err;
  PlateGeom<V> g;
  find_rotation_matrix_elements_of_quadrilateral(p2, p3, p4, rel_tol, abs_tol, g.q);
  for (int i = 0; i < 9; ++i) out.q[i] = g.q[i];
  extract_transformed_directions_of_nodes(p1, p2, p3, p4, g);
  // plate.rs:1066-1091 — the abscissa is sqrt of the *f32* value of 1/3
  V gp = std::sqrt(V(1.0f / 3.0f));
  const V one = V(1.0f), m1 = V(-1.0f);
  V ips[4][4] = {{gp * one, gp * one, one, one},
                 {gp * m1, gp * one, one, one},
                 {gp * m1, gp * m1, one, one},
                 {gp * one, gp * m1, one, one}};
  Mat<V> k(24, 24);
  for (auto& ip : ips)
    k = k.add(plate_local_stiffness_matrix_at_ip(g, young_modulus, poisson_ratio, thickness,
                                                 shear_factor, ip[0], ip[1], ip[2] * ip[3]));
  for (int i = 0; i < 4; ++i) k.at(6 * i + 5, 6 * i + 5) += V(1.0f);  // KROT6, plate.rs:25,716-721
  out.k_local = k;
  Mat<V> rm = compose_rotation_matrix_6dof(out.q, 4);
  out.k_global = rm.transpose().multiply(k).multiply(rm);
  return OK;
}

// --------------------------------------------------------------------------
// Uniformly distributed loads -> nodal loads
// --------------------------------------------------------------------------

// Beam::convert_uniformly_distributed_line_load_to_nodal_loads (beam.rs:775-797) over the beam's
// integration points [(r = 0, alpha = 2)] (beam.rs:729): f += (h(r) * q) * (det J(r) * alpha)
template <typename V>
inline void beam_line_load_nodal(const V p1[3], const V p2[3], V q, V f[2]) {
  f[0] = f[1] = V(0.0f);
  const V r = V(0.0f), alpha = V(2.0f);
  const V det = bar_determinant_of_jacobian_at_r(p1, p2, r);
  f[0] = f[0] + (h1_r(r) * q) * (det * alpha);
  f[1] = f[1] + (h2_r(r) * q) * (det * alpha);
}

// Plate::convert_uniformly_distributed_surface_load_to_nodal_loads (plate.rs:1145-1185) over the four
// Gauss points of plate.rs:1066-1091: f += (h_a(r, s) * q) * (det J(r, s) * alpha_r * alpha_s),
// h_a = quadrilateral_4n_element_functions.rs:585-611, det J = :477-503
template <typename V>
inline void plate_surface_load_nodal(const V p1[3], const V p2[3], const V p3[3], const V p4[3], V q,
                                     V rel_tol, V abs_tol, V f[4]) {
  PlateGeom<V> g;
  find_rotation_matrix_elements_of_quadrilateral(p2, p3, p4, rel_tol, abs_tol, g.q);
  extract_transformed_directions_of_nodes(p1, p2, p3, p4, g);
  V gp = std::sqrt(V(1.0f / 3.0f));
  const V one = V(1.0f), m1 = V(-1.0f), quarter = V(0.25f);
  V ips[4][4] = {{gp * one, gp * one, one, one},
                 {gp * m1, gp * one, one, one},
                 {gp * m1, gp * m1, one, one},
                 {gp * one, gp * m1, one, one}};
  for (int a = 0; a < 4; ++a) f[a] = V(0.0f);
  for (auto& ip : ips) {
    const V r = ip[0], s = ip[1];
    V j[4];
    quad_jacobian_at_r_s(g, r, s, j);
    const V scale = quad_determinant_of_jacobian(j) * ip[2] * ip[3];
    const V h[4] = {quarter * (one + r) * (one + s), quarter * (one - r) * (one + s),
                    quarter * (one - r) * (one - s), quarter * (one + r) * (one - s)};
    for (int a = 0; a < 4; ++a) f[a] = f[a] + (h[a] * q) * scale;
  }
}

// --------------------------------------------------------------------------
// Element result recovery: forces / moments from the nodal displacements.
// `u` below is the element's slice of the global displacement vector, node by node, as the
// reference gathers it (displacements[node_index * NODE_DOF + i]).
// --------------------------------------------------------------------------

// Truss::extract_element_analysis_result (truss.rs:281-333): u = [u1 (3), u2 (3)] -> ForceR
template <typename V>
inline V truss_element_result(const V p1[3], const V p2[3], V young_modulus, V area, bool has_area_2,
                              V area_2, V rel_tol, V abs_tol, const V u[6]) {
  V q[9];
  truss_find_rotation_matrix_elements(p1, p2, rel_tol, abs_tol, q);  // stored at creation, truss.rs:238
  Mat<V> ug(6, 1);
  for (int i = 0; i < 6; ++i) ug.at(i, 0) = u[i];
  Mat<V> ul = compose_rotation_matrix_3dof(q, 2).multiply(ug);
  const V ips[1][2] = {{V(0.0f), V(2.0f)}};  // truss.rs:244
  Mat<V> b(1, 6);
  V area_sum = V(0.0f);
  for (auto& ip : ips) {
    b = b.add(truss_strain_displacement_matrix_at_r(p1, p2, ip[0]));
    area_sum += truss_area_at_r(area, has_area_2, area_2, ip[0]);
  }
  Mat<V> strain = b.multiply(ul);
  Mat<V> force = strain.multiply_by_scalar(young_modulus * area_sum / V(1.0f));  // / n_ip
  return force.at(0, 0);
}

// Beam::extract_element_analysis_result (beam.rs:803-993): u = [node 1 (6), node 2 (6)] ->
// ForceR, ForceS, ForceT, MomentR, MomentS (node 1, average, node 2), MomentT (node 1, average, node 2)
template <typename V>
inline void beam_element_result(const V p1[3], const V p2[3], V young_modulus, V poisson_ratio, V area,
                                V i11, V i22, V i12, V it, V shear_factor, const V axis1[3], V rel_tol,
                                V abs_tol, const V u[12], V out[10]) {
  V i11_p, i22_p, angle, q[9];
  find_principal_moments_of_inertia(i11, i22, i12, rel_tol, i11_p, i22_p, angle);  // beam.rs:711
  beam_find_rotation_matrix_elements(p1, p2, axis1, angle, rel_tol, abs_tol, q);
  Mat<V> ug(12, 1);
  for (int i = 0; i < 12; ++i) ug.at(i, 0) = u[i];
  Mat<V> ul = compose_rotation_matrix_6dof(q, 2).multiply(ug);
  const V ips[1][2] = {{V(0.0f), V(2.0f)}};  // beam.rs:729
  const V n_ip = V(1.0f);
  V shear_modulus = young_modulus / (V(2.0f) * (V(1.0f) + poisson_ratio));
  const V cs[6] = {young_modulus * area / n_ip,
                   shear_modulus * area * shear_factor / n_ip,
                   shear_modulus * area * shear_factor / n_ip,
                   shear_modulus * it / n_ip,
                   young_modulus * i22_p / n_ip,
                   young_modulus * i11_p / n_ip};
  V f[6];
  for (int w = 0; w < 6; ++w) {
    Mat<V> b(1, 12);
    for (auto& ip : ips) b = b.add(beam_b_row(p1, p2, ip[0], w));
    f[w] = b.multiply(ul).multiply_by_scalar(cs[w]).at(0, 0);
  }
  V len = v3_norm(find_2n_element_vector(p1, p2));
  const V force_s = f[1], force_t = f[2], moment_s_average = f[4], moment_t_average = f[5];
  out[0] = f[0];
  out[1] = force_s;
  out[2] = force_t;
  out[3] = f[3];
  out[4] = moment_s_average + len * force_t / V(2.0f);
  out[5] = moment_s_average;
  out[6] = moment_s_average - len * force_t / V(2.0f);
  out[7] = moment_t_average + len * force_s / V(2.0f);
  out[8] = moment_t_average;
  out[9] = moment_t_average - len * force_s / V(2.0f);
}

// Plate::extract_element_analysis_result (plate.rs:1196-1409): u = 4 nodes x 6 ->
// MembraneForceR, MembraneForceS, MembraneForceRS, BendingMomentR, BendingMomentS, BendingMomentRS,
// ShearForceRT, ShearForceST. The strain-displacement matrices are summed over the four NODES
// (r, s = +-1), not the Gauss points.
template <typename V>
inline void plate_element_result(const V p1[3], const V p2[3], const V p3[3], const V p4[3],
                                 V young_modulus, V poisson_ratio, V thickness, V shear_factor, V rel_tol,
                                 V abs_tol, const V u[24], V out[8]) {
  PlateGeom<V> g;
  find_rotation_matrix_elements_of_quadrilateral(p2, p3, p4, rel_tol, abs_tol, g.q);
  extract_transformed_directions_of_nodes(p1, p2, p3, p4, g);
  Mat<V> ug(24, 1);
  for (int i = 0; i < 24; ++i) ug.at(i, 0) = u[i];
  Mat<V> ul = compose_rotation_matrix_6dof(g.q, 4).multiply(ug);
  const V one = V(1.0f), two = V(2.0f), zero = V(0.0f), m1 = V(-1.0f);
  const V rs[4][2] = {{one, one}, {m1, one}, {m1, m1}, {one, m1}};
  const V n_nodes = V(4.0f);

  V c_multiplier_mem = young_modulus / (one - poisson_ratio * poisson_ratio);
  Mat<V> c_mem = Mat<V>(3, 3, {one, poisson_ratio, zero, poisson_ratio, one, zero, zero, zero,
                               (one - poisson_ratio) / two})
                     .multiply_by_scalar(c_multiplier_mem);
  Mat<V> b_mem(3, 24);
  for (auto& n : rs) b_mem = b_mem.add(plate_b_mem(g, n[0], n[1]));
  Mat<V> f_mem = c_mem.multiply(b_mem.multiply(ul)).multiply_by_scalar(thickness / n_nodes);

  V c_multiplier_bend = young_modulus * thickness / (two * (one - poisson_ratio * poisson_ratio));
  Mat<V> c_bend = Mat<V>(3, 3, {one, poisson_ratio, zero, poisson_ratio, one, zero, zero, zero,
                                (one - poisson_ratio) / two})
                      .multiply_by_scalar(c_multiplier_bend);
  Mat<V> b_bend(3, 24);
  for (auto& n : rs) b_bend = b_bend.add(plate_b_bend(g, n[0], n[1]));
  Mat<V> f_bend = c_bend.multiply(b_bend.multiply(ul)).multiply_by_scalar(thickness * thickness / V(24.0f));

  V c_multiplier_shear = young_modulus / (two * (one + poisson_ratio));
  Mat<V> c_shear = Mat<V>(2, 2, {one, zero, zero, one}).multiply_by_scalar(c_multiplier_shear);
  Mat<V> b_shear(2, 24);
  for (auto& n : rs) b_shear = b_shear.add(plate_b_shear(g, n[0], n[1]));
  Mat<V> f_shear = c_shear.multiply(b_shear.multiply(ul)).multiply_by_scalar(thickness * shear_factor / n_nodes);

  out[0] = f_mem.at(0, 0);
  out[1] = f_mem.at(1, 0);
  out[2] = f_mem.at(2, 0);
  out[3] = f_bend.at(1, 0);  // BendingMomentR takes row 1, BendingMomentS row 0 (plate.rs:1384-1391)
  out[4] = f_bend.at(0, 0);
  out[5] = f_bend.at(2, 0);
  out[6] = f_shear.at(0, 0);
  out[7] = f_shear.at(1, 0);
}

// --------------------------------------------------------------------------
// Global matrix: position-keyed map + the zero-skip block scatter
// (methods_for_truss_data_handle.rs:93-123, methods_for_beam_data_handle.rs:107-137,
//  methods_for_plate_data_handle.rs:134-212; add_value = entry(pos).or_insert(0) += v)
// --------------------------------------------------------------------------
template <typename V>
struct GlobalK {
  std::unordered_map<uint64_t, V> e;
  static uint64_t key(uint64_t row, uint64_t col) { return (row << 32) | col; }
  void add_value(uint64_t row, uint64_t col, V v) { e[key(row, col)] += v; }
};

constexpr int NODE_DOF = 6;  // structs/node.rs:8

// n_nodes element nodes with `dof` (3 truss / 6 beam, plate) local dofs each;
// idx[] are the 0-based node indices in insertion order (methods_for_node_data_handle.rs:66-78)
template <typename V>
inline void scatter_blocks(GlobalK<V>& K, const Mat<V>& kg, const uint32_t* idx, int n_nodes,
                           int dof) {
  for (int a = 0; a < n_nodes; ++a)
    for (int b = 0; b < n_nodes; ++b)
      for (int i = 0; i < dof; ++i)
        for (int j = 0; j < dof; ++j) {
          V v = kg.at(a * dof + i, b * dof + j);
          if (v != V(0.0f))
            K.add_value(uint64_t(idx[a]) * NODE_DOF + i, uint64_t(idx[b]) * NODE_DOF + j, v);
        }
}

}  // namespace oracle
