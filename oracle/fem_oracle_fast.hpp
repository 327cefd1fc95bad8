// ORACLE — TEST INFRASTRUCTURE ONLY (see fem_oracle.hpp).
//
// "fast" CPU baseline: the same algorithm as the faithful restatement, but the way a careful CPU
// implementer would write it — fixed-size stack matrices, block-diagonal R applied block-wise,
// OpenMP across all host cores, and a prebuilt structural block-CSR (3x3 per truss-only node
// pair, 6x6 per beam/plate node pair) filled by an owner-computes gather, so it is deterministic.
// Used for bench.py's cpu_baseline / --impl reference legs and cross-checked against the faithful
// path in tests/test_oracle_fast.py. It shares the scalar geometry helpers of fem_oracle.hpp so the
// element numbers are the same to rounding.
#pragma once
#include <omp.h>

#include <chrono>
#include <numeric>

#include "fem_oracle.hpp"

namespace oracle {
namespace fast {

struct Mesh {
  int64_t n_nodes;
  const double *x, *y, *z;
  int64_t n_truss;
  const uint32_t *t_n1, *t_n2;
  const double *t_E, *t_A, *t_A2;
  int64_t n_beam;
  const uint32_t *b_n1, *b_n2;
  const double* b_props;  // [8][n_beam]
  const double* b_axis;   // [3][n_beam]
  int64_t n_plate;
  const uint32_t* p_n;    // [4][n_plate]
  const double* p_props;  // [4][n_plate]
  double rel_tol, abs_tol;
};

inline void xyz(const Mesh& m, uint32_t i, double p[3]) {
  p[0] = m.x[i];
  p[1] = m.y[i];
  p[2] = m.z[i];
}

// out(3x3) = Q^T * X(3x3) * Q
inline void sandwich3(const double q[9], const double* X, int ldx, double* out, int ldo) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += q[3 * k + i] * X[k * ldx + j];
      t[3 * i + j] = acc;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += t[3 * i + k] * q[3 * k + j];
      out[i * ldo + j] = acc;
    }
}

// kg (n x n, n = 3*nb) = diag(Q..)^T k diag(Q..)
inline void rotate_blocks(const double q[9], const double* k, double* kg, int nb) {
  int n = 3 * nb;
  for (int a = 0; a < nb; ++a)
    for (int b = 0; b < nb; ++b) sandwich3(q, k + (3 * a) * n + 3 * b, n, kg + (3 * a) * n + 3 * b, n);
}

inline int truss_kg(const Mesh& m, int64_t e, double kg[36]) {
  double p1[3], p2[3];
  xyz(m, m.t_n1[e], p1);
  xyz(m, m.t_n2[e], p2);
  bool has2 = m.t_A2 && !std::isnan(m.t_A2[e]);
  int err = check_truss_properties(m.t_E[e], m.t_A[e], has2, has2 ? m.t_A2[e] : 0.0);
  if (err) return err;
  double q[9];
  truss_find_rotation_matrix_elements(p1, p2, m.rel_tol, m.abs_tol, q);
  double r = 0.0, alpha = 2.0;
  double inv_j = bar_inverse_jacobian_at_r(p1, p2, r);
  double det = bar_determinant_of_jacobian_at_r(p1, p2, r);
  double b[6] = {dh1_dr(r) * inv_j, 0, 0, dh2_dr(r) * inv_j, 0, 0};
  double c = truss_area_at_r(m.t_A[e], has2, has2 ? m.t_A2[e] : 0.0, r) * m.t_E[e];
  double sc = c * det * alpha;
  double k[36];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) k[6 * i + j] = (b[i] * b[j]) * sc;
  rotate_blocks(q, k, kg, 2);
  return 0;
}

inline int beam_kg(const Mesh& m, int64_t e, double kg[144]) {
  int64_t n = m.n_beam;
  double p1[3], p2[3];
  xyz(m, m.b_n1[e], p1);
  xyz(m, m.b_n2[e], p2);
  const double* P = m.b_props;
  double E = P[e], nu = P[n + e], A = P[2 * n + e], I11 = P[3 * n + e], I22 = P[4 * n + e],
         I12 = P[5 * n + e], It = P[6 * n + e], ks = P[7 * n + e];
  double ax[3] = {m.b_axis[e], m.b_axis[n + e], m.b_axis[2 * n + e]};
  int err = check_beam_properties(E, nu, A, I11, I22, It, ks, p1, p2, ax);
  if (err) return err;
  double i11p, i22p, angle, q[9];
  find_principal_moments_of_inertia(I11, I22, I12, m.rel_tol, i11p, i22p, angle);
  beam_find_rotation_matrix_elements(p1, p2, ax, angle, m.rel_tol, m.abs_tol, q);
  double r = 0.0, alpha = 2.0;
  double inv_j = bar_inverse_jacobian_at_r(p1, p2, r);
  double det = bar_determinant_of_jacobian_at_r(p1, p2, r);
  double G = E / (2.0 * (1.0 + nu));
  double cs[6] = {A * E, G * A * ks, G * A * ks, G * It, E * i22p, E * i11p};
  double k[144];
  for (int i = 0; i < 144; ++i) k[i] = 0.0;
  for (int w = 0; w < 6; ++w) {
    double b[12] = {0};
    b[w] = dh1_dr(r) * inv_j;
    b[w + 6] = dh2_dr(r) * inv_j;
    if (w == 1 || w == 2) {
      int col = (w == 1) ? 5 : 4;
      b[col] = 0.0 - h1_r(r);
      b[col + 6] = 0.0 - h2_r(r);
    }
    double sc = cs[w] * det * alpha;
    for (int i = 0; i < 12; ++i) {
      if (b[i] == 0.0) continue;
      for (int j = 0; j < 12; ++j) k[12 * i + j] += (b[i] * b[j]) * sc;
    }
  }
  rotate_blocks(q, k, kg, 4);
  return 0;
}

inline int plate_kg(const Mesh& m, int64_t e, double kg[576], bool validate) {
  int64_t n = m.n_plate;
  double p1[3], p2[3], p3[3], p4[3];
  xyz(m, m.p_n[e], p1);
  xyz(m, m.p_n[n + e], p2);
  xyz(m, m.p_n[2 * n + e], p3);
  xyz(m, m.p_n[3 * n + e], p4);
  const double* P = m.p_props;
  double E = P[e], nu = P[n + e], t = P[2 * n + e], ks = P[3 * n + e];
  if (validate) {
    int err = check_plate_properties(E, nu, t, ks, p1, p2, p3, p4, m.rel_tol, m.abs_tol);
    if (err) return err;
  }
  PlateGeom<double> g;
  find_rotation_matrix_elements_of_quadrilateral(p2, p3, p4, m.rel_tol, m.abs_tol, g.q);
  extract_transformed_directions_of_nodes(p1, p2, p3, p4, g);
  double gp = std::sqrt(double(1.0f / 3.0f));
  const double ips[4][2] = {{gp, gp}, {-gp, gp}, {-gp, -gp}, {gp, -gp}};
  double cm = E * t / (1.0 - nu * nu), cb = E * (t * t * t) / (12.0 * (1.0 - nu * nu));
  double csh = E * t * ks / (2.0 * (1.0 + nu));
  double C3[9] = {1.0, nu, 0, nu, 1.0, 0, 0, 0, (1.0 - nu) / 2.0};
  double k[576];
  for (int i = 0; i < 576; ++i) k[i] = 0.0;
  for (auto& ip : ips) {
    double d[8], det;
    quad_dh_dx_dh_dy_raw(g, ip[0], ip[1], d, det);
    double B[8][24];
    for (auto& row : B)
      for (double& v : row) v = 0.0;
    plate_b_mem_raw(d, &B[0][0]);
    plate_b_bend_raw(d, &B[3][0]);
    plate_b_shear_raw(g, ip[0], ip[1], det, &B[6][0]);
    // D = blockdiag(cm*C3, cb*C3, csh*I2); k += B^T D B * det
    double DB[8][24];
    for (int j = 0; j < 24; ++j) {
      for (int i = 0; i < 3; ++i) {
        double am = 0, ab = 0;
        for (int l = 0; l < 3; ++l) {
          am += (C3[3 * i + l] * cm) * B[l][j];
          ab += (C3[3 * i + l] * cb) * B[3 + l][j];
        }
        DB[i][j] = am;
        DB[3 + i][j] = ab;
      }
      DB[6][j] = csh * B[6][j];
      DB[7][j] = csh * B[7][j];
    }
    for (int i = 0; i < 24; ++i)
      for (int j = 0; j < 24; ++j) {
        double acc = 0;
        for (int l = 0; l < 8; ++l) acc += B[l][i] * DB[l][j];
        k[24 * i + j] += acc * det;
      }
  }
  for (int i = 0; i < 4; ++i) k[24 * (6 * i + 5) + 6 * i + 5] += 1.0;
  rotate_blocks(g.q, k, kg, 8);
  return 0;
}

// Structural block pattern + owner-computes gather lists (built once, untimed).
struct Pattern {
  std::vector<int64_t> node_blk_ptr;   // [n_nodes+1] -> blocks
  std::vector<uint32_t> blk_col;       // neighbour node
  std::vector<uint8_t> blk_full;       // 1 = 6x6, 0 = 3x3 (truss only)
  std::vector<int64_t> blk_val;        // value offset of the block's first row segment
  std::vector<int64_t> blk_cptr;       // [n_blk+1] -> contributions
  std::vector<uint64_t> contrib;       // (family<<60) | (pair<<56) | elem
  std::vector<int64_t> node_base;      // value offset of the node's first row
  std::vector<int32_t> len03, len35;   // row lengths for dofs 0-2 / 3-5
  std::vector<int32_t> blk_off03, blk_off35;
  int64_t nnz = 0;
};

inline Pattern build_pattern(const Mesh& m) {
  struct C {
    uint64_t key;
    uint64_t payload;
  };
  std::vector<C> cs;
  cs.reserve(size_t(m.n_plate) * 16 + size_t(m.n_beam + m.n_truss) * 4);
  auto push = [&](uint64_t fam, int64_t e, const uint32_t* nd, int nn) {
    for (int a = 0; a < nn; ++a)
      for (int b = 0; b < nn; ++b)
        cs.push_back({(uint64_t(nd[a]) << 32) | nd[b],
                      (fam << 60) | (uint64_t(a * nn + b) << 56) | uint64_t(e)});
  };
  for (int64_t e = 0; e < m.n_plate; ++e) {
    uint32_t nd[4] = {m.p_n[e], m.p_n[m.n_plate + e], m.p_n[2 * m.n_plate + e],
                      m.p_n[3 * m.n_plate + e]};
    push(2, e, nd, 4);
  }
  for (int64_t e = 0; e < m.n_beam; ++e) {
    uint32_t nd[2] = {m.b_n1[e], m.b_n2[e]};
    push(1, e, nd, 2);
  }
  for (int64_t e = 0; e < m.n_truss; ++e) {
    uint32_t nd[2] = {m.t_n1[e], m.t_n2[e]};
    push(0, e, nd, 2);
  }
  std::stable_sort(cs.begin(), cs.end(), [](const C& a, const C& b) { return a.key < b.key; });
  Pattern p;
  p.node_blk_ptr.assign(m.n_nodes + 1, 0);
  p.contrib.resize(cs.size());
  for (size_t i = 0; i < cs.size(); ++i) {
    p.contrib[i] = cs[i].payload;
    if (i == 0 || cs[i].key != cs[i - 1].key) {
      p.blk_col.push_back(uint32_t(cs[i].key & 0xffffffffu));
      p.blk_full.push_back(0);
      p.blk_cptr.push_back(int64_t(i));
      p.node_blk_ptr[(cs[i].key >> 32) + 1]++;
    }
    if ((cs[i].payload >> 60) != 0) p.blk_full.back() = 1;
  }
  p.blk_cptr.push_back(int64_t(cs.size()));
  for (int64_t a = 0; a < m.n_nodes; ++a) p.node_blk_ptr[a + 1] += p.node_blk_ptr[a];
  size_t nb = p.blk_col.size();
  p.blk_off03.resize(nb);
  p.blk_off35.resize(nb);
  p.node_base.assign(m.n_nodes + 1, 0);
  p.len03.assign(m.n_nodes, 0);
  p.len35.assign(m.n_nodes, 0);
  int64_t off = 0;
  for (int64_t a = 0; a < m.n_nodes; ++a) {
    int l03 = 0, l35 = 0;
    for (int64_t b = p.node_blk_ptr[a]; b < p.node_blk_ptr[a + 1]; ++b) {
      p.blk_off03[b] = l03;
      p.blk_off35[b] = p.blk_full[b] ? l35 : -1;
      l03 += p.blk_full[b] ? 6 : 3;
      l35 += p.blk_full[b] ? 6 : 0;
    }
    p.len03[a] = l03;
    p.len35[a] = l35;
    p.node_base[a] = off;
    off += 3 * int64_t(l03) + 3 * int64_t(l35);
  }
  p.node_base[m.n_nodes] = off;
  p.nnz = off;
  return p;
}

// returns best-of-`repeats` seconds of the numeric part
inline double assemble(const Mesh& m, int n_threads, int repeats, int64_t* nnz_out,
                       double* checksum_out, int64_t* coo_rows = nullptr,
                       int64_t* coo_cols = nullptr, double* coo_vals = nullptr) {
  if (n_threads > 0) omp_set_num_threads(n_threads);
  Pattern p = build_pattern(m);
  std::vector<double> values(size_t(p.nnz));
  std::vector<double> ke_t(size_t(m.n_truss) * 36), ke_b(size_t(m.n_beam) * 144),
      ke_p(size_t(m.n_plate) * 576);
  double best = 1e300;
  for (int rep = 0; rep < std::max(1, repeats); ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    int bad = 0;
#pragma omp parallel
    {
#pragma omp for schedule(static) nowait
      for (int64_t e = 0; e < m.n_plate; ++e)
        if (plate_kg(m, e, &ke_p[size_t(e) * 576], true)) bad = 1;
#pragma omp for schedule(static) nowait
      for (int64_t e = 0; e < m.n_beam; ++e)
        if (beam_kg(m, e, &ke_b[size_t(e) * 144])) bad = 1;
#pragma omp for schedule(static)
      for (int64_t e = 0; e < m.n_truss; ++e)
        if (truss_kg(m, e, &ke_t[size_t(e) * 36])) bad = 1;
#pragma omp for schedule(static)
      for (int64_t a = 0; a < m.n_nodes; ++a) {
        for (int64_t b = p.node_blk_ptr[a]; b < p.node_blk_ptr[a + 1]; ++b) {
          double acc[36];
          for (double& v : acc) v = 0.0;
          for (int64_t c = p.blk_cptr[b]; c < p.blk_cptr[b + 1]; ++c) {
            uint64_t pl = p.contrib[c];
            int fam = int(pl >> 60), pair = int((pl >> 56) & 0xf);
            int64_t e = int64_t(pl & 0x00ffffffffffffffull);
            if (fam == 0) {
              int la = pair / 2, lb = pair % 2;
              const double* k = &ke_t[size_t(e) * 36];
              for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) acc[6 * i + j] += k[(3 * la + i) * 6 + 3 * lb + j];
            } else if (fam == 1) {
              int la = pair / 2, lb = pair % 2;
              const double* k = &ke_b[size_t(e) * 144];
              for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) acc[6 * i + j] += k[(6 * la + i) * 12 + 6 * lb + j];
            } else {
              int la = pair / 4, lb = pair % 4;
              const double* k = &ke_p[size_t(e) * 576];
              for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) acc[6 * i + j] += k[(6 * la + i) * 24 + 6 * lb + j];
            }
          }
          int w = p.blk_full[b] ? 6 : 3;
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j < w; ++j)
              values[size_t(p.node_base[a] + int64_t(i) * p.len03[a] + p.blk_off03[b] + j)] =
                  acc[6 * i + j];
          if (p.blk_full[b])
            for (int i = 3; i < 6; ++i)
              for (int j = 0; j < 6; ++j)
                values[size_t(p.node_base[a] + 3 * int64_t(p.len03[a]) +
                              int64_t(i - 3) * p.len35[a] + p.blk_off35[b] + j)] = acc[6 * i + j];
        }
      }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (bad) return -1.0;
    best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
  }
  if (nnz_out) *nnz_out = p.nnz;
  if (coo_rows && coo_cols && coo_vals) {
    for (int64_t a = 0; a < m.n_nodes; ++a)
      for (int64_t b = p.node_blk_ptr[a]; b < p.node_blk_ptr[a + 1]; ++b) {
        int w = p.blk_full[b] ? 6 : 3, h = p.blk_full[b] ? 6 : 3;
        for (int i = 0; i < h; ++i)
          for (int j = 0; j < w; ++j) {
            int64_t pos = (i < 3) ? p.node_base[a] + int64_t(i) * p.len03[a] + p.blk_off03[b] + j
                                  : p.node_base[a] + 3 * int64_t(p.len03[a]) +
                                        int64_t(i - 3) * p.len35[a] + p.blk_off35[b] + j;
            coo_rows[pos] = 6 * a + i;
            coo_cols[pos] = 6 * int64_t(p.blk_col[b]) + j;
            coo_vals[pos] = values[size_t(pos)];
          }
      }
  }
  if (checksum_out) {
    double s = 0.0;
    for (double v : values) s += v;
    *checksum_out = s;
  }
  return best;
}

// ---- sampled block rows at full size ---------------------------------------------------------------
// For each node in `sample` (any order, no duplicates): every node-pair block (a, b) of its six rows, computed
// from scratch — adjacency from a scan over all elements, element matrices by the functions above — and summed
// in the reference's accumulation order for a model loaded family by family (plates, beams, trusses; insertion
// order inside a family: methods_for_*_data_handle.rs add_value loops). This is how tests/ compare the GPU
// matrix with the oracle at BASELINE.json's full sizes without holding 18 GB of element matrices.
// Output, per sampled node i: blocks [blk_ptr[i], blk_ptr[i+1]) sorted by column node; blk_col[], blk_full[]
// (1 = 6x6 beam/plate pair, 0 = 3x3 truss-only pair), blk_val[36 per block] row-major 6x6 (3x3 zero-padded).
struct SampledRows {
  std::vector<int64_t> blk_ptr;
  std::vector<uint32_t> blk_col;
  std::vector<uint8_t> blk_full;
  std::vector<double> blk_val;
};

// faithful = true: element matrices by the operation-by-operation restatement (truss_element / beam_element /
// plate_element of fem_oracle.hpp: dense (R^T k) R on Mat), i.e. the definition of parity; false: the fast ones above.
inline int element_kg(const Mesh& m, int order, int64_t e, bool faithful, double* kg) {
  if (!faithful) return order == 0 ? plate_kg(m, e, kg, false) : order == 1 ? beam_kg(m, e, kg) : truss_kg(m, e, kg);
  double p[4][3];
  if (order == 0) {
    const int64_t n = m.n_plate;
    for (int a = 0; a < 4; ++a) xyz(m, m.p_n[a * n + e], p[a]);
    PlateOut<double> o;
    const double* P = m.p_props;
    int err = plate_element<double>(p[0], p[1], p[2], p[3], P[e], P[n + e], P[2 * n + e], P[3 * n + e], m.rel_tol,
                                    m.abs_tol, o);
    if (err) return err;
    for (int i = 0; i < 24; ++i)
      for (int j = 0; j < 24; ++j) kg[24 * i + j] = o.k_global.at(i, j);
    return 0;
  }
  if (order == 1) {
    const int64_t n = m.n_beam;
    xyz(m, m.b_n1[e], p[0]);
    xyz(m, m.b_n2[e], p[1]);
    const double* P = m.b_props;
    const double ax[3] = {m.b_axis[e], m.b_axis[n + e], m.b_axis[2 * n + e]};
    BeamOut<double> o;
    int err = beam_element<double>(p[0], p[1], P[e], P[n + e], P[2 * n + e], P[3 * n + e], P[4 * n + e], P[5 * n + e],
                                   P[6 * n + e], P[7 * n + e], ax, m.rel_tol, m.abs_tol, o);
    if (err) return err;
    for (int i = 0; i < 12; ++i)
      for (int j = 0; j < 12; ++j) kg[12 * i + j] = o.k_global.at(i, j);
    return 0;
  }
  xyz(m, m.t_n1[e], p[0]);
  xyz(m, m.t_n2[e], p[1]);
  const bool has2 = m.t_A2 && !std::isnan(m.t_A2[e]);
  TrussOut<double> o;
  int err = truss_element<double>(p[0], p[1], m.t_E[e], m.t_A[e], has2, has2 ? m.t_A2[e] : 0.0, m.rel_tol, m.abs_tol, o);
  if (err) return err;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) kg[6 * i + j] = o.k_global.at(i, j);
  return 0;
}

inline int sample_rows(const Mesh& m, const uint32_t* sample, int64_t n_sample, int n_threads, bool faithful,
                       SampledRows& out) {
  if (n_threads > 0) omp_set_num_threads(n_threads);
  std::vector<int32_t> slot(size_t(m.n_nodes), -1);
  for (int64_t i = 0; i < n_sample; ++i) slot[sample[i]] = int32_t(i);
  struct C {
    uint32_t col;
    uint32_t order;  // family-major rank: plates 0, beams 1, trusses 2
    int64_t elem;
    int pair;
  };
  std::vector<std::vector<C>> lists;
  lists.resize(static_cast<size_t>(n_sample));
  auto push = [&](uint32_t order, int64_t e, const uint32_t* nd, int nn) {
    for (int a = 0; a < nn; ++a) {
      const int32_t s = slot[nd[a]];
      if (s < 0) continue;
      for (int b = 0; b < nn; ++b) lists[size_t(s)].push_back({nd[b], order, e, a * nn + b});
    }
  };
  for (int64_t e = 0; e < m.n_plate; ++e) {
    uint32_t nd[4] = {m.p_n[e], m.p_n[m.n_plate + e], m.p_n[2 * m.n_plate + e], m.p_n[3 * m.n_plate + e]};
    push(0, e, nd, 4);
  }
  for (int64_t e = 0; e < m.n_beam; ++e) {
    uint32_t nd[2] = {m.b_n1[e], m.b_n2[e]};
    push(1, e, nd, 2);
  }
  for (int64_t e = 0; e < m.n_truss; ++e) {
    uint32_t nd[2] = {m.t_n1[e], m.t_n2[e]};
    push(2, e, nd, 2);
  }
  out.blk_ptr.assign(size_t(n_sample) + 1, 0);
  for (int64_t i = 0; i < n_sample; ++i) {
    auto& L = lists[size_t(i)];
    std::stable_sort(L.begin(), L.end(), [](const C& a, const C& b) { return a.col < b.col; });  // keeps plates, beams, trusses / insertion order
    int64_t nb = 0;
    for (size_t k = 0; k < L.size(); ++k)
      if (k == 0 || L[k].col != L[k - 1].col) ++nb;
    out.blk_ptr[size_t(i) + 1] = out.blk_ptr[size_t(i)] + nb;
  }
  const int64_t total = out.blk_ptr[size_t(n_sample)];
  out.blk_col.assign(size_t(total), 0);
  out.blk_full.assign(size_t(total), 0);
  out.blk_val.assign(size_t(total) * 36, 0.0);
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n_sample; ++i) {
    const auto& L = lists[size_t(i)];
    int64_t blk = out.blk_ptr[size_t(i)] - 1;
    std::vector<double> kg(576);
    for (size_t k = 0; k < L.size(); ++k) {
      if (k == 0 || L[k].col != L[k - 1].col) {
        ++blk;
        out.blk_col[size_t(blk)] = L[k].col;
      }
      double* acc = &out.blk_val[size_t(blk) * 36];
      const C& c = L[k];
      if (element_kg(m, int(c.order), c.elem, faithful, kg.data())) bad = 1;
      if (c.order == 0) {
        const int la = c.pair / 4, lb = c.pair % 4;
        for (int r = 0; r < 6; ++r)
          for (int q = 0; q < 6; ++q) acc[6 * r + q] += kg[size_t((6 * la + r) * 24 + 6 * lb + q)];
        out.blk_full[size_t(blk)] = 1;
      } else if (c.order == 1) {
        const int la = c.pair / 2, lb = c.pair % 2;
        for (int r = 0; r < 6; ++r)
          for (int q = 0; q < 6; ++q) acc[6 * r + q] += kg[size_t((6 * la + r) * 12 + 6 * lb + q)];
        out.blk_full[size_t(blk)] = 1;
      } else {
        const int la = c.pair / 2, lb = c.pair % 2;
        for (int r = 0; r < 3; ++r)
          for (int q = 0; q < 3; ++q) acc[6 * r + q] += kg[size_t((3 * la + r) * 6 + 3 * lb + q)];
      }
    }
  }
  return bad ? -1 : 0;
}

}  // namespace fast
}  // namespace oracle
