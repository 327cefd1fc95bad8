// ORACLE — TEST INFRASTRUCTURE ONLY (see fem_oracle.hpp). C entry points for ctypes.
// Loaded only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
#include <chrono>
#include <cstring>
#include <limits>

#include "fem_oracle.hpp"
#include "fem_oracle_fast.hpp"

using namespace oracle;

extern "C" {

const char* oracle_error_text(int code) {
  switch (code) {
    case OK: return "ok";
    case E_YOUNG: return "Young's modulus is less or equal to zero!";
    case E_POISSON: return "Poisson's ratio is less or equal to zero!";
    case E_AREA: return "Area is less or equal to zero!";
    case E_AREA2: return "Area2 is less or equal to zero!";
    case E_I11: return "I11 is less or equal to zero!";
    case E_I22: return "I22 is less or equal to zero!";
    case E_IT: return "It is less or equal to zero!";
    case E_SHEAR_FACTOR: return "Shear factor is less or equal to zero!";
    case E_PARALLEL_AXIS: return "Local axis 1 direction parallel to element!";
    case E_THICKNESS: return "Thickness is less or equal to zero!";
    case E_ON_LINE: return "Some nodes of element lie on the line!";
    case E_NOT_ON_PLANE: return "Not all nodes of element lie on the plane!";
    case E_NOT_CONVEX: return "Element non-convex!";
    default: return "unknown";
  }
}

// test-only switches of the restatement (fem_oracle.hpp Variants); returns the anti-parallel branch hits so far
long oracle_set_variants(int antiparallel, int inverse2, int reset_hits) {
  Variants& v = variants();
  if (antiparallel >= 0) v.antiparallel = antiparallel;
  if (inverse2 >= 0) v.inverse2 = inverse2;
  long hits = v.antiparallel_hits;
  if (reset_hits) v.antiparallel_hits = 0;
  return hits;
}

// ---------------------------------------------------------------- element level
int oracle_truss_f64(const double* p1, const double* p2, double E, double A, double A2,
                     double rel_tol, double abs_tol, double* q, double* k_local, double* k_global) {
  TrussOut<double> o;
  bool has2 = !std::isnan(A2);
  int err = truss_element<double>(p1, p2, E, A, has2, has2 ? A2 : 0.0, rel_tol, abs_tol, o);
  if (err) return err;
  std::memcpy(q, o.q, sizeof(o.q));
  std::memcpy(k_local, o.k_local.a.data(), 36 * sizeof(double));
  std::memcpy(k_global, o.k_global.a.data(), 36 * sizeof(double));
  return OK;
}

int oracle_truss_f32(const float* p1, const float* p2, float E, float A, float A2, float rel_tol,
                     float abs_tol, float* q, float* k_local, float* k_global) {
  TrussOut<float> o;
  bool has2 = !std::isnan(A2);
  int err = truss_element<float>(p1, p2, E, A, has2, has2 ? A2 : 0.0f, rel_tol, abs_tol, o);
  if (err) return err;
  std::memcpy(q, o.q, sizeof(o.q));
  std::memcpy(k_local, o.k_local.a.data(), 36 * sizeof(float));
  std::memcpy(k_global, o.k_global.a.data(), 36 * sizeof(float));
  return OK;
}

int oracle_beam_f64(const double* p1, const double* p2, double E, double nu, double A, double I11,
                    double I22, double I12, double It, double ks, const double* axis1,
                    double rel_tol, double abs_tol, double* q, double* principal /*I11p,I22p,angle*/,
                    double* k_local, double* k_global) {
  BeamOut<double> o;
  int err = beam_element<double>(p1, p2, E, nu, A, I11, I22, I12, It, ks, axis1, rel_tol, abs_tol, o);
  if (err) return err;
  std::memcpy(q, o.q, sizeof(o.q));
  principal[0] = o.i11_p;
  principal[1] = o.i22_p;
  principal[2] = o.angle;
  std::memcpy(k_local, o.k_local.a.data(), 144 * sizeof(double));
  std::memcpy(k_global, o.k_global.a.data(), 144 * sizeof(double));
  return OK;
}

int oracle_plate_f64(const double* p1, const double* p2, const double* p3, const double* p4,
                     double E, double nu, double t, double ks, double rel_tol, double abs_tol,
                     double* q, double* k_local, double* k_global) {
  PlateOut<double> o;
  int err = plate_element<double>(p1, p2, p3, p4, E, nu, t, ks, rel_tol, abs_tol, o);
  if (err) return err;
  std::memcpy(q, o.q, sizeof(o.q));
  std::memcpy(k_local, o.k_local.a.data(), 576 * sizeof(double));
  std::memcpy(k_global, o.k_global.a.data(), 576 * sizeof(double));
  return OK;
}

void oracle_beam_line_load_f64(const double* p1, const double* p2, double q, double* f) {
  beam_line_load_nodal<double>(p1, p2, q, f);
}

void oracle_plate_surface_load_f64(const double* p1, const double* p2, const double* p3, const double* p4,
                                   double q, double rel_tol, double abs_tol, double* f) {
  plate_surface_load_nodal<double>(p1, p2, p3, p4, q, rel_tol, abs_tol, f);
}

// The reference's only test model, replayed in f32 end to end
// (tests/fem/test_fem.rs:5-64): 2 nodes (0,0,0),(30,0,0); truss E=1e6, A=2; u1x=0; F2x=100.
// K_aa is 1x1, so colsol's LDL^T solve is u = r / k; the reaction is K_ba*u_a + K_bb*u_b - R_b
// (methods_for_global_analysis.rs:277-333) and the element force follows truss.rs:281-333.
int oracle_reference_truss_test_f32(float* k00, float* u2x, float* r1x, float* force_r) {
  const float p1[3] = {0.0f, 0.0f, 0.0f}, p2[3] = {30.0f, 0.0f, 0.0f};
  TrussOut<float> o;
  int err = truss_element<float>(p1, p2, 1e6f, 2.0f, false, 0.0f, 1e-4f, 1e-12f, o);
  if (err) return err;
  GlobalK<float> K;
  const uint32_t idx[2] = {0, 1};
  scatter_blocks(K, o.k_global, idx, 2, 3);
  float kaa = K.e[GlobalK<float>::key(6, 6)];
  float kba = K.e[GlobalK<float>::key(0, 6)];
  float ra = 100.0f;
  float ua = ra / kaa;           // 1x1 LDL^T
  float rr = kba * ua;           // u_b = 0, R_b = 0
  // truss.rs:281-333 (truss_element_result): local displacements = R * [u1; u2], strain = B * u_local,
  // force = E * A * strain
  const float ue[6] = {0.0f, 0.0f, 0.0f, ua, 0.0f, 0.0f};
  const float force_value = truss_element_result<float>(p1, p2, 1e6f, 2.0f, false, 0.0f, 1e-4f, 1e-12f, ue);
  *k00 = kaa;
  *u2x = ua;
  *r1x = rr;
  *force_r = force_value;
  return OK;
}

// ---------------------------------------------------------------- model level (faithful)
struct OracleModel {
  double rel_tol, abs_tol;
  uint32_t nodes_number;
  std::vector<double> x, y, z;
  GlobalK<double> K;
  int64_t n_elements = 0;
};

void* oracle_model_create(double rel_tol, double abs_tol, uint32_t nodes_number) {
  auto* m = new OracleModel();
  m->rel_tol = rel_tol;
  m->abs_tol = abs_tol;
  m->nodes_number = nodes_number;
  return m;
}
void oracle_model_destroy(void* h) { delete static_cast<OracleModel*>(h); }

// node index = position in these arrays (methods_for_node_data_handle.rs:66-78)
int oracle_model_set_nodes(void* h, int64_t n, const double* x, const double* y, const double* z) {
  auto* m = static_cast<OracleModel*>(h);
  if (uint64_t(n) > m->nodes_number) return -1;
  m->x.assign(x, x + n);
  m->y.assign(y, y + n);
  m->z.assign(z, z + n);
  return 0;
}

static inline void node_xyz(const OracleModel* m, uint32_t i, double p[3]) {
  p[0] = m->x[i];
  p[1] = m->y[i];
  p[2] = m->z[i];
}

// returns 0, or the element error code with *fail_at = index of the offending element
int oracle_model_add_truss(void* h, int64_t n, const uint32_t* n1, const uint32_t* n2,
                           const double* E, const double* A, const double* A2, int64_t* fail_at) {
  auto* m = static_cast<OracleModel*>(h);
  for (int64_t e = 0; e < n; ++e) {
    double p1[3], p2[3];
    node_xyz(m, n1[e], p1);
    node_xyz(m, n2[e], p2);
    TrussOut<double> o;
    bool has2 = A2 && !std::isnan(A2[e]);
    int err = truss_element<double>(p1, p2, E[e], A[e], has2, has2 ? A2[e] : 0.0, m->rel_tol,
                                    m->abs_tol, o);
    if (err) {
      if (fail_at) *fail_at = e;
      return err;
    }
    const uint32_t idx[2] = {n1[e], n2[e]};
    scatter_blocks(m->K, o.k_global, idx, 2, 3);
    m->n_elements++;
  }
  return 0;
}

int oracle_model_add_beam(void* h, int64_t n, const uint32_t* n1, const uint32_t* n2,
                          const double* E, const double* nu, const double* A, const double* I11,
                          const double* I22, const double* I12, const double* It, const double* ks,
                          const double* axis1 /* [3][n] SoA */, int64_t* fail_at) {
  auto* m = static_cast<OracleModel*>(h);
  for (int64_t e = 0; e < n; ++e) {
    double p1[3], p2[3];
    node_xyz(m, n1[e], p1);
    node_xyz(m, n2[e], p2);
    const double ax[3] = {axis1[e], axis1[n + e], axis1[2 * n + e]};
    BeamOut<double> o;
    int err = beam_element<double>(p1, p2, E[e], nu[e], A[e], I11[e], I22[e], I12[e], It[e], ks[e],
                                   ax, m->rel_tol, m->abs_tol, o);
    if (err) {
      if (fail_at) *fail_at = e;
      return err;
    }
    const uint32_t idx[2] = {n1[e], n2[e]};
    scatter_blocks(m->K, o.k_global, idx, 2, 6);
    m->n_elements++;
  }
  return 0;
}

int oracle_model_add_plate(void* h, int64_t n, const uint32_t* n1, const uint32_t* n2,
                           const uint32_t* n3, const uint32_t* n4, const double* E,
                           const double* nu, const double* t, const double* ks, int64_t* fail_at) {
  auto* m = static_cast<OracleModel*>(h);
  for (int64_t e = 0; e < n; ++e) {
    double p1[3], p2[3], p3[3], p4[3];
    node_xyz(m, n1[e], p1);
    node_xyz(m, n2[e], p2);
    node_xyz(m, n3[e], p3);
    node_xyz(m, n4[e], p4);
    PlateOut<double> o;
    int err = plate_element<double>(p1, p2, p3, p4, E[e], nu[e], t[e], ks[e], m->rel_tol,
                                    m->abs_tol, o);
    if (err) {
      if (fail_at) *fail_at = e;
      return err;
    }
    const uint32_t idx[4] = {n1[e], n2[e], n3[e], n4[e]};
    scatter_blocks(m->K, o.k_global, idx, 4, 6);
    m->n_elements++;
  }
  return 0;
}

int64_t oracle_model_nnz(void* h) { return int64_t(static_cast<OracleModel*>(h)->K.e.size()); }

// stored entries (including ones that summed to exactly 0), sorted by (row, col)
void oracle_model_get_coo(void* h, int64_t* rows, int64_t* cols, double* vals) {
  auto* m = static_cast<OracleModel*>(h);
  std::vector<std::pair<uint64_t, double>> v(m->K.e.begin(), m->K.e.end());
  std::sort(v.begin(), v.end(), [](auto& a, auto& b) { return a.first < b.first; });
  for (size_t i = 0; i < v.size(); ++i) {
    rows[i] = int64_t(v[i].first >> 32);
    cols[i] = int64_t(v[i].first & 0xffffffffu);
    vals[i] = v[i].second;
  }
}

// ---------------------------------------------------------------- element result recovery
// extract_elements_analysis_result (methods_for_element_analysis.rs:27-58) over a whole mesh: `u` is the
// global displacement vector (6 per node, node index order). out_t [n_truss], out_b [n_beam][10],
// out_p [n_plate][8] in the component order of truss.rs:325-328, beam.rs:967-987, plate.rs:1368-1401.
void oracle_element_results(int64_t n_nodes, const double* x, const double* y, const double* z,
                            int64_t n_truss, const uint32_t* t_n1, const uint32_t* t_n2,
                            const double* t_E, const double* t_A, const double* t_A2,
                            int64_t n_beam, const uint32_t* b_n1, const uint32_t* b_n2,
                            const double* b_props, const double* b_axis, int64_t n_plate,
                            const uint32_t* p_n, const double* p_props, double rel_tol,
                            double abs_tol, const double* u, double* out_t, double* out_b, double* out_p) {
  auto xyz = [&](uint32_t i, double p[3]) { p[0] = x[i]; p[1] = y[i]; p[2] = z[i]; };
  for (int64_t e = 0; e < n_truss; ++e) {
    double p1[3], p2[3], ue[6];
    xyz(t_n1[e], p1);
    xyz(t_n2[e], p2);
    for (int i = 0; i < 3; ++i) {
      ue[i] = u[size_t(t_n1[e]) * NODE_DOF + i];
      ue[3 + i] = u[size_t(t_n2[e]) * NODE_DOF + i];
    }
    const bool has2 = t_A2 && !std::isnan(t_A2[e]);
    out_t[e] = truss_element_result<double>(p1, p2, t_E[e], t_A[e], has2, has2 ? t_A2[e] : 0.0, rel_tol, abs_tol, ue);
  }
  for (int64_t e = 0; e < n_beam; ++e) {
    double p1[3], p2[3], ue[12];
    xyz(b_n1[e], p1);
    xyz(b_n2[e], p2);
    for (int i = 0; i < 6; ++i) {
      ue[i] = u[size_t(b_n1[e]) * NODE_DOF + i];
      ue[6 + i] = u[size_t(b_n2[e]) * NODE_DOF + i];
    }
    const double ax[3] = {b_axis[e], b_axis[n_beam + e], b_axis[2 * n_beam + e]};
    const double* bp = b_props;
    beam_element_result<double>(p1, p2, bp[e], bp[n_beam + e], bp[2 * n_beam + e], bp[3 * n_beam + e],
                                bp[4 * n_beam + e], bp[5 * n_beam + e], bp[6 * n_beam + e], bp[7 * n_beam + e],
                                ax, rel_tol, abs_tol, ue, out_b + 10 * e);
  }
  for (int64_t e = 0; e < n_plate; ++e) {
    double p[4][3], ue[24];
    for (int a = 0; a < 4; ++a) {
      const uint32_t n = p_n[a * n_plate + e];
      xyz(n, p[a]);
      for (int i = 0; i < 6; ++i) ue[6 * a + i] = u[size_t(n) * NODE_DOF + i];
    }
    plate_element_result<double>(p[0], p[1], p[2], p[3], p_props[e], p_props[n_plate + e],
                                 p_props[2 * n_plate + e], p_props[3 * n_plate + e], rel_tol, abs_tol, ue,
                                 out_p + 8 * e);
  }
}

// ---------------------------------------------------------------- fast multi-core baseline
// See fem_oracle_fast.hpp. Returns seconds spent in the numeric part (element matrices +
// accumulation into a prebuilt block-CSR), pattern construction excluded — the same split the GPU
// numbers use. `values_out` (may be null) receives the structural-pattern CSR values.
double oracle_fast_assemble(int64_t n_nodes, const double* x, const double* y, const double* z,
                            int64_t n_truss, const uint32_t* t_n1, const uint32_t* t_n2,
                            const double* t_E, const double* t_A, const double* t_A2,
                            int64_t n_beam, const uint32_t* b_n1, const uint32_t* b_n2,
                            const double* b_props /* [8][n]: E,nu,A,I11,I22,I12,It,ks */,
                            const double* b_axis /* [3][n] */, int64_t n_plate,
                            const uint32_t* p_n /* [4][n] */,
                            const double* p_props /* [4][n]: E,nu,t,ks */, double rel_tol,
                            double abs_tol, int n_threads, int repeats, int64_t* nnz_out,
                            double* checksum_out, int64_t* coo_rows, int64_t* coo_cols,
                            double* coo_vals) {
  fast::Mesh mesh{n_nodes, x, y, z, n_truss, t_n1, t_n2, t_E, t_A, t_A2, n_beam, b_n1, b_n2,
                  b_props, b_axis, n_plate, p_n, p_props, rel_tol, abs_tol};
  return fast::assemble(mesh, n_threads, repeats, nnz_out, checksum_out, coo_rows, coo_cols, coo_vals);
}


// Sampled block rows of the global matrix at any mesh size (fem_oracle_fast.hpp sample_rows). Two calls: with
// blk_col == nullptr it returns the number of blocks (and fills blk_ptr [n_sample + 1]); the second call fills
// blk_col / blk_full / blk_val (36 per block). Returns < 0 when an element is invalid.
int64_t oracle_sample_rows(int64_t n_nodes, const double* x, const double* y, const double* z,
                           int64_t n_truss, const uint32_t* t_n1, const uint32_t* t_n2,
                           const double* t_E, const double* t_A, const double* t_A2,
                           int64_t n_beam, const uint32_t* b_n1, const uint32_t* b_n2,
                           const double* b_props, const double* b_axis, int64_t n_plate,
                           const uint32_t* p_n, const double* p_props, double rel_tol, double abs_tol,
                           int n_threads, int faithful, int64_t n_sample, const uint32_t* sample, int64_t* blk_ptr,
                           uint32_t* blk_col, uint8_t* blk_full, double* blk_val) {
  fast::Mesh mesh{n_nodes, x, y, z, n_truss, t_n1, t_n2, t_E, t_A, t_A2, n_beam, b_n1, b_n2,
                  b_props, b_axis, n_plate, p_n, p_props, rel_tol, abs_tol};
  static thread_local fast::SampledRows cache;  // the second call returns what the first one computed
  if (!blk_col) {
    if (fast::sample_rows(mesh, sample, n_sample, n_threads, faithful != 0, cache)) return -1;
    std::memcpy(blk_ptr, cache.blk_ptr.data(), cache.blk_ptr.size() * sizeof(int64_t));
    return int64_t(cache.blk_col.size());
  }
  std::memcpy(blk_col, cache.blk_col.data(), cache.blk_col.size() * sizeof(uint32_t));
  std::memcpy(blk_full, cache.blk_full.data(), cache.blk_full.size());
  std::memcpy(blk_val, cache.blk_val.data(), cache.blk_val.size() * sizeof(double));
  const int64_t n = int64_t(cache.blk_col.size());
  cache = fast::SampledRows();
  return n;
}

// single-thread faithful timing on the same arrays: seconds for `add_*` of everything
double oracle_faithful_time(int64_t n_nodes, const double* x, const double* y, const double* z,
                            int64_t n_truss, const uint32_t* t_n1, const uint32_t* t_n2,
                            const double* t_E, const double* t_A, const double* t_A2,
                            int64_t n_beam, const uint32_t* b_n1, const uint32_t* b_n2,
                            const double* b_props, const double* b_axis, int64_t n_plate,
                            const uint32_t* p_n, const double* p_props, double rel_tol,
                            double abs_tol) {
  void* h = oracle_model_create(rel_tol, abs_tol, uint32_t(n_nodes));
  oracle_model_set_nodes(h, n_nodes, x, y, z);
  auto t0 = std::chrono::steady_clock::now();
  int64_t fail = -1;
  if (n_plate)
    oracle_model_add_plate(h, n_plate, p_n, p_n + n_plate, p_n + 2 * n_plate, p_n + 3 * n_plate,
                           p_props, p_props + n_plate, p_props + 2 * n_plate,
                           p_props + 3 * n_plate, &fail);
  if (n_beam)
    oracle_model_add_beam(h, n_beam, b_n1, b_n2, b_props, b_props + n_beam, b_props + 2 * n_beam,
                          b_props + 3 * n_beam, b_props + 4 * n_beam, b_props + 5 * n_beam,
                          b_props + 6 * n_beam, b_props + 7 * n_beam, b_axis, &fail);
  if (n_truss) oracle_model_add_truss(h, n_truss, t_n1, t_n2, t_E, t_A, t_A2, &fail);
  auto t1 = std::chrono::steady_clock::now();
  oracle_model_destroy(h);
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
