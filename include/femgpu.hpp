// C++ host mirror of the reference's FEM<f64> API for the stiffness-assembly path, header-only over
// the C ABI (include/femgpu.h). Same method names and argument order as the Rust crate
// (fem.rs:34,155,171-202; methods_for_{node,truss,beam,plate}_data_handle.rs); `Result<(), String>`
// becomes "returns, or throws femgpu::Error carrying the reference's message".
#pragma once
#include <array>
#include <cmath>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "femgpu.h"

namespace femgpu {

struct Error : std::runtime_error {
  int32_t code;
  Error(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};

class FEM {
 public:
  // FEM::create(rel_tol, abs_tol, nodes_number)                                        fem.rs:34
  static FEM create(double rel_tol, double abs_tol, uint32_t nodes_number, int device = 0) {
    FEM f;
    int32_t st = femgpu_create(&f.h_, rel_tol, abs_tol, nodes_number, device);
    if (st) throw Error(st, femgpu_last_error(nullptr));
    return f;
  }
  FEM(FEM&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  FEM(const FEM&) = delete;
  ~FEM() { femgpu_destroy(h_); }

  void reset(uint32_t nodes_number) { check(femgpu_reset(h_, nodes_number)); }  // fem.rs:155
  void add_node(uint32_t number, double x, double y, double z) {               // ..node..:66
    check(femgpu_add_nodes(h_, 1, &number, &x, &y, &z));
  }
  void add_truss(uint32_t number, uint32_t node_1_number, uint32_t node_2_number, double young_modulus,
                 double area, std::optional<double> optional_area_2 = std::nullopt) {  // ..truss..:49
    double a2 = optional_area_2.value_or(NAN);
    check(femgpu_add_truss(h_, 1, &number, &node_1_number, &node_2_number, &young_modulus, &area, &a2));
    check(femgpu_validate(h_, nullptr, nullptr, nullptr));
  }
  void add_beam(uint32_t number, uint32_t node_1_number, uint32_t node_2_number, double young_modulus,
                double poisson_ratio, double area, double i11, double i22, double i12, double it,
                double shear_factor, std::array<double, 3> local_axis_1_direction) {   // ..beam..:49
    check(femgpu_add_beam(h_, 1, &number, &node_1_number, &node_2_number, &young_modulus, &poisson_ratio, &area,
                          &i11, &i22, &i12, &it, &shear_factor, local_axis_1_direction.data()));
    check(femgpu_validate(h_, nullptr, nullptr, nullptr));
  }
  void add_plate(uint32_t number, uint32_t node_1_number, uint32_t node_2_number, uint32_t node_3_number,
                 uint32_t node_4_number, double young_modulus, double poisson_ratio, double thickness,
                 double shear_factor) {                                                // ..plate..:62
    check(femgpu_add_plate(h_, 1, &number, &node_1_number, &node_2_number, &node_3_number, &node_4_number,
                           &young_modulus, &poisson_ratio, &thickness, &shear_factor));
    check(femgpu_validate(h_, nullptr, nullptr, nullptr));
  }
  std::array<double, 9> get_truss_rotation_matrix_elements(uint32_t n) { return rot(FEMGPU_TRUSS, n); }  // fem.rs:171
  std::array<double, 9> get_beam_rotation_matrix_elements(uint32_t n) { return rot(FEMGPU_BEAM, n); }    // fem.rs:182
  std::array<double, 9> get_plate_rotation_matrix_elements(uint32_t n) { return rot(FEMGPU_PLATE, n); }  // fem.rs:193

  struct Csr {
    std::vector<int64_t> row_ptr;
    std::vector<int32_t> col_idx;
    std::vector<double> values;
  };
  // the assembled global stiffness matrix (self.stiffness_matrix of fem.rs:17) on the structural pattern
  Csr assemble() {
    int64_t n_rows = 0, nnz = 0;
    check(femgpu_assemble(h_, &n_rows, &nnz));
    Csr m{std::vector<int64_t>(size_t(n_rows) + 1), std::vector<int32_t>(size_t(nnz)), std::vector<double>(size_t(nnz))};
    check(femgpu_get_csr(h_, m.row_ptr.data(), m.col_idx.data(), m.values.data()));
    return m;
  }
  // ---- boundary conditions, separation, solve, results: the reference's sparse-iterative flow (tests/fem/test_fem.rs:83-225)
  void add_displacement(uint32_t node_number, int dof_parameter, double value) {       // methods_for_bc_data_handle.rs:175
    check(femgpu_add_displacement(h_, 1, &node_number, &dof_parameter, &value));
  }
  void add_concentrated_load(uint32_t node_number, int dof_parameter, double value) {  // methods_for_bc_data_handle.rs:31
    check(femgpu_add_concentrated_load(h_, 1, &node_number, &dof_parameter, &value));
  }
  struct Separated {
    int64_t n_aa = 0, n_bb = 0, nnz[4] = {0, 0, 0, 0};
  };
  // methods_for_separate_stiffness_matrix.rs:217; the quadrants stay in HBM (femgpu_get_separated_csr copies them out)
  Separated separate_stiffness_matrix_sparse_iterative() {
    Separated s;
    check(femgpu_separate_sparse(h_, &s.n_aa, &s.n_bb, s.nnz));
    sep_ = s;
    return s;
  }
  struct Skyline {
    std::vector<int64_t> k_aa_skyline, maxa;
    std::vector<double> a;
  };
  // diagnostics: did the last separation take the opt-in one-pass kernel (FEMGPU_SEP_ONE_PASS=1)?
  bool last_separation_read_k_once() {
    int32_t flag = 0;
    check(femgpu_last_separate_path(h_, &flag));
    return flag != 0;
  }

  // methods_for_separate_stiffness_matrix.rs:63 without the dense detour: K_aa as (a, maxa), the compacted column form
  // of methods_for_global_analysis.rs:50-80
  Skyline separate_stiffness_matrix_direct() {
    Separated s;
    int64_t n_val = 0;
    check(femgpu_separate_direct(h_, &s.n_aa, &s.n_bb, &n_val));
    sep_ = s;
    Skyline k{std::vector<int64_t>(size_t(s.n_aa)), std::vector<int64_t>(size_t(s.n_aa) + 1), std::vector<double>(size_t(n_val))};
    check(femgpu_get_skyline(h_, k.k_aa_skyline.data(), k.a.data(), k.maxa.data()));
    return k;
  }
  // k_aa_matrix / k_ab_matrix / k_ba_matrix / k_bb_matrix of the reference's SeparatedStiffnessMatrix
  // (structs/separated_stiffness_matrix.rs:8-16) as a dense row-major matrix; which = 0 aa, 1 ab, 2 ba, 3 bb
  std::vector<double> separated_dense(int which) {
    const int64_t rows = which < 2 ? sep_.n_aa : sep_.n_bb, cols = (which == 0 || which == 2) ? sep_.n_aa : sep_.n_bb;
    std::vector<double> m(size_t(rows) * size_t(cols));
    if (!m.empty()) check(femgpu_get_separated_dense(h_, which, m.data()));
    return m;
  }
  // methods_for_global_analysis.rs:161 (skyline LDL^T, COLSOL)
  std::vector<double> find_ua_vector_direct() {
    check(femgpu_solve_direct(h_));
    std::vector<double> u(size_t(sep_.n_aa));
    check(femgpu_get_ua(h_, u.data(), nullptr));
    return u;
  }
  // methods_for_global_analysis.rs:189 / :235 -> (u_a, iterations)
  std::pair<std::vector<double>, int64_t> find_ua_vector_iterative_pcg_jacobi_sparse(int64_t max_iter) {
    return solve(FEMGPU_PCG_JACOBI, max_iter);
  }
  std::pair<std::vector<double>, int64_t> find_ua_vector_iterative_pcg_block_jacobi_sparse(int64_t max_iter) {
    return solve(FEMGPU_PCG_BLOCK_JACOBI, max_iter);
  }
  // find_r_r_vector_sparse (:334) + compose_global_analysis_result (:362)
  std::vector<double> find_r_r_vector_sparse() {
    check(femgpu_global_analysis(h_));
    std::vector<double> r(size_t(sep_.n_bb));
    check(femgpu_get_reactions(h_, r.data(), nullptr));
    return r;
  }
  // extract_elements_analysis_result (methods_for_element_analysis.rs:27): values of one family, element-major
  // (1 / 10 / 8 per element, component order of femgpu.h)
  std::vector<double> element_results(int family) {
    uint64_t n[4] = {0, 0, 0, 0};
    check(femgpu_counts(h_, &n[0], &n[1], &n[2], &n[3]));
    static const size_t comps[3] = {1, 10, 8};
    std::vector<double> out(size_t(n[1 + family]) * comps[family]);
    check(femgpu_element_results(h_, family, out.data(), nullptr));
    return out;
  }
  femgpu_t* handle() { return h_; }

 private:
  FEM() = default;
  void check(int32_t st) {
    if (st) throw Error(st, femgpu_last_error(h_));
  }
  std::array<double, 9> rot(int family, uint32_t number) {
    std::array<double, 9> out{};
    check(femgpu_rotation_elements(h_, family, number, out.data()));
    return out;
  }
  std::pair<std::vector<double>, int64_t> solve(int preconditioner, int64_t max_iter) {
    int64_t it = 0;
    check(femgpu_solve_pcg(h_, preconditioner, max_iter, &it));
    std::vector<double> u(size_t(sep_.n_aa));
    check(femgpu_get_ua(h_, u.data(), nullptr));
    return {std::move(u), it};
  }
  femgpu_t* h_ = nullptr;
  Separated sep_;
};

}  // namespace femgpu
