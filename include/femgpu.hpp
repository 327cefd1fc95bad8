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
  femgpu_t* h_ = nullptr;
};

}  // namespace femgpu
