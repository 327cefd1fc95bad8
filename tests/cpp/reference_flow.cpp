// Replays the reference crate's own integration tests (src/tests/fem/test_fem.rs:5-64 direct, :83-225 sparse
// iterative: nodes (0,0,0) and (30,0,0), one truss E = 1e6, A = 2, node 1 clamped, 100 on node 2 X) through the C++
// host mirror include/femgpu.hpp, in f64. Built and run by tests/test_cpp_header.py.
//   reference_flow <device>     device >= 0: the whole flow on that GPU; device < 0: staging-only handle, host checks only
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "femgpu.hpp"

#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                             \
    }                                                                       \
  } while (0)

static bool near(double a, double b, double rel) { return std::fabs(a - b) <= rel * std::fabs(b); }

int main(int argc, char** argv) {
  const int device = argc > 1 ? std::atoi(argv[1]) : 0;
  using femgpu::FEM;
  FEM fem = FEM::create(1e-4, 1e-12, 3, device < 0 ? FEMGPU_DEVICE_NONE : device);
  fem.add_node(1, 0.0, 0.0, 0.0);
  fem.add_node(2, 30.0, 0.0, 0.0);
  // Result<(), String> -> exception carrying the reference's text (methods_for_node_data_handle.rs:16-40)
  try {
    fem.add_node(2, 5.0, 0.0, 0.0);
    EXPECT(!"duplicate node number accepted");
  } catch (const femgpu::Error& e) {
    EXPECT(e.code == FEMGPU_E_NODE_NUMBER_EXISTS);
    EXPECT(std::string(e.what()) == "Node with number 2 already exists!");
  }
  fem.reset(2);  // FEM::reset (fem.rs:155): an empty model for two nodes — the reference's test model
  fem.add_node(1, 0.0, 0.0, 0.0);
  fem.add_node(2, 30.0, 0.0, 0.0);
  try {
    fem.add_node(3, 1.0, 1.0, 1.0);
    EXPECT(!"node limit ignored");
  } catch (const femgpu::Error& e) {
    EXPECT(std::string(e.what()) == "Nodes number could not be greater than 2!");
  }
  if (device < 0) {
    // no CPU compute path: the element is recorded by the host checks, its device validation refuses
    try {
      fem.add_truss(1, 1, 2, 1e6, 2.0);
      EXPECT(!"a staging-only handle validated an element");
    } catch (const femgpu::Error& e) {
      EXPECT(e.code == FEMGPU_ERR_NO_DEVICE);
    }
    std::printf("CPP_HEADER_OK staging-only\n");
    return 0;
  }
  fem.add_truss(1, 1, 2, 1e6, 2.0);
  try {
    fem.add_truss(2, 2, 1, 1e6, 2.0);
    EXPECT(!"duplicate node pair accepted");
  } catch (const femgpu::Error& e) {
    EXPECT(std::string(e.what()) == "Truss element with node number 2 and 1 already exists!");
  }
  const auto q = fem.get_truss_rotation_matrix_elements(1);
  EXPECT(q[0] == 1.0 && q[4] == 1.0 && q[8] == 1.0 && q[1] == 0.0);
  FEM::Csr K = fem.assemble();
  EXPECT(K.row_ptr.size() == 13 && K.values.size() == 36);  // two nodes, 3x3 blocks: rows 0-2 and 6-8 hold 6 entries
  EXPECT(near(K.values[0], 66666.66666666667, 1e-15));
  // a truss only has axial stiffness: every other DOF is inactive, and constraining one of them is an error
  // (methods_for_separate_stiffness_matrix.rs:233-262) — the reference's test constrains node 1 X only
  fem.add_displacement(1, 0, 0.0);
  fem.add_concentrated_load(2, 0, 100.0);
  // sparse iterative flow (test_fem.rs:83-225)
  auto sep = fem.separate_stiffness_matrix_sparse_iterative();
  EXPECT(sep.n_aa == 1 && sep.nnz[0] == 1);
  for (int block = 0; block < 2; ++block) {
    auto [u, iterations] = block ? fem.find_ua_vector_iterative_pcg_block_jacobi_sparse(1000)
                                 : fem.find_ua_vector_iterative_pcg_jacobi_sparse(1000);
    EXPECT(iterations == 1 && u.size() == 1 && near(u[0], 0.0015, 1e-14));
  }
  auto r = fem.find_r_r_vector_sparse();
  EXPECT(near(r[0], -100.0, 1e-13));
  auto force = fem.element_results(FEMGPU_TRUSS);
  EXPECT(force.size() == 1 && near(force[0], 100.0, 1e-13));
  // direct flow (test_fem.rs:5-64)
  auto sky = fem.separate_stiffness_matrix_direct();
  EXPECT(sky.a.size() == 1 && near(sky.a[0], 66666.66666666667, 1e-15));
  auto kaa = fem.separated_dense(0);   // SeparatedStiffnessMatrix::get_k_aa_matrix of the reference: 1 x 1 here
  auto kab = fem.separated_dense(1);
  EXPECT(kaa.size() == 1 && near(kaa[0], 66666.66666666667, 1e-15) && kab.size() == 1 && near(kab[0], -66666.66666666667, 1e-15));
  auto ud = fem.find_ua_vector_direct();
  EXPECT(near(ud[0], 0.0015, 1e-14));
  std::printf("CPP_HEADER_OK device %d\n", device);
  return 0;
}
