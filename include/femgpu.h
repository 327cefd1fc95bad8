/*
 * femgpu — C ABI of the B200-native structural-FEM stiffness assembly path.
 *
 * Drop-in boundary for ONE hot path of RomanShushakov/finite_element_method v0.9.12: per-element
 * local stiffness (truss / beam / plate) + scatter-assembly into the global FP64 stiffness matrix.
 * The reference has no FFI of its own; the seam is the inherent-method API of `FEM<V>`. Every entry
 * point below cites the reference item it replaces (paths relative to /root/reference/src/fem/).
 * A Rust host binds these with an `extern "C"` block; plain pointers and
 * sizes only, no torch / C++ types.
 *
 * Conventions
 *   - every function returns int32: 0 = ok, >0 = reference-style validation error (FEMGPU_E_*),
 *     <0 = CUDA / NCCL / usage failure. femgpu_last_error() gives the text; for validation errors
 *     it is the exact `Err(String)` the reference would return.
 *   - host arrays are caller-owned and copied during the call; the handle owns all device memory.
 *   - node / element *numbers* are user labels (u32); node *indices* are 0-based insertion order
 *     (methods_for_node_data_handle.rs:66-78); global row = 6*index + dof (structs/node.rs:8).
 *   - batched adds are "prefix-atomic": elements before the first failing one are accepted, the
 *     failing one and everything after it in the batch are not — what a sequential `add_*?` loop
 *     over the reference would leave behind.
 *   - one handle <-> one `FEM` <-> one GPU of one process; not thread-safe (the reference is not
 *     either). Multi-GPU = one process per GPU, each with a handle, joined by femgpu_dist_init().
 *   - FP64 only. There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef FEMGPU_H
#define FEMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct femgpu_handle femgpu_t;

/* pass as `device` to femgpu_create for a staging-only handle: host bookkeeping (numbering,
 * duplicate checks, error texts) works, every call that needs the GPU returns
 * FEMGPU_ERR_NO_DEVICE. There is no CPU compute path behind it. */
#define FEMGPU_DEVICE_NONE (-1)

/* element families */
#define FEMGPU_TRUSS 0
#define FEMGPU_BEAM 1
#define FEMGPU_PLATE 2

/* validation error codes (>0). Texts follow the reference's compose_error_message() functions. */
enum {
  FEMGPU_OK = 0,
  /* methods_for_node_data_handle.rs:16-32 */
  FEMGPU_E_NODE_NUMBER_EXISTS = 1,
  FEMGPU_E_NODE_NOT_EXIST = 2,
  FEMGPU_E_NODE_INDEX_EXISTS = 3,
  FEMGPU_E_NODE_COORDINATES_EXIST = 4,
  FEMGPU_E_NODE_LIMIT = 5,
  /* methods_for_{truss,beam,plate}_data_handle.rs error enums */
  FEMGPU_E_ELEMENT_NUMBER_EXISTS = 10,
  FEMGPU_E_ELEMENT_SAME_NODES = 11,
  FEMGPU_E_ELEMENT_NOT_EXIST = 12,
  /* structs/truss.rs:23-42, structs/beam.rs:26-61, structs/plate.rs:27-56 */
  FEMGPU_E_YOUNG_MODULUS = 20,
  FEMGPU_E_POISSON_RATIO = 21,
  FEMGPU_E_AREA = 22,
  FEMGPU_E_AREA2 = 23,
  FEMGPU_E_I11 = 24,
  FEMGPU_E_I22 = 25,
  FEMGPU_E_IT = 26,
  FEMGPU_E_SHEAR_FACTOR = 27,
  FEMGPU_E_PARALLEL_LOCAL_AXIS = 28,
  FEMGPU_E_THICKNESS = 29,
  FEMGPU_E_NODES_ON_LINE = 30,
  FEMGPU_E_NODES_NOT_ON_PLANE = 31,
  FEMGPU_E_NOT_CONVEX = 32,
  /* methods_for_bc_data_handle.rs:190-194 */
  FEMGPU_E_DISPLACEMENT_EXISTS = 40,
  /* methods_for_separate_stiffness_matrix.rs:233-243, :257-259, :303-307 */
  FEMGPU_E_NO_STIFFNESS_FOR_DISPLACEMENT = 41,
  FEMGPU_E_NO_RESTRAINTS = 42,
  FEMGPU_E_KAA_EMPTY = 43,
  /* methods_for_separate_stiffness_matrix.rs:49-58 (direct separation only) */
  FEMGPU_E_NO_STIFFNESS_FOR_LOAD = 44,
  /* find_ua_vector_iterative_*: "PCG failed" (methods_for_global_analysis.rs:228,272) */
  FEMGPU_E_SOLVER = 50
};

/* failures (<0) */
enum {
  FEMGPU_ERR_CUDA = -1,
  FEMGPU_ERR_USAGE = -2,
  FEMGPU_ERR_NCCL = -3,
  FEMGPU_ERR_LIMIT = -4,
  FEMGPU_ERR_NO_DEVICE = -5
};

/* ---- lifecycle ---------------------------------------------------------------------------- */

/* FEM::create(rel_tol, abs_tol, nodes_number)                                   fem.rs:34-65
 * `device` is the CUDA ordinal this handle computes on. The matrix order is fixed at
 * 6*nodes_number (fem.rs:37) whether or not that many nodes are ever added. */
int32_t femgpu_create(femgpu_t** out, double rel_tol, double abs_tol, uint32_t nodes_number,
                      int32_t device);
/* FEM::reset(nodes_number): drops nodes, elements, pattern and values            fem.rs:155-169 */
int32_t femgpu_reset(femgpu_t* h, uint32_t nodes_number);
void femgpu_destroy(femgpu_t* h);
/* text of the last non-zero status returned on this handle (NULL handle: creation errors) */
const char* femgpu_last_error(const femgpu_t* h);

/* ---- model definition --------------------------------------------------------------------- */

/* n x FEM::add_node(number, x, y, z), in array order     methods_for_node_data_handle.rs:66-78
 * Checks, per node and in this order: nodes_number limit, duplicate number, duplicate
 * coordinates (:42-64). The reference scans all nodes per call; here it is a hash lookup. */
int32_t femgpu_add_nodes(femgpu_t* h, size_t n, const uint32_t* number, const double* x,
                         const double* y, const double* z);

/* n x FEM::add_truss(number, node_1, node_2, young_modulus, area, optional_area_2)
 *                                                    methods_for_truss_data_handle.rs:49-128
 * area_2 may be NULL (all None); a NaN entry means None for that element. */
int32_t femgpu_add_truss(femgpu_t* h, size_t n, const uint32_t* number, const uint32_t* node_1,
                         const uint32_t* node_2, const double* young_modulus, const double* area,
                         const double* area_2);

/* n x FEM::add_beam(number, node_1, node_2, young_modulus, poisson_ratio, area, i11, i22, i12, it,
 *                   shear_factor, local_axis_1_direction)   methods_for_beam_data_handle.rs:49-142
 * local_axis_1 is struct-of-arrays: [3][n] (all x, then all y, then all z). */
int32_t femgpu_add_beam(femgpu_t* h, size_t n, const uint32_t* number, const uint32_t* node_1,
                        const uint32_t* node_2, const double* young_modulus,
                        const double* poisson_ratio, const double* area, const double* i11,
                        const double* i22, const double* i12, const double* it,
                        const double* shear_factor, const double* local_axis_1);

/* n x FEM::add_plate(number, node_1..4, young_modulus, poisson_ratio, thickness, shear_factor)
 *                                                    methods_for_plate_data_handle.rs:62-217 */
int32_t femgpu_add_plate(femgpu_t* h, size_t n, const uint32_t* number, const uint32_t* node_1,
                         const uint32_t* node_2, const uint32_t* node_3, const uint32_t* node_4,
                         const double* young_modulus, const double* poisson_ratio,
                         const double* thickness, const double* shear_factor);

/* Runs the element-level checks of Truss/Beam/Plate::create (structs/truss.rs:44-64,
 * structs/beam.rs:63-117, structs/plate.rs:58-158) on the device for every element added since the
 * last validation. On failure the offending element and everything added after it are rolled
 * back (prefix semantics) and the reference's message is returned. femgpu_symbolic() calls this
 * implicitly; the single-element host mirrors call it after each add so errors surface eagerly.
 * The three out-pointers may be NULL. */
int32_t femgpu_validate(femgpu_t* h, int32_t* family, uint32_t* number, int32_t* code);

/* counts of accepted entities */
int32_t femgpu_counts(const femgpu_t* h, uint64_t* nodes, uint64_t* truss, uint64_t* beam,
                      uint64_t* plate);
/* user labels in insertion order (index i = the i-th node / element added): family -1 = nodes
 * [nodes], 0 / 1 / 2 = truss / beam / plate element numbers. Host bookkeeping only. */
int32_t femgpu_get_numbers(const femgpu_t* h, int32_t family, uint32_t* out);

/* ---- assembly ----------------------------------------------------------------------------- */

/* One-time symbolic pass: DOF numbering (6*index+dof), node-pair block list, CSR row_ptr/col_idx
 * on the structural pattern (3x3 per truss-only node pair, 6x6 per beam/plate pair, union when
 * families share a pair), and the CSR-slot gather map used by the numeric pass. Replaces the
 * position-keyed `SquareMatrix` of fem.rs:17,37 and the start_positions arithmetic of
 * methods_for_truss_data_handle.rs:83-109 / ..beam..:96-123 / ..plate..:113-132. */
int32_t femgpu_symbolic(femgpu_t* h, int64_t* n_rows, int64_t* nnz);

/* Numeric pass (re-runnable): rotation matrices, local stiffness, R^T k R and deterministic
 * accumulation into the CSR values — what add_truss/add_beam/add_plate do per element
 * (methods_for_truss_data_handle.rs:64-123, ..beam..:77-137, ..plate..:91-212). Asynchronous on
 * the handle's stream; femgpu_synchronize() or any copy-out waits for it. */
int32_t femgpu_numeric(femgpu_t* h);
int32_t femgpu_synchronize(femgpu_t* h);

/* convenience: validate + symbolic (if stale) + numeric + synchronize */
int32_t femgpu_assemble(femgpu_t* h, int64_t* n_rows, int64_t* nnz);

/* ---- results ------------------------------------------------------------------------------ */

/* Host copy-out of the CSR on the structural pattern; any pointer may be NULL to skip it.
 * Read-side contract of the reference (methods_for_separate_stiffness_matrix.rs:223,277-278):
 * an unordered set of (row, col, value) where an absent entry and a stored 0.0 are equivalent. */
int32_t femgpu_get_csr(femgpu_t* h, int64_t* row_ptr, int32_t* col_idx, double* values);

/* Device pointers (valid until the next symbolic pass / reset / destroy) for a downstream GPU
 * consumer. Rows [row_begin, row_end) are the ones this handle owns (all rows on one GPU). */
int32_t femgpu_get_csr_device(femgpu_t* h, const int64_t** row_ptr, const int32_t** col_idx,
                              const double** values, int64_t* row_begin, int64_t* row_end);

/* The assembled matrix compacted to the entries that are != 0.0 — exactly the set the reference's position-keyed map
 * holds (its add_* loops skip exact zeros, methods_for_truss/beam/plate_data_handle.rs:115 / :129 / :204) — as CSR:
 * row_ptr [n_rows + 1], col_idx / values [count], columns ascending inside a row. The compaction runs on the device once
 * per numeric pass and stays in HBM; the call copies what is asked for (any pointer may be NULL; NULL arrays = count
 * only). For a flat plate mesh this is 36 % of the structural pattern: the cheap way to bring K to the host.
 * Multi-GPU: the rows this rank owns (after the interface exchange); all other rows are empty. */
int32_t femgpu_get_nonzero_csr(femgpu_t* h, int64_t* count, int64_t* row_ptr, int32_t* col_idx, double* values);

/* Compacts to the reference's value-dependent pattern (entries that are != 0.0) as sorted COO.
 * Call with NULL arrays to get the count. */
int32_t femgpu_get_nonzero_coo(femgpu_t* h, int64_t* count, int64_t* rows, int64_t* cols,
                               double* values);

/* FEM::get_{truss,beam,plate}_rotation_matrix_elements(number)                 fem.rs:171-202 */
int32_t femgpu_rotation_elements(femgpu_t* h, int32_t family, uint32_t number, double out[9]);

/* Test hook: the transformed element matrix (R^T k R) of one element, row-major,
 * 36 | 144 | 576 doubles — `transformed_local_stiffness_matrix` in add_truss/add_beam/add_plate. */
int32_t femgpu_element_matrix(femgpu_t* h, int32_t family, uint32_t number, double* out);

/* element -> CSR-slot scatter map of one element: for every (local row, local col) of its
 * transformed matrix the index into `values`, or -1 where the structural pattern has no slot
 * (never happens for entries the reference would store). 36 | 144 | 576 int64. */
int32_t femgpu_element_slots(femgpu_t* h, int32_t family, uint32_t number, int64_t* out);

/* ---- boundary conditions and separation of K (first consumer of the assembled matrix) ------- */

/* DOF parameters: DOFParameter of methods_for_bc_data_handle.rs:6-13 */
#define FEMGPU_DOF_X 0
#define FEMGPU_DOF_Y 1
#define FEMGPU_DOF_Z 2
#define FEMGPU_DOF_THX 3
#define FEMGPU_DOF_THY 4
#define FEMGPU_DOF_THZ 5

/* n x FEM::add_displacement(node_number, dof_parameter, value)   methods_for_bc_data_handle.rs:175-203
 * marks global index 6*node_index+dof as constrained; a second displacement on the same index is
 * the reference's "Displacement {dof:?} already applied to node {n}!". Prefix semantics. */
int32_t femgpu_add_displacement(femgpu_t* h, size_t n, const uint32_t* node_number, const int32_t* dof,
                                const double* value);
/* n x FEM::add_concentrated_load(node_number, dof_parameter, value): += into the forces vector
 *                                                               methods_for_bc_data_handle.rs:31-56 */
int32_t femgpu_add_concentrated_load(femgpu_t* h, size_t n, const uint32_t* node_number, const int32_t* dof,
                                     const double* value);

/* n x FEM::add_uniformly_distributed_line_load(beam_element_number, dof_parameter, value)
 *                                                              methods_for_bc_data_handle.rs:58-102
 * n x FEM::add_uniformly_distributed_surface_load(plate_element_number, dof_parameter, value)    :104-173
 * The nodal equivalents (Beam::convert_uniformly_distributed_line_load_to_nodal_loads structs/beam.rs:775-797,
 * Plate::convert_uniformly_distributed_surface_load_to_nodal_loads structs/plate.rs:1145-1185) are
 * evaluated on the device and added to the forces vector, per DOF in call order — across all three load kinds:
 * a concentrated load added after distributed ones is queued behind them — when the forces are next needed
 * (femgpu_get_forces, femgpu_separate_sparse). */
int32_t femgpu_add_line_load(femgpu_t* h, size_t n, const uint32_t* beam_number, const int32_t* dof,
                             const double* value);
int32_t femgpu_add_surface_load(femgpu_t* h, size_t n, const uint32_t* plate_number, const int32_t* dof,
                                const double* value);
/* forces vector [6 * nodes_number] = concentrated + distributed loads (fem.rs:19); either pointer may be NULL */
int32_t femgpu_get_forces(femgpu_t* h, double* forces, const double** forces_device);

/* FEM::separate_stiffness_matrix_sparse_iterative()   methods_for_separate_stiffness_matrix.rs:217-320
 * on the device, from the CSR values of the last numeric pass: DOFs with a zero diagonal are
 * inactive (error when one of them is constrained), constrained ones form the b-set, the rest the
 * a-set, both numbered in ascending global index; K_aa, K_ab, K_ba, K_bb hold the entries != 0.0
 * between active DOFs. The reference returns unordered triplets; here each quadrant is a CSR matrix
 * in local indices (rows and columns ascending) — the same set of triplets. Also evaluates
 * b = R_a - K_ab u_b (find_b_sparse, methods_for_global_analysis.rs:28-48).
 * nnz order: aa, ab, ba, bb. Out-pointers may be NULL. Single GPU. */
int32_t femgpu_separate_sparse(femgpu_t* h, int64_t* n_aa, int64_t* n_bb, int64_t nnz[4]);
/* FEM::separate_stiffness_matrix_direct()              methods_for_separate_stiffness_matrix.rs:63-215
 * Same classification as the sparse separation plus the direct variant's extra check (a load on a DOF without
 * stiffness: "There are no stiffness to withstand load ..."; no restraints: "There are no restraints applied!").
 * The reference makes K dense to collect k_aa_skyline and dense quadrants whose only consumer turns K_aa into the
 * compacted column form of its skyline solver (convert_k_aa_into_compacted_form, methods_for_global_analysis.rs:50-80);
 * here k_aa_skyline and (a, maxa) are built straight from the CSR quadrant on the device, and K_ab / K_ba / K_bb
 * stay available as the CSR quadrants of femgpu_get_separated_csr — or, for a caller that wants the reference's
 * SeparatedStiffnessMatrix as it is, as dense matrices through femgpu_get_separated_dense.
 * skyline_values = length of `a`. */
int32_t femgpu_separate_direct(femgpu_t* h, int64_t* n_aa, int64_t* n_bb, int64_t* skyline_values);
/* k_aa_skyline [n_aa], a [skyline_values], maxa [n_aa + 1]: column j of K_aa = a[maxa[j]] (diagonal),
 * a[maxa[j] + k] = K_aa[j - k, j] for k <= k_aa_skyline[j]. Pointers may be NULL. */
int32_t femgpu_get_skyline(femgpu_t* h, int64_t* k_aa_skyline, double* a, int64_t* maxa);
/* k_aa_indexes [n_aa] / k_bb_indexes [n_bb]: global DOF index of every local row */
int32_t femgpu_get_separated_indexes(femgpu_t* h, int64_t* k_aa_indexes, int64_t* k_bb_indexes);
/* quadrant `which` (0 aa, 1 ab, 2 ba, 3 bb): row_ptr [rows+1], col_idx / values [nnz] */
int32_t femgpu_get_separated_csr(femgpu_t* h, int32_t which, int64_t* row_ptr, int32_t* col_idx, double* values);
int32_t femgpu_get_separated_csr_device(femgpu_t* h, int32_t which, const int64_t** row_ptr,
                                        const int32_t** col_idx, const double** values);
/* Quadrant `which` as a dense row-major matrix [rows x cols] — k_aa_matrix / k_ab_matrix / k_ba_matrix / k_bb_matrix of
 * the reference's SeparatedStiffnessMatrix (structs/separated_stiffness_matrix.rs:8-16), which
 * separate_stiffness_matrix_direct returns next to k_aa_indexes, k_bb_indexes and k_aa_skyline. Densified on the device
 * from the CSR quadrant; refused with FEMGPU_ERR_LIMIT above 2^28 entries (the dense form is O((6N)^2): it exists for
 * callers of the reference signature on the model sizes the reference's own dense separation could handle). */
int32_t femgpu_get_separated_dense(femgpu_t* h, int32_t which, double* out);
/* b = R_a - K_ab u_b [n_aa]; either pointer may be NULL */
int32_t femgpu_separated_rhs(femgpu_t* h, double* b, const double** b_device);
/* device milliseconds of the last femgpu_separate_sparse() */
int32_t femgpu_last_separate_ms(femgpu_t* h, float* ms);
/* *one_pass = 1 when the last separation read K from HBM once (FEMGPU_SEP_ONE_PASS=1: tiles of 128 rows chained by a
 * scan over the tiles, outputs sized by upper bounds), 0 for count + fill (the default: faster as measured; also the
 * fallback when the bounds do not fit in memory). Same result either way. */
int32_t femgpu_last_separate_path(femgpu_t* h, int32_t* one_pass);

/* ---- global analysis and element results (downstream of the separated matrix, all in HBM) ---- */

#define FEMGPU_PCG_JACOBI 0
#define FEMGPU_PCG_BLOCK_JACOBI 1

/* FEM::find_ua_vector_iterative_pcg_jacobi_sparse / ..._pcg_block_jacobi_sparse
 *                                                        methods_for_global_analysis.rs:189-275
 * K_aa u_a = b (b = R_a - K_ab u_b of the last femgpu_separate_sparse), preconditioned conjugate
 * gradients from u_a = 0 on the device CSR; block Jacobi groups the rows of one node
 * (build_block_starts_from_k_aa_indexes, :121-147). Stops when ||r||_2 <= max(rel_tol ||b||_2, abs_tol);
 * `iterations` = search directions used. The reference's arithmetic lives in the un-vendored crate
 * iterative_solvers_smpl: parity is pinned on the reference's own test only (one iteration, u = 0.0015).
 * Deterministic (fixed reduction trees, no atomics). When max_iter search directions did not reach the stopping
 * test the call returns FEMGPU_E_SOLVER (the reference crate's behaviour in that case is not pinned: its solver is
 * un-vendored); `iterations`, femgpu_solve_info and femgpu_get_ua still report the last iterate. Block Jacobi is
 * limited to n_aa < 2^28 (FEMGPU_ERR_LIMIT). */
int32_t femgpu_solve_pcg(femgpu_t* h, int32_t preconditioner, int64_t max_iter, int64_t* iterations);
/* FEM::find_ua_vector_direct                              methods_for_global_analysis.rs:161-187
 * K_aa u_a = b by the active-column LDL^T of the skyline form femgpu_separate_direct left on the device (the
 * reference hands (a, maxa) to colsol::factorization / find_unknown, i.e. Bathe's COLSOL). One warp walks the
 * columns in COLSOL's order, its lanes share the dot products: meant for the model sizes the reference's dense
 * separation could handle. FEMGPU_E_SOLVER when a pivot is not positive. */
int32_t femgpu_solve_direct(femgpu_t* h);
/* u_a [n_aa] of the last solve; femgpu_set_ua installs the result of an external solver instead */
int32_t femgpu_get_ua(femgpu_t* h, double* u_a, const double** u_a_device);
int32_t femgpu_set_ua(femgpu_t* h, const double* u_a);
/* iterations, final ||r||_2 and device milliseconds of the last femgpu_solve_pcg (pointers may be NULL) */
int32_t femgpu_solve_info(femgpu_t* h, int64_t* iterations, double* residual, float* ms);

/* FEM::find_r_r_vector_sparse (r_r = K_ba u_a + K_bb u_b - R_b, methods_for_global_analysis.rs:100-137,
 * :334-360) followed by FEM::compose_global_analysis_result (:362-385): displacements[k_aa_indexes] = u_a,
 * forces[k_bb_indexes] = r_r. */
int32_t femgpu_global_analysis(femgpu_t* h);
/* r_r [n_bb]; either pointer may be NULL */
int32_t femgpu_get_reactions(femgpu_t* h, double* r_r, const double** r_r_device);
/* FEM::extract_global_analysis_result (:387-...): displacements / forces vectors [6 * nodes_number]
 * (row 6 * node_index + dof); either pointer may be NULL */
int32_t femgpu_get_global_result(femgpu_t* h, double* displacements, double* forces);
/* installs a displacements vector [6 * nodes_number] from elsewhere (an external solver, a test) */
int32_t femgpu_set_displacements(femgpu_t* h, const double* displacements);

/* FEM::extract_elements_analysis_result            methods_for_element_analysis.rs:27-58
 * for every element of `family`, in insertion order, from the displacements vector:
 *   truss  1 value : ForceR                                              structs/truss.rs:281-333
 *   beam  10 values: ForceR, ForceS, ForceT, MomentR, MomentS (node 1, average, node 2),
 *                    MomentT (node 1, average, node 2)                   structs/beam.rs:803-993
 *   plate  8 values: MembraneForceR, MembraneForceS, MembraneForceRS, BendingMomentR, BendingMomentS,
 *                    BendingMomentRS, ShearForceRT, ShearForceST        structs/plate.rs:1196-1409
 * out [n_elements][values] (element-major); either pointer may be NULL. */
int32_t femgpu_element_results(femgpu_t* h, int32_t family, double* out, const double** out_device);

/* ---- instrumentation ---------------------------------------------------------------------- */

/* kernels launched by this handle since creation / since the last call with reset != 0 */
int32_t femgpu_launch_count(femgpu_t* h, int32_t reset, uint64_t* launches);
/* device milliseconds of the last femgpu_numeric() (CUDA events on the handle's stream), split as
 * [0] total, [1] element-record kernels, [2] assembly kernel, [3] interface exchange */
int32_t femgpu_last_numeric_ms(femgpu_t* h, float out[4]);
/* same for the pass issued `passes_back` passes before the last one (0 = last; the handle keeps
 * the events of the most recent 64 passes), so a timed loop needs no sync between passes */
int32_t femgpu_numeric_ms_history(femgpu_t* h, uint32_t passes_back, float out[4]);
/* sum of the assemble_kernel launch durations of that pass (the pass may cut the slabs into several ranges, each one
 * launch, while the element records of the next range are computed on a second stream) and the number of launches */
int32_t femgpu_numeric_kernel_ms(femgpu_t* h, uint32_t passes_back, float* assemble_ms, int32_t* launches);
/* bytes of device memory held by the handle */
int32_t femgpu_device_bytes(const femgpu_t* h, uint64_t* bytes);
/* the CUDA stream handle (cudaStream_t) work is issued on, for callers that time with events */
int32_t femgpu_stream(femgpu_t* h, void** stream);

/* measured FP64 FMA throughput of the handle's device in TFLOP/s (micro-benchmark, ~10 ms): the denominator of
 * the FP64-pipe fraction bench.py reports next to the HBM roofline (no reference counterpart) */
int32_t femgpu_fp64_fma_peak(femgpu_t* h, double* tflops);

/* ---- multi-GPU (one process per GPU) ------------------------------------------------------- */

/* Joins `world` handles (one per process/GPU) into one assembly. `nccl_unique_id` is the 128-byte
 * ncclUniqueId created by rank 0 (femgpu_dist_unique_id) and broadcast by the host. Must be called
 * before femgpu_symbolic(). Rank g owns the contiguous node-index range given by
 * femgpu_dist_set_ownership(); each rank is given the elements whose lowest-index node it owns and either every
 * node or — femgpu_dist_set_node_window() — just the window of nodes those elements touch. femgpu_symbolic() is collective (NCCL all-gathers and a key exchange; it also maps
 * the neighbours' receive windows with CUDA IPC). femgpu_numeric() is NOT a host-level collective: contributions
 * to rows owned by another rank are summed locally, stored into the owner's HBM over NVLink by the sender's pack
 * kernel and added by the owner in (source rank, slot) order once the sender's flag is up; every rank must run the
 * same number of passes, and a rank that waits in vain for 20 s (FEMGPU_P2P_TIMEOUT_MS) reports FEMGPU_ERR_NCCL
 * "timed out" from femgpu_synchronize / femgpu_get_csr instead of hanging. Without peer access the exchange is one
 * grouped ncclSend/ncclRecv per neighbour. No rank may still be running passes when a joined handle is destroyed. */
int32_t femgpu_dist_unique_id(uint8_t out[128]);
int32_t femgpu_dist_init(femgpu_t* h, int32_t rank, int32_t world, const uint8_t nccl_unique_id[128]);
int32_t femgpu_dist_set_ownership(femgpu_t* h, uint32_t node_index_begin, uint32_t node_index_end);
/* Node window of a rank: the nodes added after this call get the global insertion indices first_node_index,
 * first_node_index + 1, ... instead of 0, 1, ... — a rank then only has to be given the nodes of its own rows plus the
 * halo its elements touch (host ingest and node memory O(nodes / world) instead of O(nodes)); the global row of a
 * node stays 6 * index + dof (structs/node.rs:8) and nodes_number stays the size of the WHOLE model. An element that
 * names a node outside the window fails like one naming an unknown node. Call after femgpu_create / femgpu_reset,
 * before the first femgpu_add_nodes. */
int32_t femgpu_dist_set_node_window(femgpu_t* h, uint32_t first_node_index);
/* interface traffic of the last numeric pass: bytes sent / received by this rank */
int32_t femgpu_dist_last_exchange_bytes(femgpu_t* h, uint64_t* sent, uint64_t* received);
/* how the last symbolic pass set the exchange up: *p2p = 1 peer windows over NVLink, 0 ncclSend/ncclRecv;
 * *passes = numeric passes issued since (pointers may be NULL) */
int32_t femgpu_dist_info(femgpu_t* h, int32_t* p2p, uint64_t* passes);

#ifdef __cplusplus
}
#endif
#endif /* FEMGPU_H */
