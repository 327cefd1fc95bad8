//! `FEM` with the method signatures of `finite_element_method::FEM<f64>` for the stiffness-assembly
//! path (fem.rs:34,155,171-202; methods_for_{node,truss,beam,plate}_data_handle.rs), backed by the
//! femgpu C ABI (include/femgpu.h). `Result<(), String>` carries the reference's own messages.
//! NOT compiled here (no Rust toolchain in the build image).
use std::ffi::CStr;
use std::os::raw::c_char;

#[repr(C)]
pub struct FemGpu { _private: [u8; 0] }

extern "C" {
    fn femgpu_create(out: *mut *mut FemGpu, rel_tol: f64, abs_tol: f64, nodes_number: u32, device: i32) -> i32;
    fn femgpu_reset(h: *mut FemGpu, nodes_number: u32) -> i32;
    fn femgpu_destroy(h: *mut FemGpu);
    fn femgpu_last_error(h: *const FemGpu) -> *const c_char;
    fn femgpu_add_nodes(h: *mut FemGpu, n: usize, number: *const u32, x: *const f64, y: *const f64, z: *const f64) -> i32;
    fn femgpu_add_truss(h: *mut FemGpu, n: usize, number: *const u32, n1: *const u32, n2: *const u32,
                        e: *const f64, a: *const f64, a2: *const f64) -> i32;
    fn femgpu_add_beam(h: *mut FemGpu, n: usize, number: *const u32, n1: *const u32, n2: *const u32,
                       e: *const f64, nu: *const f64, a: *const f64, i11: *const f64, i22: *const f64,
                       i12: *const f64, it: *const f64, ks: *const f64, axis1: *const f64) -> i32;
    fn femgpu_add_plate(h: *mut FemGpu, n: usize, number: *const u32, n1: *const u32, n2: *const u32,
                        n3: *const u32, n4: *const u32, e: *const f64, nu: *const f64, t: *const f64,
                        ks: *const f64) -> i32;
    fn femgpu_validate(h: *mut FemGpu, family: *mut i32, number: *mut u32, code: *mut i32) -> i32;
    fn femgpu_assemble(h: *mut FemGpu, n_rows: *mut i64, nnz: *mut i64) -> i32;
    fn femgpu_get_csr(h: *mut FemGpu, row_ptr: *mut i64, col_idx: *mut i32, values: *mut f64) -> i32;
    fn femgpu_get_nonzero_csr(h: *mut FemGpu, count: *mut i64, row_ptr: *mut i64, col_idx: *mut i32, values: *mut f64) -> i32;
    fn femgpu_rotation_elements(h: *mut FemGpu, family: i32, number: u32, out: *mut f64) -> i32;
    fn femgpu_add_displacement(h: *mut FemGpu, n: usize, node: *const u32, dof: *const i32, value: *const f64) -> i32;
    fn femgpu_add_concentrated_load(h: *mut FemGpu, n: usize, node: *const u32, dof: *const i32, value: *const f64) -> i32;
    fn femgpu_add_line_load(h: *mut FemGpu, n: usize, beam: *const u32, dof: *const i32, value: *const f64) -> i32;
    fn femgpu_add_surface_load(h: *mut FemGpu, n: usize, plate: *const u32, dof: *const i32, value: *const f64) -> i32;
    fn femgpu_separate_sparse(h: *mut FemGpu, n_aa: *mut i64, n_bb: *mut i64, nnz: *mut i64) -> i32;
    fn femgpu_get_separated_indexes(h: *mut FemGpu, k_aa_indexes: *mut i64, k_bb_indexes: *mut i64) -> i32;
    fn femgpu_get_separated_csr(h: *mut FemGpu, which: i32, row_ptr: *mut i64, col_idx: *mut i32, values: *mut f64) -> i32;
    fn femgpu_separated_rhs(h: *mut FemGpu, b: *mut f64, b_device: *mut *const f64) -> i32;
    fn femgpu_get_separated_dense(h: *mut FemGpu, which: i32, out: *mut f64) -> i32;
    fn femgpu_last_separate_path(h: *mut FemGpu, one_pass: *mut i32) -> i32;
    fn femgpu_separate_direct(h: *mut FemGpu, n_aa: *mut i64, n_bb: *mut i64, skyline_values: *mut i64) -> i32;
    fn femgpu_solve_direct(h: *mut FemGpu) -> i32;
    fn femgpu_get_skyline(h: *mut FemGpu, k_aa_skyline: *mut i64, a: *mut f64, maxa: *mut i64) -> i32;
    fn femgpu_counts(h: *const FemGpu, nodes: *mut u64, truss: *mut u64, beam: *mut u64, plate: *mut u64) -> i32;
    fn femgpu_get_numbers(h: *const FemGpu, family: i32, out: *mut u32) -> i32;
    fn femgpu_solve_pcg(h: *mut FemGpu, preconditioner: i32, max_iter: i64, iterations: *mut i64) -> i32;
    fn femgpu_get_ua(h: *mut FemGpu, u_a: *mut f64, u_a_device: *mut *const f64) -> i32;
    fn femgpu_global_analysis(h: *mut FemGpu) -> i32;
    fn femgpu_get_reactions(h: *mut FemGpu, r_r: *mut f64, r_r_device: *mut *const f64) -> i32;
    fn femgpu_get_global_result(h: *mut FemGpu, displacements: *mut f64, forces: *mut f64) -> i32;
    fn femgpu_element_results(h: *mut FemGpu, family: i32, out: *mut f64, out_device: *mut *const f64) -> i32;
}

/// methods_for_element_analysis.rs:5-21
#[derive(Debug, Clone, Copy, PartialEq, Eq, PartialOrd, Ord)]
pub enum ElementForceComponent {
    ForceR, ForceS, ForceT, MembraneForceR, MembraneForceS, MembraneForceRS, ShearForceRT, ShearForceST,
    MomentR, MomentS, MomentT, BendingMomentR, BendingMomentS, BendingMomentRS,
}
use ElementForceComponent::*;
/// component of every value femgpu_element_results returns per element (truss.rs:325-328, beam.rs:967-987,
/// plate.rs:1368-1401)
const COMPONENTS: [&[ElementForceComponent]; 3] = [
    &[ForceR],
    &[ForceR, ForceS, ForceT, MomentR, MomentS, MomentS, MomentS, MomentT, MomentT, MomentT],
    &[MembraneForceR, MembraneForceS, MembraneForceRS, BendingMomentR, BendingMomentS, BendingMomentRS,
      ShearForceRT, ShearForceST],
];

/// methods_for_bc_data_handle.rs:6-13
#[derive(Debug, Clone, Copy, PartialEq, Eq, PartialOrd, Ord)]
pub enum DOFParameter { X = 0, Y = 1, Z = 2, ThX = 3, ThY = 4, ThZ = 5 }

/// structs/separated_stiffness_matrix_sparse.rs, with each quadrant as CSR in local indices
/// (row_ptr, col_idx, values) instead of an unordered triplet list, plus b = R_a - K_ab u_b.
pub struct SeparatedStiffnessMatrixSparse {
    pub k_aa_indexes: Vec<i64>,
    pub k_bb_indexes: Vec<i64>,
    pub quadrants: [(Vec<i64>, Vec<i32>, Vec<f64>); 4], // aa, ab, ba, bb
    pub b: Vec<f64>,
}

pub struct FEM { h: *mut FemGpu, nodes_number: u32 }

impl FEM {
    fn check(&self, st: i32) -> Result<(), String> {
        if st == 0 { return Ok(()); }
        Err(unsafe { CStr::from_ptr(femgpu_last_error(self.h)) }.to_string_lossy().into_owned())
    }
    /// fem.rs:34
    pub fn create(rel_tol: f64, abs_tol: f64, nodes_number: u32) -> Self {
        let mut h = std::ptr::null_mut();
        let st = unsafe { femgpu_create(&mut h, rel_tol, abs_tol, nodes_number, 0) };
        assert!(st == 0, "femgpu_create failed: no CUDA device (there is no CPU fallback)");
        FEM { h, nodes_number }
    }
    /// fem.rs:155
    pub fn reset(&mut self, nodes_number: u32) { unsafe { femgpu_reset(self.h, nodes_number); } self.nodes_number = nodes_number; }
    /// methods_for_node_data_handle.rs:66
    pub fn add_node(&mut self, number: u32, x: f64, y: f64, z: f64) -> Result<(), String> {
        self.check(unsafe { femgpu_add_nodes(self.h, 1, &number, &x, &y, &z) })
    }
    /// methods_for_truss_data_handle.rs:49
    pub fn add_truss(&mut self, number: u32, node_1_number: u32, node_2_number: u32, young_modulus: f64,
                     area: f64, optional_area_2: Option<f64>) -> Result<(), String> {
        let a2 = optional_area_2.unwrap_or(f64::NAN);
        self.check(unsafe { femgpu_add_truss(self.h, 1, &number, &node_1_number, &node_2_number,
                                             &young_modulus, &area, &a2) })?;
        self.check(unsafe { femgpu_validate(self.h, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut()) })
    }
    /// methods_for_beam_data_handle.rs:49
    #[allow(clippy::too_many_arguments)]
    pub fn add_beam(&mut self, number: u32, node_1_number: u32, node_2_number: u32, young_modulus: f64,
                    poisson_ratio: f64, area: f64, i11: f64, i22: f64, i12: f64, it: f64, shear_factor: f64,
                    local_axis_1_direction: [f64; 3]) -> Result<(), String> {
        self.check(unsafe { femgpu_add_beam(self.h, 1, &number, &node_1_number, &node_2_number, &young_modulus,
                                            &poisson_ratio, &area, &i11, &i22, &i12, &it, &shear_factor,
                                            local_axis_1_direction.as_ptr()) })?;
        self.check(unsafe { femgpu_validate(self.h, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut()) })
    }
    /// methods_for_plate_data_handle.rs:62
    #[allow(clippy::too_many_arguments)]
    pub fn add_plate(&mut self, number: u32, node_1_number: u32, node_2_number: u32, node_3_number: u32,
                     node_4_number: u32, young_modulus: f64, poisson_ratio: f64, thickness: f64,
                     shear_factor: f64) -> Result<(), String> {
        self.check(unsafe { femgpu_add_plate(self.h, 1, &number, &node_1_number, &node_2_number, &node_3_number,
                                             &node_4_number, &young_modulus, &poisson_ratio, &thickness,
                                             &shear_factor) })?;
        self.check(unsafe { femgpu_validate(self.h, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut()) })
    }
    /// Bulk load: struct-of-arrays slices, prefix-atomic (see include/femgpu.h).
    pub fn add_plates(&mut self, number: &[u32], n1: &[u32], n2: &[u32], n3: &[u32], n4: &[u32], e: &[f64],
                      nu: &[f64], t: &[f64], ks: &[f64]) -> Result<(), String> {
        // the C ABI takes ONE count for all nine arrays: a shorter slice would be read past its end
        let n = number.len();
        if [n1.len(), n2.len(), n3.len(), n4.len(), e.len(), nu.len(), t.len(), ks.len()].iter().any(|&l| l != n) {
            return Err(format!("add_plates: every slice must hold {} entries", n));
        }
        self.check(unsafe { femgpu_add_plate(self.h, number.len(), number.as_ptr(), n1.as_ptr(), n2.as_ptr(),
                                             n3.as_ptr(), n4.as_ptr(), e.as_ptr(), nu.as_ptr(), t.as_ptr(), ks.as_ptr()) })
    }
    /// fem.rs:171 / :182 / :193 (family 0 = truss, 1 = beam, 2 = plate)
    pub fn get_rotation_matrix_elements(&self, family: i32, number: u32) -> Result<[f64; 9], String> {
        let mut out = [0f64; 9];
        self.check(unsafe { femgpu_rotation_elements(self.h, family, number, out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// The entries != 0.0 of the assembled matrix — exactly what `self.stiffness_matrix` holds in the reference, whose
    /// add_* skip exact zeros (methods_for_truss/beam/plate_data_handle.rs:115/129/204) — as CSR, compacted on the device.
    pub fn assemble_nonzero_csr(&mut self) -> Result<(Vec<i64>, Vec<i32>, Vec<f64>), String> {
        let (mut n_rows, mut nnz) = (0i64, 0i64);
        self.check(unsafe { femgpu_assemble(self.h, &mut n_rows, &mut nnz) })?;
        let mut count = 0i64;
        self.check(unsafe { femgpu_get_nonzero_csr(self.h, &mut count, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut()) })?;
        let mut rp = vec![0i64; n_rows as usize + 1];
        let mut ci = vec![0i32; count as usize];
        let mut v = vec![0f64; count as usize];
        self.check(unsafe { femgpu_get_nonzero_csr(self.h, &mut count, rp.as_mut_ptr(), ci.as_mut_ptr(), v.as_mut_ptr()) })?;
        Ok((rp, ci, v))
    }
    /// The assembled global stiffness matrix as CSR on the structural pattern (what
    /// `self.stiffness_matrix` holds in the reference, fem.rs:17). Stored zeros == absent entries.
    pub fn assemble_csr(&mut self) -> Result<(Vec<i64>, Vec<i32>, Vec<f64>), String> {
        let (mut n_rows, mut nnz) = (0i64, 0i64);
        self.check(unsafe { femgpu_assemble(self.h, &mut n_rows, &mut nnz) })?;
        let mut rp = vec![0i64; n_rows as usize + 1];
        let mut ci = vec![0i32; nnz as usize];
        let mut v = vec![0f64; nnz as usize];
        self.check(unsafe { femgpu_get_csr(self.h, rp.as_mut_ptr(), ci.as_mut_ptr(), v.as_mut_ptr()) })?;
        Ok((rp, ci, v))
    }
}

impl FEM {
    /// methods_for_bc_data_handle.rs:175
    pub fn add_displacement(&mut self, node_number: u32, dof_parameter: DOFParameter, value: f64) -> Result<(), String> {
        let dof = dof_parameter as i32;
        self.check(unsafe { femgpu_add_displacement(self.h, 1, &node_number, &dof, &value) })
    }
    /// methods_for_bc_data_handle.rs:31
    pub fn add_concentrated_load(&mut self, node_number: u32, dof_parameter: DOFParameter, value: f64) -> Result<(), String> {
        let dof = dof_parameter as i32;
        self.check(unsafe { femgpu_add_concentrated_load(self.h, 1, &node_number, &dof, &value) })
    }
    /// methods_for_bc_data_handle.rs:58
    pub fn add_uniformly_distributed_line_load(&mut self, beam_element_number: u32, dof_parameter: DOFParameter,
                                               value: f64) -> Result<(), String> {
        let dof = dof_parameter as i32;
        self.check(unsafe { femgpu_add_line_load(self.h, 1, &beam_element_number, &dof, &value) })
    }
    /// methods_for_bc_data_handle.rs:104
    pub fn add_uniformly_distributed_surface_load(&mut self, plate_element_number: u32, dof_parameter: DOFParameter,
                                                  value: f64) -> Result<(), String> {
        let dof = dof_parameter as i32;
        self.check(unsafe { femgpu_add_surface_load(self.h, 1, &plate_element_number, &dof, &value) })
    }
    /// methods_for_separate_stiffness_matrix.rs:217 — runs on the device on the assembled CSR values
    pub fn separate_stiffness_matrix_sparse_iterative(&mut self) -> Result<SeparatedStiffnessMatrixSparse, String> {
        let (mut n_aa, mut n_bb, mut nnz) = (0i64, 0i64, [0i64; 4]);
        self.check(unsafe { femgpu_separate_sparse(self.h, &mut n_aa, &mut n_bb, nnz.as_mut_ptr()) })?;
        let mut ia = vec![0i64; n_aa as usize];
        let mut ib = vec![0i64; n_bb as usize];
        self.check(unsafe { femgpu_get_separated_indexes(self.h, ia.as_mut_ptr(), ib.as_mut_ptr()) })?;
        let mut quadrants: [(Vec<i64>, Vec<i32>, Vec<f64>); 4] = Default::default();
        for q in 0..4 {
            let rows = if q < 2 { n_aa } else { n_bb } as usize;
            let (mut rp, mut ci, mut v) = (vec![0i64; rows + 1], vec![0i32; nnz[q] as usize], vec![0f64; nnz[q] as usize]);
            self.check(unsafe { femgpu_get_separated_csr(self.h, q as i32, rp.as_mut_ptr(), ci.as_mut_ptr(), v.as_mut_ptr()) })?;
            quadrants[q] = (rp, ci, v);
        }
        let mut b = vec![0f64; n_aa as usize];
        self.check(unsafe { femgpu_separated_rhs(self.h, b.as_mut_ptr(), std::ptr::null_mut()) })?;
        Ok(SeparatedStiffnessMatrixSparse { k_aa_indexes: ia, k_bb_indexes: ib, quadrants, b })
    }
}

impl FEM {
    /// diagnostics: did the last separation take the opt-in one-pass kernel (FEMGPU_SEP_ONE_PASS=1)?
    pub fn last_separation_read_k_once(&mut self) -> Result<bool, String> {
        let mut flag = 0i32;
        self.check(unsafe { femgpu_last_separate_path(self.h, &mut flag) })?;
        Ok(flag != 0)
    }

    /// methods_for_separate_stiffness_matrix.rs:63 without the dense detour: (k_aa_indexes, k_bb_indexes,
    /// k_aa_skyline, a, maxa) — K_aa already in the compacted column form colsol::factorization takes
    /// (methods_for_global_analysis.rs:50-80, :183)
    pub fn separate_stiffness_matrix_direct(&mut self) -> Result<(Vec<i64>, Vec<i64>, Vec<i64>, Vec<f64>, Vec<i64>), String> {
        let (mut n_aa, mut n_bb, mut n_val) = (0i64, 0i64, 0i64);
        self.check(unsafe { femgpu_separate_direct(self.h, &mut n_aa, &mut n_bb, &mut n_val) })?;
        let (mut ia, mut ib) = (vec![0i64; n_aa as usize], vec![0i64; n_bb as usize]);
        self.check(unsafe { femgpu_get_separated_indexes(self.h, ia.as_mut_ptr(), ib.as_mut_ptr()) })?;
        let (mut sky, mut a, mut maxa) = (vec![0i64; n_aa as usize], vec![0f64; n_val as usize], vec![0i64; n_aa as usize + 1]);
        self.check(unsafe { femgpu_get_skyline(self.h, sky.as_mut_ptr(), a.as_mut_ptr(), maxa.as_mut_ptr()) })?;
        Ok((ia, ib, sky, a, maxa))
    }
    /// methods_for_global_analysis.rs:161 — skyline LDL^T on the form separate_stiffness_matrix_direct left on the device
    pub fn find_ua_vector_direct(&mut self, n_aa: usize) -> Result<Vec<f64>, String> {
        self.check(unsafe { femgpu_solve_direct(self.h) })?;
        let mut u_a = vec![0f64; n_aa];
        self.check(unsafe { femgpu_get_ua(self.h, u_a.as_mut_ptr(), std::ptr::null_mut()) })?;
        Ok(u_a)
    }
    /// methods_for_global_analysis.rs:189 / :235 — K_aa, r_a and u_b never left the device, so the three
    /// arguments of the reference collapse into the handle. Returns (u_a, iterations).
    pub fn find_ua_vector_iterative_pcg_jacobi_sparse(&mut self, n_aa: usize, max_iter: usize) -> Result<(Vec<f64>, usize), String> {
        self.solve(0, n_aa, max_iter)
    }
    pub fn find_ua_vector_iterative_pcg_block_jacobi_sparse(&mut self, n_aa: usize, max_iter: usize) -> Result<(Vec<f64>, usize), String> {
        self.solve(1, n_aa, max_iter)
    }
    fn solve(&mut self, preconditioner: i32, n_aa: usize, max_iter: usize) -> Result<(Vec<f64>, usize), String> {
        let mut it = 0i64;
        self.check(unsafe { femgpu_solve_pcg(self.h, preconditioner, max_iter as i64, &mut it) })?;
        let mut u_a = vec![0f64; n_aa];
        self.check(unsafe { femgpu_get_ua(self.h, u_a.as_mut_ptr(), std::ptr::null_mut()) })?;
        Ok((u_a, it as usize))
    }
    /// methods_for_global_analysis.rs:334 (find_r_r_vector_sparse) + :362 (compose_global_analysis_result)
    pub fn find_r_r_vector_sparse(&mut self, n_bb: usize) -> Result<Vec<f64>, String> {
        self.check(unsafe { femgpu_global_analysis(self.h) })?;
        let mut r_r = vec![0f64; n_bb];
        self.check(unsafe { femgpu_get_reactions(self.h, r_r.as_mut_ptr(), std::ptr::null_mut()) })?;
        Ok(r_r)
    }
    /// methods_for_global_analysis.rs:387 — (node_number, dof, displacement, load) per node and DOF
    pub fn extract_global_analysis_result(&mut self) -> Result<Vec<(u32, DOFParameter, f64, f64)>, String> {
        let mut n = [0u64; 4];
        self.check(unsafe { femgpu_counts(self.h, &mut n[0], &mut n[1], &mut n[2], &mut n[3]) })?;
        let mut numbers = vec![0u32; n[0] as usize];
        self.check(unsafe { femgpu_get_numbers(self.h, -1, numbers.as_mut_ptr()) })?;
        let rows = 6 * self.nodes_number as usize;
        let (mut d, mut f) = (vec![0f64; rows], vec![0f64; rows]);
        self.check(unsafe { femgpu_get_global_result(self.h, d.as_mut_ptr(), f.as_mut_ptr()) })?;
        const DOFS: [DOFParameter; 6] = [DOFParameter::X, DOFParameter::Y, DOFParameter::Z, DOFParameter::ThX,
                                         DOFParameter::ThY, DOFParameter::ThZ];
        let mut out = Vec::with_capacity(6 * numbers.len());
        for (i, number) in numbers.iter().enumerate() {
            for k in 0..6 { out.push((*number, DOFS[k], d[6 * i + k], f[6 * i + k])); }
        }
        Ok(out)
    }
    /// methods_for_element_analysis.rs:27 — trusses, beams, plates; each family in insertion order
    pub fn extract_elements_analysis_result(&mut self) -> Result<Vec<(u32, Vec<(ElementForceComponent, f64)>)>, String> {
        let mut n = [0u64; 4];
        self.check(unsafe { femgpu_counts(self.h, &mut n[0], &mut n[1], &mut n[2], &mut n[3]) })?;
        let mut out = Vec::new();
        for family in 0..3usize {
            let (count, comps) = (n[1 + family] as usize, COMPONENTS[family]);
            let mut numbers = vec![0u32; count];
            let mut vals = vec![0f64; count * comps.len()];
            self.check(unsafe { femgpu_get_numbers(self.h, family as i32, numbers.as_mut_ptr()) })?;
            self.check(unsafe { femgpu_element_results(self.h, family as i32, vals.as_mut_ptr(), std::ptr::null_mut()) })?;
            for (e, number) in numbers.iter().enumerate() {
                out.push((*number, comps.iter().cloned().zip(vals[e * comps.len()..(e + 1) * comps.len()].iter().cloned()).collect()));
            }
        }
        Ok(out)
    }
}

impl Drop for FEM {
    fn drop(&mut self) { unsafe { femgpu_destroy(self.h) } }
}
