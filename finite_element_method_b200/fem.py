"""Host-side mirror of the reference's `FEM<V>` API for the stiffness-assembly path.

Same method names, argument order and error texts as the Rust crate
(/root/reference/src/fem/fem.rs:34-65,155-202, methods_for_node_data_handle.rs:66-78,
methods_for_truss_data_handle.rs:49-128, methods_for_beam_data_handle.rs:49-142,
methods_for_plate_data_handle.rs:62-217), where `Result<(), String>` becomes "returns None or
raises FemError(str)". Everything is a thin call into the C ABI of include/femgpu.h; all numerics
run in libfemgpu.so on the GPU. `V = f64` only.

Differences a caller can observe, by design:
  * add_* record the element and validate it on the device right away, but the stiffness numbers
    are produced by `assemble()` (symbolic once, numeric re-runnable) instead of inside each call;
  * bulk variants (`add_nodes`, `add_trusses`, `add_beams`, `add_plates`) take arrays and are
    prefix-atomic: everything before the first failing element is kept.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

TRUSS, BEAM, PLATE = 0, 1, 2

# ElementForceComponent (methods_for_element_analysis.rs:5-21) of every value extract_elements_analysis_result
# returns per element: truss.rs:325-328, beam.rs:967-987, plate.rs:1368-1401
ELEMENT_RESULT_COMPONENTS = {
    TRUSS: ("ForceR",),
    BEAM: ("ForceR", "ForceS", "ForceT", "MomentR", "MomentS", "MomentS", "MomentS", "MomentT", "MomentT", "MomentT"),
    PLATE: ("MembraneForceR", "MembraneForceS", "MembraneForceRS", "BendingMomentR", "BendingMomentS",
            "BendingMomentRS", "ShearForceRT", "ShearForceST"),
}


class DOFParameter:
    """methods_for_bc_data_handle.rs:6-13"""
    X, Y, Z, ThX, ThY, ThZ = range(6)


class SeparatedStiffnessMatrixSparse:
    """structs/separated_stiffness_matrix_sparse.rs. The four quadrants are CSR matrices in local
    indices (row_ptr, col_idx, values); `triplets(q)` expands one to the reference's (i, j, value) list."""

    def __init__(self, k_aa_indexes, k_bb_indexes, quadrants, rhs, device_ms):
        self.k_aa_indexes, self.k_bb_indexes = k_aa_indexes, k_bb_indexes
        self.n_aa, self.n_bb = len(k_aa_indexes), len(k_bb_indexes)
        self.k_aa, self.k_ab, self.k_ba, self.k_bb = quadrants
        self.b = rhs                      # R_a - K_ab u_b (find_b_sparse)
        self.device_ms = device_ms

    def triplets(self, quadrant):
        rp, ci, v = quadrant
        rows = np.repeat(np.arange(len(rp) - 1, dtype=np.int64), np.diff(rp))
        return rows, ci.astype(np.int64), v


class SeparatedStiffnessMatrix:
    """structs/separated_stiffness_matrix.rs:8-65 — what `separate_stiffness_matrix_direct` returns in the reference, with
    its getters. The dense quadrants are fetched from the device on first use (`femgpu_get_separated_dense`); `a`, `maxa`
    are K_aa in the compacted column form the skyline solver consumes (convert_k_aa_into_compacted_form). Unpacks like
    the tuple (k_aa_indexes, k_bb_indexes, k_aa_skyline, a, maxa) earlier versions returned."""

    def __init__(self, fem, k_aa_indexes, k_bb_indexes, k_aa_skyline, a, maxa):
        self._fem = fem
        self.k_aa_indexes, self.k_bb_indexes, self.k_aa_skyline = k_aa_indexes, k_bb_indexes, k_aa_skyline
        self.a, self.maxa = a, maxa
        self._dense = {}

    def __iter__(self):
        return iter((self.k_aa_indexes, self.k_bb_indexes, self.k_aa_skyline, self.a, self.maxa))

    def get_k_aa_indexes(self):
        return self.k_aa_indexes

    def get_k_bb_indexes(self):
        return self.k_bb_indexes

    def get_k_aa_skyline(self):
        return self.k_aa_skyline

    def _quadrant(self, which):
        if which not in self._dense:
            self._dense[which] = self._fem.separated_dense(which)
        return self._dense[which]

    def get_k_aa_matrix(self):
        return self._quadrant(0)

    def get_k_ab_matrix(self):
        return self._quadrant(1)

    def get_k_ba_matrix(self):
        return self._quadrant(2)

    def get_k_bb_matrix(self):
        return self._quadrant(3)


class FemError(Exception):
    """`Err(String)` of the reference. `.code` is the FEMGPU_E_* / FEMGPU_ERR_* status."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = int(code)
        self.message = message


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


def _same_length(what, n, **arrays):
    """The C ABI takes one count for all arrays of a batch: a short array would be read past its end."""
    for name, a in arrays.items():
        if a is not None and len(a) != n:
            raise ValueError(f"{what}: `{name}` has {len(a)} entries, expected {n}")


class FEM:
    """`FEM::create(rel_tol, abs_tol, nodes_number)` — fem.rs:34."""

    def __init__(self, rel_tol: float, abs_tol: float, nodes_number: int, device: int = 0):
        self._L = _lib.load()
        self._h = _lib.H()
        st = self._L.femgpu_create(C.byref(self._h), rel_tol, abs_tol, nodes_number, device)
        if st:
            msg = self._L.femgpu_last_error(None).decode()
            self._h = None
            raise FemError(st, msg)
        self.rel_tol, self.abs_tol, self.nodes_number, self.device = rel_tol, abs_tol, nodes_number, device

    @classmethod
    def create(cls, rel_tol: float, abs_tol: float, nodes_number: int, device: int = 0) -> "FEM":
        return cls(rel_tol, abs_tol, nodes_number, device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.femgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st:
            raise FemError(st, self._L.femgpu_last_error(self._h).decode())

    # ------------------------------------------------------------------ reference API
    def reset(self, nodes_number: int) -> None:
        """fem.rs:155"""
        self._check(self._L.femgpu_reset(self._h, nodes_number))
        self.nodes_number = nodes_number

    def add_node(self, number: int, x: float, y: float, z: float) -> None:
        """methods_for_node_data_handle.rs:66"""
        self.add_nodes([number], [x], [y], [z])

    def add_truss(self, number, node_1_number, node_2_number, young_modulus, area, optional_area_2=None) -> None:
        """methods_for_truss_data_handle.rs:49"""
        self.add_trusses([number], [node_1_number], [node_2_number], [young_modulus], [area],
                         None if optional_area_2 is None else [optional_area_2])
        self.validate()

    def add_beam(self, number, node_1_number, node_2_number, young_modulus, poisson_ratio, area, i11, i22, i12,
                 it, shear_factor, local_axis_1_direction) -> None:
        """methods_for_beam_data_handle.rs:49"""
        ax = np.asarray(local_axis_1_direction, np.float64).reshape(3, 1)
        self.add_beams([number], [node_1_number], [node_2_number], [young_modulus], [poisson_ratio], [area],
                       [i11], [i22], [i12], [it], [shear_factor], ax)
        self.validate()

    def add_plate(self, number, node_1_number, node_2_number, node_3_number, node_4_number, young_modulus,
                  poisson_ratio, thickness, shear_factor) -> None:
        """methods_for_plate_data_handle.rs:62"""
        self.add_plates([number], [node_1_number], [node_2_number], [node_3_number], [node_4_number],
                        [young_modulus], [poisson_ratio], [thickness], [shear_factor])
        self.validate()

    def get_truss_rotation_matrix_elements(self, number: int):
        """fem.rs:171"""
        return self._rotation(TRUSS, number)

    def get_beam_rotation_matrix_elements(self, number: int):
        """fem.rs:182"""
        return self._rotation(BEAM, number)

    def get_plate_rotation_matrix_elements(self, number: int):
        """fem.rs:193"""
        return self._rotation(PLATE, number)

    def _rotation(self, family, number):
        out = np.zeros(9)
        self._check(self._L.femgpu_rotation_elements(self._h, family, number, _p(out, _lib.dp)))
        return out

    # ------------------------------------------------------------------ bulk API
    def add_nodes(self, number, x, y, z) -> None:
        number, x, y, z = _u32(number), _f64(x), _f64(y), _f64(z)
        _same_length("add_nodes", len(number), x=x, y=y, z=z)
        self._check(self._L.femgpu_add_nodes(self._h, len(number), _p(number, _lib.u32p), _p(x, _lib.dp),
                                             _p(y, _lib.dp), _p(z, _lib.dp)))

    def add_trusses(self, number, node_1, node_2, young_modulus, area, area_2=None) -> None:
        number, n1, n2 = _u32(number), _u32(node_1), _u32(node_2)
        E, A = _f64(young_modulus), _f64(area)
        A2 = None
        if area_2 is None:
            a2p = C.cast(None, _lib.dp)
        else:
            A2 = _f64([np.nan if v is None else v for v in area_2] if isinstance(area_2, (list, tuple)) else area_2)
            a2p = _p(A2, _lib.dp)
        _same_length("add_trusses", len(number), node_1=n1, node_2=n2, young_modulus=E, area=A, area_2=A2)
        self._check(self._L.femgpu_add_truss(self._h, len(number), _p(number, _lib.u32p), _p(n1, _lib.u32p),
                                             _p(n2, _lib.u32p), _p(E, _lib.dp), _p(A, _lib.dp), a2p))

    def add_beams(self, number, node_1, node_2, young_modulus, poisson_ratio, area, i11, i22, i12, it,
                  shear_factor, local_axis_1) -> None:
        """local_axis_1: array of shape (3, n) — struct of arrays."""
        number, n1, n2 = _u32(number), _u32(node_1), _u32(node_2)
        arrs = [_f64(v) for v in (young_modulus, poisson_ratio, area, i11, i22, i12, it, shear_factor)]
        ax = _f64(np.asarray(local_axis_1, np.float64).reshape(3, -1))
        if ax.shape[1] != len(number):
            raise ValueError("local_axis_1 must have shape (3, n)")
        _same_length("add_beams", len(number), node_1=n1, node_2=n2,
                     **dict(zip(("young_modulus", "poisson_ratio", "area", "i11", "i22", "i12", "it", "shear_factor"), arrs)))
        self._check(self._L.femgpu_add_beam(self._h, len(number), _p(number, _lib.u32p), _p(n1, _lib.u32p),
                                            _p(n2, _lib.u32p), *[_p(a, _lib.dp) for a in arrs], _p(ax, _lib.dp)))

    def add_plates(self, number, node_1, node_2, node_3, node_4, young_modulus, poisson_ratio, thickness,
                   shear_factor) -> None:
        number = _u32(number)
        ns = [_u32(v) for v in (node_1, node_2, node_3, node_4)]
        arrs = [_f64(v) for v in (young_modulus, poisson_ratio, thickness, shear_factor)]
        _same_length("add_plates", len(number), **dict(zip(("node_1", "node_2", "node_3", "node_4"), ns)),
                     **dict(zip(("young_modulus", "poisson_ratio", "thickness", "shear_factor"), arrs)))
        self._check(self._L.femgpu_add_plate(self._h, len(number), _p(number, _lib.u32p),
                                             *[_p(a, _lib.u32p) for a in ns], *[_p(a, _lib.dp) for a in arrs]))

    def validate(self) -> None:
        """Device-side checks of *::create for everything added since the last call."""
        self._check(self._L.femgpu_validate(self._h, None, None, None))

    def counts(self):
        v = [C.c_uint64() for _ in range(4)]
        self._check(self._L.femgpu_counts(self._h, *[C.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)

    def node_numbers(self):
        """user labels of the nodes, index = insertion order"""
        out = np.empty(self.counts()[0], np.uint32)
        self._check(self._L.femgpu_get_numbers(self._h, -1, _p(out, _lib.u32p) if len(out) else None))
        return out

    def element_numbers(self, family: int):
        out = np.empty(self.counts()[1 + family], np.uint32)
        self._check(self._L.femgpu_get_numbers(self._h, family, _p(out, _lib.u32p) if len(out) else None))
        return out

    # ------------------------------------------------------------------ assembly
    def symbolic(self):
        """One-time pattern + gather-map construction. Returns (n_rows, nnz)."""
        nr, nz = C.c_int64(), C.c_int64()
        self._check(self._L.femgpu_symbolic(self._h, C.byref(nr), C.byref(nz)))
        self.n_rows, self.nnz = int(nr.value), int(nz.value)
        return self.n_rows, self.nnz

    def numeric(self) -> None:
        """Element matrices + deterministic accumulation into the CSR values (asynchronous)."""
        self._check(self._L.femgpu_numeric(self._h))

    def synchronize(self) -> None:
        self._check(self._L.femgpu_synchronize(self._h))

    def assemble(self):
        nr, nz = C.c_int64(), C.c_int64()
        self._check(self._L.femgpu_assemble(self._h, C.byref(nr), C.byref(nz)))
        self.n_rows, self.nnz = int(nr.value), int(nz.value)
        return self.n_rows, self.nnz

    def csr(self, values_only: bool = False, out=None):
        """Host copy of the assembled matrix on the structural pattern."""
        n_rows, nnz = self.symbolic()
        vals = out if out is not None else np.empty(nnz, np.float64)
        if values_only:
            self._check(self._L.femgpu_get_csr(self._h, None, None, _p(vals, _lib.dp)))
            return vals
        rp = np.empty(n_rows + 1, np.int64)
        ci = np.empty(nnz, np.int32)
        self._check(self._L.femgpu_get_csr(self._h, _p(rp, _lib.i64p), _p(ci, _lib.i32p), _p(vals, _lib.dp)))
        return rp, ci, vals

    def csr_device(self):
        """(row_ptr, col_idx, values) device addresses and the owned row range — zero-copy hand-off."""
        p = [C.c_void_p() for _ in range(3)]
        rb, re = C.c_int64(), C.c_int64()
        self._check(self._L.femgpu_get_csr_device(self._h, *[C.byref(x) for x in p], C.byref(rb), C.byref(re)))
        return tuple(x.value for x in p), (int(rb.value), int(re.value))

    def nonzero_csr(self, out=None):
        """The reference's stored set — entries != 0.0 — as CSR (row_ptr, col_idx, values); compacted on the device.
        out = (row_ptr, col_idx, values) host arrays to fill (values / col_idx at least `count` long; e.g. pinned)."""
        cnt = C.c_int64()
        self._check(self._L.femgpu_get_nonzero_csr(self._h, C.byref(cnt), None, None, None))
        n = int(cnt.value)
        if out is None:
            out = (np.empty(self.n_rows + 1, np.int64), np.empty(n, np.int32), np.empty(n, np.float64))
        rp, ci, v = out
        assert len(rp) >= self.n_rows + 1 and len(ci) >= n and len(v) >= n
        self._check(self._L.femgpu_get_nonzero_csr(self._h, C.byref(cnt), _p(rp, _lib.i64p), _p(ci, _lib.i32p), _p(v, _lib.dp)))
        return rp[:self.n_rows + 1], ci[:n], v[:n]

    def nonzero_coo(self):
        """The reference's value-dependent pattern: entries != 0.0, sorted by (row, col)."""
        cnt = C.c_int64()
        self._check(self._L.femgpu_get_nonzero_coo(self._h, C.byref(cnt), None, None, None))
        n = int(cnt.value)
        r = np.empty(n, np.int64); c = np.empty(n, np.int64); v = np.empty(n, np.float64)
        if n:
            self._check(self._L.femgpu_get_nonzero_coo(self._h, C.byref(cnt), _p(r, _lib.i64p), _p(c, _lib.i64p),
                                                       _p(v, _lib.dp)))
        return r, c, v

    # ------------------------------------------------------------------ boundary conditions, separation
    def add_displacement(self, node_number, dof_parameter, value) -> None:
        """methods_for_bc_data_handle.rs:175 (arrays are accepted: n calls in order)"""
        nn, dd, vv = _u32(np.atleast_1d(node_number)), np.ascontiguousarray(np.atleast_1d(dof_parameter), np.int32), \
            _f64(np.atleast_1d(value))
        _same_length("add_displacement", len(nn), dof_parameter=dd, value=vv)
        self._check(self._L.femgpu_add_displacement(self._h, len(nn), _p(nn, _lib.u32p), _p(dd, _lib.i32p),
                                                    _p(vv, _lib.dp)))

    def add_concentrated_load(self, node_number, dof_parameter, value) -> None:
        """methods_for_bc_data_handle.rs:31"""
        nn, dd, vv = _u32(np.atleast_1d(node_number)), np.ascontiguousarray(np.atleast_1d(dof_parameter), np.int32), \
            _f64(np.atleast_1d(value))
        _same_length("add_concentrated_load", len(nn), dof_parameter=dd, value=vv)
        self._check(self._L.femgpu_add_concentrated_load(self._h, len(nn), _p(nn, _lib.u32p), _p(dd, _lib.i32p),
                                                         _p(vv, _lib.dp)))

    def add_uniformly_distributed_line_load(self, beam_element_number, dof_parameter, value) -> None:
        """methods_for_bc_data_handle.rs:58 (arrays are accepted: n calls in order)"""
        nn, dd, vv = _u32(np.atleast_1d(beam_element_number)), np.ascontiguousarray(np.atleast_1d(dof_parameter), np.int32), \
            _f64(np.atleast_1d(value))
        _same_length("add_uniformly_distributed_line_load", len(nn), dof_parameter=dd, value=vv)
        self._check(self._L.femgpu_add_line_load(self._h, len(nn), _p(nn, _lib.u32p), _p(dd, _lib.i32p), _p(vv, _lib.dp)))

    def add_uniformly_distributed_surface_load(self, plate_element_number, dof_parameter, value) -> None:
        """methods_for_bc_data_handle.rs:104"""
        nn, dd, vv = _u32(np.atleast_1d(plate_element_number)), np.ascontiguousarray(np.atleast_1d(dof_parameter), np.int32), \
            _f64(np.atleast_1d(value))
        _same_length("add_uniformly_distributed_surface_load", len(nn), dof_parameter=dd, value=vv)
        self._check(self._L.femgpu_add_surface_load(self._h, len(nn), _p(nn, _lib.u32p), _p(dd, _lib.i32p), _p(vv, _lib.dp)))

    def forces_vector(self, copy_out: bool = True):
        """forces_vector of fem.rs:19: concentrated loads + nodal equivalents of the distributed loads.
        copy_out=False evaluates it on the device only and returns its device address."""
        if not copy_out:
            dev = C.c_void_p()
            self._check(self._L.femgpu_get_forces(self._h, None, C.byref(dev)))
            return int(dev.value or 0)
        out = np.zeros(6 * self.nodes_number, np.float64)
        self._check(self._L.femgpu_get_forces(self._h, _p(out, _lib.dp), None))
        return out

    def last_separation_read_k_once(self) -> bool:
        """True when the last separation took the one-pass path (femgpu_last_separate_path)"""
        flag = C.c_int32()
        self._check(self._L.femgpu_last_separate_path(self._h, C.byref(flag)))
        return bool(flag.value)

    def separate_stiffness_matrix_sparse_iterative(self, copy_out: bool = True):
        """methods_for_separate_stiffness_matrix.rs:217, on the device. copy_out=False leaves the
        quadrants in HBM (returns counts only): (n_aa, n_bb, [nnz_aa, nnz_ab, nnz_ba, nnz_bb], device_ms)."""
        na, nb = C.c_int64(), C.c_int64()
        nnz = np.zeros(4, np.int64)
        self._check(self._L.femgpu_separate_sparse(self._h, C.byref(na), C.byref(nb), _p(nnz, _lib.i64p)))
        ms = C.c_float()
        self._check(self._L.femgpu_last_separate_ms(self._h, C.byref(ms)))
        self._sep_counts = (int(na.value), int(nb.value))
        if not copy_out:
            return int(na.value), int(nb.value), [int(x) for x in nnz], float(ms.value)
        ia, ib = np.empty(na.value, np.int64), np.empty(nb.value, np.int64)
        self._check(self._L.femgpu_get_separated_indexes(self._h, _p(ia, _lib.i64p), _p(ib, _lib.i64p)))
        quads = []
        for q in range(4):
            rows = na.value if q < 2 else nb.value
            rp, ci, v = np.empty(rows + 1, np.int64), np.empty(nnz[q], np.int32), np.empty(nnz[q], np.float64)
            self._check(self._L.femgpu_get_separated_csr(self._h, q, _p(rp, _lib.i64p), _p(ci, _lib.i32p),
                                                         _p(v, _lib.dp)))
            quads.append((rp, ci, v))
        b = np.empty(na.value, np.float64)
        self._check(self._L.femgpu_separated_rhs(self._h, _p(b, _lib.dp), None))
        return SeparatedStiffnessMatrixSparse(ia, ib, quads, b, float(ms.value))

    def separate_stiffness_matrix_direct(self):
        """methods_for_separate_stiffness_matrix.rs:63-215 -> SeparatedStiffnessMatrix: k_aa_indexes, k_bb_indexes,
        k_aa_skyline, the dense quadrants through its get_k_*_matrix() getters (fetched on demand), and K_aa in the
        compacted column form (a, maxa) of the skyline solver (convert_k_aa_into_compacted_form,
        methods_for_global_analysis.rs:50-80) — built on the device without the reference's dense detour. K_ab / K_ba /
        K_bb also stay available as the CSR quadrants of the handle."""
        na, nb, nv = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._L.femgpu_separate_direct(self._h, C.byref(na), C.byref(nb), C.byref(nv)))
        self._sep_counts = (int(na.value), int(nb.value))
        ia, ib = np.empty(na.value, np.int64), np.empty(nb.value, np.int64)
        self._check(self._L.femgpu_get_separated_indexes(self._h, _p(ia, _lib.i64p), _p(ib, _lib.i64p)))
        sky, a, maxa = np.empty(na.value, np.int64), np.empty(nv.value, np.float64), np.empty(na.value + 1, np.int64)
        self._check(self._L.femgpu_get_skyline(self._h, _p(sky, _lib.i64p), _p(a, _lib.dp), _p(maxa, _lib.i64p)))
        return SeparatedStiffnessMatrix(self, ia, ib, sky, a, maxa)

    def separated_dense(self, which: int):
        """quadrant `which` (0 aa, 1 ab, 2 ba, 3 bb) of the last separation as a dense matrix"""
        na, nb = self._n_aa_bb()
        rows, cols = (na if which < 2 else nb), (na if which in (0, 2) else nb)
        out = np.zeros((rows, cols), np.float64)
        if rows and cols:
            self._check(self._L.femgpu_get_separated_dense(self._h, which, _p(out, _lib.dp)))
        return out

    # ------------------------------------------------------------------ global analysis, element results
    def _solve(self, preconditioner: int, max_iter: int, copy_out: bool):
        it = C.c_int64()
        self._check(self._L.femgpu_solve_pcg(self._h, preconditioner, int(max_iter), C.byref(it)))
        if not copy_out:
            return None, int(it.value)
        return self.u_a_vector(), int(it.value)

    def find_ua_vector_iterative_pcg_jacobi_sparse(self, max_iter: int, copy_out: bool = True):
        """methods_for_global_analysis.rs:189-233 on the separated matrix the handle holds (the reference passes
        it back in together with r_a / u_b; here they never left the device). Returns (u_a, iterations)."""
        return self._solve(0, max_iter, copy_out)

    def find_ua_vector_iterative_pcg_block_jacobi_sparse(self, max_iter: int, copy_out: bool = True):
        """methods_for_global_analysis.rs:235-275 (blocks = the rows of one node, :121-147)"""
        return self._solve(1, max_iter, copy_out)

    def find_ua_vector_direct(self, copy_out: bool = True):
        """methods_for_global_analysis.rs:161-187: skyline LDL^T (COLSOL) on the form separate_stiffness_matrix_direct
        left on the device. Returns u_a."""
        self._check(self._L.femgpu_solve_direct(self._h))
        return self.u_a_vector() if copy_out else None

    def find_r_r_vector(self, copy_out: bool = True):
        """methods_for_global_analysis.rs:277-309 (the direct flow's name for the reactions): the same sums as the
        sparse variant, on the CSR quadrants"""
        return self.find_r_r_vector_sparse(copy_out)

    def u_a_vector(self):
        n = self._n_aa_bb()[0]
        out = np.empty(n, np.float64)
        self._check(self._L.femgpu_get_ua(self._h, _p(out, _lib.dp), None))
        return out

    def set_u_a_vector(self, u_a) -> None:
        """install the solution of an external solver (e.g. a direct solve of K_aa) as u_a"""
        u = _f64(u_a)
        if len(u) != self._n_aa_bb()[0]:
            raise ValueError("u_a has the wrong length")
        self._check(self._L.femgpu_set_ua(self._h, _p(u, _lib.dp)))

    def _n_aa_bb(self):
        if getattr(self, "_sep_counts", None) is None:
            raise FemError(-2, "no separated matrix: call separate_stiffness_matrix_sparse_iterative first")
        return self._sep_counts

    def solve_info(self):
        it, res, ms = C.c_int64(), C.c_double(), C.c_float()
        self._check(self._L.femgpu_solve_info(self._h, C.byref(it), C.byref(res), C.byref(ms)))
        return int(it.value), float(res.value), float(ms.value)

    def find_r_r_vector_sparse(self, copy_out: bool = True):
        """methods_for_global_analysis.rs:334-360 (+ compose_global_analysis_result, :362-385, in the same call)"""
        self._check(self._L.femgpu_global_analysis(self._h))
        if not copy_out:
            return None
        out = np.empty(self._n_aa_bb()[1], np.float64)
        self._check(self._L.femgpu_get_reactions(self._h, _p(out, _lib.dp), None))
        return out

    def compose_global_analysis_result(self) -> None:
        """methods_for_global_analysis.rs:362-385: displacements[k_aa_indexes] = u_a, forces[k_bb_indexes] = r_r"""
        self._check(self._L.femgpu_global_analysis(self._h))

    def global_analysis_vectors(self):
        """(displacements, forces), each [6 * nodes_number], after compose_global_analysis_result"""
        n = 6 * self.nodes_number
        d, f = np.empty(n, np.float64), np.empty(n, np.float64)
        self._check(self._L.femgpu_get_global_result(self._h, _p(d, _lib.dp), _p(f, _lib.dp)))
        return d, f

    def extract_global_analysis_result(self):
        """methods_for_global_analysis.rs:387-...: [(node_number, dof_parameter, displacement, load)], sorted by
        (node insertion order, dof) — the reference iterates a HashMap, its order is unspecified"""
        d, f = self.global_analysis_vectors()
        out = []
        for i, number in enumerate(self.node_numbers()):
            for k in range(6):
                out.append((int(number), k, float(d[6 * i + k]), float(f[6 * i + k])))
        return out

    def set_displacements_vector(self, displacements) -> None:
        u = _f64(displacements)
        if len(u) != 6 * self.nodes_number:
            raise ValueError("displacements must hold 6 * nodes_number values")
        self._check(self._L.femgpu_set_displacements(self._h, _p(u, _lib.dp)))

    def element_results(self, family: int, copy_out: bool = True):
        """rows = elements of `family` in insertion order, columns = the reference's components:
        truss [n, 1], beam [n, 10], plate [n, 8]. copy_out=False leaves them in HBM (returns the device address)."""
        n = self.counts()[1 + family]
        k = ELEMENT_RESULT_COMPONENTS[family]
        if not copy_out:
            dev = C.c_void_p()
            self._check(self._L.femgpu_element_results(self._h, family, None, C.byref(dev)))
            return int(dev.value or 0)
        out = np.empty((n, len(k)), np.float64)
        self._check(self._L.femgpu_element_results(self._h, family, _p(out, _lib.dp) if n else None, None))
        return out

    def extract_elements_analysis_result(self):
        """methods_for_element_analysis.rs:27-58: [(element_number, [(component, value), ...])], trusses, then
        beams, then plates, each family in insertion order"""
        out = []
        for family in (TRUSS, BEAM, PLATE):
            vals = self.element_results(family)
            names = ELEMENT_RESULT_COMPONENTS[family]
            for number, row in zip(self.element_numbers(family), vals):
                out.append((int(number), [(c, float(v)) for c, v in zip(names, row)]))
        return out

    # ------------------------------------------------------------------ hooks
    def element_matrix(self, family: int, number: int):
        n = {TRUSS: 6, BEAM: 12, PLATE: 24}[family]
        out = np.zeros(n * n)
        self._check(self._L.femgpu_element_matrix(self._h, family, number, _p(out, _lib.dp)))
        return out.reshape(n, n)

    def element_slots(self, family: int, number: int):
        n = {TRUSS: 6, BEAM: 12, PLATE: 24}[family]
        out = np.zeros(n * n, np.int64)
        self._check(self._L.femgpu_element_slots(self._h, family, number, _p(out, _lib.i64p)))
        return out.reshape(n, n)

    def launch_count(self, reset: bool = False) -> int:
        v = C.c_uint64()
        self._check(self._L.femgpu_launch_count(self._h, int(reset), C.byref(v)))
        return int(v.value)

    def last_numeric_ms(self):
        out = np.zeros(4, np.float32)
        self._check(self._L.femgpu_last_numeric_ms(self._h, _p(out, _lib.fp)))
        return [float(x) for x in out]

    def numeric_ms_history(self, passes_back: int = 0):
        out = np.zeros(4, np.float32)
        self._check(self._L.femgpu_numeric_ms_history(self._h, passes_back, _p(out, _lib.fp)))
        return [float(x) for x in out]

    def numeric_kernel_ms(self, passes_back: int = 0):
        """(sum of the assemble_kernel launch durations of that pass in ms, number of launches)"""
        ms, n = C.c_float(), C.c_int32()
        self._check(self._L.femgpu_numeric_kernel_ms(self._h, passes_back, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def device_bytes(self) -> int:
        v = C.c_uint64()
        self._check(self._L.femgpu_device_bytes(self._h, C.byref(v)))
        return int(v.value)

    def fp64_fma_peak(self) -> float:
        """measured FP64 FMA throughput of the device, TFLOP/s (micro-benchmark)"""
        v = C.c_double()
        self._check(self._L.femgpu_fp64_fma_peak(self._h, C.byref(v)))
        return float(v.value)

    def stream(self) -> int:
        v = C.c_void_p()
        self._check(self._L.femgpu_stream(self._h, C.byref(v)))
        return int(v.value or 0)

    # ------------------------------------------------------------------ multi-GPU
    @staticmethod
    def dist_unique_id() -> bytes:
        L = _lib.load()
        buf = (C.c_uint8 * 128)()
        st = L.femgpu_dist_unique_id(buf)
        if st:
            raise FemError(st, "ncclGetUniqueId failed")
        return bytes(buf)

    def dist_init(self, rank: int, world: int, unique_id: bytes) -> None:
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._L.femgpu_dist_init(self._h, rank, world, buf))

    def dist_set_ownership(self, node_index_begin: int, node_index_end: int) -> None:
        self._check(self._L.femgpu_dist_set_ownership(self._h, node_index_begin, node_index_end))

    def dist_set_node_window(self, first_node_index: int) -> None:
        """the nodes added next get the global insertion indices first_node_index, first_node_index + 1, ..."""
        self._check(self._L.femgpu_dist_set_node_window(self._h, first_node_index))

    def dist_last_exchange_bytes(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._L.femgpu_dist_last_exchange_bytes(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def dist_info(self):
        """(p2p, passes): 1 when ghost rows travel through peer windows over NVLink (0: ncclSend/ncclRecv), and the
        numeric passes issued since the last symbolic pass"""
        a, b = C.c_int32(), C.c_uint64()
        self._check(self._L.femgpu_dist_info(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    # ------------------------------------------------------------------ convenience
    _API_SOURCES = ("x", "y", "z", "p_n", "p_props", "b_n1", "b_n2", "b_props", "b_axis", "t_n1", "t_n2", "t_E", "t_A", "t_A2",
                    "node_number_offset", "element_number_offset", "node_window_begin")

    @classmethod
    def api_arrays(cls, mesh: dict, cache: bool = False) -> dict:
        """The arrays the bulk API takes for a mesh dict of finite_element_method_b200.meshes (0-based node indices,
        implicit labels): node / element numbers and node numbers per element, as a caller of the reference holds
        them. With `cache` they are built once per mesh dict and kept in it (key "_api") as long as the dict still
        holds the very same source objects; arrays changed in place need a fresh dict (or cache=False)."""
        sources = tuple(mesh.get(k) for k in cls._API_SOURCES)
        api = mesh.get("_api") if cache else None
        if api is not None and len(api["sources"]) == len(sources) and all(a is b for a, b in zip(api["sources"], sources)):
            return api
        n = len(mesh["x"])
        first = mesh.get("node_number_offset", 1)
        w0 = int(mesh.get("node_window_begin", 0))   # x / y / z hold the nodes [w0, w0 + n) of the whole model only
        num = mesh.get("element_number_offset", 1)
        api = {"window": w0, "sources": sources,
               "nodes": (np.arange(first + w0, first + w0 + n, dtype=np.uint32), _f64(mesh["x"]), _f64(mesh["y"]), _f64(mesh["z"]))}
        pn = np.asarray(mesh["p_n"], np.uint32).reshape(4, -1)
        if pn.shape[1]:
            pp = np.asarray(mesh["p_props"], np.float64).reshape(4, -1)
            api["plates"] = (np.arange(num, num + pn.shape[1], dtype=np.uint32), *[_u32(pn[i] + first) for i in range(4)],
                             *[_f64(pp[i]) for i in range(4)])
        nb = len(mesh["b_n1"])
        if nb:
            bp = np.asarray(mesh["b_props"], np.float64).reshape(8, -1)
            api["beams"] = (np.arange(num, num + nb, dtype=np.uint32), _u32(np.asarray(mesh["b_n1"], np.uint32) + first),
                            _u32(np.asarray(mesh["b_n2"], np.uint32) + first), *[_f64(bp[i]) for i in range(8)],
                            _f64(mesh["b_axis"]))
        nt = len(mesh["t_n1"])
        if nt:
            a2 = mesh.get("t_A2")
            api["trusses"] = (np.arange(num, num + nt, dtype=np.uint32), _u32(np.asarray(mesh["t_n1"], np.uint32) + first),
                              _u32(np.asarray(mesh["t_n2"], np.uint32) + first), _f64(mesh["t_E"]), _f64(mesh["t_A"]),
                              a2 if a2 is None or isinstance(a2, (list, tuple)) else _f64(a2))
        if cache:
            mesh["_api"] = api
        return api

    def load_mesh(self, mesh: dict, cache: bool = False) -> None:
        """Bulk-load a mesh dict from finite_element_method_b200.meshes (node number = index + 1;
        insertion order plates -> beams -> trusses). `cache`: see api_arrays."""
        api = self.api_arrays(mesh, cache)
        if api["window"]:
            self.dist_set_node_window(api["window"])
        self.add_nodes(*api["nodes"])
        if "plates" in api:
            self.add_plates(*api["plates"])
        if "beams" in api:
            self.add_beams(*api["beams"])
        if "trusses" in api:
            self.add_trusses(*api["trusses"])
