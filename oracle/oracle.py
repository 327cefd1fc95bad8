"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for oracle/_build/liboracle.so, the CPU restatement of the reference's stiffness
hot path (see fem_oracle.hpp for the file:line citations). Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module. The product package
(finite_element_method_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

ERROR_TEXT = {}

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_i64p = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ only; no reference sources involved)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "fem_oracle.hpp", "fem_oracle_fast.hpp")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_error_text.restype = C.c_char_p
        _lib.oracle_model_create.restype = C.c_void_p
        _lib.oracle_model_create.argtypes = [C.c_double, C.c_double, C.c_uint32]
        _lib.oracle_model_destroy.argtypes = [C.c_void_p]
        _lib.oracle_model_nnz.restype = C.c_int64
        _lib.oracle_model_nnz.argtypes = [C.c_void_p]
        _lib.oracle_fast_assemble.restype = C.c_double
        _lib.oracle_faithful_time.restype = C.c_double
    return _lib


def set_variants(antiparallel=-1, inverse2=-1, reset_hits=False):
    """Test-only switches of the restatement (fem_oracle.hpp `Variants`): antiparallel 0 zero axis (Q = -I, default) /
    1 pi about y / 2 pi about z / 3 identity; inverse2 0 closed form (default) / 1 LU / 2 pivoted elimination.
    -1 leaves a switch alone. Returns how often the anti-parallel branch has been taken so far."""
    f = lib().oracle_set_variants
    f.restype = C.c_long
    return int(f(C.c_int(antiparallel), C.c_int(inverse2), C.c_int(int(reset_hits))))


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _u(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(_u32p)


class OracleError(Exception):
    def __init__(self, code, at=None):
        self.code = int(code)
        self.at = at
        super().__init__(f"{lib().oracle_error_text(int(code)).decode()} (code {code}, element {at})")


# ------------------------------------------------------------------ element level
def truss(p1, p2, E, A, A2=None, rel_tol=1e-4, abs_tol=1e-12):
    q = np.zeros(9); kl = np.zeros(36); kg = np.zeros(36)
    _, a = _d(p1); _, b = _d(p2)
    err = lib().oracle_truss_f64(a, b, C.c_double(E), C.c_double(A),
                                 C.c_double(float("nan") if A2 is None else A2),
                                 C.c_double(rel_tol), C.c_double(abs_tol),
                                 q.ctypes.data_as(_dp), kl.ctypes.data_as(_dp), kg.ctypes.data_as(_dp))
    if err:
        raise OracleError(err)
    return q.reshape(3, 3), kl.reshape(6, 6), kg.reshape(6, 6)


def truss_f32(p1, p2, E, A, A2=None, rel_tol=1e-4, abs_tol=1e-12):
    q = np.zeros(9, np.float32); kl = np.zeros(36, np.float32); kg = np.zeros(36, np.float32)
    a = np.ascontiguousarray(p1, np.float32); b = np.ascontiguousarray(p2, np.float32)
    err = lib().oracle_truss_f32(a.ctypes.data_as(_fp), b.ctypes.data_as(_fp), C.c_float(E), C.c_float(A),
                                 C.c_float(float("nan") if A2 is None else A2),
                                 C.c_float(rel_tol), C.c_float(abs_tol),
                                 q.ctypes.data_as(_fp), kl.ctypes.data_as(_fp), kg.ctypes.data_as(_fp))
    if err:
        raise OracleError(err)
    return q.reshape(3, 3), kl.reshape(6, 6), kg.reshape(6, 6)


def beam(p1, p2, E, nu, A, I11, I22, I12, It, ks, axis1, rel_tol=1e-4, abs_tol=1e-12):
    q = np.zeros(9); pr = np.zeros(3); kl = np.zeros(144); kg = np.zeros(144)
    _, a = _d(p1); _, b = _d(p2); _, ax = _d(axis1)
    err = lib().oracle_beam_f64(a, b, *[C.c_double(v) for v in (E, nu, A, I11, I22, I12, It, ks)], ax,
                                C.c_double(rel_tol), C.c_double(abs_tol), q.ctypes.data_as(_dp),
                                pr.ctypes.data_as(_dp), kl.ctypes.data_as(_dp), kg.ctypes.data_as(_dp))
    if err:
        raise OracleError(err)
    return q.reshape(3, 3), pr, kl.reshape(12, 12), kg.reshape(12, 12)


def plate(p1, p2, p3, p4, E, nu, t, ks, rel_tol=1e-4, abs_tol=1e-12):
    q = np.zeros(9); kl = np.zeros(576); kg = np.zeros(576)
    ps = [_d(p)[1] for p in (p1, p2, p3, p4)]
    err = lib().oracle_plate_f64(*ps, *[C.c_double(v) for v in (E, nu, t, ks)],
                                 C.c_double(rel_tol), C.c_double(abs_tol), q.ctypes.data_as(_dp),
                                 kl.ctypes.data_as(_dp), kg.ctypes.data_as(_dp))
    if err:
        raise OracleError(err)
    return q.reshape(3, 3), kl.reshape(24, 24), kg.reshape(24, 24)


def beam_line_load(p1, p2, q):
    """Beam::convert_uniformly_distributed_line_load_to_nodal_loads (beam.rs:775-797) -> f[2]"""
    f = np.zeros(2)
    lib().oracle_beam_line_load_f64(_d(p1)[1], _d(p2)[1], C.c_double(q), f.ctypes.data_as(_dp))
    return f


def plate_surface_load(p1, p2, p3, p4, q, rel_tol=1e-4, abs_tol=1e-12):
    """Plate::convert_uniformly_distributed_surface_load_to_nodal_loads (plate.rs:1145-1185) -> f[4]"""
    f = np.zeros(4)
    lib().oracle_plate_surface_load_f64(*[_d(p)[1] for p in (p1, p2, p3, p4)], C.c_double(q),
                                        C.c_double(rel_tol), C.c_double(abs_tol), f.ctypes.data_as(_dp))
    return f


def reference_truss_test_f32():
    """Replays the reference's own f32 test model; returns (k00, u2x, r1x, force_r) as float32."""
    v = [C.c_float() for _ in range(4)]
    err = lib().oracle_reference_truss_test_f32(*[C.byref(x) for x in v])
    if err:
        raise OracleError(err)
    return tuple(np.float32(x.value) for x in v)


# ------------------------------------------------------------------ model level (faithful)
class Model:
    """Faithful single-thread assembly into a position-keyed map, in the order add_* is called.

    Node arguments are 0-based node *indices* (insertion order), not user numbers.
    """

    def __init__(self, rel_tol, abs_tol, nodes_number):
        self._h = C.c_void_p(lib().oracle_model_create(rel_tol, abs_tol, nodes_number))
        self.n_rows = 6 * nodes_number

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_model_destroy(self._h)
            self._h = None

    def set_nodes(self, x, y, z):
        x, px = _d(x); y, py = _d(y); z, pz = _d(z)
        if lib().oracle_model_set_nodes(self._h, C.c_int64(len(x)), px, py, pz):
            raise ValueError("more nodes than nodes_number")

    def add_truss(self, n1, n2, E, A, A2=None):
        n1, p1 = _u(n1); n2, p2 = _u(n2); E, pE = _d(E); A, pA = _d(A)
        if A2 is None:
            A2 = np.full(len(n1), np.nan)
        A2, pA2 = _d(A2)
        at = C.c_int64(-1)
        err = lib().oracle_model_add_truss(self._h, C.c_int64(len(n1)), p1, p2, pE, pA, pA2, C.byref(at))
        if err:
            raise OracleError(err, at.value)

    def add_beam(self, n1, n2, E, nu, A, I11, I22, I12, It, ks, axis1):
        n1, p1 = _u(n1); n2, p2 = _u(n2)
        arrs = [_d(v) for v in (E, nu, A, I11, I22, I12, It, ks)]
        ax = np.ascontiguousarray(np.asarray(axis1, np.float64).reshape(3, -1))
        at = C.c_int64(-1)
        err = lib().oracle_model_add_beam(self._h, C.c_int64(len(n1)), p1, p2, *[a[1] for a in arrs],
                                          ax.ctypes.data_as(_dp), C.byref(at))
        if err:
            raise OracleError(err, at.value)

    def add_plate(self, n1, n2, n3, n4, E, nu, t, ks):
        ns = [_u(v) for v in (n1, n2, n3, n4)]
        arrs = [_d(v) for v in (E, nu, t, ks)]
        at = C.c_int64(-1)
        err = lib().oracle_model_add_plate(self._h, C.c_int64(len(ns[0][0])), *[a[1] for a in ns],
                                           *[a[1] for a in arrs], C.byref(at))
        if err:
            raise OracleError(err, at.value)

    def coo(self):
        """Stored entries sorted by (row, col): (rows, cols, vals)."""
        nnz = lib().oracle_model_nnz(self._h)
        r = np.zeros(nnz, np.int64); c = np.zeros(nnz, np.int64); v = np.zeros(nnz, np.float64)
        lib().oracle_model_get_coo(self._h, r.ctypes.data_as(_i64p), c.ctypes.data_as(_i64p),
                                   v.ctypes.data_as(_dp))
        return r, c, v


# ------------------------------------------------------------------ mesh-level helpers
def _mesh_args(mesh):
    """mesh: dict produced by finite_element_method_b200.meshes generators (plain numpy arrays)."""
    keep = []

    def d(a):
        a, p = _d(a); keep.append(a); return p

    def u(a):
        a, p = _u(a); keep.append(a); return p

    nt, nb, npl = len(mesh["t_n1"]), len(mesh["b_n1"]), len(mesh["p_n"][0]) if len(mesh["p_n"]) else 0
    t_A2 = mesh.get("t_A2")
    if t_A2 is None:
        t_A2 = np.full(nt, np.nan)
    args = [C.c_int64(len(mesh["x"])), d(mesh["x"]), d(mesh["y"]), d(mesh["z"]),
            C.c_int64(nt), u(mesh["t_n1"]), u(mesh["t_n2"]), d(mesh["t_E"]), d(mesh["t_A"]), d(t_A2),
            C.c_int64(nb), u(mesh["b_n1"]), u(mesh["b_n2"]),
            d(np.asarray(mesh["b_props"], np.float64).reshape(8, -1)),
            d(np.asarray(mesh["b_axis"], np.float64).reshape(3, -1)),
            C.c_int64(npl), u(np.asarray(mesh["p_n"], np.uint32).reshape(4, -1)),
            d(np.asarray(mesh["p_props"], np.float64).reshape(4, -1)),
            C.c_double(mesh["rel_tol"]), C.c_double(mesh["abs_tol"])]
    return args, keep


def fast_assemble(mesh, n_threads=0, repeats=1, want_coo=False):
    """Multi-core owner-computes CPU assembly. Returns dict(seconds, nnz, checksum[, coo])."""
    args, keep = _mesh_args(mesh)
    nnz = C.c_int64(0); cs = C.c_double(0.0)
    null_i = C.cast(None, _i64p); null_d = C.cast(None, _dp)
    sec = lib().oracle_fast_assemble(*args, C.c_int(n_threads), C.c_int(repeats), C.byref(nnz),
                                     C.byref(cs), null_i, null_i, null_d)
    out = {"seconds": sec, "nnz": nnz.value, "checksum": cs.value}
    if sec < 0:
        raise OracleError(-1)
    if want_coo:
        r = np.zeros(nnz.value, np.int64); c = np.zeros(nnz.value, np.int64); v = np.zeros(nnz.value)
        lib().oracle_fast_assemble(*args, C.c_int(n_threads), C.c_int(1), C.byref(nnz), C.byref(cs),
                                   r.ctypes.data_as(_i64p), c.ctypes.data_as(_i64p), v.ctypes.data_as(_dp))
        out["coo"] = (r, c, v)
    return out


def sample_rows(mesh, nodes, n_threads=0, faithful=True):
    """Block rows of the sampled nodes, computed from scratch at any mesh size (fem_oracle_fast.hpp sample_rows):
    returns (blk_ptr [n+1], blk_col, blk_full, blk_val [nblk, 6, 6]); blocks of a node sorted by column node,
    contributions summed plates -> beams -> trusses in insertion order (the reference's add_value sequence for a
    model loaded family by family). faithful=True: element matrices by the operation-by-operation restatement
    (dense (R^T k) R on Mat, fem_oracle.hpp); False: the multi-core baseline's block-wise ones."""
    args, keep = _mesh_args(mesh)
    nodes = np.ascontiguousarray(nodes, np.uint32)
    assert len(np.unique(nodes)) == len(nodes)
    ptr = np.zeros(len(nodes) + 1, np.int64)
    f = lib().oracle_sample_rows
    f.restype = C.c_int64
    u8p = C.POINTER(C.c_uint8)
    nb = f(*args, C.c_int(n_threads), C.c_int(int(faithful)), C.c_int64(len(nodes)), nodes.ctypes.data_as(_u32p), ptr.ctypes.data_as(_i64p),
           C.cast(None, _u32p), C.cast(None, u8p), C.cast(None, _dp))
    if nb < 0:
        raise OracleError(-1)
    col = np.zeros(nb, np.uint32); full = np.zeros(nb, np.uint8); val = np.zeros((nb, 6, 6))
    f(*args, C.c_int(n_threads), C.c_int(int(faithful)), C.c_int64(len(nodes)), nodes.ctypes.data_as(_u32p), ptr.ctypes.data_as(_i64p),
      col.ctypes.data_as(_u32p), full.ctypes.data_as(u8p), val.ctypes.data_as(_dp))
    return ptr, col, full, val


def element_results(mesh, u):
    """extract_elements_analysis_result (methods_for_element_analysis.rs:27-58) of a mesh dict for the global
    displacement vector u (6 per node): (truss [nt], beam [nb, 10], plate [np, 8]) in the reference's
    component order (truss.rs:325-328, beam.rs:967-987, plate.rs:1368-1401)."""
    args, keep = _mesh_args(mesh)
    nt, nb = len(mesh["t_n1"]), len(mesh["b_n1"])
    npl = len(mesh["p_n"][0]) if len(mesh["p_n"]) else 0
    u, pu = _d(u)
    ot = np.zeros(nt); ob = np.zeros((nb, 10)); op = np.zeros((npl, 8))
    lib().oracle_element_results(*args, pu, ot.ctypes.data_as(_dp), ob.ctypes.data_as(_dp), op.ctypes.data_as(_dp))
    return ot, ob, op


def faithful_time(mesh):
    """Seconds for the single-thread faithful add_* loop over the whole mesh (plates, beams, trusses)."""
    args, keep = _mesh_args(mesh)
    return lib().oracle_faithful_time(*args)


def faithful_coo(mesh):
    """Faithful assembly of a mesh dict in the order plates -> beams -> trusses."""
    m = Model(mesh["rel_tol"], mesh["abs_tol"], mesh.get("nodes_number", len(mesh["x"])))
    m.set_nodes(mesh["x"], mesh["y"], mesh["z"])
    if len(mesh["p_n"]) and len(mesh["p_n"][0]):
        pn = np.asarray(mesh["p_n"]).reshape(4, -1); pp = np.asarray(mesh["p_props"]).reshape(4, -1)
        m.add_plate(pn[0], pn[1], pn[2], pn[3], pp[0], pp[1], pp[2], pp[3])
    if len(mesh["b_n1"]):
        bp = np.asarray(mesh["b_props"]).reshape(8, -1)
        m.add_beam(mesh["b_n1"], mesh["b_n2"], *[bp[i] for i in range(8)], mesh["b_axis"])
    if len(mesh["t_n1"]):
        m.add_truss(mesh["t_n1"], mesh["t_n2"], mesh["t_E"], mesh["t_A"], mesh.get("t_A2"))
    return m.coo()


# ---------------------------------------------------------------------------------------------
# Sparse separation of K (restatement, numpy; integer/index work only — values are copied).
# Follows FEM::separate_stiffness_matrix_sparse_iterative, methods_for_separate_stiffness_matrix.rs:217-320,
# find_b_sparse, methods_for_global_analysis.rs:28-48, compose_r_a_vector / compose_u_b_vector :322-343.
# PINNED against the reference's own sparse-separation test model (src/tests/fem/test_fem.rs:65-80:
# n_aa > 0, K_aa non-empty) and the end-to-end numbers of :83-150 (u = b / K_aa = 0.0015, one PCG
# iteration) in tests/test_separation.py; larger meshes are checked against this restatement.
# ---------------------------------------------------------------------------------------------
DOF_NAMES = ("X", "Y", "Z", "ThX", "ThY", "ThZ")


class SeparationError(Exception):
    pass


def separate_sparse(n_dof, rows, cols, vals, constrained, node_numbers=None, forces=None, displacements=None):
    """rows/cols/vals: the position-keyed map of K as COO (one entry per position).
    constrained: bool[n_dof] imposed_constraints. Returns a dict with k_aa_indexes, k_bb_indexes, the
    four triplet lists sorted by (i, j) (the reference's order is hash-map order, i.e. unspecified)
    and b = R_a - K_ab u_b."""
    rows = np.asarray(rows, np.int64); cols = np.asarray(cols, np.int64); vals = np.asarray(vals, np.float64)
    constrained = np.asarray(constrained, bool)
    diag = np.zeros(n_dof)                                   # :222-227
    d = rows == cols
    diag[rows[d]] = vals[d]
    zero = diag == 0.0
    bad = np.nonzero(zero & constrained)[0]                  # :232-243, first index in ascending order
    if len(bad):
        idx = int(bad[0])
        number = 0 if node_numbers is None or idx // 6 >= len(node_numbers) else int(node_numbers[idx // 6])
        raise SeparationError(f"There are no stiffness to withstand displacement {DOF_NAMES[idx % 6]} applied to node {number}!")
    k_bb = np.nonzero(~zero & constrained)[0]                # :250-254
    k_aa = np.nonzero(~zero & ~constrained)[0]
    if len(k_bb) == 0:
        raise SeparationError("No restraints")               # :257-259
    aa_pos = np.full(n_dof, -1, np.int64); aa_pos[k_aa] = np.arange(len(k_aa))   # :264-271
    bb_pos = np.full(n_dof, -1, np.int64); bb_pos[k_bb] = np.arange(len(k_bb))
    nz = vals != 0.0                                         # :277-279
    r, c, v = rows[nz], cols[nz], vals[nz]
    out = {"k_aa_indexes": k_aa, "k_bb_indexes": k_bb, "n_aa": len(k_aa), "n_bb": len(k_bb)}
    for name, rp, cp in (("k_aa", aa_pos, aa_pos), ("k_ab", aa_pos, bb_pos), ("k_ba", bb_pos, aa_pos), ("k_bb", bb_pos, bb_pos)):
        m = (rp[r] >= 0) & (cp[c] >= 0)                      # :288-299
        i, j, x = rp[r][m], cp[c][m], v[m]
        o = np.lexsort((j, i))
        out[name] = (i[o], j[o], x[o])
    if len(out["k_aa"][0]) == 0:                             # :303-307
        raise SeparationError("Sparse separation: K_aa is empty (structure has no free stiffness?)")
    if forces is not None and displacements is not None:
        b = np.asarray(forces, np.float64)[k_aa].copy()      # compose_r_a_vector
        u_b = np.asarray(displacements, np.float64)[k_bb]    # compose_u_b_vector
        i, j, x = out["k_ab"]
        np.subtract.at(b, i, x * u_b[j])                     # find_b_sparse (row order; the reference's is unspecified)
        out["b"] = b
    return out


# ---------------------------------------------------------------------------------------------
# Global analysis downstream of the separation (restatement, numpy): the iterative solve, the reactions
# and the composed result vectors.
#   find_ua_vector_iterative_pcg_jacobi_sparse / ..._block_jacobi_sparse  methods_for_global_analysis.rs:189-275
#   build_block_starts_from_k_aa_indexes :121-147, find_r_r_sparse :100-137, compose_global_analysis_result :362-385
# The PCG arithmetic itself lives in the un-vendored crate iterative_solvers_smpl 0.1.5: what is restated is
# the textbook preconditioned conjugate gradient from x0 = 0, stopping when ||r||_2 <= max(rel_tol ||b||_2,
# abs_tol), the count being the number of search directions used. PARITY UNPINNED beyond the reference's own
# test model (iterations == 1, u = 0.0015; src/tests/fem/test_fem.rs:83-225), which tests/ check.
# ---------------------------------------------------------------------------------------------
def block_starts_from_k_aa_indexes(k_aa_indexes):
    """methods_for_global_analysis.rs:121-147: local row where the rows of the next node begin"""
    starts = []
    current = None
    for i, g in enumerate(k_aa_indexes):
        node = int(g) // 6
        if node != current:
            starts.append(i)
            current = node
    return starts


def pcg(n, k_aa, b, max_iter, rel_tol, abs_tol, block_starts=None):
    """k_aa = (i, j, value) triplets of K_aa. Returns (u_a, iterations). Jacobi when block_starts is None."""
    import scipy.sparse as sp
    i, j, v = k_aa
    A = sp.csr_matrix((np.asarray(v, np.float64), (np.asarray(i), np.asarray(j))), shape=(n, n))
    b = np.asarray(b, np.float64)
    if block_starts is None:
        dinv = 1.0 / A.diagonal()

        def minv(r):
            return dinv * r
    else:
        Ad = A.toarray() if n <= 4096 else None
        bounds = list(block_starts) + [n]
        invs = []
        for s, e in zip(bounds[:-1], bounds[1:]):
            blk = Ad[s:e, s:e] if Ad is not None else A[s:e, s:e].toarray()
            invs.append((s, e, np.linalg.inv(blk)))

        def minv(r):
            z = np.empty_like(r)
            for s, e, m in invs:
                z[s:e] = m @ r[s:e]
            return z
    x = np.zeros(n)
    r = b.copy()
    tol = max(rel_tol * np.sqrt(b @ b), abs_tol)
    if np.sqrt(r @ r) <= tol:
        return x, 0
    z = minv(r)
    p = z.copy()
    rz = r @ z
    for k in range(max_iter):
        ap = A @ p
        alpha = rz / (p @ ap)
        x = x + alpha * p
        r = r - alpha * ap
        if np.sqrt(r @ r) <= tol:
            return x, k + 1
        z = minv(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    raise SeparationError(f"PCG did not converge in {max_iter} iterations")


def reactions(sep, u_a, forces, displacements):
    """find_r_r_sparse: r_r = K_ba u_a + K_bb u_b - R_b (each product summed in (row, column) order)"""
    n_bb = sep["n_bb"]
    u_b = np.asarray(displacements, np.float64)[sep["k_bb_indexes"]]
    y_ba = np.zeros(n_bb); y_bb = np.zeros(n_bb)
    i, j, x = sep["k_ba"]
    np.add.at(y_ba, i, x * np.asarray(u_a)[j])
    i, j, x = sep["k_bb"]
    np.add.at(y_bb, i, x * u_b[j])
    return y_ba + y_bb - np.asarray(forces, np.float64)[sep["k_bb_indexes"]]


def compose_global_analysis_result(sep, u_a, r_r, forces, displacements):
    d = np.asarray(displacements, np.float64).copy()
    f = np.asarray(forces, np.float64).copy()
    d[sep["k_aa_indexes"]] = u_a
    f[sep["k_bb_indexes"]] = r_r
    return d, f


# ---------------------------------------------------------------------------------------------
# Direct separation (restatement, numpy; index work and copied values).
# FEM::separate_stiffness_matrix_direct, methods_for_separate_stiffness_matrix.rs:63-215, with check_excluded_index
# :36-61, and the compacted column form its consumer builds (convert_k_aa_into_compacted_form,
# methods_for_global_analysis.rs:50-80). Pinned on the reference's direct test model (src/tests/fem/test_fem.rs:5-64:
# one free DOF, K_aa = [EA/L], skyline [0]) in tests/test_separation.py.
# ---------------------------------------------------------------------------------------------
def separate_direct(n_dof, rows, cols, vals, constrained, node_numbers=None, forces=None):
    rows = np.asarray(rows, np.int64); cols = np.asarray(cols, np.int64); vals = np.asarray(vals, np.float64)
    constrained = np.asarray(constrained, bool)
    forces = np.zeros(n_dof) if forces is None else np.asarray(forces, np.float64)
    diag = np.zeros(n_dof)
    d = rows == cols
    diag[rows[d]] = vals[d]
    k_aa, k_bb = [], []
    for index in range(n_dof):                                   # :70-86, ascending
        if diag[index] == 0.0:
            number = 0 if node_numbers is None or index // 6 >= len(node_numbers) else int(node_numbers[index // 6])
            if constrained[index]:                               # check_excluded_index :43-48
                raise SeparationError(f"There are no stiffness to withstand displacement {DOF_NAMES[index % 6]} applied to node {number}!")
            if forces[index] != 0.0:                             # :50-58
                raise SeparationError(f"There are no stiffness to withstand load {DOF_NAMES[index % 6]} applied to node {number}!")
        elif constrained[index]:
            k_bb.append(index)
        else:
            k_aa.append(index)
    if not k_bb:
        raise SeparationError("There are no restraints applied!")  # :88-90
    k_aa = np.asarray(k_aa, np.int64); k_bb = np.asarray(k_bb, np.int64)
    pos = np.full(n_dof, -1, np.int64); pos[k_aa] = np.arange(len(k_aa))
    m = (pos[rows] >= 0) & (pos[cols] >= 0) & (vals != 0.0)
    i, j, v = pos[rows][m], pos[cols][m], vals[m]
    skyline = np.zeros(len(k_aa), np.int64)                      # :117-129: j - i over the non-zero entries above the diagonal
    up = j > i
    np.maximum.at(skyline, j[up], (j - i)[up])
    maxa = np.concatenate([[0], np.cumsum(skyline + 1)])         # methods_for_global_analysis.rs:50-80
    a = np.zeros(int(maxa[-1]))
    on = j >= i
    a[maxa[j[on]] + (j - i)[on]] = v[on]
    return {"k_aa_indexes": k_aa, "k_bb_indexes": k_bb, "k_aa_skyline": skyline, "a": a, "maxa": maxa}


# ---------------------------------------------------------------------------------------------
# Direct solve (restatement, numpy): FEM::find_ua_vector_direct, methods_for_global_analysis.rs:161-187, hands the
# compacted column form (a, maxa) to the un-vendored crate colsol 1.0.1 (`factorization`, `find_unknown`) — Bathe's
# COLSOL active-column LDL^T, restated here column by column in its order (the dot products through numpy).
# Pinned on the reference's direct test model (src/tests/fem/test_fem.rs:5-64: a = [EA/L], b = [100] -> u = 0.0015);
# PARITY UNPINNED beyond that (the crate's source is not available).
# ---------------------------------------------------------------------------------------------
def colsol(a, maxa, b):
    a = np.array(a, np.float64); v = np.array(b, np.float64); maxa = np.asarray(maxa, np.int64)
    nn = len(maxa) - 1
    for n in range(nn):                                   # factorisation: K = L D L^T
        kn = maxa[n]; kl = kn + 1; ku = maxa[n + 1] - 1; kh = ku - kl
        if kh > 0:
            k = n - kh; klt = ku
            for ic in range(1, kh + 1):
                klt -= 1
                ki = maxa[k]; nd = maxa[k + 1] - ki - 1
                if nd > 0:
                    kk = min(ic, nd)
                    a[klt] -= a[ki + 1:ki + kk + 1] @ a[klt + 1:klt + kk + 1]
                k += 1
        if kh >= 0:
            rows = n - 1 - np.arange(ku - kl + 1)
            c = a[kl:ku + 1] / a[maxa[rows]]
            a[kn] -= c @ a[kl:ku + 1]
            a[kl:ku + 1] = c
        if not a[kn] > 0.0:
            raise SeparationError(f"stiffness matrix not positive definite (equation {n + 1})")
    for n in range(nn):                                   # forward reduction
        kl = maxa[n] + 1; ku = maxa[n + 1] - 1
        if ku - kl >= 0:
            v[n] -= a[kl:ku + 1] @ v[n - 1 - np.arange(ku - kl + 1)]
    v /= a[maxa[:-1]]
    for n in range(nn - 1, 0, -1):                        # back-substitution
        kl = maxa[n] + 1; ku = maxa[n + 1] - 1
        if ku - kl >= 0:
            v[n - 1 - np.arange(ku - kl + 1)] -= a[kl:ku + 1] * v[n]
    return v
