"""Sparse separation of the assembled matrix: FEM::add_displacement,
add_concentrated_load, separate_stiffness_matrix_sparse_iterative, find_b_sparse.

CPU tests pin the oracle restatement on the reference's own test model
(/root/reference/src/tests/fem/test_fem.rs:65-150) and cover the host-side checks; GPU tests compare the
device separation with the oracle, bit for bit (index work and copied values)."""
import numpy as np
import pytest
import scipy.sparse as sp

from finite_element_method_b200 import FEM, DOFParameter, FemError, meshes
from oracle import oracle as O


# ---------------------------------------------------------------------------- oracle, pinned
def test_oracle_separation_on_the_reference_test_model():
    """test_fem.rs:65-80 / :83-150: nodes (0,0,0), (30,0,0), truss E=1e6 A=2, u1x = 0, F2x = 100.
    K has +-EA/L at rows/cols {0, 6}; every other DOF has a zero diagonal and is inactive."""
    k = 1e6 * 2.0 / 30.0
    rows, cols, vals = [0, 0, 6, 6], [0, 6, 0, 6], [k, -k, -k, k]
    constrained = np.zeros(12, bool); constrained[0] = True
    forces = np.zeros(12); forces[6] = 100.0
    sep = O.separate_sparse(12, rows, cols, vals, constrained, [1, 2], forces, np.zeros(12))
    assert sep["n_aa"] == 1 and sep["n_bb"] == 1          # assert!(sep.get_n_aa() > 0)
    assert list(sep["k_aa_indexes"]) == [6] and list(sep["k_bb_indexes"]) == [0]
    assert len(sep["k_aa"][0]) == 1                        # assert!(!sep.get_k_aa_triplets().is_empty())
    assert sep["k_aa"][2][0] == k and sep["k_ab"][2][0] == -k and sep["k_ba"][2][0] == -k and sep["k_bb"][2][0] == k
    # the reference's expected result: u2x = 0.0014999999 (its f32 chain), one Jacobi-PCG iteration on a
    # 1x1 system; in f64 the same quotient is 0.0015 to within one f32 ulp (1.2e-10) of that literal
    u = sep["b"][0] / sep["k_aa"][2][0]
    assert abs(u - 0.0014999999) < 2e-10
    # reaction R1x = K_ba u_a + K_bb u_b - F = -100 (find_r_r_sparse)
    assert np.isclose(sep["k_ba"][2][0] * u, -100.0, rtol=1e-15)


def test_oracle_separation_errors():
    k = 5.0
    with pytest.raises(O.SeparationError, match="No restraints"):
        O.separate_sparse(12, [0, 6], [0, 6], [k, k], np.zeros(12, bool))
    c = np.zeros(12, bool); c[1] = True
    with pytest.raises(O.SeparationError, match="There are no stiffness to withstand displacement Y applied to node 7!"):
        O.separate_sparse(12, [0, 6], [0, 6], [k, k], c, [7, 8])
    c = np.zeros(12, bool); c[0] = c[6] = True
    with pytest.raises(O.SeparationError, match="K_aa is empty"):
        O.separate_sparse(12, [0, 6], [0, 6], [k, k], c)


# ---------------------------------------------------------------------------- host logic (no GPU)
def test_add_displacement_host_checks():
    fem = FEM(1e-4, 1e-12, 2, device=-1)
    fem.add_node(1, 0.0, 0.0, 0.0)
    fem.add_node(2, 30.0, 0.0, 0.0)
    fem.add_displacement(1, DOFParameter.X, 0.0)
    with pytest.raises(FemError, match="Displacement X already applied to node 1!"):
        fem.add_displacement(1, DOFParameter.X, 0.5)
    with pytest.raises(FemError, match="Node with number 9 does not exist!"):
        fem.add_displacement(9, DOFParameter.ThZ, 0.0)
    with pytest.raises(FemError, match="Node with number 9 does not exist!"):
        fem.add_concentrated_load(9, DOFParameter.X, 1.0)
    # prefix semantics of the batched form: the first two are kept, the duplicate stops the batch
    with pytest.raises(FemError, match="Displacement Y already applied to node 2!"):
        fem.add_displacement([2, 2, 2, 1], [1, 2, 1, 3], [0.0, 0.0, 0.0, 0.0])
    with pytest.raises(FemError, match="Displacement Z already applied to node 2!"):
        fem.add_displacement(2, DOFParameter.Z, 0.0)
    fem.add_displacement(1, DOFParameter.ThX, 0.0)         # was after the failing entry: not applied yet
    with pytest.raises(FemError) as e:
        fem.separate_stiffness_matrix_sparse_iterative()
    assert "no CPU fallback" in str(e.value)
    fem.reset(2)
    fem.add_node(1, 0.0, 0.0, 0.0)
    fem.add_displacement(1, DOFParameter.X, 0.0)           # reset dropped the constraints
    fem.close()


# ---------------------------------------------------------------------------- GPU parity
def _oracle_for(mesh, fem, constrained, forces, disp):
    n_dof = 6 * len(mesh["x"])
    r, c, v = O.faithful_coo(mesh)
    numbers = np.arange(1, len(mesh["x"]) + 1)
    return O.separate_sparse(n_dof, r, c, v, constrained, numbers, forces, disp)


def _check_against_oracle(mesh, fixed_nodes, fixed_dofs, loads, values=None):
    n_dof = 6 * len(mesh["x"])
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], len(mesh["x"]), device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    values = np.zeros(len(fixed_nodes)) if values is None else np.asarray(values, float)
    fem.add_displacement(np.asarray(fixed_nodes) + 1, fixed_dofs, values)
    constrained = np.zeros(n_dof, bool); disp = np.zeros(n_dof); forces = np.zeros(n_dof)
    constrained[6 * np.asarray(fixed_nodes) + np.asarray(fixed_dofs)] = True
    disp[6 * np.asarray(fixed_nodes) + np.asarray(fixed_dofs)] = values
    for node, dof, val in loads:
        fem.add_concentrated_load(node + 1, dof, val)
        forces[6 * node + dof] += val
    sep = fem.separate_stiffness_matrix_sparse_iterative()
    ref = _oracle_for(mesh, fem, constrained, forces, disp)
    assert np.array_equal(sep.k_aa_indexes, ref["k_aa_indexes"]) and np.array_equal(sep.k_bb_indexes, ref["k_bb_indexes"])
    # K itself carries the assembly's rounding (1e-12 bar of test_parity_gpu); the separation must route
    # every entry to the oracle's (quadrant, i, j). Compare positions exactly and values with that bar.
    _, _, gv = fem.csr()
    scale = np.abs(gv).max()
    for name, quad in (("k_aa", sep.k_aa), ("k_ab", sep.k_ab), ("k_ba", sep.k_ba), ("k_bb", sep.k_bb)):
        i, j, x = sep.triplets(quad)
        ri, rj, rx = ref[name]
        # the GPU assembly may keep entries that are exactly cancelled on one side only: compare on the union
        key = lambda a, b: a * (max(sep.n_aa, sep.n_bb) + 1) + b
        g = dict(zip(key(i, j).tolist(), x.tolist())); o = dict(zip(key(ri, rj).tolist(), rx.tolist()))
        for k_ in set(g) | set(o):
            assert abs(g.get(k_, 0.0) - o.get(k_, 0.0)) <= 1e-12 * scale, (name, k_)
        assert len(set(g) ^ set(o)) <= 0.001 * max(1, len(o)), (name, len(g), len(o))
        assert np.all(np.diff(quad[0]) >= 0) and quad[0][0] == 0 and quad[0][-1] == len(quad[1])
    assert np.allclose(sep.b, ref["b"], rtol=1e-12, atol=1e-12 * max(1.0, np.abs(ref["b"]).max()))
    fem.close()
    return sep


@pytest.mark.gpu
def test_gpu_separation_reference_truss_model():
    mesh = meshes.reference_truss_model()
    fem = FEM(1e-4, 1e-12, 2, device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    fem.add_displacement(1, DOFParameter.X, 0.0)
    fem.add_concentrated_load(2, DOFParameter.X, 100.0)
    sep = fem.separate_stiffness_matrix_sparse_iterative()
    k = 1e6 * 2.0 / 30.0
    assert sep.n_aa == 1 and sep.n_bb == 1 and list(sep.k_aa_indexes) == [6] and list(sep.k_bb_indexes) == [0]
    assert sep.k_aa[2][0] == k and sep.k_ab[2][0] == -k and sep.k_ba[2][0] == -k and sep.k_bb[2][0] == k
    assert abs(sep.b[0] / sep.k_aa[2][0] - 0.0014999999) < 2e-10
    fem.close()


@pytest.mark.gpu
def test_gpu_separation_matches_oracle_truss_cube():
    mesh = meshes.truss_cube(3)
    # pin node 0 fully (translations), node 2 in y/z, node 6 in z; prescribed non-zero settlement on one DOF
    _check_against_oracle(mesh, [0, 0, 0, 2, 2, 6], [0, 1, 2, 1, 2, 2], [(26, 0, 1e3), (13, 2, -5e2), (26, 0, 2.5e2)],
                          values=[0, 0, 0, 0, 1e-3, 0])


@pytest.mark.gpu
def test_gpu_separation_matches_oracle_mixed():
    mesh = meshes.mixed_structure(12, 9)
    w = 13
    fixed = [(n, d) for n in range(w) for d in range(6)]              # clamp the first grid line
    fixed += [(w * 9 + 3, 2), (w * 9 + 7, 4)]
    loads = [(w * 5 + 6, 2, -1e4), (w * 9 + 12, 0, 3e3)]
    sep = _check_against_oracle(mesh, [f[0] for f in fixed], [f[1] for f in fixed], loads,
                                values=[0.0] * (len(fixed) - 2) + [2e-3, -1e-3])
    assert sep.n_aa + sep.n_bb == 6 * len(mesh["x"])                 # plates + beams activate all 6 DOFs


@pytest.mark.gpu
def test_gpu_separation_errors_follow_the_reference():
    mesh = meshes.reference_truss_model()
    fem = FEM(1e-4, 1e-12, 2, device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    with pytest.raises(FemError, match="No restraints"):
        fem.separate_stiffness_matrix_sparse_iterative()
    fem.add_displacement(2, DOFParameter.Y, 0.0)                      # a truss along x has no stiffness in y
    with pytest.raises(FemError, match="There are no stiffness to withstand displacement Y applied to node 2!"):
        fem.separate_stiffness_matrix_sparse_iterative()
    fem.close()
    fem = FEM(1e-4, 1e-12, 2, device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    fem.add_displacement([1, 2], [0, 0], [0.0, 0.0])
    with pytest.raises(FemError, match="K_aa is empty"):
        fem.separate_stiffness_matrix_sparse_iterative()
    fem.close()


@pytest.mark.gpu
def test_gpu_separation_fullsize_properties():
    """Config P (4M plates): size-independent properties — every non-zero of K lands in exactly one
    quadrant (checksum of checksums), the quadrants' shapes add up, K_ab = K_ba^T in count."""
    mesh = meshes.plate_grid(600, 400, "flat")
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], len(mesh["x"]), device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    w = 601
    nodes = np.repeat(np.arange(w), 6); dofs = np.tile(np.arange(6), w)
    fem.add_displacement(nodes + 1, dofs, np.zeros(len(nodes)))
    sep = fem.separate_stiffness_matrix_sparse_iterative()
    rp, ci, v = fem.csr()
    assert sep.n_aa + sep.n_bb == 6 * len(mesh["x"]) and sep.n_bb == 6 * w
    nnz = [len(q[2]) for q in (sep.k_aa, sep.k_ab, sep.k_ba, sep.k_bb)]
    assert sum(nnz) == int(np.count_nonzero(v)) and nnz[1] == nnz[2]
    total = sum(float(np.sum(q[2])) for q in (sep.k_aa, sep.k_ab, sep.k_ba, sep.k_bb))
    assert abs(total - float(np.sum(v))) <= 1e-9 * float(np.sum(np.abs(v)))
    assert np.array_equal(np.sort(np.concatenate([sep.k_aa_indexes, sep.k_bb_indexes])), np.arange(6 * len(mesh["x"])))
    fem.close()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["plates", "mixed", "long-rows"])
def test_one_pass_separation_equals_count_and_fill(which, monkeypatch):
    """FEMGPU_SEP_ONE_PASS=1: the separation reads K from HBM once (tiles of 128 rows chained by a scan over the
    tiles); count + fill is the default. Both must give the same four CSR quadrants, indexes and right-hand side, bit for bit — on meshes of
    tens of thousands of tiles, with constrained DOFs spread over the model (large K_ab / K_ba / K_bb), and with rows
    longer than the 64 entries a warp keeps in registers."""
    if which == "plates":
        mesh = meshes.plate_grid(600, 400, "flat")
    elif which == "mixed":
        mesh = meshes.mixed_structure(300, 250, variant="x0")
    else:
        mesh = meshes.truss_lattice(40, 10**9, jitter=True)     # lattice + body diagonals: up to 8 neighbours x 3 ... and
        # a hub: every 50th node (53, 103, ...) tied to node 0 -> one row of several thousand entries, its neighbours above 64
        hub = np.arange(53, len(mesh["x"]), 50, dtype=np.uint32)
        mesh["t_n1"] = np.concatenate([mesh["t_n1"], np.zeros(len(hub), np.uint32)])
        mesh["t_n2"] = np.concatenate([mesh["t_n2"], hub])
        mesh["t_E"] = np.concatenate([mesh["t_E"], np.full(len(hub), 2.1e11)])
        mesh["t_A"] = np.concatenate([mesh["t_A"], np.full(len(hub), 1e-4)])
    n = len(mesh["x"])
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], n, device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    rng = np.random.default_rng(3)
    nodes = np.sort(rng.choice(n, n // 7, replace=False))
    ndof = 3 if which == "long-rows" else 6                      # trusses carry no rotational stiffness
    fem.add_displacement(np.repeat(nodes + 1, ndof), np.tile(np.arange(ndof), len(nodes)), rng.normal(0, 1e-3, ndof * len(nodes)))
    free = np.setdiff1d(np.arange(n), nodes)[::11]
    fem.add_concentrated_load(free + 1, np.full(len(free), 2), rng.normal(0, 1e3, len(free)))
    monkeypatch.setenv("FEMGPU_SEP_ONE_PASS", "1")
    one = fem.separate_stiffness_matrix_sparse_iterative()
    assert fem.last_separation_read_k_once()
    monkeypatch.delenv("FEMGPU_SEP_ONE_PASS")
    two = fem.separate_stiffness_matrix_sparse_iterative()
    assert not fem.last_separation_read_k_once()
    assert np.array_equal(one.k_aa_indexes, two.k_aa_indexes) and np.array_equal(one.k_bb_indexes, two.k_bb_indexes)
    for a, b in zip((one.k_aa, one.k_ab, one.k_ba, one.k_bb), (two.k_aa, two.k_ab, two.k_ba, two.k_bb)):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert np.array_equal(one.b, two.b)
    assert min(len(q[2]) for q in (one.k_aa, one.k_ab, one.k_ba, one.k_bb)) > 1000
    if which == "long-rows":
        rp = fem.csr()[0]
        assert np.diff(rp).max() > 1000
    monkeypatch.setenv("FEMGPU_SEP_ONE_PASS", "1")
    again = fem.separate_stiffness_matrix_sparse_iterative()     # and deterministic
    assert fem.last_separation_read_k_once() and np.array_equal(again.k_aa[2], one.k_aa[2]) and np.array_equal(again.k_aa[0], one.k_aa[0])
    fem.close()


# ---------------------------------------------------------------------------- distributed loads (§8f rank 2)
def test_oracle_distributed_loads_analytic():
    """beam.rs:775-797 / plate.rs:1145-1185: a uniform load q on a straight member of length L gives
    qL/2 per node; on a quadrilateral of area A the four nodal loads add up to qA (and are qA/4 each on
    a rectangle)."""
    assert np.allclose(O.beam_line_load([0, 0, 0], [3, 4, 0], 10.0), [25.0, 25.0], rtol=1e-15)
    assert np.allclose(O.plate_surface_load([2, 1.5, 0], [0, 1.5, 0], [0, 0, 0], [2, 0, 0], 4.0), 3.0, rtol=1e-15)
    p = [[2.2, 1.4, 0.3], [0.1, 1.6, 0.3], [0, 0, 0.3], [2.0, -0.1, 0.3]]
    x, y = np.array([q[0] for q in p]), np.array([q[1] for q in p])
    area = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
    f = O.plate_surface_load(*p, 4.0)
    assert abs(f.sum() - 4.0 * area) < 1e-13 and f.min() > 0


def test_distributed_load_host_checks():
    fem = FEM(1e-4, 1e-12, 4, device=-1)
    for i, (x, y) in enumerate([(1, 1), (0, 1), (0, 0), (1, 0)]):
        fem.add_node(i + 1, float(x), float(y), 0.0)
    with pytest.raises(FemError, match="Beam element with number 3 does not exist!"):
        fem.add_uniformly_distributed_line_load(3, DOFParameter.Y, 1.0)
    with pytest.raises(FemError, match="Plate element with number 1 does not exist!"):
        fem.add_uniformly_distributed_surface_load(1, DOFParameter.Z, 1.0)
    fem.close()


def _expected_forces(mesh, line, surface, point):
    """The reference's sequence of `+=` into the forces vector, with the oracle's nodal loads."""
    F = np.zeros(6 * len(mesh["x"]))
    P = np.stack([mesh["x"], mesh["y"], mesh["z"]], axis=1)
    for node, dof, val in point:
        F[6 * node + dof] += val
    for e, dof, q in line:
        a, b = int(mesh["b_n1"][e]), int(mesh["b_n2"][e])
        f = O.beam_line_load(P[a], P[b], q)
        F[6 * a + dof] += f[0]; F[6 * b + dof] += f[1]
    pn = np.asarray(mesh["p_n"]).reshape(4, -1)
    for e, dof, q in surface:
        n = [int(pn[k][e]) for k in range(4)]
        f = O.plate_surface_load(*[P[k] for k in n], q, mesh["rel_tol"], mesh["abs_tol"])
        for k in range(4):
            F[6 * n[k] + dof] += f[k]
    return F


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["mixed", "jitter", "x0"])
def test_gpu_distributed_loads_match_oracle(which):
    rng = np.random.default_rng(7)
    if which == "mixed":
        mesh = meshes.mixed_structure(10, 8)
    else:
        mesh = meshes.plate_grid(9, 7, which)
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], len(mesh["x"]), device=0)
    fem.load_mesh(mesh)
    npl = np.asarray(mesh["p_n"]).reshape(4, -1).shape[1]
    nb = len(mesh["b_n1"])
    # every plate gets a pressure, a third of them a second (in-plane) load; every other beam a line load
    surface = [(e, 2, float(rng.uniform(-5e3, 5e3))) for e in range(npl)] + \
              [(e, 0, float(rng.uniform(-1e3, 1e3))) for e in range(0, npl, 3)]
    line = [(e, int(rng.integers(0, 6)), float(rng.uniform(-2e3, 2e3))) for e in range(0, nb, 2)]
    point = [(3, 1, 750.0), (len(mesh["x"]) - 1, 2, -125.0), (3, 1, 250.0)]
    first_plate = 1                              # load_mesh numbers every family from 1
    for node, dof, val in point:
        fem.add_concentrated_load(node + 1, dof, val)
    if line:
        fem.add_uniformly_distributed_line_load([e + 1 for e, _, _ in line], [d for _, d, _ in line], [q for _, _, q in line])
    fem.add_uniformly_distributed_surface_load([e + first_plate for e, _, _ in surface], [d for _, d, _ in surface],
                                               [q for _, _, q in surface])
    F = fem.forces_vector()
    ref = _expected_forces(mesh, line, surface, point)
    scale = np.abs(ref).max()
    assert np.abs(F - ref).max() <= 1e-12 * scale, np.abs(F - ref).max() / scale
    assert np.array_equal(F, fem.forces_vector())                     # re-evaluation is bit-identical
    # the distributed loads reach the right-hand side of the separated system
    fem.assemble()
    w = int(round(mesh["x"].max())) + 1 if which != "x0" else 10
    fixed = np.repeat(np.arange(w), 6)
    fem.add_displacement(fixed + 1, np.tile(np.arange(6), w), np.zeros(len(fixed)))
    sep = fem.separate_stiffness_matrix_sparse_iterative()
    assert np.abs(sep.b - ref[sep.k_aa_indexes]).max() <= 1e-12 * scale   # u_b = 0: b = R_a
    fem.close()


@pytest.mark.gpu
def test_loads_of_all_kinds_sum_in_call_order():
    """methods_for_bc_data_handle.rs:47-53, 81-98, 126-172: every load kind does `+=` into the forces vector when it is
    added, so a DOF's value is the left-to-right sum over ALL calls, whatever their kind. Values of very different
    magnitude make the order visible in the last bits (concentrated loads used to be summed first)."""
    mesh = meshes.mixed_structure(4, 3)
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], len(mesh["x"]), device=0)
    fem.load_mesh(mesh)
    P = np.stack([mesh["x"], mesh["y"], mesh["z"]], axis=1)
    pn = np.asarray(mesh["p_n"]).reshape(4, -1)
    node = int(pn[0][0])                                   # node 1 of plate 0: receives every load below
    beam = int(np.flatnonzero(mesh["b_n1"] == node)[0]) if np.any(mesh["b_n1"] == node) else int(np.flatnonzero(mesh["b_n2"] == node)[0])
    calls = [("surface", 0, 1.0e-3), ("point", node, 1.0e13), ("line", beam, 0.7), ("point", node, -1.0e13),
             ("surface", 0, 3.0e-3), ("point", node, 0.1), ("line", beam, -0.3)]
    F = np.zeros(6 * len(mesh["x"]))
    for kind, who, val in calls:                           # the reference's sequence of `+=`
        if kind == "point":
            fem.add_concentrated_load(who + 1, 2, val)
            F[6 * who + 2] += val
        elif kind == "line":
            fem.add_uniformly_distributed_line_load(who + 1, 2, val)
            a, b = int(mesh["b_n1"][who]), int(mesh["b_n2"][who])
            f = O.beam_line_load(P[a], P[b], val)
            F[6 * a + 2] += f[0]; F[6 * b + 2] += f[1]
        else:
            fem.add_uniformly_distributed_surface_load(who + 1, 2, val)
            n = [int(pn[k][who]) for k in range(4)]
            f = O.plate_surface_load(*[P[k] for k in n], val, mesh["rel_tol"], mesh["abs_tol"])
            for k in range(4):
                F[6 * n[k] + 2] += f[k]
    got = fem.forces_vector()
    assert np.abs(got - F).max() <= 1e-10, (got[6 * node + 2], F[6 * node + 2])
    # the test is sensitive to the order: with the concentrated loads summed first (1e13 - 1e13 + 0.1 exactly, then the
    # distributed shares) this DOF would come out differently, far above the bar
    alt = 0.0
    for kind, who, val in calls:
        if kind == "point":
            alt += val
    dist_share = 0.0
    for kind, who, val in calls:
        if kind == "line":
            a, b = int(mesh["b_n1"][who]), int(mesh["b_n2"][who])
            f = O.beam_line_load(P[a], P[b], val)
            dist_share += f[0] if a == node else f[1]
        elif kind == "surface":
            f = O.plate_surface_load(*[P[int(pn[k][who])] for k in range(4)], val, mesh["rel_tol"], mesh["abs_tol"])
            dist_share += f[0]
    assert abs((alt + dist_share) - F[6 * node + 2]) > 1e-6
    fem.close()


# ---------------------------------------------------------------------------- direct separation (skyline)
def test_oracle_direct_separation_on_the_reference_test_model():
    """test_fem.rs:5-64 (the direct path): one free DOF -> K_aa = [EA/L], skyline [0], a = [EA/L], maxa = [0, 1]"""
    k = 1e6 * 2.0 / 30.0
    c = np.zeros(12, bool); c[0] = True
    f = np.zeros(12); f[6] = 100.0
    d = O.separate_direct(12, [0, 0, 6, 6], [0, 6, 0, 6], [k, -k, -k, k], c, [1, 2], f)
    assert list(d["k_aa_indexes"]) == [6] and list(d["k_bb_indexes"]) == [0]
    assert list(d["k_aa_skyline"]) == [0] and list(d["a"]) == [k] and list(d["maxa"]) == [0, 1]
    assert d["a"][0] and 100.0 / d["a"][0] == 0.0014999999999999998        # colsol on a 1 x 1 system
    f[7] = 1.0                                                               # a load on a DOF without stiffness
    with pytest.raises(O.SeparationError, match="There are no stiffness to withstand load Y applied to node 2!"):
        O.separate_direct(12, [0, 0, 6, 6], [0, 6, 0, 6], [k, -k, -k, k], c, [1, 2], f)
    with pytest.raises(O.SeparationError, match="There are no restraints applied!"):
        O.separate_direct(12, [0, 0, 6, 6], [0, 6, 0, 6], [k, -k, -k, k], np.zeros(12, bool), [1, 2])


def test_oracle_skyline_form():
    # 4 free DOFs, entries (0,2) and (1,3) above the diagonal -> heights [0, 0, 2, 2]
    r = [0, 1, 2, 3, 0, 2, 1, 3, 4]; c_ = [0, 1, 2, 3, 2, 0, 3, 1, 4]; v = [4., 5., 6., 7., 1., 1., 2., 2., 9.]
    con = np.zeros(5, bool); con[4] = True
    d = O.separate_direct(5, r, c_, v, con)
    assert list(d["k_aa_skyline"]) == [0, 0, 2, 2]
    assert list(d["maxa"]) == [0, 1, 2, 5, 8]
    assert list(d["a"]) == [4., 5., 6., 0., 1., 7., 0., 2.]


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["reference", "mixed", "beams"])
def test_direct_separation_matches_oracle(which):
    if which == "reference":
        mesh = meshes.reference_truss_model()
        fixed_nodes, fixed_dofs = [0], [0]
    else:
        mesh = meshes.mixed_structure(7, 5) if which == "mixed" else meshes.beam_frame(4, 10**9)
        y0 = np.flatnonzero(np.asarray(mesh["y"]) == np.min(mesh["y"]))
        fixed_nodes, fixed_dofs = np.repeat(y0, 6), np.tile(np.arange(6), len(y0))
    n = len(mesh["x"])
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], n, device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    fem.add_displacement(np.asarray(fixed_nodes) + 1, fixed_dofs, np.zeros(len(fixed_nodes)))
    fem.add_concentrated_load(n, DOFParameter.X, 100.0)
    ia, ib, sky, a, maxa = fem.separate_stiffness_matrix_direct()
    rp, ci, v = fem.csr()
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    con = np.zeros(6 * n, bool); con[6 * np.asarray(fixed_nodes) + np.asarray(fixed_dofs)] = True
    f = np.zeros(6 * n); f[6 * (n - 1)] = 100.0
    ref = O.separate_direct(6 * n, rows, ci, v, con, np.arange(1, n + 1), f)       # same K: bit-for-bit
    assert np.array_equal(ia, ref["k_aa_indexes"]) and np.array_equal(ib, ref["k_bb_indexes"])
    assert np.array_equal(sky, ref["k_aa_skyline"]) and np.array_equal(maxa, ref["maxa"])
    assert np.array_equal(a, ref["a"])
    if which == "reference":
        assert list(sky) == [0] and a[0] == 66666.66666666667
    # the reference's SeparatedStiffnessMatrix (structs/separated_stiffness_matrix.rs): same getters, dense quadrants
    sep = fem.separate_stiffness_matrix_direct()
    assert np.array_equal(sep.get_k_aa_indexes(), ia) and np.array_equal(sep.get_k_bb_indexes(), ib)
    assert np.array_equal(sep.get_k_aa_skyline(), sky)
    K = sp.csr_matrix((v, ci, rp), shape=(6 * n, 6 * n)).toarray()
    for got, (r_idx, c_idx) in ((sep.get_k_aa_matrix(), (ia, ia)), (sep.get_k_ab_matrix(), (ia, ib)),
                                (sep.get_k_ba_matrix(), (ib, ia)), (sep.get_k_bb_matrix(), (ib, ib))):
        assert got.shape == (len(r_idx), len(c_idx))
        assert np.array_equal(got, K[np.ix_(r_idx, c_idx)])        # the values of K itself, bit for bit
    fem.close()


@pytest.mark.gpu
def test_direct_separation_errors():
    mesh = meshes.reference_truss_model()
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], 2, device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    with pytest.raises(FemError, match="There are no restraints applied!"):
        fem.separate_stiffness_matrix_direct()
    fem.add_displacement(1, DOFParameter.X, 0.0)
    fem.add_concentrated_load(2, DOFParameter.Y, 5.0)        # the truss has no stiffness in y
    with pytest.raises(FemError, match="There are no stiffness to withstand load Y applied to node 2!"):
        fem.separate_stiffness_matrix_direct()
    fem.separate_stiffness_matrix_sparse_iterative()         # the sparse variant does not look at the loads
    fem.close()


# ---------------------------------------------------------------------------- direct solve (skyline LDL^T)
def test_oracle_colsol():
    """the reference's direct test (test_fem.rs:5-64) is a 1 x 1 system; larger ones against a dense solve"""
    k = 66666.66666666667
    assert O.colsol([k], [0, 1], [100.0])[0] == 100.0 / k == 0.0014999999999999998
    rng = np.random.default_rng(11)
    n = 40
    A = np.zeros((n, n))
    for i in range(n):
        for j in range(max(0, i - rng.integers(0, 6)), i):
            A[i, j] = A[j, i] = rng.normal()
    A += np.diag(np.abs(A).sum(axis=1) + 1.0)              # SPD, variable column heights
    sky = np.array([max([j - i for i in range(j) if A[i, j] != 0.0] + [0]) for j in range(n)])
    maxa = np.concatenate([[0], np.cumsum(sky + 1)])
    a = np.concatenate([[A[j - m, j] for m in range(sky[j] + 1)] for j in range(n)])
    b = rng.normal(size=n)
    u = O.colsol(a, maxa, b)
    assert np.linalg.norm(u - np.linalg.solve(A, b)) <= 1e-12 * np.linalg.norm(u)
    with pytest.raises(O.SeparationError, match="not positive definite"):
        O.colsol([1.0, 1.0, 2.0], [0, 1, 3], [1.0, 1.0])   # [[1, 2], [2, 1]]: column 1 = (diagonal 1, above it 2)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["reference", "mixed", "beams"])
def test_direct_solve_matches_oracle_and_dense(which):
    """the reference's direct flow (test_fem.rs:5-64): separate_stiffness_matrix_direct -> find_ua_vector_direct ->
    find_r_r_vector -> compose_global_analysis_result"""
    if which == "reference":
        mesh = meshes.reference_truss_model()
        fixed_nodes, fixed_dofs = [0], [0]
    else:
        mesh = meshes.mixed_structure(7, 5) if which == "mixed" else meshes.beam_frame(4, 10**9)
        y0 = np.flatnonzero(np.asarray(mesh["y"]) == np.min(mesh["y"]))
        fixed_nodes, fixed_dofs = np.repeat(y0, 6), np.tile(np.arange(6), len(y0))
    n = len(mesh["x"])
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], n, device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    fem.add_displacement(np.asarray(fixed_nodes) + 1, fixed_dofs, np.zeros(len(fixed_nodes)))
    fem.add_concentrated_load(n, DOFParameter.X, 100.0)
    ia, ib, sky, a, maxa = fem.separate_stiffness_matrix_direct()
    sep = fem.separate_stiffness_matrix_sparse_iterative()    # the same quadrants, for b and a dense check
    fem.separate_stiffness_matrix_direct()
    u = fem.find_ua_vector_direct()
    u_ref = O.colsol(a, maxa, sep.b)
    assert np.linalg.norm(u - u_ref) <= 1e-9 * np.linalg.norm(u_ref)          # same algorithm, other dot order
    i, j, v = sep.triplets(sep.k_aa)
    A = sp.csr_matrix((v, (i, j)), shape=(sep.n_aa, sep.n_aa))
    assert np.linalg.norm(A @ u - sep.b) <= 1e-9 * np.linalg.norm(sep.b)
    assert np.array_equal(u, fem.find_ua_vector_direct())                     # deterministic
    r_r = fem.find_r_r_vector()
    d, f = fem.global_analysis_vectors()
    assert abs(f[0::6].sum()) <= 1e-7 * 100.0 and np.array_equal(d[ia], u)
    if which == "reference":
        assert u[0] == 100.0 / 66666.66666666667 and abs(r_r[0] + 100.0) < 1e-12   # 0.0015, -100 (test_fem.rs:40-58)
        assert abs(fem.extract_elements_analysis_result()[0][1][0][1] - 100.0) < 1e-11
    fem.close()
