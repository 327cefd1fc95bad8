"""Global analysis downstream of the separated matrix: the PCG solve, reactions,
composed result vectors and the element result recovery.

CPU tests cover the host-side usage errors; GPU tests replay the reference's own end-to-end test
(/root/reference/src/tests/fem/test_fem.rs:83-225: u2x = 0.0015, R1x = -100, ForceR = 100, one iteration for
both preconditioners) through the C ABI and compare larger models with the oracle restatement
(oracle/fem_oracle.hpp *_element_result, oracle/oracle.py pcg / reactions).

Tolerances: element forces within 1e-12 of the largest force of that component in the model (they are
differences of nearly equal displacements: the bar is relative to the scale, like the assembly's block floor);
PCG solutions are compared through the residual the stopping test uses and, against the oracle's PCG, by
iteration count (+-1: the reductions are summed in a different order) and solution difference."""
import numpy as np
import pytest

from finite_element_method_b200 import BEAM, PLATE, TRUSS, FEM, DOFParameter, FemError, meshes
from oracle import oracle as O


# ---------------------------------------------------------------------------- host logic (no GPU)
def test_analysis_needs_a_device_and_a_separated_matrix():
    fem = FEM(1e-4, 1e-12, 2, device=-1)
    fem.add_node(1, 0.0, 0.0, 0.0)
    fem.add_node(2, 30.0, 0.0, 0.0)
    with pytest.raises(FemError) as e:
        fem._solve(0, 10, True)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(FemError) as e:
        fem.element_results(TRUSS)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(FemError) as e:
        fem.compose_global_analysis_result()
    assert "no CPU fallback" in str(e.value)
    assert list(fem.node_numbers()) == [1, 2]
    fem.close()


# ---------------------------------------------------------------------------- GPU
def _reference_model():
    fem = FEM(1e-4, 1e-12, 2, device=0)
    fem.add_node(1, 0.0, 0.0, 0.0)
    fem.add_node(2, 30.0, 0.0, 0.0)
    fem.add_truss(1, 1, 2, 1e6, 2.0, None)
    fem.add_displacement(1, DOFParameter.X, 0.0)
    fem.add_concentrated_load(2, DOFParameter.X, 100.0)
    fem.assemble()
    return fem


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["jacobi", "block_jacobi"])
def test_reference_integration_test(solver):
    """test_fem.rs:83-150 (Jacobi) and :153-225 (block Jacobi), in f64"""
    fem = _reference_model()
    sep = fem.separate_stiffness_matrix_sparse_iterative()
    assert sep.n_aa == 1 and sep.n_bb == 1
    solve = fem.find_ua_vector_iterative_pcg_jacobi_sparse if solver == "jacobi" else \
        fem.find_ua_vector_iterative_pcg_block_jacobi_sparse
    u_a, iterations = solve(1000)
    assert iterations == 1                                        # assert_eq!(iterations, 1)
    assert u_a[0] == 100.0 / 66666.66666666667                    # 0.0015 (the reference's f32 chain prints 0.0014999999)
    assert abs(u_a[0] - 0.0014999999) < 2e-10
    r_r = fem.find_r_r_vector_sparse()
    assert abs(r_r[0] + 100.0) < 1e-12                            # (1, X, 0.0, -100.0)
    fem.compose_global_analysis_result()
    res = fem.extract_global_analysis_result()
    expected = {(1, 0): (0.0, -100.0), (2, 0): (u_a[0], 100.0)}
    for number, dof, d, f in res:
        ed, ef = expected.get((number, dof), (0.0, 0.0))
        assert d == ed and abs(f - ef) < 1e-12, (number, dof, d, f)
    elems = fem.extract_elements_analysis_result()
    assert len(elems) == 1 and elems[0][0] == 1 and elems[0][1][0][0] == "ForceR"
    assert abs(elems[0][1][0][1] - 100.0) < 1e-11                 # (1, [(ForceR, 100.0)])
    fem.close()


def _compare_results(mesh, seed):
    n = len(mesh["x"])
    rng = np.random.default_rng(seed)
    u = rng.normal(size=6 * n) * 1e-3
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], n, device=0)
    fem.load_mesh(mesh)
    fem.set_displacements_vector(u)
    ot, ob, op = O.element_results(mesh, u)
    for family, ref in ((TRUSS, ot.reshape(-1, 1)), (BEAM, ob), (PLATE, op)):
        got = fem.element_results(family)
        assert got.shape == ref.shape
        if ref.size == 0:
            continue
        scale = np.abs(ref).max(axis=0)
        err = np.abs(got - ref).max(axis=0)
        assert np.all(err <= 1e-12 * np.maximum(scale, 1e-300)), (mesh["name"], family, err / scale)
        again = fem.element_results(family)
        assert np.array_equal(got, again)                        # deterministic
    fem.close()


@pytest.mark.gpu
def test_element_results_match_oracle():
    for k, mesh in enumerate([meshes.truss_cube(3), meshes.truss_lattice(6, 10**9, jitter=True),
                              meshes.beam_frame(5, 10**9), meshes.beam_frame(5, 10**9, jitter=True),
                              meshes.plate_grid(6, 5, "flat"), meshes.plate_grid(6, 5, "jitter"),
                              meshes.plate_grid(6, 5, "x0"), meshes.mixed_structure(12, 9)]):
        _compare_results(mesh, 100 + k)


@pytest.mark.gpu
def test_element_results_numbers_follow_insertion_order():
    mesh = meshes.mixed_structure(4, 3)
    n = len(mesh["x"])
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], n, device=0)
    fem.load_mesh(mesh)
    fem.set_displacements_vector(np.zeros(6 * n))
    out = fem.extract_elements_analysis_result()
    nt, nb, npl = len(mesh["t_n1"]), len(mesh["b_n1"]), len(mesh["p_n"][0])
    assert len(out) == nt + nb + npl
    assert [len(v) for _, v in out] == [1] * nt + [10] * nb + [8] * npl
    assert all(val == 0.0 for _, v in out for _, val in v)
    fem.close()


def _clamped_model(mesh, rel_tol):
    """clamp the first grid line of nodes, load the last node in z and x"""
    n = len(mesh["x"])
    fem = FEM(rel_tol, mesh["abs_tol"], n, device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    y0 = np.flatnonzero(np.asarray(mesh["y"]) == np.min(mesh["y"]))
    nodes = np.repeat(y0, 6)
    dofs = np.tile(np.arange(6), len(y0))
    vals = np.zeros(len(nodes)); vals[2::6] = 1e-3                # a prescribed settlement in z
    fem.add_displacement(nodes + 1, dofs, vals)
    fem.add_concentrated_load(n, DOFParameter.Z, -500.0)
    fem.add_concentrated_load(n, DOFParameter.X, 250.0)
    n_dof = 6 * n
    constrained = np.zeros(n_dof, bool); constrained[6 * nodes + dofs] = True
    disp = np.zeros(n_dof); disp[6 * nodes + dofs] = vals
    forces = np.zeros(n_dof); forces[6 * (n - 1) + 2] = -500.0; forces[6 * (n - 1)] = 250.0
    return fem, constrained, disp, forces


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["mixed", "plates_x0", "beams"])
def test_pcg_solve_reactions_and_results(which):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    mesh = {"mixed": lambda: meshes.mixed_structure(10, 8), "plates_x0": lambda: meshes.plate_grid(8, 6, "x0"),
            "beams": lambda: meshes.beam_frame(4, 10**9)}[which]()
    rel_tol = 1e-10
    # the drilling stiffness of a plate node is 1.0 next to membrane terms of 1e9: the plate systems are
    # ill-conditioned, so the distance to a direct solve is only a sanity bound there; the contract is the residual
    direct_tol = 1e-5 if which == "beams" else 1e-2
    fem, constrained, disp, forces = _clamped_model(mesh, rel_tol)
    sep = fem.separate_stiffness_matrix_sparse_iterative()
    i, j, v = sep.triplets(sep.k_aa)
    A = sp.csr_matrix((v, (i, j)), shape=(sep.n_aa, sep.n_aa))
    u_direct = spl.spsolve(A.tocsc(), sep.b)
    starts = O.block_starts_from_k_aa_indexes(sep.k_aa_indexes)
    for name, solve, st in (("jacobi", fem.find_ua_vector_iterative_pcg_jacobi_sparse, None),
                            ("block", fem.find_ua_vector_iterative_pcg_block_jacobi_sparse, starts)):
        u_a, it = solve(20000)
        res = np.linalg.norm(A @ u_a - sep.b)
        assert res <= 1.01 * max(rel_tol * np.linalg.norm(sep.b), mesh["abs_tol"]), (name, res)
        u_ref, it_ref = O.pcg(sep.n_aa, (i, j, v), sep.b, 20000, rel_tol, mesh["abs_tol"], st)
        assert abs(it - it_ref) <= 2 + it_ref // 10, (name, it, it_ref)
        assert np.linalg.norm(u_a - u_direct) <= direct_tol * np.linalg.norm(u_direct), name
        u2, it2 = solve(20000)
        assert it2 == it and np.array_equal(u2, u_a)              # deterministic, bit for bit
        assert fem.solve_info()[0] == it
    # reactions and the composed vectors against the restatement, from the same u_a
    r_r = fem.find_r_r_vector_sparse()
    osep = {"n_bb": sep.n_bb, "k_bb_indexes": sep.k_bb_indexes, "k_aa_indexes": sep.k_aa_indexes,
            "k_ba": sep.triplets(sep.k_ba), "k_bb": sep.triplets(sep.k_bb)}
    rr_ref = O.reactions(osep, u_a, forces, disp)
    # a reaction is a sum of products ~1e6 that cancel to ~1e2: the bar is relative to the size of the terms
    bi, bj, bv = osep["k_ba"]
    terms = np.zeros(sep.n_bb); np.add.at(terms, bi, np.abs(bv * u_a[bj]))
    assert np.abs(r_r - rr_ref).max() <= 1e-12 * max(terms.max(), np.abs(rr_ref).max())
    d, f = fem.global_analysis_vectors()
    d_ref, f_ref = O.compose_global_analysis_result(osep, u_a, r_r, forces, disp)
    assert np.array_equal(d, d_ref) and np.array_equal(f, f_ref)
    # global equilibrium: applied loads + reactions sum to zero in every translational direction
    for k in range(3):
        assert abs(f[k::6].sum()) <= 1e-6 * np.abs(f).max()
    # element results from the composed displacements
    ot, ob, op = O.element_results(mesh, d)
    for family, ref in ((TRUSS, ot.reshape(-1, 1)), (BEAM, ob), (PLATE, op)):
        got = fem.element_results(family)
        if ref.size:
            assert np.all(np.abs(got - ref).max(axis=0) <= 1e-12 * np.abs(ref).max(axis=0) + 1e-300), family
    fem.close()


@pytest.mark.gpu
def test_external_solution_and_errors():
    fem = _reference_model()
    with pytest.raises(FemError, match="no separated matrix"):
        fem._solve(0, 10, True)
    fem.separate_stiffness_matrix_sparse_iterative()
    with pytest.raises(FemError, match="no u_a"):
        fem.compose_global_analysis_result()
    fem.set_u_a_vector([0.0015])                                  # e.g. from a direct solver
    r_r = fem.find_r_r_vector_sparse()
    assert abs(r_r[0] + 100.0) < 1e-9
    with pytest.raises(FemError, match="did not converge"):
        fem._solve(0, 0, True)
    # a new boundary condition invalidates the separated matrix and everything derived from it
    fem.add_concentrated_load(2, DOFParameter.X, 1.0)
    with pytest.raises(FemError, match="no separated matrix"):
        fem._solve(0, 10, True)
    fem.close()
