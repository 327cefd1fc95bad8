"""The oracle against the reference's own known answers and analytic checks (CPU only).

Golden values: /root/reference/src/tests/fem/test_fem.rs:42-58 — in f32 the 2-node truss model gives
u2x = 0.0014999999, reaction -100, ForceR = 100. tests/golden/reference_truss.json additionally pins
the f64 numbers the oracle produced in the build container (generator: tests/golden/make_golden.py).
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_truss_model_f32_known_answer():
    k00, u2x, r1x, force_r = O.reference_truss_test_f32()
    assert u2x == np.float32(0.0014999999)          # test_fem.rs:49
    assert r1x == np.float32(-100.0)                # test_fem.rs:42
    assert force_r == np.float32(100.0)             # test_fem.rs:58
    assert abs(float(k00) - 2e6 / 30.0) < 1e-2


def test_reference_truss_model_f64():
    q, kl, kg = O.truss([0, 0, 0], [30, 0, 0], 1e6, 2.0)
    assert np.array_equal(q, np.eye(3))
    assert kg[0, 0] == 66666.66666666667 and kg[0, 3] == -66666.66666666667
    assert 100.0 / kg[0, 0] == 0.0014999999999999998
    assert (kg != 0).sum() == 4                     # only the axial terms survive the zero-skip


def test_golden_fixtures():
    with open(os.path.join(GOLD, "element_golden.json")) as f:
        gold = json.load(f)
    for case in gold["truss"]:
        q, kl, kg = O.truss(case["p1"], case["p2"], case["E"], case["A"], case["A2"])
        assert np.array_equal(kg, np.array(case["kg"]))
    for case in gold["beam"]:
        q, pr, kl, kg = O.beam(case["p1"], case["p2"], *case["props"], case["axis"])
        assert np.array_equal(kg, np.array(case["kg"]))
        assert np.array_equal(q, np.array(case["q"]))
    for case in gold["plate"]:
        q, kl, kg = O.plate(*case["p"], *case["props"])
        assert np.array_equal(kg, np.array(case["kg"]))


def test_plate_gauss_abscissa_is_sqrt_of_f32_third():
    # plate.rs:1066-1091: V::from(1f32/3f32).my_sqrt(), not 1/sqrt(3)
    g = np.sqrt(np.float64(np.float32(1.0) / np.float32(3.0)))
    assert g == 0.5773502777928151
    # a unit square plate integrates x^2 exactly only with the true abscissa; with the f32-rounded
    # one the membrane term differs at the 1e-8 level from the textbook matrix
    p = [[1, 1, 0], [0, 1, 0], [0, 0, 0], [1, 0, 0]]
    E, nu, t = 1.0, 0.25, 1.0
    _, kl, _ = O.plate(*p, E, nu, t, 5 / 6)
    c = E * t / (1 - nu * nu)
    # k[u1,u1] = c*(a + (1-nu)/2*a) with a = integral of (dh1/dx)^2 = (1+3g^2)/6... evaluated at g
    a = 0.25 * (1 + g) ** 2 / 2 + 0.25 * (1 - g) ** 2 / 2
    assert abs(kl[0, 0] - c * (a + (1 - nu) / 2 * a)) < 1e-15


def test_truss_rotation_invariance_and_symmetry():
    rng = np.random.default_rng(1)
    for _ in range(20):
        p1, p2 = rng.normal(size=3), rng.normal(size=3)
        q, kl, kg = O.truss(p1, p2, 2.1e11, 1e-4, 2e-4)
        L = np.linalg.norm(p2 - p1)
        n = (p2 - p1) / L
        A_mid = 1.5e-4   # tapered area at r = 0
        ref = 2.1e11 * A_mid / L * np.outer(n, n)
        assert np.allclose(kg[:3, :3], ref, rtol=1e-12, atol=1e-12 * abs(ref).max())
        assert np.allclose(kg, kg.T, rtol=0, atol=1e-15 * abs(kg).max())
        assert np.allclose(q @ q.T, np.eye(3), atol=1e-12)


def test_beam_cantilever_reduced_integration():
    E, nu, A, I11, I22, It, ks, L = 2.1e11, 0.3, 1e-2, 8e-6, 4e-6, 1e-5, 5 / 6, 2.0
    q, pr, kl, kg = O.beam([0, 0, 0], [L, 0, 0], E, nu, A, I11, I22, 0.0, It, ks, [0, 0, 1])
    assert (kl != 0).sum() == 40
    G = E / (2 * (1 + nu))
    Kff = kg[6:, 6:]
    u = np.linalg.solve(Kff, np.eye(6)[1])
    assert abs(u[1] - (L ** 3 / (4 * E * I11) + L / (ks * G * A))) < 1e-12 * abs(u[1]) * 10
    u = np.linalg.solve(Kff, np.eye(6)[2])
    assert abs(u[2] - (L ** 3 / (4 * E * I22) + L / (ks * G * A))) < 1e-12 * abs(u[2]) * 10
    u = np.linalg.solve(Kff, np.eye(6)[0])
    assert abs(u[0] - L / (E * A)) < 1e-12 * abs(u[0]) * 10


def test_beam_principal_axes_swap():
    # I22 > I11 forces the pi/2 loop (beam.rs:147-157), with PI taken from f32
    q, pr, kl, kg = O.beam([0, 0, 0], [1, 0, 0], 1.0, 0.3, 1.0, 1.0, 3.0, 0.0, 1.0, 1.0, [0, 0, 1])
    assert pr[0] >= pr[1] and abs(pr[0] - 3.0) < 1e-6 and abs(pr[1] - 1.0) < 1e-6
    assert abs(pr[2] - 3.1415927410125732 / 2) < 1e-15


def test_plate_rigid_body_modes_and_symmetry():
    rng = np.random.default_rng(2)
    base = np.array([[1, 0.75, 0], [0, 0.75, 0], [0, 0, 0], [1, 0, 0]], float)
    base[:, :2] += rng.uniform(-0.1, 0.1, (4, 2))
    q, kl, kg = O.plate(*base, 2.1e11, 0.3, 0.01, 5 / 6)
    assert (kl != 0).sum() == 212 and np.array_equal(q, np.eye(3))
    assert abs(kg - kg.T).max() <= 1e-15 * abs(kg).max()
    # translations are exact null vectors (the drilling penalty only touches theta_z)
    for d in range(3):
        v = np.zeros(24); v[d::6] = 1.0
        assert abs(kg @ v).max() <= 1e-12 * abs(kg).max()
    w = np.linalg.eigvalsh((kl + kl.T) / 2)
    assert (np.abs(w) < 1e-3).sum() == 6                  # rigid-body modes
    assert (np.abs(w - 1.0) < 1e-6).sum() == 4            # the four drilling penalties (KROT6 = 1)


def test_plate_in_other_plane_matches_rotated_flat_plate():
    p = np.array([[1, 0.75, 0], [0, 0.75, 0], [0, 0, 0], [1, 0, 0]], float)
    q0, kl0, kg0 = O.plate(*p, 2.1e11, 0.3, 0.01, 5 / 6)
    # same plate in the x = 0 plane: (x, y, 0) -> (0, x, y)
    p2 = np.stack([np.zeros(4), p[:, 0], p[:, 1]], axis=1)
    q, kl, kg = O.plate(*p2, 2.1e11, 0.3, 0.01, 5 / 6)
    assert not np.array_equal(q, np.eye(3))
    # eigenvalues are invariant under the rigid rotation
    assert np.allclose(np.linalg.eigvalsh((kg + kg.T) / 2), np.linalg.eigvalsh((kg0 + kg0.T) / 2),
                       rtol=0, atol=1e-9 * abs(kg0).max())


def test_validation_codes():
    with pytest.raises(O.OracleError) as e:
        O.truss([0, 0, 0], [1, 0, 0], -1.0, 1.0)
    assert e.value.code == 1
    with pytest.raises(O.OracleError) as e:
        O.beam([0, 0, 0], [1, 0, 0], 1, .3, 1, 1, 1, 0, 1, 1, [2, 0, 0])   # axis parallel to element
    assert e.value.code == 9
    flat = [[1, 1, 0], [0, 1, 0], [0, 0, 0], [1, 0, 0]]
    with pytest.raises(O.OracleError) as e:
        O.plate([1, 1, 0.1], *flat[1:], 1, .3, 1, 1)                        # node 1 off the plane
    assert e.value.code == 12
    with pytest.raises(O.OracleError) as e:
        O.plate([2, 0, 0], *flat[1:], 1, .3, 1, 1)                          # nodes 3, 4, 1 collinear
    assert e.value.code == 11
    with pytest.raises(O.OracleError) as e:
        O.plate([0.2, 0.2, 0], *flat[1:], 1, .3, 1, 1)                      # re-entrant corner
    assert e.value.code == 13


# ---- element result recovery and the iterative solve -------------------------------
def _mesh_one(kind, p, props, axis=None):
    from finite_element_method_b200 import meshes  # noqa: F401  (mesh dict layout)
    p = np.asarray(p, float)
    m = {"name": kind, "rel_tol": 1e-4, "abs_tol": 1e-12, "x": p[:, 0], "y": p[:, 1], "z": p[:, 2],
         "t_n1": np.zeros(0, np.uint32), "t_n2": np.zeros(0, np.uint32), "t_E": np.zeros(0), "t_A": np.zeros(0),
         "b_n1": np.zeros(0, np.uint32), "b_n2": np.zeros(0, np.uint32), "b_props": np.zeros((8, 0)),
         "b_axis": np.zeros((3, 0)), "p_n": np.zeros((4, 0), np.uint32), "p_props": np.zeros((4, 0))}
    if kind == "truss":
        m.update(t_n1=np.array([0], np.uint32), t_n2=np.array([1], np.uint32), t_E=np.array([props[0]]),
                 t_A=np.array([props[1]]))
        if len(props) > 2:
            m["t_A2"] = np.array([props[2]])
    elif kind == "beam":
        m.update(b_n1=np.array([0], np.uint32), b_n2=np.array([1], np.uint32),
                 b_props=np.asarray(props, float).reshape(8, 1), b_axis=np.asarray(axis, float).reshape(3, 1))
    else:
        m.update(p_n=np.arange(4, dtype=np.uint32).reshape(4, 1), p_props=np.asarray(props, float).reshape(4, 1))
    return m


def test_truss_result_is_axial_force():
    # truss.rs:281-333: N = E A (u2 - u1).n / L for any orientation (tapered: area at r = 0)
    rng = np.random.default_rng(5)
    for tapered in (False, True):
        p = rng.normal(size=(2, 3))
        u = rng.normal(size=12) * 1e-3
        props = (2.1e11, 1e-4, 2e-4) if tapered else (2.1e11, 1e-4)
        ft, _, _ = O.element_results(_mesh_one("truss", p, props), u)
        L = np.linalg.norm(p[1] - p[0]); n = (p[1] - p[0]) / L
        ref = 2.1e11 * (1.5e-4 if tapered else 1e-4) * ((u[6:9] - u[0:3]) @ n) / L
        assert abs(ft[0] - ref) <= 1e-12 * abs(ref) * 10


def test_beam_result_cantilever_tip_load():
    # solve the one-element cantilever for a tip load P along local v and recover the section forces:
    # ForceS = P (constant shear), MomentT (average) = P L / 2 = the moment at mid-span, node values +- P L / 2
    E, nu, A, I11, I22, It, ks, L, P = 2.1e11, 0.3, 1e-2, 8e-6, 4e-6, 1e-5, 5 / 6, 2.0, 1000.0
    props = [E, nu, A, I11, I22, 0.0, It, ks]
    q, pr, kl, kg = O.beam([0, 0, 0], [L, 0, 0], *props, [0, 0, 1])
    f = np.zeros(6); f[1] = P
    u = np.zeros(12); u[6:] = np.linalg.solve(kg[6:, 6:], f)
    _, fb, _ = O.element_results(_mesh_one("beam", [[0, 0, 0], [L, 0, 0]], props, [0, 0, 1]), u)
    fb = fb[0]
    assert abs(fb[0]) < 1e-6 * P and abs(abs(fb[1]) - P) < 1e-9 * P and abs(fb[2]) < 1e-6 * P
    assert abs(abs(fb[8]) - P * L / 2) < 1e-9 * P * L                 # MomentT average
    assert abs(fb[7] - fb[9]) == pytest.approx(abs(L * fb[1]), rel=1e-12)   # node 1 / node 2 differ by L * ForceS
    assert min(abs(fb[7]), abs(fb[9])) < 1e-9 * P * L                 # free end carries no moment


def test_plate_result_constant_membrane_strain():
    # u = eps_x * x on a skewed flat quad: N_x = E t eps / (1 - nu^2), N_y = nu N_x, N_xy = 0, no bending/shear
    rng = np.random.default_rng(6)
    p = np.array([[1, 0.75, 0], [0, 0.75, 0], [0, 0, 0], [1, 0, 0]], float)
    p[:, :2] += rng.uniform(-0.1, 0.1, (4, 2))
    E, nu, t, eps = 2.1e11, 0.3, 0.01, 1e-4
    u = np.zeros(24); u[0::6] = eps * p[:, 0]
    _, _, fp = O.element_results(_mesh_one("plate", p, [E, nu, t, 5 / 6]), u)
    nx = E * t * eps / (1 - nu * nu)
    assert abs(fp[0, 0] - nx) < 1e-10 * nx and abs(fp[0, 1] - nu * nx) < 1e-10 * nx
    assert np.all(np.abs(fp[0, 2:]) < 1e-9 * nx)


def test_plate_result_constant_curvature():
    # theta_y = -kappa * x (w = kappa x^2 / 2 would add shear; a pure rotation field gives curvature kappa_x and
    # transverse shear -theta): BendingMoment rows follow plate.rs:1294-1332 with c_bend = E t / (2 (1 - nu^2)), * t^2 / 24
    p = np.array([[1, 0.75, 0], [0, 0.75, 0], [0, 0, 0], [1, 0, 0]], float)
    E, nu, t, kappa = 2.1e11, 0.3, 0.01, 1e-3
    u = np.zeros(24); u[4::6] = -kappa * p[:, 0]
    _, _, fp = O.element_results(_mesh_one("plate", p, [E, nu, t, 5 / 6]), u)
    # c_bend * t^2 / 24 = E t^3 / (48 (1 - nu^2)), times the SUM over the four nodes = the textbook D = E t^3 / (12 (1 - nu^2))
    d = E * t * t * t / (12 * (1 - nu * nu))
    assert abs(abs(fp[0, 4]) - d * kappa) < 1e-10 * d * kappa          # BendingMomentS takes row 0 (kappa_x)
    assert abs(abs(fp[0, 3]) - nu * d * kappa) < 1e-10 * d * kappa     # BendingMomentR takes row 1


def test_pcg_reference_model_one_iteration():
    # src/tests/fem/test_fem.rs:83-225 in f64: K_aa = [66666.67], b = [100] -> one iteration, u = 0.0015, R = -100
    for starts in (None, [0]):
        u, it = O.pcg(1, ([0], [0], [66666.66666666667]), [100.0], 1000, 1e-4, 1e-12, starts)
        assert it == 1 and u[0] == 100.0 / 66666.66666666667
    sep = {"n_bb": 1, "k_bb_indexes": np.array([0]), "k_aa_indexes": np.array([6]),
           "k_ba": (np.array([0]), np.array([0]), np.array([-66666.66666666667])),
           "k_bb": (np.array([0]), np.array([0]), np.array([66666.66666666667]))}
    forces = np.zeros(12); forces[6] = 100.0
    rr = O.reactions(sep, u, forces, np.zeros(12))
    assert abs(rr[0] + 100.0) < 1e-12
    d, f = O.compose_global_analysis_result(sep, u, rr, forces, np.zeros(12))
    assert d[6] == u[0] and f[0] == rr[0] and f[6] == 100.0


def test_block_starts():
    assert O.block_starts_from_k_aa_indexes([6, 7, 8, 12, 14, 18]) == [0, 3, 5]
    assert O.block_starts_from_k_aa_indexes([]) == []


def test_element_results_golden_fixture():
    """tests/golden/element_results_golden.json (generator: make_golden.py) pins the oracle's element results bit
    for bit across machines / compilers (-ffp-contract=off); the GPU path is compared with the oracle in
    tests/test_analysis.py"""
    from finite_element_method_b200 import meshes
    with open(os.path.join(GOLD, "element_results_golden.json")) as f:
        gold = json.load(f)
    for name, mesh in (("truss_cube", meshes.truss_cube(3)), ("beam_frame_jitter", meshes.beam_frame(3, 10 ** 9, jitter=True)),
                       ("plate_x0", meshes.plate_grid(3, 2, "x0")), ("mixed", meshes.mixed_structure(3, 2))):
        u = np.random.default_rng(20241018).normal(size=6 * len(mesh["x"])) * 1e-3
        ot, ob, op = O.element_results(mesh, u)
        assert np.array_equal(ot, np.array(gold[name]["truss"]))
        assert np.array_equal(ob, np.array(gold[name]["beam"]).reshape(ob.shape))
        assert np.array_equal(op, np.array(gold[name]["plate"]).reshape(op.shape))
