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
