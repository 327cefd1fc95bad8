"""Parity of the CUDA path (through the C ABI) against the oracle. Bar: every CSR entry within 1e-12
relative on the union of patterns (absent == 0), with the 6x6-block-scale floor of SURVEY.md §8c for
entries that are pure rounding noise. All of these need a B200."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import parity_report
from finite_element_method_b200 import BEAM, FEM, PLATE, TRUSS, FemError, meshes

pytestmark = pytest.mark.gpu
RTOL = 1e-12   # BASELINE.json north_star: "matches the reference within 1e-12 relative per nonzero"


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def assemble(mesh):
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], mesh["nodes_number"])
    fem.load_mesh(mesh)
    n_rows, nnz = fem.assemble()
    return fem, n_rows, nnz


SMALL = {
    "reference-model": lambda: meshes.reference_truss_model(),
    "truss-cube-27-nodes": lambda: meshes.truss_cube(3),
    "truss-lattice-jitter": lambda: meshes.truss_lattice(8, 10 ** 9, jitter=True),
    "truss-lattice-axis-aligned": lambda: meshes.truss_lattice(8, 10 ** 9),
    "beam-frame": lambda: meshes.beam_frame(6, 10 ** 9),
    "beam-frame-jitter": lambda: meshes.beam_frame(6, 10 ** 9, jitter=True),
    "plate-flat": lambda: meshes.plate_grid(12, 9, "flat"),
    "plate-jitter": lambda: meshes.plate_grid(12, 9, "jitter"),
    "plate-x0-plane": lambda: meshes.plate_grid(12, 9, "x0"),
    "mixed": lambda: meshes.mixed_structure(12, 10),
    "folded-plate-flat-and-tilted": lambda: meshes.folded_plate(12, 8),
    "hub-star-unstaged-slab": lambda: meshes.hub_star(700, 7),
}


@pytest.mark.parametrize("name", list(SMALL))
def test_csr_matches_oracle_nonzero_by_nonzero(name, O):
    mesh = SMALL[name]()
    fem, n_rows, nnz = assemble(mesh)
    assert n_rows == 6 * mesh["nodes_number"]
    assert nnz == meshes.algorithmic_bytes(mesh)["nnz"]
    rep = parity_report(n_rows, fem.csr(), O.faithful_coo(mesh), RTOL)
    assert rep["n_fail"] == 0, rep
    assert rep["max_block_rel"] < 1e-14, rep
    fem.close()


@pytest.mark.parametrize("threads", [32, 64])
@pytest.mark.parametrize("name", ["plate-flat", "plate-x0-plane", "mixed", "beam-frame-jitter", "truss-cube-27-nodes",
                                  "folded-plate-flat-and-tilted", "hub-star-unstaged-slab"])
def test_both_cta_shapes_match_oracle(name, threads, O, monkeypatch):
    """The assembly kernel exists in a one-warp and a two-warp-per-slab shape (the symbolic pass picks one from
    the family mix); both must give the oracle's matrix on every kind of mesh, and the same values as each other
    up to the rounding of a different (still fixed) split of heavy blocks."""
    monkeypatch.setenv("FEMGPU_ASM_THREADS", str(threads))
    mesh = SMALL[name]()
    fem, n_rows, nnz = assemble(mesh)
    rep = parity_report(n_rows, fem.csr(), O.faithful_coo(mesh), RTOL)
    assert rep["n_fail"] == 0, rep
    v1 = fem.csr(values_only=True).copy()
    fem.numeric(); fem.synchronize()
    assert np.array_equal(v1, fem.csr(values_only=True))
    fem.close()


@pytest.mark.parametrize("threads", [32, 64])
@pytest.mark.parametrize("name", ["plate-jitter", "mixed", "beam-frame", "truss-lattice-jitter"])
def test_record_staging_variants_agree_bit_for_bit(name, threads, monkeypatch):
    """Element records reach shared memory either run-wise with TMA bulk copies (consecutively numbered meshes) or
    per element with cp.async (the symbolic pass measures the runs and picks one); same records, same arithmetic:
    the matrices must be identical. Also with the bank-spreading lane order, which only permutes lanes."""
    mesh = SMALL[name]()
    monkeypatch.setenv("FEMGPU_ASM_THREADS", str(threads))
    vals = []
    for bulk, spread in (("0", "0"), ("1", "0"), ("1", "1")):
        monkeypatch.setenv("FEMGPU_ASM_BULK", bulk)
        monkeypatch.setenv("FEMGPU_SPREAD_BANKS", spread)
        fem, n_rows, nnz = assemble(mesh)
        vals.append(fem.csr(values_only=True).copy())
        fem.close()
    assert np.array_equal(vals[0], vals[1]) and np.array_equal(vals[0], vals[2])


def test_shuffled_numbering_takes_the_per_element_staging(O, monkeypatch):
    """a mesh whose elements are inserted in random order has no runs: the symbolic pass must still give the
    oracle's matrix (accumulation order = insertion order) through the cp.async staging it selects"""
    mesh = meshes.mixed_structure(12, 9)
    rng = np.random.default_rng(3)
    pp = rng.permutation(len(mesh["p_n"][0])); pb = rng.permutation(len(mesh["b_n1"])); pt = rng.permutation(len(mesh["t_n1"]))
    mesh["p_n"] = np.asarray(mesh["p_n"]).reshape(4, -1)[:, pp]
    mesh["p_props"] = np.asarray(mesh["p_props"]).reshape(4, -1)[:, pp]
    mesh["b_n1"], mesh["b_n2"] = np.asarray(mesh["b_n1"])[pb], np.asarray(mesh["b_n2"])[pb]
    mesh["b_props"] = np.asarray(mesh["b_props"]).reshape(8, -1)[:, pb]
    mesh["b_axis"] = np.asarray(mesh["b_axis"]).reshape(3, -1)[:, pb]
    mesh["t_n1"], mesh["t_n2"] = np.asarray(mesh["t_n1"])[pt], np.asarray(mesh["t_n2"])[pt]
    mesh["t_E"], mesh["t_A"] = np.asarray(mesh["t_E"])[pt], np.asarray(mesh["t_A"])[pt]
    if mesh.get("t_A2") is not None:
        mesh["t_A2"] = np.asarray(mesh["t_A2"])[pt]
    fem, n_rows, nnz = assemble(mesh)
    rep = parity_report(n_rows, fem.csr(), O.faithful_coo(mesh), RTOL)
    assert rep["n_fail"] == 0, rep
    fem.close()


def test_reference_model_known_answer():
    """config 1(i): K entries +-66666.66666666667 at rows/cols {0, 6}, nothing else."""
    fem, n_rows, nnz = assemble(meshes.reference_truss_model())
    r, c, v = fem.nonzero_coo()
    assert sorted(zip(r.tolist(), c.tolist())) == [(0, 0), (0, 6), (6, 0), (6, 6)]
    K = sp.coo_matrix((v, (r, c)), shape=(n_rows, n_rows)).toarray()
    assert K[0, 0] == 66666.66666666667 and K[6, 6] == 66666.66666666667
    assert K[0, 6] == -66666.66666666667 and K[6, 0] == -66666.66666666667
    assert 100.0 / K[6, 6] == 0.0014999999999999998      # the reference's u2x in f64
    assert np.array_equal(fem.get_truss_rotation_matrix_elements(1), np.eye(3).ravel())
    fem.close()


def test_reference_pattern_is_reproduced_by_compaction(O):
    """The reference stores only contributions != 0.0; compaction of the structural pattern gives the
    same (row, col) set wherever the oracle's stored value is non-zero."""
    for make in (SMALL["truss-cube-27-nodes"], SMALL["plate-flat"], SMALL["beam-frame"]):
        mesh = make()
        fem, n_rows, _ = assemble(mesh)
        r, c, v = fem.nonzero_coo()
        orr, oc, ov = O.faithful_coo(mesh)
        keep = ov != 0.0
        got = set(zip(r.tolist(), c.tolist()))
        want = set(zip(orr[keep].tolist(), oc[keep].tolist()))
        # entries present on one side only must be rounding noise (|v| <= 1e-12 * block scale)
        K = sp.coo_matrix((v, (r, c)), shape=(n_rows, n_rows)).tocsr()
        Ko = sp.coo_matrix((ov, (orr, oc)), shape=(n_rows, n_rows)).tocsr()
        scale = abs(Ko).max()
        for (i, j) in got ^ want:
            assert abs(K[i, j]) <= 1e-12 * scale and abs(Ko[i, j]) <= 1e-12 * scale
        assert len(got & want) >= 0.99 * len(want)
        fem.close()


@pytest.mark.parametrize("name", ["truss-cube-27-nodes", "plate-flat", "folded-plate-flat-and-tilted", "mixed", "hub-star-unstaged-slab"])
def test_nonzero_csr_is_the_structural_csr_without_its_zeros(name):
    """femgpu_get_nonzero_csr (compaction on the device, what the e2e path reads back) against the structural CSR
    filtered on the host: identical row pointers, columns and values, bit for bit; recomputed after a new pass."""
    if name not in SMALL:
        pytest.skip(f"no mesh {name}")
    mesh = SMALL[name]()
    fem, n_rows, nnz = assemble(mesh)
    rp, ci, v = fem.csr()
    keep = v != 0.0
    rows = np.repeat(np.arange(n_rows), np.diff(rp))
    want_rp = np.zeros(n_rows + 1, np.int64)
    np.add.at(want_rp, rows[keep] + 1, 1)
    want_rp = np.cumsum(want_rp)
    g_rp, g_ci, g_v = fem.nonzero_csr()
    assert np.array_equal(g_rp, want_rp)
    assert np.array_equal(g_ci, ci[keep]) and np.array_equal(g_v, v[keep])
    assert 0 < len(g_v) <= nnz
    r, c, vv = fem.nonzero_coo()                      # the older COO form agrees
    assert np.array_equal(c, g_ci) and np.array_equal(vv, g_v) and np.array_equal(r, rows[keep])
    fem.numeric()                                     # a new pass invalidates the cached compaction
    g2 = fem.nonzero_csr()
    assert np.array_equal(g2[0], g_rp) and np.array_equal(g2[2], g_v)
    fem.close()


def test_element_matrices_match_oracle_and_golden(O):
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "element_golden.json")))
    for case in gold["truss"]:
        f = FEM(1e-4, 1e-12, 2)
        f.add_nodes([1, 2], *np.array([case["p1"], case["p2"]]).T)
        f.add_truss(1, 1, 2, case["E"], case["A"], case["A2"])
        kg = f.element_matrix(TRUSS, 1)
        ref = np.array(case["kg"])
        assert np.allclose(kg, ref, rtol=RTOL, atol=RTOL * abs(ref).max())
        q, _, _ = O.truss(case["p1"], case["p2"], case["E"], case["A"], case["A2"])
        assert np.allclose(f.get_truss_rotation_matrix_elements(1).reshape(3, 3), q, rtol=0, atol=1e-15)
        f.close()
    for case in gold["beam"]:
        f = FEM(1e-4, 1e-12, 2)
        f.add_nodes([1, 2], *np.array([case["p1"], case["p2"]]).T)
        f.add_beam(1, 1, 2, *case["props"], case["axis"])
        kg = f.element_matrix(BEAM, 1)
        ref = np.array(case["kg"])
        assert np.allclose(kg, ref, rtol=RTOL, atol=RTOL * abs(ref).max())
        assert np.allclose(f.get_beam_rotation_matrix_elements(1).reshape(3, 3), np.array(case["q"]), rtol=0, atol=1e-15)
        f.close()
    for case in gold["plate"]:
        f = FEM(1e-4, 1e-12, 4)
        f.add_nodes([1, 2, 3, 4], *np.array(case["p"]).T)
        f.add_plate(1, 1, 2, 3, 4, *case["props"])
        kg = f.element_matrix(PLATE, 1)
        ref = np.array(case["kg"])
        blockmax = np.abs(ref).reshape(4, 6, 4, 6).max(axis=(1, 3))
        tol = RTOL * np.repeat(np.repeat(blockmax, 6, 0), 6, 1)
        assert np.all((np.abs(kg - ref) <= RTOL * np.abs(ref)) | (np.abs(kg - ref) <= tol))
        q, _, _ = O.plate(*case["p"], *case["props"])
        assert np.allclose(f.get_plate_rotation_matrix_elements(1).reshape(3, 3), q, rtol=0, atol=1e-15)
        f.close()


def test_scatter_map_matches_start_positions():
    """element -> CSR slot map against the reference's start_positions arithmetic
    (methods_for_plate_data_handle.rs:113-132): slot (i, j) of local block (la, lb) must be the CSR
    entry (6*idx[la]+i, 6*idx[lb]+j)."""
    mesh = meshes.mixed_structure(5, 4)
    fem, n_rows, nnz = assemble(mesh)
    rp, ci, v = fem.csr()
    rows = np.repeat(np.arange(n_rows), np.diff(rp))
    pn = np.asarray(mesh["p_n"]).reshape(4, -1)
    for e in (0, 7, pn.shape[1] - 1):
        slots = fem.element_slots(PLATE, e + 1)
        idx = pn[:, e]
        for la in range(4):
            for lb in range(4):
                for i in range(6):
                    for j in range(6):
                        s = slots[6 * la + i, 6 * lb + j]
                        assert s >= 0 and rows[s] == 6 * idx[la] + i and ci[s] == 6 * idx[lb] + j
    # truss: only the 3x3 translational slots exist... inside a 6x6 block here because plates share the pair
    slots = fem.element_slots(TRUSS, 1)
    a, b = mesh["t_n1"][0], mesh["t_n2"][0]
    assert rows[slots[0, 3]] == 6 * a and ci[slots[0, 3]] == 6 * b
    fem.close()
    # truss-only mesh: rotational rows of truss-only nodes stay empty (methods_for_truss_data_handle.rs:93-123)
    fem, n_rows, nnz = assemble(meshes.truss_cube(3))
    rp, _, _ = fem.csr()
    lens = np.diff(rp).reshape(-1, 6)
    assert np.all(lens[:, 3:] == 0) and np.all(lens[:, :3] > 0)
    fem.close()


def test_deterministic_and_rerunnable():
    mesh = meshes.mixed_structure(40, 30)
    fem, n_rows, nnz = assemble(mesh)
    v1 = fem.csr(values_only=True).copy()
    for _ in range(3):
        fem.numeric()
    fem.synchronize()
    v2 = fem.csr(values_only=True)
    assert np.array_equal(v1, v2)          # bitwise: no atomics, fixed summation order
    fem2, _, _ = assemble(mesh)
    assert np.array_equal(v1, fem2.csr(values_only=True))
    assert fem.launch_count() > 0
    fem.close(); fem2.close()


def test_accumulation_follows_insertion_order(O):
    """Two assemblies of the same mesh with families inserted in different orders agree to rounding and
    each matches the oracle run in its own order."""
    mesh = meshes.mixed_structure(8, 6)
    fem, n_rows, _ = assemble(mesh)                      # plates, beams, trusses
    rep = parity_report(n_rows, fem.csr(), O.faithful_coo(mesh), RTOL)
    assert rep["n_fail"] == 0
    f2 = FEM(mesh["rel_tol"], mesh["abs_tol"], mesh["nodes_number"])
    n = len(mesh["x"])
    f2.add_nodes(np.arange(1, n + 1), mesh["x"], mesh["y"], mesh["z"])
    t = mesh
    f2.add_trusses(np.arange(1, len(t["t_n1"]) + 1), t["t_n1"] + 1, t["t_n2"] + 1, t["t_E"], t["t_A"])
    bp = t["b_props"]
    f2.add_beams(np.arange(1, len(t["b_n1"]) + 1), t["b_n1"] + 1, t["b_n2"] + 1, *[bp[i] for i in range(8)], t["b_axis"])
    pn, pp = t["p_n"], t["p_props"]
    f2.add_plates(np.arange(1, pn.shape[1] + 1), pn[0] + 1, pn[1] + 1, pn[2] + 1, pn[3] + 1, pp[0], pp[1], pp[2], pp[3])
    f2.assemble()
    a, b = fem.csr(values_only=True), f2.csr(values_only=True)
    assert np.allclose(a, b, rtol=1e-13, atol=1e-13 * np.abs(a).max())
    fem.close(); f2.close()


def test_device_validation_errors_and_rollback():
    f = FEM(1e-4, 1e-12, 8)
    f.add_nodes([1, 2, 3, 4, 5, 6, 7], [0, 1, 1, 0, 2, 0.2, 3], [0, 0, 1, 1, 0, 0.2, 0], [0, 0, 0, 0, 0, 0, 0.5])
    with pytest.raises(FemError) as e:
        f.add_beam(1, 1, 2, 2e11, .3, 1e-2, 8e-6, 4e-6, 0, 1e-5, 5 / 6, [3.0, 0, 0])
    assert e.value.code == 28 and str(e.value) == "Local axis 1 direction [3.0, 0.0, 0.0] parallel to element 1!"
    with pytest.raises(FemError) as e:
        f.add_plate(1, 5, 4, 1, 2, 2e11, .3, .01, 5 / 6)       # nodes 1, 2, 5 on a line
    assert e.value.code == 30 and str(e.value) == "Some nodes of 1 element lie on the line!"
    with pytest.raises(FemError) as e:
        f.add_plate(1, 7, 4, 1, 2, 2e11, .3, .01, 5 / 6)       # node 7 is off the plane
    assert e.value.code == 31 and str(e.value) == "Not all nodes of element 1 lie on the plane!"
    with pytest.raises(FemError) as e:
        f.add_plate(1, 6, 4, 1, 2, 2e11, .3, .01, 5 / 6)       # re-entrant corner
    assert e.value.code == 32 and str(e.value) == "Element 1 non-convex!"
    assert f.counts() == (7, 0, 0, 0)                            # failed elements left nothing behind
    f.add_plate(1, 3, 4, 1, 2, 2e11, .3, .01, 5 / 6)
    # batch: geometry error in the middle keeps the prefix, drops the rest
    with pytest.raises(FemError) as e:
        f.add_beams([1, 2, 3], [1, 2, 3], [2, 3, 4], [2e11] * 3, [.3] * 3, [1e-2] * 3, [8e-6] * 3, [4e-6] * 3, [0] * 3,
                    [1e-5] * 3, [5 / 6] * 3, np.array([[0, 0, 0], [0, 1, 0], [1, 0, 1.0]]))
        f.validate()
    assert e.value.code == 28 and "parallel to element 2!" in str(e.value)
    assert f.counts() == (7, 0, 1, 1)
    with pytest.raises(FemError) as e:
        f.get_beam_rotation_matrix_elements(2)
    assert str(e.value) == "Beam element with number 2 does not exist!"
    n_rows, nnz = f.assemble()
    assert nnz == 36 * 16                                         # the 4-node plate; the beam pair is inside it
    f.close()


def test_empty_and_ragged_models():
    f = FEM(1e-4, 1e-12, 5)
    assert f.assemble() == (30, 0)                                # no nodes, no elements
    rp, ci, v = f.csr()
    assert len(ci) == 0 and np.all(rp == 0)
    f.add_nodes([10, 20, 30], [0, 1, 5], [0, 0, 5], [0, 0, 5])    # node 30 stays unconnected, 2 slots unused
    f.add_truss(1, 10, 20, 1.0, 1.0)
    assert f.assemble() == (30, 36)
    rp, ci, v = f.csr()
    assert np.all(np.diff(rp)[12:] == 0)
    f.close()


def test_incremental_adds_invalidate_the_pattern(O):
    mesh = meshes.plate_grid(6, 4)
    f = FEM(mesh["rel_tol"], mesh["abs_tol"], mesh["nodes_number"])
    n = len(mesh["x"])
    f.add_nodes(np.arange(1, n + 1), mesh["x"], mesh["y"], mesh["z"])
    pn, pp = mesh["p_n"], mesh["p_props"]
    half = pn.shape[1] // 2
    f.add_plates(np.arange(1, half + 1), *(pn[:, :half] + 1), *pp[:, :half])
    f.assemble()
    f.add_plates(np.arange(half + 1, pn.shape[1] + 1), *(pn[:, half:] + 1), *pp[:, half:])
    n_rows, nnz = f.assemble()
    rep = parity_report(n_rows, f.csr(), O.faithful_coo(mesh), RTOL)
    assert rep["n_fail"] == 0, rep
    f.close()
