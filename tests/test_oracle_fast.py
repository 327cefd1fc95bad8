"""The multi-core "fast" CPU baseline agrees with the faithful restatement (CPU only)."""
import numpy as np
import pytest
import scipy.sparse as sp

from finite_element_method_b200 import meshes
from oracle import oracle as O

CASES = [
    lambda: meshes.reference_truss_model(),
    lambda: meshes.truss_cube(3),
    lambda: meshes.truss_lattice(6, 10 ** 9, jitter=True),
    lambda: meshes.beam_frame(5, 10 ** 9),
    lambda: meshes.beam_frame(5, 10 ** 9, jitter=True),
    lambda: meshes.plate_grid(6, 5, "flat"),
    lambda: meshes.plate_grid(6, 5, "jitter"),
    lambda: meshes.plate_grid(6, 5, "x0"),
    lambda: meshes.mixed_structure(6, 4),
]


@pytest.mark.parametrize("make", CASES)
def test_fast_matches_faithful(make):
    mesh = make()
    n = 6 * len(mesh["x"])
    r, c, v = O.faithful_coo(mesh)
    A = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
    out = O.fast_assemble(mesh, n_threads=2, want_coo=True)
    fr, fc, fv = out["coo"]
    B = sp.coo_matrix((fv, (fr, fc)), shape=(n, n)).tocsr()
    assert out["nnz"] == meshes.algorithmic_bytes(mesh)["nnz"]
    assert abs(A - B).max() <= 1e-14 * abs(A).max()


def test_structural_nnz_closed_form():
    m = meshes.plate_grid(7, 5)
    assert meshes.algorithmic_bytes(m)["nnz"] == meshes.grid_nnz_fast(7, 5)
    m = meshes.mixed_structure(8, 6)
    assert meshes.algorithmic_bytes(m)["nnz"] == meshes.grid_nnz_fast(8, 6)
    assert meshes.grid_nnz_fast(2000, 2000) == 1_296_432_036


def test_partition_owns_every_element_once():
    m = meshes.mixed_structure(10, 12)
    for world in (2, 3, 4):
        parts = meshes.partition_rows(m, world, 11)
        assert parts[0][0] == 0 and parts[-1][1] == len(m["x"])
        tot = sum(meshes.n_elements(meshes.local_part(m, b, e)) for b, e in parts)
        assert tot == meshes.n_elements(m)


@pytest.mark.parametrize("make", [lambda: meshes.mixed_structure(9, 7), lambda: meshes.truss_cube(4),
                                  lambda: meshes.plate_grid(6, 5, "x0")])
def test_sampled_rows_harness_on_cpu(make):
    """The full-size GPU parity tests compare sampled block rows (tests/fullsize_common.py). Here the same harness
    runs on CPU tensors: the fast baseline's structural CSR against the faithful sampled rows, and a corrupted
    copy must be caught."""
    import sys, os
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fullsize_common import compare_sampled_rows, sample_nodes
    mesh = make()
    n = len(mesh["x"])
    out = O.fast_assemble(mesh, n_threads=2, want_coo=True)
    r, c, v = out["coo"]                       # structural layout order = CSR order
    assert np.all(np.diff(r) >= 0)
    rp = np.zeros(6 * n + 1, np.int64)
    np.add.at(rp, r + 1, 1)
    rp = np.cumsum(rp)
    csr = (torch.as_tensor(rp), torch.as_tensor(c.astype(np.int32)), torch.as_tensor(v.copy()))
    nodes = sample_nodes(n, n_random=50)
    rep = compare_sampled_rows(None, mesh, nodes, device="cpu", csr=csr)
    assert rep["n_fail"] == 0 and rep["entries"] > 0 and rep["max_block_rel"] < 1e-13, rep
    bad = v.copy()
    k = int(rp[6 * int(nodes[len(nodes) // 2])])
    bad[k] = bad[k] * (1 + 1e-9) + 1e-3
    rep = compare_sampled_rows(None, mesh, nodes, device="cpu", csr=(csr[0], csr[1], torch.as_tensor(bad)))
    assert rep["n_fail"] == 1, rep


def test_node_windows_cover_what_a_rank_needs():
    """meshes.with_node_window: one contiguous node range per rank holding its own nodes and every node its elements
    touch; grid strips need one more grid line (+1 node), a 3D lattice one more plane."""
    for mesh, width in ((meshes.mixed_structure(10, 12), 11), (meshes.truss_lattice(6, 10 ** 9), None),
                        (meshes.beam_frame(5, 10 ** 9), None)):
        n = len(mesh["x"])
        for world in (2, 3, 4):
            for begin, end in meshes.partition_rows(mesh, world, width):
                part = meshes.with_node_window(meshes.local_part(mesh, begin, end), begin, end)
                w0, w1 = part["node_window_begin"], part["node_window_begin"] + len(part["x"])
                assert w0 == begin and w1 >= end and w1 <= n and part["nodes_number"] == n
                used = [np.asarray(part[k]).ravel() for k in ("t_n1", "t_n2", "b_n1", "b_n2", "p_n") if np.asarray(part[k]).size]
                if used:
                    u = np.concatenate(used)
                    assert u.min() >= w0 and u.max() < w1
                assert np.array_equal(part["x"], mesh["x"][w0:w1]) and np.array_equal(part["z"], mesh["z"][w0:w1])
                if width and end < n:
                    assert w1 - end <= width + 1            # halo: one more grid line
