"""BASELINE.json's full-size configurations, VALUE parity: >= 10 k sampled node rows (60 k matrix rows, incl. the
first and last slab, slab boundaries and the grid lines where 2- and 8-rank partitions cut) of the GPU matrix
against the oracle's faithful element matrices summed in the reference's order — bit-exact pattern, values
within 1e-12. Index-width or slab-table bugs that only show at 10^7 elements would fail here."""
import pytest

from finite_element_method_b200 import FEM, meshes

from fullsize_common import compare_sampled_rows, sample_nodes

pytestmark = pytest.mark.gpu

CASES = {
    "T-1M-truss": (lambda: meshes.truss_lattice(64), None),
    "T-1M-truss-jitter": (lambda: meshes.truss_lattice(64, jitter=True), None),
    "B-2M-beam": (lambda: meshes.beam_frame(88), None),
    "B-2M-beam-jitter": (lambda: meshes.beam_frame(88, jitter=True), None),
    "P-4M-plate": (lambda: meshes.plate_grid(2000, 2000), 2001),
    "P-4M-plate-jitter": (lambda: meshes.plate_grid(2000, 2000, "jitter"), 2001),
    "P-4M-plate-x0": (lambda: meshes.plate_grid(2000, 2000, "x0"), 2001),
    "M-10M-mixed": (lambda: meshes.mixed_structure(2000, 2000), 2001),
    "M-10M-mixed-x0": (lambda: meshes.mixed_structure(2000, 2000, variant="x0"), 2001),
}


@pytest.mark.parametrize("name", list(CASES))
def test_full_size_values_match_oracle_on_sampled_rows(name):
    make, grid_w = CASES[name]
    mesh = make()
    n = len(mesh["x"])
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], n)
    fem.load_mesh(mesh)
    fem.assemble()
    lines = (250, 1000, 1750) if grid_w else ()
    nodes = sample_nodes(n, grid_w, n_random=12_000, lines=lines)     # >= 10 k distinct nodes after de-duplication
    rep = compare_sampled_rows(fem, mesh, nodes, rtol=1e-12, faithful=True)
    print(name, rep)
    assert rep["nodes"] >= 10_000
    assert rep["n_fail"] == 0, rep
    assert rep["max_block_rel"] <= 1e-12, rep
    fem.close()
