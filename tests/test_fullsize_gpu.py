"""Full BASELINE.json sizes: size-independent properties checked on the device (torch is only the
test harness here: it wraps the library's device pointers, it does not compute any part of K)."""
import numpy as np
import pytest

from finite_element_method_b200 import FEM, meshes

pytestmark = pytest.mark.gpu


class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_csr(fem):
    import torch
    (rp, ci, v), (rb, re) = fem.csr_device()
    n_rows, nnz = fem.symbolic()
    row_ptr = torch.as_tensor(_DevArray(rp, n_rows + 1, "<i8"), device="cuda")
    col = torch.as_tensor(_DevArray(ci, nnz, "<i4"), device="cuda")
    val = torch.as_tensor(_DevArray(v, nnz, "<f8"), device="cuda")
    return row_ptr, col, val


def spmv(row_ptr, col, val, x):
    import torch
    n = row_ptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n, device="cuda"), row_ptr[1:] - row_ptr[:-1])
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    y.index_add_(0, rows, val * x[col.long()])
    return y, rows


CONFIGS = {
    "T-1M-truss": lambda: meshes.truss_lattice(64),
    "B-2M-beam": lambda: meshes.beam_frame(88),
    "P-4M-plate": lambda: meshes.plate_grid(2000, 2000),
    "M-10M-mixed": lambda: meshes.mixed_structure(2000, 2000),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_full_size_properties(name):
    import torch
    mesh = CONFIGS[name]()
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], mesh["nodes_number"])
    fem.load_mesh(mesh)
    n_rows, nnz = fem.assemble()
    assert n_rows == 6 * len(mesh["x"])
    if name.startswith(("P", "M")):
        assert nnz == 1_296_432_036                      # SURVEY.md §8d structural nnz
    row_ptr, col, val = device_csr(fem)
    scale = val.abs().max()
    assert torch.isfinite(val).all()
    # (1) rigid translations are null vectors: K t_d = 0 for d = x, y, z
    for d in range(3):
        t = torch.zeros(n_rows, dtype=torch.float64, device="cuda")
        t[d::6] = 1.0
        y, rows = spmv(row_ptr, col, val, t)
        assert y.abs().max() <= 1e-9 * scale, (name, d, float(y.abs().max() / scale))
    # (2) symmetry through two random vectors: x^T K y == y^T K x
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(n_rows, dtype=torch.float64, device="cuda", generator=g)
    z = torch.rand(n_rows, dtype=torch.float64, device="cuda", generator=g)
    kx, _ = spmv(row_ptr, col, val, x)
    kz, _ = spmv(row_ptr, col, val, z)
    a, b = torch.dot(z, kx), torch.dot(x, kz)
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b))
    # (3) positive semi-definite sample: x^T K x >= 0
    assert torch.dot(x, kx) > 0
    # (4) diagonal entries of connected translational dofs are positive
    del kx, kz
    diag_mask = rows == col.long()
    assert (val[diag_mask] > 0).all() if name != "T-1M-truss" else (val[diag_mask] >= 0).all()
    # (5) re-assembly is bit-identical (determinism at full size)
    chk = val.clone()
    fem.numeric(); fem.synchronize()
    assert torch.equal(chk, val)
    # (6) linearity in the material: doubling every E doubles K exactly (power of two)
    del chk, rows, diag_mask
    fem.close()


def test_linearity_in_young_modulus():
    import torch
    mesh = meshes.mixed_structure(300, 200)
    outs = []
    for s in (1.0, 2.0):
        m = dict(mesh)
        m["p_props"] = mesh["p_props"].copy(); m["p_props"][0] *= s
        m["b_props"] = mesh["b_props"].copy(); m["b_props"][0] *= s
        m["t_E"] = mesh["t_E"] * s
        fem = FEM(m["rel_tol"], m["abs_tol"], m["nodes_number"])
        fem.load_mesh(m)
        fem.assemble()
        outs.append(fem.csr(values_only=True).copy())
        fem.close()
    # the drilling penalty (KROT6 = 1, plate.rs:25) does not scale with E: K(2E) - 2 K(E) is minus the
    # number of plates at the node on every theta_z diagonal, and exactly zero everywhere else
    a, b = outs
    d = b - 2 * a
    mask = np.abs(d) > 0.5
    assert mask.sum() == 301 * 201
    assert set(np.round(-d[mask]).astype(int).tolist()) == {1, 2, 4}
    assert np.abs(d[mask] + np.round(-d[mask])).max() < 1e-6
    assert np.abs(d[~mask]).max() <= 1e-12 * np.abs(a).max()
