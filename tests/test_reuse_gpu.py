"""A re-used instance (FEM::reset, fem.rs:155, then the next model) registers its host staging with CUDA and uploads
each accepted batch as it is added (api.cu: pin_take / make_room / upload_early). Nothing of that may show in the
results: every model assembled on a re-used handle must equal, bit for bit, the same model on a fresh one."""
import numpy as np
import pytest

from finite_element_method_b200 import FEM, FemError, meshes

pytestmark = pytest.mark.gpu


def fresh(mesh):
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], len(mesh["x"]), device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    out = [a.copy() for a in fem.csr()]
    fem.close()
    return out


def same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


def test_reused_handle_uploads_from_registered_staging_with_identical_results():
    small = meshes.mixed_structure(700, 400)          # 0.28M plates + 0.28M beams + 0.14M trusses: staging of 1-5 MB per array
    large = meshes.mixed_structure(900, 700)          # every staging vector has to grow (and be registered again)
    ref_small, ref_large = fresh(small), fresh(large)
    fem = FEM(small["rel_tol"], small["abs_tol"], len(small["x"]), device=0)
    fem.load_mesh(small)
    fem.assemble()
    assert same(fem.csr(), ref_small)                 # first model: staged upload
    for mesh, ref in ((small, ref_small), (small, ref_small), (large, ref_large), (small, ref_small), (large, ref_large)):
        fem.reset(len(mesh["x"]))
        fem.load_mesh(mesh)                            # registered staging, uploads queued per batch
        fem.assemble()
        assert same(fem.csr(), ref)
    # a batch that fails in the middle on a re-used handle: the prefix stays, the tail can be added afterwards
    fem.reset(len(small["x"]))
    api = FEM.api_arrays(small)
    fem.add_nodes(*api["nodes"])
    num, n1, n2, n3, n4, *props = api["plates"]
    n1b = n1.copy(); k = 200_000
    n1b[k] = 0                                         # a node number that does not exist
    with pytest.raises(FemError, match="Node with number 0 does not exist!"):
        fem.add_plates(num, n1b, n2, n3, n4, *props)
    assert fem.counts()[3] == k
    fem.add_plates(num[k:], n1[k:], n2[k:], n3[k:], n4[k:], *[p[k:] for p in props])
    fem.add_beams(*api["beams"])
    fem.add_trusses(*api["trusses"])
    fem.assemble()
    assert same(fem.csr(), ref_small)
    fem.close()


def test_registration_can_be_forced_or_switched_off(monkeypatch):
    mesh = meshes.plate_grid(900, 700)
    ref = fresh(mesh)
    for flag in ("1", "0"):
        monkeypatch.setenv("FEMGPU_PIN_HOST", flag)
        fem = FEM(mesh["rel_tol"], mesh["abs_tol"], len(mesh["x"]), device=0)
        for _ in range(2):
            fem.load_mesh(mesh)
            fem.assemble()
            assert same(fem.csr(), ref)
            fem.reset(len(mesh["x"]))
        fem.close()
