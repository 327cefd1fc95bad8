"""Small models through every kernel variant, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finite_element_method_b200 import FEM, meshes

for mesh in (meshes.mixed_structure(24, 18), meshes.plate_grid(20, 15, "jitter"), meshes.beam_frame(6, 10**9), meshes.truss_lattice(8, 10**9)):
    n = len(mesh["x"])
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], n, device=0)
    fem.load_mesh(mesh)
    fem.assemble()
    fem.numeric(); fem.synchronize()
    v = fem.csr(values_only=True)
    y0 = np.flatnonzero(np.asarray(mesh["y"]) == np.min(mesh["y"]))
    ndof = 3 if len(mesh["b_n1"]) == 0 and len(mesh["p_n"][0]) == 0 else 6
    fem.add_displacement(np.repeat(y0, ndof) + 1, np.tile(np.arange(ndof), len(y0)), np.zeros(ndof * len(y0)))
    fem.add_concentrated_load(n, 2, -10.0)
    fem.separate_stiffness_matrix_sparse_iterative()
    try:
        fem.find_ua_vector_iterative_pcg_block_jacobi_sparse(50, copy_out=False)
    except Exception as e:   # not converged in 50 iterations: fine here
        pass
    fem.set_displacements_vector(np.zeros(6 * n))
    for f in range(3):
        fem.element_results(f)
    print(mesh["name"], "ok", float(np.abs(v).max()), flush=True)
    fem.close()
