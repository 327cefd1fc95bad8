#!/usr/bin/env python
"""Config M assembled once, then the K separation three ways in one process (for `ncu -k regex:quadrant`): count + fill
(the default), the one-pass kernel (FEMGPU_SEP_ONE_PASS=1), each preceded by a warm-up call that sizes the buffers."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finite_element_method_b200 import FEM, meshes

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
mesh = meshes.mixed_structure(nx, nx)
n = len(mesh["x"])
fem = FEM(mesh["rel_tol"], mesh["abs_tol"], n, device=0)
fem.load_mesh(mesh)
fem.assemble()
w = nx + 1
fixed = np.arange(0, w, max(1, w // 64))[:64]
fem.add_displacement(np.repeat(fixed + 1, 6), np.tile(np.arange(6), len(fixed)), np.zeros(6 * len(fixed)))
for one_pass in (False, True):
    if one_pass:
        os.environ["FEMGPU_SEP_ONE_PASS"] = "1"
    for rep in range(2):
        t = time.perf_counter()
        n_aa, n_bb, nnz, ms = fem.separate_stiffness_matrix_sparse_iterative(copy_out=False)
        print(f"one_pass={one_pass} rep={rep}: {ms:.3f} ms (device), wall {1e3 * (time.perf_counter() - t):.1f} ms, read K once: {fem.last_separation_read_k_once()}, nnz {nnz}", flush=True)
fem.close()
