"""Quick GPU sanity run (not a test): parity of element matrices and small assemblies vs the oracle."""
import os, sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finite_element_method_b200 import FEM, FemError, meshes, TRUSS, BEAM, PLATE
from oracle import oracle as O


def check(mesh):
    t0 = time.time()
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], len(mesh["x"]))
    fem.load_mesh(mesh)
    n_rows, nnz = fem.assemble()
    rp, ci, v = fem.csr()
    K = sp.csr_matrix((v, ci, rp), shape=(n_rows, n_rows))
    r, c, ov = O.faithful_coo(mesh)
    Ko = sp.coo_matrix((ov, (r, c)), shape=(n_rows, n_rows)).tocsr()
    err = abs(K - Ko).max() / abs(Ko).max()
    v2 = fem.csr(values_only=True)
    fem.numeric(); fem.synchronize()
    v3 = fem.csr(values_only=True)
    print(f"{mesh['name']:40s} elems={meshes.n_elements(mesh):8d} nnz={nnz:10d} err={err:.2e} "
          f"deterministic={np.array_equal(v2, v3)} ms={fem.last_numeric_ms()} t={time.time()-t0:.2f}s", flush=True)
    fem.close()
    return err


if __name__ == "__main__":
    for m in [meshes.reference_truss_model(), meshes.truss_cube(3), meshes.truss_lattice(6, 10**9, jitter=True),
              meshes.beam_frame(5, 10**9), meshes.beam_frame(5, 10**9, jitter=True), meshes.plate_grid(6, 5, "flat"),
              meshes.plate_grid(6, 5, "jitter"), meshes.plate_grid(6, 5, "x0"), meshes.mixed_structure(6, 4),
              meshes.mixed_structure(40, 30), meshes.truss_lattice(20, 10**9), meshes.plate_grid(100, 80, "jitter")]:
        check(m)
    for big in [meshes.truss_lattice(64), meshes.beam_frame(88), meshes.plate_grid(2000, 2000), meshes.mixed_structure(2000, 2000)]:
        t0 = time.time()
        fem = FEM(big["rel_tol"], big["abs_tol"], len(big["x"]))
        fem.load_mesh(big); t1 = time.time()
        n_rows, nnz = fem.symbolic(); t2 = time.time()
        for _ in range(3):
            fem.numeric(); fem.synchronize()
        ms = fem.last_numeric_ms()
        ab = 8 * nnz
        print(f"{big['name']:40s} elems={meshes.n_elements(big)} nnz={nnz} load={t1-t0:.2f}s symbolic={t2-t1:.2f}s "
              f"numeric ms={ms} -> {meshes.n_elements(big)/ms[0]/1e6:.3f} Gelem/s, value-write {ab/ms[0]/1e6:.1f} GB/s "
              f"dev_bytes={fem.device_bytes()/1e9:.2f} GB", flush=True)
        fem.close()
