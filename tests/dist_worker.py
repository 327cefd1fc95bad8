"""Worker for the multi-GPU test: launched by torch.distributed.run, one rank per GPU.
Each rank assembles its row strip of a mixed mesh with the NCCL interface exchange and checks its
owned rows against a single-GPU assembly of the whole mesh made on the same device."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from finite_element_method_b200 import FEM, meshes

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    which = sys.argv[1] if len(sys.argv) > 1 else "mixed"
    if which == "fullsize":
        return fullsize(rank, world, local, dist, torch)
    if which in ("mixed", "mismatch"):
        mesh, width = meshes.mixed_structure(40, 36), 41
    elif which == "plate":
        mesh, width = meshes.plate_grid(50, 31, "jitter"), 51
    else:
        mesh, width = meshes.truss_lattice(10, 10 ** 9, jitter=True), None
    n = len(mesh["x"])
    begin, end = meshes.partition_rows(mesh, world, width)[rank]
    # the rank is given its own elements and only the window of nodes they touch (own rows + halo)
    part = meshes.with_node_window(meshes.local_part(mesh, begin, end), begin, end)
    assert len(part["x"]) < n or world == 1

    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], n, device=local)
    uid = [FEM.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    fem.dist_init(rank, world, uid[0])
    fem.dist_set_ownership(begin, end)
    fem.load_mesh(part)
    n_rows, nnz = fem.assemble()
    rp, ci, v = fem.csr()
    v1 = v.copy()
    fem.numeric(); fem.synchronize()
    assert np.array_equal(v1, fem.csr(values_only=True)), "dist re-assembly is not bit-identical"
    sent, recv = fem.dist_last_exchange_bytes()
    p2p, passes = fem.dist_info()
    assert passes == 2, passes
    if which == "mismatch":
        # A rank that runs passes its neighbours never run must get an error, not a hung GPU (the
        # 8-GPU bench hung in an unmatched ncclSend). Rank 0 only sends; its ring of two slots is full after two
        # unanswered passes and the third one times out (FEMGPU_P2P_TIMEOUT_MS is set short by the test).
        from finite_element_method_b200 import FemError
        msg = "ncclSend/ncclRecv exchange (no peer access): nothing to test"
        if p2p and rank == 0 and world > 1:
            for _ in range(3):
                fem.numeric()
            try:
                fem.synchronize()
                raise AssertionError("three unanswered passes did not time out")
            except FemError as e:
                assert e.code == -3 and "timed out" in str(e), (e.code, str(e))
                msg = str(e)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            print(f"DIST_OK world={world} p2p={p2p} mismatch -> {msg}")
        fem.close()
        dist.destroy_process_group()
        return
    # FEM::reset keeps the communicator: the same handle assembles the model again from scratch
    fem.reset(n)
    fem.dist_set_ownership(begin, end)
    fem.load_mesh(part)
    assert fem.assemble() == (n_rows, nnz)
    assert np.array_equal(v1, fem.csr(values_only=True)), "assembly after reset differs"

    ref = FEM(mesh["rel_tol"], mesh["abs_tol"], n, device=local)
    ref.load_mesh(mesh)
    ref.assemble()
    rrp, rci, rv = ref.csr()
    lo, hi = 6 * begin, 6 * end
    assert np.array_equal(np.diff(rp[lo:hi + 1]), np.diff(rrp[lo:hi + 1])), "owned row lengths differ"
    a0, a1, b0, b1 = rp[lo], rp[hi], rrp[lo], rrp[hi]
    assert np.array_equal(ci[a0:a1], rci[b0:b1]), "owned column indices differ"
    scale = np.abs(rv).max()
    err = np.abs(v[a0:a1] - rv[b0:b1]).max() / scale
    assert err < 1e-14, err
    tot = torch.tensor([float(meshes.n_elements(part)), float(sent), float(recv)], device="cuda", dtype=torch.float64)
    dist.all_reduce(tot)
    if rank == 0:
        assert int(tot[0].item()) == meshes.n_elements(mesh)
        assert tot[1].item() == tot[2].item() and (world == 1 or tot[1].item() > 0)
        print(f"DIST_OK world={world} p2p={p2p} mesh={mesh['name']} max_err={err:.2e} exchanged_bytes={int(tot[1].item())}")
    torch.cuda.synchronize()
    dist.barrier()
    fem.close(); ref.close()
    dist.destroy_process_group()


def fullsize(rank, world, local, dist, torch):
    """BASELINE.json config 5 as the north star states it: ONE 10M-element mixed mesh partitioned into `world`
    contiguous row strips. Every rank compares sampled rows of the strip it owns — among them its first grid line
    (which received the lower neighbour's ghost contributions) and its last one — with the oracle."""
    import faulthandler
    faulthandler.dump_traceback_later(360, exit=True)    # 10M elements per rank to build and check: allow more time
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from finite_element_method_b200 import FEM, meshes
    from fullsize_common import compare_sampled_rows, sample_nodes
    nx = ny = 2000
    w = nx + 1
    mesh = meshes.mixed_structure(nx, ny)
    n = len(mesh["x"])
    begin, end = meshes.partition_rows(mesh, world, w)[rank]
    part = meshes.with_node_window(meshes.mixed_structure(nx, ny, rows=(begin // w, end // w)), begin, end)
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], n, device=local)
    uid = [FEM.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    fem.dist_init(rank, world, uid[0])
    fem.dist_set_ownership(begin, end)
    fem.load_mesh(part)
    fem.assemble()
    j0, j1 = begin // w, end // w
    nodes = sample_nodes(n, w, lo=begin, hi=end, n_random=max(2000, 10_000 // world), lines=(j0, j0 + 1, j1 - 1))
    rep = compare_sampled_rows(fem, mesh, nodes, rtol=1e-12, faithful=True, device=f"cuda:{local}")
    assert rep["n_fail"] == 0 and rep["max_block_rel"] <= 1e-12, (rank, rep)
    # the instance re-used (FEM::reset): staging registered with CUDA, every batch uploaded as it is added — with a
    # node window and ghost rows in play. Same strip, same bits.
    v1 = fem.csr(values_only=True).copy()
    for _ in range(2):
        fem.reset(n)
        fem.dist_set_ownership(begin, end)
        fem.load_mesh(part, cache=True)
        fem.assemble()
        assert np.array_equal(v1, fem.csr(values_only=True)), "assembly of the re-used handle differs"
    del v1
    tot = torch.tensor([float(rep["nodes"]), float(rep["entries"]), rep["max_block_rel"]], device="cuda", dtype=torch.float64)
    mx = tot.clone()
    dist.all_reduce(tot)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"DIST_OK world={world} p2p={fem.dist_info()[0]} mesh={mesh['name']} strips: sampled nodes={int(tot[0].item())} "
              f"entries={int(tot[1].item())} max_block_rel={mx[2].item():.2e}")
    fem.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    import faulthandler
    import traceback
    faulthandler.dump_traceback_later(150, exit=True)   # a stuck rank must not hold the test (and the GPU box) for long
    try:
        main()
    except BaseException:
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)                                      # no atexit / NCCL teardown that could wait for the other ranks
