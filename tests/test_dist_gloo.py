"""N > 1 host-side logic on CPU (gloo, world_size 2): the row-strip partition, the
lowest-node ownership rule and the interface exchange plan, checked with the oracle doing the
arithmetic. Mirrors what libfemgpu does with NCCL: every rank assembles its own elements, sends the
rows it does not own to their owner, the owner adds partials in rank order."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from finite_element_method_b200 import meshes
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh = meshes.mixed_structure(10, 9)
    n = len(mesh["x"])
    parts = meshes.partition_rows(mesh, world, 11)
    begin, end = parts[rank]
    part = meshes.local_part(mesh, begin, end)
    n_rows = 6 * n
    r, c, v = O.faithful_coo(part)
    K = sp.coo_matrix((v, (r, c)), shape=(n_rows, n_rows)).tocsr()
    # ghost rows: everything outside my row range goes to its owner, dense per destination
    out = [None] * world
    for dst, (b, e) in enumerate(parts):
        if dst == rank:
            continue
        blk = K[6 * b:6 * e].tocoo()
        out[dst] = (blk.row + 6 * b, blk.col, blk.data)
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    mine = K[6 * begin:6 * end].tocsr()
    full = sp.csr_matrix((n_rows, n_rows))
    full = sp.vstack([sp.csr_matrix((6 * begin, n_rows)), mine, sp.csr_matrix((n_rows - 6 * end, n_rows))]).tocsr()
    exchanged = 0
    for src in range(world):               # fixed source-rank order
        msg = gathered[src][rank] if src != rank else None
        if msg is None:
            continue
        rr, cc, vv = msg
        exchanged += len(vv)
        assert np.all((rr >= 6 * begin) & (rr < 6 * end))
        full = full + sp.coo_matrix((vv, (rr, cc)), shape=(n_rows, n_rows)).tocsr()
    # with lowest-node ownership the exchange is one-directional: lower strip -> upper strip
    sent_to_lower = any(out[d] is not None and len(out[d][2]) for d in range(rank))
    fr, fc, fv = O.faithful_coo(mesh)
    ref = sp.coo_matrix((fv, (fr, fc)), shape=(n_rows, n_rows)).tocsr()[6 * begin:6 * end]
    err = abs(full[6 * begin:6 * end] - ref).max() / abs(ref).max()
    q.put((rank, float(err), int(exchanged), bool(sent_to_lower), meshes.n_elements(part)))
    dist.destroy_process_group()


def test_world2_partition_and_exchange_plan():
    import torch.multiprocessing as mp
    from finite_element_method_b200 import meshes
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29533
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = sum(r[4] for r in res)
    assert total == meshes.n_elements(meshes.mixed_structure(10, 9))
    for rank, err, exchanged, sent_to_lower, _ in res:
        assert err < 1e-13, (rank, err)
        assert not sent_to_lower
    assert res[0][2] == 0 and res[1][2] > 0          # rank 1 receives the interface rows of rank 0
