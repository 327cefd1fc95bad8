"""Shared by the full-size parity tests: sampled block rows of the GPU matrix against the oracle.

torch is only the test harness here (it wraps the library's device pointers and gathers the sampled rows on the
device); it computes no part of K."""
import numpy as np


class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_csr(fem, device="cuda"):
    import torch
    (rp, ci, v), (rb, re) = fem.csr_device()
    n_rows, nnz = fem.symbolic()
    row_ptr = torch.as_tensor(_DevArray(rp, n_rows + 1, "<i8"), device=device)
    col = torch.as_tensor(_DevArray(ci, nnz, "<i4"), device=device)
    val = torch.as_tensor(_DevArray(v, nnz, "<f8"), device=device)
    return row_ptr, col, val


def sample_nodes(n_nodes, grid_w=None, lo=0, hi=None, n_random=10_000, seed=20240701, lines=()):
    """Node sample inside [lo, hi): the first / last 72 nodes (first and last slab), `n_random` random ones (slab
    boundaries fall among them: a slab is ~8 nodes) and, for grid meshes, whole grid lines `lines` (rank cuts)."""
    hi = n_nodes if hi is None else hi
    parts = [np.arange(lo, min(hi, lo + 72)), np.arange(max(lo, hi - 72), hi)]
    rng = np.random.default_rng(seed)
    parts.append(rng.integers(lo, hi, size=min(n_random, hi - lo)))
    for j in lines:
        line = np.arange(j * grid_w, (j + 1) * grid_w)
        parts.append(line[(line >= lo) & (line < hi)])
    return np.unique(np.concatenate(parts)).astype(np.uint32)


def compare_sampled_rows(fem, mesh, nodes, rtol=1e-12, faithful=True, device="cuda", csr=None):
    """GPU CSR rows of `nodes` vs oracle.sample_rows: pattern (row lengths, column indices) identical, values within
    the SURVEY §8c bar: |a-b| <= rtol*max(|a|,|b|) or <= rtol * (largest entry of the 6x6 node-pair block).
    Returns a report dict; raises AssertionError on the first structural mismatch."""
    import torch
    from oracle import oracle as O
    ptr, bcol, bfull, bval = O.sample_rows(mesh, nodes, faithful=faithful)
    row_ptr, col, val = device_csr(fem, device) if csr is None else csr   # csr: torch tensors (harness self-test)
    rows = (6 * nodes.astype(np.int64)[:, None] + np.arange(7)[None, :]).ravel()
    rp = row_ptr[torch.as_tensor(rows, device=device)].cpu().numpy().reshape(-1, 7)
    # gather every value / column of the sampled rows on the device
    start, stop = rp[:, 0], rp[:, 6]
    counts = (stop - start).astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(counts)])
    idx = np.repeat(start - offs[:-1], counts) + np.arange(offs[-1])
    idx_t = torch.as_tensor(idx, device=device)
    g_val = val[idx_t].cpu().numpy()
    g_col = col[idx_t].cpu().numpy()
    n_bad = n_floor = n_total = 0
    worst_rel = worst_blk = 0.0
    for i in range(len(nodes)):
        b0, b1 = ptr[i], ptr[i + 1]
        full = bfull[b0:b1].astype(bool)
        V = bval[b0:b1]                                    # [nb, 6, 6]
        C = 6 * bcol[b0:b1].astype(np.int64)[:, None] + np.arange(6)[None, :]
        w_mask = np.where(full[:, None], True, np.arange(6)[None, :] < 3)
        scale = np.abs(V).reshape(len(V), -1).max(axis=1) if len(V) else np.zeros(0)
        S = np.broadcast_to(scale[:, None], (len(V), 6))
        seg = rp[i] - rp[i, 0]
        base = offs[i]
        for r in range(6):
            got_v = g_val[base + seg[r]:base + seg[r + 1]]
            got_c = g_col[base + seg[r]:base + seg[r + 1]]
            if r < 3:
                exp_v, exp_c, exp_s = V[:, r, :][w_mask], C[w_mask], S[w_mask]
            else:
                exp_v, exp_c, exp_s = V[full][:, r, :].ravel(), C[full].ravel(), S[full].ravel()
            assert len(got_v) == len(exp_v), (int(nodes[i]), r, len(got_v), len(exp_v))
            assert np.array_equal(got_c, exp_c), (int(nodes[i]), r)
            d = np.abs(got_v - exp_v)
            mag = np.maximum(np.abs(got_v), np.abs(exp_v))
            rel_ok = d <= rtol * mag
            blk_ok = d <= rtol * exp_s
            n_total += len(d)
            n_floor += int((~rel_ok & blk_ok).sum())
            n_bad += int((~rel_ok & ~blk_ok).sum())
            if len(d):
                with np.errstate(divide="ignore", invalid="ignore"):
                    rel = np.where(mag > 0, d / mag, 0.0)
                    blk = np.where(exp_s > 0, d / exp_s, 0.0)
                worst_rel = max(worst_rel, float(rel[rel_ok | ~blk_ok].max(initial=0.0)))
                worst_blk = max(worst_blk, float(blk.max(initial=0.0)))
    return {"nodes": int(len(nodes)), "rows": int(6 * len(nodes)), "entries": int(n_total), "n_fail": n_bad,
            "n_floor": n_floor, "max_rel": worst_rel, "max_block_rel": worst_blk}
