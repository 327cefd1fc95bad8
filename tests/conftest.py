import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once (nvcc / g++ cross-compile without a GPU)."""
    import __graft_entry__ as g
    g.build()


def parity_report(n_rows, gpu_csr, oracle_coo, rtol=1e-12):
    """SURVEY.md §8c parity metric. Compare on the union of patterns with absent == 0. An entry passes
    iff |a-b| <= rtol*max(|a|,|b|) or |a-b| <= rtol*s, s = max |entry| of its 6x6 node-pair block
    (floor for entries that are pure rounding noise). Returns dict(max_rel, n_floor, n_fail, n)."""
    import scipy.sparse as sp
    rp, ci, v = gpu_csr
    A = sp.csr_matrix((v, ci, rp), shape=(n_rows, n_rows))
    r, c, ov = oracle_coo
    B = sp.coo_matrix((ov, (r, c)), shape=(n_rows, n_rows)).tocsr()
    # union pattern
    U = (abs(A) + abs(B)).tocoo()
    rows, cols = U.row, U.col
    a = np.asarray(A[rows, cols]).ravel()
    b = np.asarray(B[rows, cols]).ravel()
    diff = np.abs(a - b)
    mag = np.maximum(np.abs(a), np.abs(b))
    # block scale: max |entry| over the 6x6 node-pair block of the oracle/gpu union
    nb = n_rows // 6
    key = (rows // 6).astype(np.int64) * nb + (cols // 6)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    blockmax = np.maximum.reduceat(mag[order], starts)
    scale = np.empty_like(mag)
    scale[order] = np.repeat(blockmax, np.diff(np.r_[starts, len(ks)]))
    rel_ok = diff <= rtol * mag
    floor_ok = diff <= rtol * scale
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(mag > 0, diff / mag, 0.0)
    return {"n": len(a), "max_rel": float(rel[rel_ok | ~floor_ok].max()) if len(a) else 0.0,
            "n_floor": int((~rel_ok & floor_ok).sum()), "n_fail": int((~rel_ok & ~floor_ok).sum()),
            "max_block_rel": float((diff / np.where(scale > 0, scale, 1)).max()) if len(a) else 0.0}
