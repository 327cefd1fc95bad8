"""Multi-GPU path on real GPUs (needs >= 2 devices): interface-row exchange (peer windows over NVLink, or
ncclSend/ncclRecv with FEMGPU_DIST_P2P=0) vs a single-GPU assembly of the same mesh, on every GPU of the box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(which, world, env_extra=None, port=29511, timeout=200):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), which]
    env = dict(os.environ, **(env_extra or {}))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout, r.stdout[-2000:]
    print(r.stdout[-400:])
    return r.stdout


@pytest.mark.parametrize("which", ["mixed", "plate", "truss"])
def test_multi_rank_assembly_matches_single_gpu(which):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(8, _n_gpus()) if which == "mixed" else 2
    _run(which, world)


def test_multi_rank_assembly_nccl_fallback_matches_single_gpu():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run("mixed", min(4, _n_gpus()), {"FEMGPU_DIST_P2P": "0"}, port=29512)
    assert "p2p=0" in out


def test_unmatched_passes_time_out_instead_of_hanging():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run("mismatch", 2, {"FEMGPU_P2P_TIMEOUT_MS": "300"}, port=29513)
    assert "timed out" in out or "nothing to test" in out


def test_full_size_mixed_mesh_in_strips_matches_oracle():
    """config 5: the 10M-element mesh split over every GPU of the box (>= 2), sampled rows vs the oracle"""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run("fullsize", min(8, _n_gpus()), port=29514, timeout=400)
    assert "sampled nodes" in out
