"""Multi-GPU path on real GPUs (needs >= 2 devices): NCCL interface-row exchange vs a single-GPU
assembly of the same mesh."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("which", ["mixed", "plate", "truss"])
def test_two_rank_assembly_matches_single_gpu(which):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(4, _n_gpus()) if which == "mixed" else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "tests", "dist_worker.py"), which]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout, r.stdout[-2000:]
