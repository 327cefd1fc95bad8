"""The parallel host structures behind the batched adds (label tables claimed by all cores, duplicate indices filled
shard by shard, the parked worker pool) under ThreadSanitizer: no data race, and the sequential answers."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "finite_element_method_b200", "csrc")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "host_structures_tsan")


def test_host_structures_are_race_free_and_keep_the_sequential_answers():
    if not shutil.which("g++") or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("needs g++ and the CUDA headers")
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-I", CSRC, "-I/usr/local/cuda/include",
           os.path.join(ROOT, "tests", "cpp", "host_structures_tsan.cpp"), "-o", EXE, "-L/usr/local/cuda/lib64", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and ("tsan" in r.stderr or "sanitize" in r.stderr):
        pytest.skip("ThreadSanitizer runtime not installed")
    assert r.returncode == 0, r.stderr[-3000:]
    for threads in ("3", "8"):
        env = dict(os.environ, FEMGPU_HOST_THREADS=threads, LD_LIBRARY_PATH="/usr/local/cuda/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
        r = subprocess.run([EXE], capture_output=True, text=True, timeout=600, env=env)
        out = r.stdout + r.stderr
        assert r.returncode == 0 and "HOST_STRUCTURES_DONE" in out, out[-3000:]
        assert "WRONG" not in out and "ThreadSanitizer" not in out, out[-3000:]
