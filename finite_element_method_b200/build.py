"""In-tree build of libfemgpu.so (nvcc, sm_100a). No JIT cache: the .so lives next to this file."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfemgpu.so")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp")) or f == "Makefile"]
    srcs.append(os.path.join(HERE, "..", "include", "femgpu.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA translation unit for sm_100a and link libfemgpu.so.

    nvcc cross-compiles without a GPU. Raises on failure; never falls back to anything."""
    if force or _stale():
        if not os.path.exists("/usr/local/cuda/bin/nvcc") and not _which("nvcc"):
            raise RuntimeError("nvcc not found and libfemgpu.so is missing or stale")
        r = subprocess.run(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4), "all"],
                           capture_output=True, text=True)
        if verbose or r.returncode:
            print(r.stdout[-4000:])
            print(r.stderr[-4000:])
        if r.returncode:
            raise RuntimeError("building libfemgpu.so failed")
    return LIB


def _which(name):
    from shutil import which
    return which(name)


if __name__ == "__main__":
    print(build(force=False, verbose=True))
