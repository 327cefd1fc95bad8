"""The C-ABI library loads and exports every symbol include/femgpu.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from finite_element_method_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "femgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(femgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    syms = _declared_symbols()
    assert len(syms) >= 25
    lib = _lib.load()
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in femgpu.h but not exported by libfemgpu.so"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(syms)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = _lib.H()
    st = lib.femgpu_create(C.byref(h), 1e-4, 1e-12, 2, 0)
    assert st == -5  # FEMGPU_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.femgpu_last_error(None)


def test_library_is_built_for_sm_100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    assert "sm_90" not in out and "sm_80" not in out


def test_product_never_references_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "finite_element_method_b200")):
        if "_obj" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"\boracle\b", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
