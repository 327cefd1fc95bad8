"""include/femgpu.hpp (the C++ host mirror) is compiled with g++ against libfemgpu.so and run: the reference crate's
own integration flow (src/tests/fem/test_fem.rs) replayed from C++. CPU: staging-only handle (host checks, error
texts, and the refusal to compute without a device); GPU: the whole flow."""
import os
import subprocess

import pytest

from finite_element_method_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "reference_flow")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "reference_flow.cpp")
    hdrs = [os.path.join(ROOT, "include", f) for f in ("femgpu.hpp", "femgpu.h")]
    libdir = os.path.dirname(_lib.LIB_PATH)
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) > os.path.getmtime(p) for p in [src, _lib.LIB_PATH] + hdrs):
        return
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
           "-L", libdir, "-lfemgpu", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def _run(device):
    _build()
    r = subprocess.run([EXE, str(device)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "CPP_HEADER_OK" in r.stdout
    return r.stdout


def test_cpp_mirror_compiles_and_runs_host_checks():
    assert "staging-only" in _run(-1)


@pytest.mark.gpu
def test_cpp_mirror_replays_the_reference_integration_tests():
    assert "device 0" in _run(0)
