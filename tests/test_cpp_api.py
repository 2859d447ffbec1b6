"""The C++ host mirror of the reference API (include/homomorph.hpp): compiled with g++ against libhmgpu.so and run —
CPU-only checks here, the reference's own end-to-end cases (tests/cpp/test_homomorph.cpp) on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "homomorph_rust_b200")
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(BUILD, "test_homomorph")


def build():
    import homomorph_rust_b200 as hm

    hm.lib()  # the library must exist (fail loudly otherwise)
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_homomorph.cpp")
    deps = [src, os.path.join(ROOT, "include", "homomorph.hpp"), os.path.join(ROOT, "include", "hmgpu.h")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-L", LIBDIR, "-lhmgpu",
                        f"-Wl,-rpath,{LIBDIR}", "-o", EXE], check=True, capture_output=True, text=True)
    return EXE


def test_cpp_mirror_cpu():
    r = subprocess.run([build()], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpu: 0 failure(s)" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_gpu():
    r = subprocess.run([build(), "--gpu"], capture_output=True, text=True, timeout=170)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gpu: 0 failure(s)" in r.stdout
