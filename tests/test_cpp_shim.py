"""The header-only C++ shim (include/sccd.hpp) compiles against the C ABI without Eigen and,
on a GPU, behaves like the reference's own tests expect."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "shim_example.cpp")
LIBDIR = os.path.join(ROOT, "scalable-ccd_b200")


def build(tmp_path):
    exe = str(tmp_path / "shim_example")
    subprocess.check_call(
        ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", f"-I{ROOT}/include", SRC, "-o", exe,
         f"-L{LIBDIR}", "-lsccd_b200", f"-Wl,-rpath,{LIBDIR}"])
    return exe


def test_shim_compiles_and_links(sccd, tmp_path):
    sccd.capi.load()
    assert os.path.exists(build(tmp_path))


@pytest.mark.gpu
def test_shim_runs_like_the_reference_tests(sccd, tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "collisions=1" in out.stdout
