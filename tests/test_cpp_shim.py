"""The header-only C++ shim (include/sccd.hpp) compiles against the C ABI without Eigen and,
on a GPU, behaves like the reference's own tests expect."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "shim_example.cpp")
LIBDIR = os.path.join(ROOT, "scalable-ccd_b200")


def build(tmp_path, use_float=False):
    exe = str(tmp_path / ("shim_example_f32" if use_float else "shim_example"))
    subprocess.check_call(
        ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", f"-I{ROOT}/include"]
        + (["-DSCCD_SHIM_USE_FLOAT"] if use_float else [])  # the reference's float build
        + [SRC, "-o", exe, f"-L{LIBDIR}", "-lsccd_b200", f"-Wl,-rpath,{LIBDIR}"])
    return exe


@pytest.mark.parametrize("use_float", [False, True])
def test_shim_compiles_and_links(sccd, tmp_path, use_float):
    sccd.capi.load()
    assert os.path.exists(build(tmp_path, use_float))


@pytest.mark.gpu
@pytest.mark.parametrize("use_float", [False, True])
def test_shim_runs_like_the_reference_tests(sccd, tmp_path, use_float):
    exe = build(tmp_path, use_float)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "collisions=1" in out.stdout and "builders=1" in out.stdout
