import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from _pkg import load_package  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def sccd():
    return load_package()


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as _orc
    _orc.lib()
    return _orc


@pytest.fixture(scope="session")
def scene_c1(sccd):
    return sccd.scenes.scene_c1()


@pytest.fixture(scope="session")
def scene_small(sccd):
    """~6K boxes: cloth 31x31 over the UV sphere (brute-force sized)."""
    return sccd.scenes.cloth_on_sphere(31, seed=7, sphere="uv")


@pytest.fixture(scope="session")
def ctx(sccd):
    c = sccd.Context(0)
    yield c
    c.close()
