import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "tests", "emu")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import loader
    loader.build()
    return loader


@pytest.fixture(scope="session")
def emu():
    """Host emulation of the kernels (test tooling, see tests/emu/emu_lib.py); gives scenes + kernel-logic checks on CPU."""
    import emu_lib
    return emu_lib.load()


@pytest.fixture(scope="session")
def gpu_solver():
    from pl_slam_plucker_b200 import solver
    s = solver.LBASolver(0)          # raises if CUDA / libplba.so is unavailable: no fallback
    yield s
    s.close()
