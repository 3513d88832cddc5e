import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def chaos_lib():
    """The product library, built in-tree if needed (nvcc cross-compiles without a GPU)."""
    from gym_lorenz_b200 import _lib, build
    build.build_library()
    return _lib.load()


@pytest.fixture(scope="session")
def oracle_api():
    from oracle import api
    api.build()
    return api
