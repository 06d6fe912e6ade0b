import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and load libb200chan.so; GPU tests fail loudly if it cannot be loaded."""
    from radiocapture_rf_b200 import build, _lib
    build.build_library()
    return _lib.load()


@pytest.fixture()
def engine(built_lib):
    from radiocapture_rf_b200.engine import Engine
    e = Engine(0)  # raises (does not skip) when there is no GPU: -m gpu tests must run the CUDA path
    yield e
    e.close()
