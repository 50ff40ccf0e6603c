import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine_lib():
    """librekf_b200.so, (re)built in-tree if stale.  Loading needs no GPU."""
    from reflector_ekf_slam_b200 import build, engine
    build.build()
    return engine.load_library()
