import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def built_library():
    """Make sure the C-ABI library (and the C oracle) are built; returns the .so path."""
    import __graft_entry__ as entry
    entry.build()
    from graphnet_b200 import _lib
    return _lib.library_path()
