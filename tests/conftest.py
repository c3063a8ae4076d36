import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def b200_lib():
    """Build (if needed) and load libfestim_b200.so."""
    import __graft_entry__ as g

    g.build()
    from festim_b200.backend import load_library

    return load_library()
