import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load libmmrca.so.  Building only needs nvcc, not a GPU."""
    from garbage_classification_rca_b200 import _native
    _native.build()
    return _native.lib()
