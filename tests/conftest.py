import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def mfhn():
    """The product package (ctypes binding of libmfhn.so); built on demand."""
    so = os.path.join(ROOT, "dealii-matrixfree-hanging-nodes_b200", "libmfhn.so")
    if not os.path.exists(so):
        import __graft_entry__

        __graft_entry__.build()
    return importlib.import_module("dealii-matrixfree-hanging-nodes_b200")
