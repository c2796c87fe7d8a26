import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _have_ref():
    from oracle import pyref
    return pyref.available()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref/libsauref.so)."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/libsauref.so not built (needs /root/reference once)")
    return pyref


@pytest.fixture(scope="session")
def port():
    """The scalar CPU restatement (oracle/_ref/liboracle.so)."""
    from oracle import pyport
    if not pyport.available():
        pytest.skip("oracle/_ref/liboracle.so not built")
    return pyport
