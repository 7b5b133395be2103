import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from tests import _oracle

    return _oracle.load_oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled in place (oracle/_ref/libgr4ref.so); skip when not built."""
    from tests import _oracle

    lib = _oracle.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libgr4ref.so not built (reference tree absent)")
    return lib
