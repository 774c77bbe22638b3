import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure), built on demand from oracle/gibbs_oracle.c."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def sg():
    """The product package; loading the shared library must not need a GPU."""
    import signer_b200
    from signer_b200 import _lib
    _lib.lib()
    return signer_b200
