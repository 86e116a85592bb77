import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE = os.path.join(ROOT, "oracle")
if ORACLE not in sys.path:
    sys.path.insert(0, ORACLE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cb():
    """the product package with its CUDA library built; never falls back to anything else."""
    import cbird_b200
    from cbird_b200 import build

    build.build()
    cbird_b200.lib()
    return cbird_b200


@pytest.fixture(scope="session")
def po():
    """the parity checkers (test infrastructure)."""
    import pyoracle

    pyoracle.build()
    return pyoracle
