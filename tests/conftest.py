import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")
    config.addinivalue_line("markers", "slow: longer CPU test")


@pytest.fixture(scope="session")
def pkg():
    from util import gb
    return gb


@pytest.fixture(scope="session")
def host(pkg):
    import __graft_entry__ as entry
    entry.build()
    return pkg.engine()
