import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import ROOT, get_oracle  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_factory():
    return get_oracle


@pytest.fixture(scope="session")
def lib():
    """The C ABI (built in-tree); loading must work without a GPU."""
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    from cuhe_b200._lib import load_library
    return load_library()
