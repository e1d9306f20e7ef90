import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    entry.build()
    return entry.load_package()


@pytest.fixture(scope="session")
def oracle():
    o = entry.load_oracle()
    o.build()
    return o


@pytest.fixture(scope="session")
def codes(pkg):
    return pkg.codes
