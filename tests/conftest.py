"""pytest configuration: `gpu` marker + shared fixtures.

CPU tests (-m "not gpu") cover the oracle against the reference's pins and golden fixtures, the
host logic and the C-ABI surface.  GPU tests (-m gpu) are the parity tests proper and call
through the C ABI (go-blosc_b200/lib/libb2b.so).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    o = entry.load_oracle()
    o.build()
    return o


@pytest.fixture(scope="session")
def pkg():
    entry.build()
    return entry.load_package()


@pytest.fixture(scope="session")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()
