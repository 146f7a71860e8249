"""pytest configuration: the `gpu` marker, and the parity checkers as session fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


@pytest.fixture(scope="session")
def po():
    import pyoracle

    pyoracle.build(ref=True)
    return pyoracle


@pytest.fixture(scope="session")
def orc(po):
    return po.Checker("orc")


@pytest.fixture(scope="session")
def ref(po):
    if not po.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here and no prebuilt copy)")
    return po.Checker("ref")


@pytest.fixture(scope="session")
def ht(po):
    return po.HostTargets()


@pytest.fixture(scope="session")
def amx():
    """The CUDA library through its C-ABI.  No fallback: missing library or GPU is a failure."""
    from automix_b200 import _lib

    _lib.lib()
    assert _lib.device_count() > 0, "gpu-marked test without a CUDA device"
    return _lib
