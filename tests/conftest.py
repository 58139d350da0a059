import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def emu():
    from emu import emulib
    return emulib.load()


@pytest.fixture(scope="session")
def cs():
    """The product package with its CUDA library built (gpu tests)."""
    import composable_sdr_b200 as cs
    cs.build.build()
    return cs
