import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _device_count():
    """CUDA devices seen by the product library (0 when it is not built yet or there is no GPU)"""
    try:
        from composable_sdr_b200 import _lib
        return int(_lib.load().csdr_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # plain `pytest tests` on a box without a GPU: the gpu-marked tests are skipped, not run into create() failures
    if _device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
