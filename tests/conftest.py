import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: BASELINE configurations at their real size (seconds each on a B200)")


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.build(ref=os.path.exists("/root/reference/cqf/gqf.c"))
    return pyoracle


@pytest.fixture(scope="session")
def shq():
    mod = importlib.import_module("sh-assembly_b200")
    if not os.path.exists(mod.LIB_PATH):
        mod.build()
    return mod
