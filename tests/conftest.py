import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "xpt-mde-2021_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_sessionfinish(session, exitstatus):
    """tests/helpers.margin() records the achieved error of every tolerance check it guards"""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import helpers
        helpers.dump_margins(os.path.join(ROOT, "gpurun_out", "parity_margins.json"))
    except Exception:
        pass
