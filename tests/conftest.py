import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def pkg():
    """The product package under its import alias (builds the native library on first use if it is missing)."""
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ge.PKG_DIR, "libcrimac_b200.so")):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
