import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library, built in-tree on demand (nvcc cross-compiles without a GPU)."""
    from flashvtg_b200 import _build, _lib
    _build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def dbg_lib():
    """libflashvtg_b200_dbg.so: the product sources compiled with -DFVTG_DEBUG_HOOKS (+ probe.cu) - the kernel-level
    unit tests reach single kernels through its fvtg_dbg_* hooks; the product library exports none of them."""
    from flashvtg_b200 import _build, _lib
    _build.build()
    return _lib.load_debug()
