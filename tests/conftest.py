import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))  # test helpers (hwb_testutil)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (runs the product library, parity vs the oracle)')


@pytest.fixture(scope='session')
def built():
    from hwang_b200 import build
    build.build_gen()
    build.build_emu()
    return True


@pytest.fixture(scope='session')
def emu(built):
    """Bind the Python API to the host-emulation build of the device code (CPU test tier)."""
    from hwang_b200 import _lib, build
    _lib.use_library(build.EMU)
    return _lib


@pytest.fixture(scope='session')
def gpu():
    """Bind the Python API to the product library; every gpu-marked test goes through the C ABI."""
    from hwang_b200 import _lib
    if not os.path.exists(_lib.PRODUCT_LIB):
        from hwang_b200 import build
        build.build_product()
    from hwang_b200 import build
    build.build_gen()
    _lib.use_library(_lib.PRODUCT_LIB)
    import hwang_b200 as hw
    assert hw.device_count() > 0, 'gpu tests need a CUDA device'
    return _lib
