import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """libtadev.so through ctypes; building it is __graft_entry__.build()'s job."""
    from tiledarray_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


@pytest.fixture(scope="session")
def dev():
    """One Device (tadev_ctx) on cuda:0 for the whole GPU session. No fallback: if the library
    or the device is missing the GPU tests fail, they are not skipped."""
    from tiledarray_b200 import Device
    d = Device(0)
    yield d
    d.close()


@pytest.fixture(scope="session")
def world(dev):
    from tiledarray_b200.tiledarray import World
    w = World(device=dev)
    w.init_comm(1, 1)
    return w
