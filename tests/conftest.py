import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# TEST INFRASTRUCTURE: tests/test_emulated_kernels.py re-runs GPU test modules in a subprocess against the CPU emulation
# of libvga_b200 (tests/emu/): the ctypes loader of that subprocess is pointed at the emulation's scratch directory here.
# The product itself has no such switch.
if os.environ.get("VGA_EMU_LIBDIR"):
    from depthmapx_b200 import capi as _capi
    _capi.LIBDIR = os.environ["VGA_EMU_LIBDIR"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libdmxref.so (the compiled reference)")
    config.addinivalue_line("markers", "slow: minutes on the GPU box (the 10^6-cell configuration)")


def pytest_collection_modifyitems(config, items):
    """Every GPU test gets a hard time limit: a kernel that never returns must end the run, not hang the GPU box."""
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(1800, method="thread"))


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the product libraries and the oracle exist (prebuilt files travel to the GPU box)."""
    from depthmapx_b200 import capi
    from oracle import pyoracle as po
    if not (os.path.exists(os.path.join(capi.LIBDIR, "libvga_b200.so")) and
            os.path.exists(os.path.join(capi.LIBDIR, "libvga_host.so"))):
        from depthmapx_b200 import build
        build.build()
    if not po.have_oracle():
        po.build(ref=False)
    yield


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


GOLDEN = ["box2x2", "oblique20", "oblique16s07", "office24"]
