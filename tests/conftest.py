import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref_data():
    return dict(np.load(os.path.join(GOLDEN, "ref_data.npz")))


@pytest.fixture(scope="session")
def ref_cases():
    return dict(np.load(os.path.join(GOLDEN, "ref_fct_cases.npz")))


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


@pytest.fixture(autouse=True)
def _guard_bands(request):
    """FCT_GUARD=1 pytest -m gpu: after every GPU test no canary band around a device buffer of libfctpdeco may have been
    written to (the stand-in for compute-sanitizer memcheck, which the GPU pool does not allow)"""
    yield
    if os.environ.get("FCT_GUARD", "0") != "1" or request.node.get_closest_marker("gpu") is None:
        return
    import ctypes as C
    from fem_fct_pdeco_b200 import _lib
    bad, live = C.c_int64(), C.c_int64()
    _lib.check(_lib.lib.fct_guard_check(C.byref(bad), C.byref(live)))
    print(f"FCT_GUARD: {live.value} live device buffers checked, {bad.value} corrupted")
    assert bad.value == 0, f"{bad.value} device buffer(s) had their canary bands overwritten"
