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


def cfg2_inputs(g, dof_xy):
    """inputs of the config-2 script golden: stored in the small fixture, regenerated for the sampled 81^2 one exactly as
    tests/golden/make_golden.py:ref_script_cfg2 draws them (u0 closed form; default_rng(seed): c first, then the target)"""
    if "sample" not in g:
        return g["u0"], g["c"], g["uhat"]
    ns = int(g["ns"][0])
    u0 = np.exp(-20 * ((dof_xy[:, 0] + 2 / 3) ** 2 + 5 * (dof_xy[:, 1] + 5 / 6) ** 2))
    rng = np.random.default_rng(int(g["seed"][0]))
    L = (ns + 1) * u0.size
    c = 0.5 + rng.random(L)
    uhat = np.tile(u0, ns + 1) * (1.0 + 0.1 * rng.random(L))
    return u0, c, uhat


def golden_field_error(g, name, traj):
    """relative L2 error of a trajectory against golden field `name`: full field, or (sampled fixtures) every sample-th DoF of
    each time level together with the per-level 2-norms of the full field"""
    ns = int(g["ns"][0])
    traj = np.asarray(traj).reshape(ns + 1, -1)
    if "sample" not in g:
        return rel_l2(traj.ravel(), g[name])
    st = int(g["sample"][0])
    e_s = rel_l2(traj[:, ::st], g[name + "_s"])
    nrm = g[name + "_norm"]
    e_n = float(np.max(np.abs(np.linalg.norm(traj, axis=1) - nrm) / np.maximum(nrm, 1e-300)))
    return max(e_s, e_n)


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
