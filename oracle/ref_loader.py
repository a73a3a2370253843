"""Oracle: import the reference's own helpers.py unmodified, with dolfin/matplotlib stubbed.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Works only where /root/reference exists (the
build container); the GPU box has no reference tree, so nothing that runs there may call this.
It is used by tests/golden/make_golden.py to generate fixtures with the reference's *own*
FCT_alg_ref / ChebSI / artificial_diffusion_mat / L2_norm_sq_Q / cost_functional, and by the
CPU tests (skipped when the tree is absent) to pin oracle/fct_numpy.py against them.

Recipe: SURVEY.md App. F.1 -- helpers.py imports `dolfin` and `matplotlib.pyplot` at module top
(helpers.py:6,10-11); empty stand-in modules make the import succeed; everything that does not
touch df.* then runs as written.
"""
import importlib.util
import os
import sys
import types
import warnings

REFERENCE_DIR = os.environ.get("FCT_REFERENCE_DIR", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "helpers.py"))


def load_reference_helpers():
    """Return the reference's helpers module (cached in sys.modules as `_fct_reference_helpers`)."""
    name = "_fct_reference_helpers"
    if name in sys.modules:
        return sys.modules[name]
    if not reference_available():
        raise FileNotFoundError(f"{REFERENCE_DIR}/helpers.py not found")
    saved = {k: sys.modules.get(k) for k in ("dolfin", "matplotlib", "matplotlib.pyplot")}
    dolfin = types.ModuleType("dolfin")
    for attr in ("dx", "dot", "grad", "assemble", "exp", "div"):
        setattr(dolfin, attr, None)
    dolfin.TrialFunction = lambda V: None
    dolfin.TestFunction = lambda V: None
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["dolfin"] = dolfin
    if saved["matplotlib"] is None:
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_DIR, "helpers.py"))
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
        sys.modules[name] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load_reference_helpers_on_fake_dolfin():
    """The reference's helpers.py, unmodified, with `dolfin` bound to oracle/fake_dolfin.py (numpy P1 assembly behind dolfin's
    names): its time loops -- solve_schnak_system, solve_adjoint_schnak_system, solve_nonlinear_equation,
    solve_adjoint_nonlinear_equation, solve_chtxs_system, solve_adjoint_chtxs_system, armijo_line_search_ref -- then run as
    written.  Cached in sys.modules as `_fct_reference_helpers_fd`."""
    from . import fake_dolfin
    name = "_fct_reference_helpers_fd"
    if name in sys.modules:
        return sys.modules[name]
    if not reference_available():
        raise FileNotFoundError(f"{REFERENCE_DIR}/helpers.py not found")
    saved = {k: sys.modules.get(k) for k in ("dolfin", "matplotlib", "matplotlib.pyplot")}
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["dolfin"] = fake_dolfin.make_module()
    if saved["matplotlib"] is None:
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_DIR, "helpers.py"))
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
        sys.modules[name] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
