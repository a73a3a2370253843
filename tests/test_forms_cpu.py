"""CPU tests of the UFL-like front end (fem-fct-pdeco_b200/forms.py): the reference's form expressions (the
call sites of helpers.py, old_helpers.py, mimura_data_helpers.py and the legacy scripts) must be recognised and turned into
the right C-ABI assembly calls.  A recording stand-in replaces the GPU context: no kernel runs here, the GPU tests
(test_gpu_forms.py) check the numbers."""
import numpy as np
import pytest

from fem_fct_pdeco_b200 import _lib as L
from fem_fct_pdeco_b200 import forms as F


class RecCtx:
    """records assemble_matrix / assemble_vector calls"""
    n, nnz = 9, 33

    def __init__(self):
        self.calls = []

    def empty(self, size):
        return ("buf", size)

    def array(self, host):
        return np.array(host, dtype=np.float64)

    def assemble_matrix(self, kind, out, **kw):
        self.calls.append(("M", kind, kw))

    def assemble_vector(self, kind, out, **kw):
        self.calls.append(("V", kind, kw))


class FakeMesh:
    nodes = 9

    def __init__(self):
        self.ctx = RecCtx()

    def context(self):
        return self.ctx


class FakeV:
    def __init__(self):
        self._m = FakeMesh()

    def mesh(self):
        return self._m

    def dim(self):
        return 9


def kinds(V, form, matrix=True):
    terms = F._terms(form)
    ctx = V.mesh().context()
    del ctx.calls[:]
    (F._assemble_matrix_terms if matrix else F._assemble_vector_terms)(ctx, terms)
    return [(c[1], c[2]) for c in ctx.calls]


def test_matrix_forms_of_the_reference_call_sites():
    V = FakeV()
    u, v = F.TrialFunction(V), F.TestFunction(V)
    f, g, h = (F.vec_to_function(np.arange(9.0) + k, V) for k in range(3))
    dx, dot, grad = F.dx, F.dot, F.grad
    assert [k for k, _ in kinds(V, u * v * dx)] == [L.FORM_MASS]                                   # helpers.py:553
    assert [k for k, _ in kinds(V, dot(grad(u), grad(v)) * dx)] == [L.FORM_STIFFNESS]            # :555
    (k, kw), = kinds(V, f * g * u * v * dx)                                                       # :591
    assert k == L.FORM_WMASS2 and kw["scale"] == 1.0 and np.array_equal(kw["c0"], f.vec) and np.array_equal(kw["c1"], g.vec)
    # a sum of two terms: the second one accumulates (Du*Ad-like combinations stay scipy arithmetic in the scripts)
    calls = kinds(V, 2.0 * u * v * dx + 0.5 * dot(grad(u), grad(v)) * dx)
    assert [c[0] for c in calls] == [L.FORM_MASS, L.FORM_STIFFNESS]
    assert [c[1]["scale"] for c in calls] == [2.0, 0.5] and [c[1]["accumulate"] for c in calls] == [False, True]
    # analytic polynomial wind, both orientations (helpers.py:581, 681)
    wind = F.Expression(("2*(x[1]-0.5)*x[0]*(1-x[0])", "-2*(x[0]-0.5)*x[1]*(1-x[1])"), degree=4)
    (k, kw), = kinds(V, dot(wind, grad(v)) * u * dx)
    assert k == L.FORM_WIND_POLY3 and kw["c0"].shape == (20,)
    (k, _), = kinds(V, dot(wind, grad(u)) * v * dx)
    assert k == L.FORM_WIND_POLY3_T
    # drift-control forms (advection_solidbody_FCT_PDECO_alltime.py:222-223)
    b = F.Constant((1.0, 1.0))
    (k, kw), = kinds(V, dot(b, grad(f)) * u * v * dx)
    assert k == L.FORM_DRIFT_MASS and (kw["s0"], kw["s1"]) == (1.0, 1.0)
    (k, kw), = kinds(V, dot(b, grad(v)) * f * u * dx)
    assert k == L.FORM_DRIFT_CONV
    # chemotaxis (helpers.py:1345-1346; old_helpers.py:102) and its adjoint (helpers.py:1494-1496: two terms, one kernel)
    (k, kw), = kinds(V, F.exp(-0.5 * f) * dot(grad(g), grad(v)) * u * dx)
    assert k == L.FORM_CHTX_EXP and kw["s0"] == 0.5
    (k, _), = kinds(V, dot(grad(g), grad(v)) * u * dx)
    assert k == L.FORM_CHTX
    (k, kw), = kinds(V, (1 - 0.5 * f) * F.exp(-0.5 * f) * dot(grad(u), grad(g)) * v * dx)
    assert k == L.FORM_CHTX_ADJ and kw["s0"] == 0.5
    with pytest.raises(NotImplementedError):
        kinds(V, f * g * h * f * u * v * dx)


def test_projected_wind_divergence_form():
    """div(w_h u) v = (w_h . grad u) v + div(w_h) u v with the P1 field of project(wind, W) (Schnak_FCT_PDECO.py:255-256)"""
    V = FakeV()
    u, v = F.TrialFunction(V), F.TestFunction(V)
    w = F.VecFunction(None, np.arange(9.0), -np.arange(9.0))
    calls = kinds(V, F.div(w * u) * v * F.dx)
    assert [c[0] for c in calls] == [L.FORM_WIND_P1_T, L.FORM_DIVW_MASS]
    assert all(np.array_equal(c[1]["c0"], w.wx) and np.array_equal(c[1]["c1"], w.wy) for c in calls)


def test_linear_forms_of_the_reference_call_sites():
    V = FakeV()
    v = F.TestFunction(V)
    f, g, h = (F.vec_to_function(np.arange(9.0) + k, V) for k in range(3))
    dx = F.dx
    (k, kw), = kinds(V, F.Constant(3.0) * v * dx, matrix=False)                                    # helpers.py:594
    assert k == L.LOAD_CONST and kw["scale"] == 3.0
    calls = kinds(V, (2.0 * f + g * g * h) * v * dx, matrix=False)                                 # :584-585
    assert [c[0] for c in calls] == [L.LOAD_P1_1, L.LOAD_P1_3] and [c[1]["scale"] for c in calls] == [2.0, 1.0]
    b = F.Constant((1.0, 1.0))
    (k, kw), = kinds(V, f * F.dot(b, F.grad(g)) * v * dx, matrix=False)            # advection_solidbody_FCT_PDECO_alltime.py:272
    assert k == L.LOAD_DRIFT_GRAD and np.array_equal(kw["c0"], f.vec) and np.array_equal(kw["c1"], g.vec)
    (k, kw), = kinds(V, 8.5 * f * F.exp(-0.5 * f) * F.dot(F.grad(g), F.grad(v)) * dx, matrix=False)   # helpers.py:1531-1532
    assert k == L.LOAD_CHTX_ADJ and kw["s0"] == 0.5 and kw["s1"] == 8.5


def test_expression_is_a_polynomial_with_settable_parameters():
    """Expression((...), degree=4, pi=np.pi, t=0) with `wind.t = t` inside the time loop (Schnak_FCT_PDECO.py:66-68,199)"""
    w = F.Expression(("-(x[1]-0.5)*sin(2*pi*t)", "(x[0]-0.5)*sin(2*pi*t)"), degree=4, pi=np.pi, t=0)
    assert np.allclose(w.coefs, 0.0)
    w.t = 0.25                                       # sin(pi/2) = 1:  wx = 0.5 - y,  wy = x - 0.5
    # monomial order 1, x, y, x^2, xy, y^2, ...   (include/fctpdeco.h, FCT_FORM_WIND_POLY3)
    assert np.allclose(w.coefs[0][:3], [0.5, 0.0, -1.0]) and np.allclose(w.coefs[1][:3], [-0.5, 1.0, 0.0])
    assert np.allclose(w.coefs[:, 3:], 0.0) and w.t == 0.25
    cubic = F.Expression(("2*(x[1]-0.5)*x[0]*(1-x[0])", "-2*(x[0]-0.5)*x[1]*(1-x[1])"), degree=4)
    # 2 (y - 1/2)(x - x^2) = -x + x^2 + 2xy - 2 x^2 y
    assert np.allclose(cubic.coefs[0], [0, -1, 0, 1, 2, 0, 0, -2, 0, 0])
    with pytest.raises(NotImplementedError):
        F.Expression("x[0]", degree=1)
