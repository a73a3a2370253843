"""Host-side entry points of the drop-in (no GPU involved) against the reference's OWN functions, imported unmodified from
/root/reference/helpers.py with dolfin/matplotlib stubbed (oracle/ref_loader.py).  Runs in the build container; skipped
where the reference tree is absent (the GPU box).  Covers the L0 row of SURVEY.md 8(a): DoF reordering, boundary nodes,
COO dump, relative error, parameter getters, trajectory import / extraction."""
import contextlib
import io
import os

import numpy as np
import pytest
import scipy.sparse as sp

from fem_fct_pdeco_b200 import helpers
from fem_fct_pdeco_b200.mesh import RectMeshP1
from oracle.ref_loader import load_reference_helpers, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return load_reference_helpers()


def _quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


@pytest.mark.parametrize("n,steps", [(5, 1), (12, 4)])
def test_reorder_vectors(ref, n, steps):
    m = RectMeshP1(n, 0.0, 1.0)
    v2d = m.vertex_to_dof
    rng = np.random.default_rng(n)
    vec = rng.random(steps * m.nodes)
    a = helpers.reorder_vector_to_dof(vec, steps, m.nodes, v2d)
    assert np.array_equal(a, ref.reorder_vector_to_dof(vec, steps, m.nodes, v2d))
    b = helpers.reorder_vector_from_dof(vec, steps, m.nodes, v2d)
    assert np.array_equal(b, ref.reorder_vector_from_dof(vec, steps, m.nodes, v2d))
    # legacy names the BASELINE-named scripts call (advection_solidbody_FCT.py:121,153)
    assert np.array_equal(helpers.reorder_vector_to_dof_time(vec, steps, m.nodes, v2d), a)
    assert np.array_equal(helpers.reorder_vector_from_dof_time(vec, steps, m.nodes, v2d), b)


def test_boundary_nodes_rel_err_sparse_nonzero(ref):
    m = RectMeshP1(9, -1.0, 1.0)
    got = helpers.generate_boundary_nodes(m.nodes, m.vertex_to_dof)
    exp = ref.generate_boundary_nodes(m.nodes, m.vertex_to_dof)
    assert len(got) == len(exp) and all(np.array_equal(np.asarray(g), np.asarray(e)) for g, e in zip(got, exp)) \
        if isinstance(exp, (tuple, list)) else np.array_equal(np.asarray(got), np.asarray(exp))
    rng = np.random.default_rng(1)
    a, b = rng.random(50), rng.random(50)
    assert helpers.rel_err(a, b) == ref.rel_err(a, b)
    H = sp.random(30, 30, density=0.2, random_state=3, format="lil") - sp.random(30, 30, density=0.1, random_state=4,
                                                                                 format="lil")
    assert np.array_equal(helpers.sparse_nonzero(sp.lil_matrix(H)), ref.sparse_nonzero(sp.lil_matrix(H)))


def test_parameter_getters(ref, monkeypatch):
    """scalar parameters equal; the winds are df.Expression objects in the reference (recorded here through a stand-in)
    and polynomial coefficient arrays in the drop-in (FCT_FORM_WIND_POLY3), checked against the expression strings"""
    monkeypatch.setattr(ref.df, "Expression", lambda *a, **k: ("Expression", a, k), raising=False)
    g, e = helpers.get_schnak_sys_params(), ref.get_schnak_sys_params()
    assert tuple(g[:7]) == tuple(e[:7])
    assert e[7][1][0] == ("1 * (x[1] - 0.5) * x[0] * (1 - x[0])", "-1 * (x[0] - 0.5) * x[1] * (1 - x[1])")
    # coefficient layout: wx then wy over the monomials 1, x, y, x^2, xy, y^2, x^3, x^2 y, x y^2, y^3
    x, y = 0.3, 0.8
    mono = np.array([1, x, y, x * x, x * y, y * y, x ** 3, x * x * y, x * y * y, y ** 3])
    w = np.asarray(g[7]).reshape(2, 10)
    assert abs(w[0] @ mono - (y - 0.5) * x * (1 - x)) < 1e-15 and abs(w[1] @ mono + (x - 0.5) * y * (1 - y)) < 1e-15
    g, e = helpers.get_nonlinear_eqns_params(), ref.get_nonlinear_eqns_params()
    assert tuple(g[:2]) == tuple(e[:2])
    w = np.asarray(g[2]).reshape(2, 10)
    exprs = e[2][1][0]
    xx = [x, y]                                          # the reference's strings are C expressions in x[0], x[1]
    ex = [eval(t.replace("x[0]", "xx[0]").replace("x[1]", "xx[1]"), {"xx": xx, "speed": e[1], **e[2][2]}) for t in exprs]
    assert abs(w[0] @ mono - ex[0]) < 1e-14 and abs(w[1] @ mono - ex[1]) < 1e-14
    assert tuple(helpers.get_chtxs_sys_params()) == tuple(ref.get_chtxs_sys_params())


def test_initial_conditions(ref, monkeypatch):
    """the np.arange-grid initial conditions bit for bit"""
    monkeypatch.setattr(ref.df, "Expression", lambda *a, **k: ("Expression", a, k), raising=False)   # schnak IC calls the getter
    for n, a1, a2 in ((10, 0.0, 1.0), (16, 0.0, 1.0)):
        m = RectMeshP1(n, a1, a2)
        dx = (a2 - a1) / n
        for name in ("schnak_sys_IC", "nonlinear_equation_IC", "chtxs_sys_IC"):
            got = getattr(helpers, name)(a1, a2, dx, m.nodes, m.vertex_to_dof)
            exp = getattr(ref, name)(a1, a2, dx, m.nodes, m.vertex_to_dof)
            if isinstance(exp, tuple):
                assert len(got) == len(exp) and all(np.array_equal(a, b) for a, b in zip(got, exp)), name
            else:
                assert np.array_equal(got, exp), name


def test_import_data_final_and_extract_data(ref, tmp_path):
    m = RectMeshP1(6, 0.0, 1.0)
    ns = 3
    rng = np.random.default_rng(2)
    traj = rng.random((ns + 1) * m.nodes)
    f = tmp_path / "traj.csv"
    traj.tofile(str(f), sep=",")                      # how the reference scripts write trajectories
    for kw in (dict(num_steps=ns, time_dep=True), dict(num_steps=2, time_dep=False), dict()):
        g_re, g = helpers.import_data_final(str(f), m.nodes, m.vertex_to_dof, **kw)
        e_re, e = ref.import_data_final(str(f), m.nodes, m.vertex_to_dof, **kw)
        assert np.array_equal(g, e) and np.array_equal(g_re, e_re)
    # extract_data writes <name>_T<T>.csv next to the input: same bytes from both
    d_ref, d_new = tmp_path / "ref", tmp_path / "new"
    for d in (d_ref, d_new):
        d.mkdir()
        traj.reshape(1, -1).tofile(str(d / "u.csv"), sep=",")
    dt, T = 0.25, 0.5
    _quiet(ref.extract_data, str(d_ref), "u", T, dt, m.nodes, m.vertex_to_dof)
    _quiet(helpers.extract_data, str(d_new), "u", T, dt, m.nodes, m.vertex_to_dof)
    assert (d_ref / f"u_T{T}.csv").read_bytes() == (d_new / f"u_T{T}.csv").read_bytes()
    # the binary path added for 4097^2-size trajectories holds the same numbers
    np.save(str(tmp_path / "traj.npy"), traj)
    b_re, b = helpers.import_data_final(str(tmp_path / "traj.npy"), m.nodes, m.vertex_to_dof, num_steps=ns, time_dep=True)
    e_re, e = ref.import_data_final(str(f), m.nodes, m.vertex_to_dof, num_steps=ns, time_dep=True)
    assert np.allclose(b, e, rtol=0, atol=0) and np.array_equal(b_re, e_re)


def test_cost_functional_error_behaviour(ref):
    """invalid `optim` raises ValueError with the reference's message in both (helpers.py:417-419)"""
    M = sp.identity(4, format="csr")
    x = np.ones(8)
    with pytest.raises(ValueError) as e_ref:
        _quiet(ref.cost_functional, x, x, x, 1, 0.1, M, 0.1, "sometimes")
    with pytest.raises(ValueError) as e_new:
        helpers.cost_functional(x, x, x, 1, 0.1, M, 0.1, "sometimes")
    assert str(e_new.value) == str(e_ref.value)


def test_mimura_initial_condition_vs_reference():
    """mimura_data_helpers.m_initial_condition (config 3, chemotaxis_mimura_FCT_PGD.py:98): the reference's function is
    executed from its source (its module imports dolfin/matplotlib at the top; only the function text is compiled)"""
    import re
    from fem_fct_pdeco_b200 import mimura_data_helpers as mdh
    from oracle.ref_loader import REFERENCE_DIR
    src = open(os.path.join(REFERENCE_DIR, "mimura_data_helpers.py")).read()
    m = re.search(r"^def m_initial_condition\(.*?(?=^def )", src, flags=re.S | re.M)
    ns = {"np": np}
    exec(compile(m.group(0), "mimura_data_helpers.py:m_initial_condition", "exec"), ns)
    for a1, a2, dx in ((0.0, 16.0, 0.125), (0.0, 1.0, 0.1)):
        assert np.array_equal(mdh.m_initial_condition(a1, a2, dx), ns["m_initial_condition"](a1, a2, dx))


def test_boundary_postprocessing_helpers_vs_reference(ref):
    """smooth_corners_on_boundary / rescale_boundary_nodes (helpers.py:2003-2121): vectorised here, compared with the
    reference's loops bit for bit; norm_true_control's error behaviour (helpers.py:1982-1984)"""
    n = 7
    m = RectMeshP1(n, 0.0, 1.0)
    rng = np.random.default_rng(4)
    vec = rng.random(m.nodes)
    got = helpers.smooth_corners_on_boundary(vec, None, m.vertex_to_dof, 0.0, 1.0, 1.0 / n)
    exp = ref.smooth_corners_on_boundary(vec, None, m.vertex_to_dof, 0.0, 1.0, 1.0 / n)
    assert np.array_equal(got, exp) and not np.array_equal(got, vec)
    got = helpers.rescale_boundary_nodes(vec, m.vertex_to_dof, a1=0.0, a2=1.0, deltax=1.0 / n)
    exp = ref.rescale_boundary_nodes(vec, m.vertex_to_dof, a1=0.0, a2=1.0, deltax=1.0 / n)
    assert np.array_equal(got, exp) and not np.array_equal(got, vec)
    with pytest.raises(ValueError) as e_ref:
        ref.norm_true_control("linear", 1, 0.1, None, None)
    with pytest.raises(ValueError) as e_new:
        helpers.norm_true_control("linear", 1, 0.1, None, None)
    assert str(e_new.value) == str(e_ref.value)
