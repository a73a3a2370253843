"""GPU parity tests of the reference-named time loops (solve_*_system, solve_adjoint_*, armijo_line_search_ref)
against the oracle restatements of helpers.py:511-698, 881-1038, 1250-1581, 1583-1713 on small meshes, and against
outputs of the reference's OWN functions (helpers.py run unmodified on oracle/fake_dolfin.py: tests/golden/ref_loops.npz).
Tolerance: 1e-12 relative L2 per time step (accumulating over the steps of a trajectory)."""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest

from conftest import rel_l2
from fem_fct_pdeco_b200 import helpers as hp
from fem_fct_pdeco_b200.mesh import FunctionSpaceP1, RectMeshP1, vertex_to_dof_map
from oracle import pdeco_systems as osys
from oracle.fct_numpy import cost_functional as o_cost

pytestmark = pytest.mark.gpu


def _quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _space(n):
    mesh = RectMeshP1(n, 0.0, 1.0)
    return mesh, FunctionSpaceP1(mesh)


def test_schnak_state_and_adjoint():
    n, ns, dt = 10, 4, 1e-3
    mesh, V = _space(n)
    nodes = V.dim()
    orc = osys.SchnakProblem(n, 0.0, 1.0)
    u0, v0 = hp.schnak_sys_IC(0.0, 1.0, 1.0 / n, nodes, vertex_to_dof_map(V))
    uo0, vo0 = orc.initial_condition()
    assert np.array_equal(u0, uo0) and np.array_equal(v0, vo0)
    rng = np.random.default_rng(2)
    c = 0.1 + 0.05 * rng.random((ns + 1) * nodes)
    u_o, v_o = orc.state(c, u0, v0, ns, dt, rescaling=1.0)
    uk = np.zeros((ns + 1) * nodes); vk = np.zeros_like(uk)
    uk[:nodes], vk[:nodes] = u0, v0
    r1, r2 = _quiet(hp.solve_schnak_system, c, uk, vk, V, nodes, ns, dt, mesh.dof_neighbors())
    assert r1 is uk and r2 is vk                                        # in place AND returned
    for i in range(1, ns + 1):
        assert rel_l2(uk.reshape(ns + 1, -1)[i], u_o[i]) < 1e-12 * i, i
        assert rel_l2(vk.reshape(ns + 1, -1)[i], v_o[i]) < 1e-12 * i, i
    uhat, vhat = u_o[-1] * 1.01 + 0.01, v_o[-1] * 0.99
    p_o, q_o = orc.adjoint(u_o, v_o, uhat, vhat, ns, dt)
    pk = np.zeros_like(uk); qk = np.zeros_like(uk)
    _quiet(hp.solve_adjoint_schnak_system, u_o.ravel().copy(), v_o.ravel().copy(), uhat, vhat, pk, qk, ns * dt, V, nodes,
           ns, dt, mesh.dof_neighbors())
    assert rel_l2(pk, p_o.ravel()) < 1e-11 and rel_l2(qk, q_o.ravel()) < 1e-11


def test_nonlinear_state_adjoint_and_armijo():
    n, ns, dt = 12, 5, 2e-3
    mesh, V = _space(n)
    nodes = V.dim()
    orc = osys.NonlinearProblem(n, 0.0, 1.0)
    u0 = hp.nonlinear_equation_IC(0.0, 1.0, 1.0 / n, nodes, vertex_to_dof_map(V))
    assert np.array_equal(u0, orc.initial_condition())
    rng = np.random.default_rng(3)
    c = rng.random((ns + 1) * nodes)
    u_o = orc.state(c, u0, ns, dt)
    uk = np.zeros((ns + 1) * nodes); uk[:nodes] = u0
    with pytest.warns(UserWarning):
        _quiet(hp.solve_nonlinear_equation, c, uk.copy(), np.zeros(3), V, nodes, ns, dt, mesh.dof_neighbors())
    out, none = _quiet(hp.solve_nonlinear_equation, c, uk, None, V, nodes, ns, dt, mesh.dof_neighbors())
    assert out is uk and none is None
    assert rel_l2(uk, u_o.ravel()) < 1e-11
    uhat_T = u_o[-1] + 0.05 * rng.random(nodes)
    p_o = orc.adjoint(u_o, uhat_T, ns, dt)
    pk = np.zeros_like(uk)
    _quiet(hp.solve_adjoint_nonlinear_equation, u_o.ravel().copy(), uhat_T, pk, ns * dt, V, nodes, ns, dt,
           mesh.dof_neighbors())
    assert rel_l2(pk, p_o.ravel()) < 1e-11
    # projected Armijo line search, final-time tracking (nonlinear_FCT_PDECO_refactored.py:148-152)
    beta = 0.1
    d = -(beta * c - pk)
    cost0 = o_cost(orc.pat, u_o.ravel(), uhat_T, c, ns, dt, orc.M, beta, "finaltime")
    solver = lambda ci: (orc.state(ci, u0, ns, dt).ravel(), None)
    v1_o, _, c_o, k_o = osys.armijo_ref(orc, solver, u_o.ravel(), c, d, uhat_T, ns, dt, 0.0, 1.0, beta, cost0, "finaltime")
    res = _quiet(hp.armijo_line_search_ref, uk.copy(), c, d, uhat_T, ns, dt, 0.0, 1.0, beta, cost0, nodes, "finaltime", V,
                 nonlinear_solver=hp.solve_nonlinear_equation, dof_neighbors=mesh.dof_neighbors())
    assert len(res) == 3                                               # (var1, c_inc, k+1) when var2 is None
    v1_g, c_g, k_g = res
    assert k_g == k_o and np.array_equal(c_g, c_o) and rel_l2(v1_g, v1_o) < 1e-11
    with pytest.raises(ValueError):
        hp.armijo_line_search_ref(uk, c, d, uhat_T, ns, dt, 0.0, 1.0, beta, cost0, nodes, "never", V)


def test_chemotaxis_state_and_adjoint(ref_data):
    n, ns, dt = 40, 3, 1e-3
    mesh, V = _space(n)
    nodes = V.dim()
    m0, f0 = hp.chtxs_sys_IC(0.0, 1.0, 1.0 / n, nodes, vertex_to_dof_map(V))
    assert np.array_equal(m0, ref_data["chtxs_m"][0])
    mk = np.zeros((ns + 1) * nodes); fk = np.zeros_like(mk)
    mk[:nodes], fk[:nodes] = m0, f0
    # the configuration that produced the reference's shipped trajectory: Constant(100), rescaling=1
    _quiet(hp.solve_chtxs_system, np.zeros(nodes), mk, fk, V, nodes, ns, dt, mesh.dof_neighbors(), control_fun=100.0,
           rescaling=1)
    for i in range(1, ns + 1):
        assert rel_l2(mk.reshape(ns + 1, -1)[i], ref_data["chtxs_m"][i]) < 1e-12
        assert rel_l2(fk.reshape(ns + 1, -1)[i], ref_data["chtxs_f"][i]) < 1e-12
    # vector control (stale step-1 slice, App. D-1) against the oracle
    orc = osys.ChemotaxisAdjoint(n, 0.0, 1.0)
    rng = np.random.default_rng(4)
    c = 50 + 10 * rng.random((ns + 1) * nodes)
    m_o, f_o = orc.forward(c, m0, f0, ns, dt)
    _quiet(hp.solve_chtxs_system, c, mk, fk, V, nodes, ns, dt, mesh.dof_neighbors())
    assert rel_l2(mk, m_o.ravel()) < 1e-12 and rel_l2(fk, f_o.ravel()) < 1e-12
    for optim, uh, vh in (("finaltime", m_o[-1] * 1.02, f_o[-1] * 0.98), ("alltime", m_o * 1.02, f_o * 0.98)):
        p_o, q_o = orc.adjoint(m_o, f_o, uh, vh, c, ns, dt, optim)
        pk = np.zeros_like(mk); qk = np.zeros_like(mk)
        _quiet(hp.solve_adjoint_chtxs_system, m_o.ravel().copy(), f_o.ravel().copy(), np.ravel(uh), np.ravel(vh), pk, qk, c,
               ns * dt, V, nodes, ns, dt, mesh.dof_neighbors(), optim)
        assert rel_l2(pk, p_o.ravel()) < 1e-11 and rel_l2(qk, q_o.ravel()) < 1e-11, optim
    with pytest.raises(ValueError):
        hp.solve_adjoint_chtxs_system(mk, fk, mk, fk, mk, fk, c, 1.0, V, nodes, ns, dt, None, "sometimes")


def test_config2_projected_gradient_iteration():
    """One projected-gradient iteration of advection_solidbody_FCT_PDECO_alltime.py:196-303 (config 2, the shape of
    the 4096^2 benchmark) on a 21x21 mesh: state, adjoint, gradient (device loops), legacy Armijo search, projection,
    cost functional -- against the oracle.  Final cost within 1e-9 (north_star)."""
    from oracle import pdeco_numpy as drv
    n, ns = 20, 6
    a1, a2 = -1.0, 1.0
    beta, c_lower, c_upper = 0.01, 0.0, 5.0
    orc = drv.AdvectionDriftPDECO(n, a1, a2, beta=beta, c_lower=c_lower, c_upper=c_upper)
    h = (a2 - a1) / n
    dt = 0.25 * h / (2 * np.sqrt(2))
    u0 = orc.gaussian_ic()
    uhat = orc.target(u0, ns, dt, c_const=2.0)
    c = np.ones((ns + 1, orc.nodes))
    u_o = orc.state(c, u0, ns, dt)
    p_o = orc.adjoint(c, u_o, uhat, ns, dt)
    d_o = orc.gradient(c, u_o, p_o, ns)
    s_o, uinc_o, k_o = orc.armijo(u0, c, d_o, uhat, ns, dt)
    c1_o = np.clip(c + s_o * d_o, c_lower, c_upper)
    J_o = orc.cost(uinc_o, uhat, c1_o, ns, dt)

    mesh, V = RectMeshP1(n, a1, a2), None
    V = FunctionSpaceP1(mesh)
    nodes = V.dim()
    ctx = mesh.context()
    M = ctx.to_scipy(ctx.static()[0].download())
    uk = np.zeros((ns + 1) * nodes); uk[:nodes] = u0
    dc, du, duh = ctx.array(c.ravel()), ctx.array(uk), ctx.array(uhat.ravel())
    dp, dd = ctx.empty(uk.size), ctx.empty(uk.size)
    ctx.advdrift_state(dc, du, ns, dt)
    ctx.advdrift_adjoint(dc, du, duh, dp, ns, dt)
    ctx.advdrift_gradient(dc, du, dp, dd, ns, beta)
    uk, pk, dk = du.download(), dp.download(), dd.download()
    assert rel_l2(uk, u_o.ravel()) < 1e-11 and rel_l2(pk, p_o.ravel()) < 1e-11 and rel_l2(dk, d_o.ravel()) < 1e-11
    sk, u_inc = _quiet(hp.armijo_line_search_sbr_drift, uk, pk, c.ravel(), dk, uhat.ravel(), 0.0, (1.0, 1.0), ns, dt, nodes,
                       M, None, None, None, c_lower, c_upper, beta, V, mesh.dof_neighbors(), optim="alltime")
    assert sk == s_o and u_inc is uk
    ckp1 = np.clip(c.ravel() + sk * dk, c_lower, c_upper)
    J_g = _quiet(hp.cost_functional, u_inc, uhat.ravel(), ckp1, ns, dt, M, beta, optim="alltime")
    assert rel_l2(u_inc, uinc_o.ravel()) < 1e-11
    assert abs(J_g / J_o - 1) < 1e-9


# ---- the reference's OWN loops (helpers.py run unmodified on oracle/fake_dolfin.py; tests/golden/ref_loops.npz) -------------
@pytest.fixture(scope="module")
def loops():
    import os
    from conftest import GOLDEN
    return dict(np.load(os.path.join(GOLDEN, "ref_loops.npz")))


def test_reference_loops_schnak(loops):
    """solve_schnak_system / solve_adjoint_schnak_system (helpers.py:511-698) against outputs of the reference's own functions"""
    g = loops
    n, ns, dt = int(g["n"][0]), int(g["ns"][0]), float(g["schnak_dt"][0])
    mesh, V = _space(n)
    nodes = V.dim()
    uk = np.zeros((ns + 1) * nodes); vk = np.zeros_like(uk)
    uk[:nodes], vk[:nodes] = g["schnak_u0"], g["schnak_v0"]
    _quiet(hp.solve_schnak_system, g["schnak_c"], uk, vk, V, nodes, ns, dt, mesh.dof_neighbors())
    ru, rv = g["schnak_u"].reshape(ns + 1, -1), g["schnak_v"].reshape(ns + 1, -1)
    for i in range(1, ns + 1):
        assert rel_l2(uk.reshape(ns + 1, -1)[i], ru[i]) < 1e-12 * i, i
        assert rel_l2(vk.reshape(ns + 1, -1)[i], rv[i]) < 1e-12 * i, i
    pk = np.zeros_like(uk); qk = np.zeros_like(uk)
    _quiet(hp.solve_adjoint_schnak_system, g["schnak_u"].copy(), g["schnak_v"].copy(), g["schnak_uhat"], g["schnak_vhat"], pk, qk,
           ns * dt, V, nodes, ns, dt, mesh.dof_neighbors())
    assert rel_l2(pk, g["schnak_p"]) < 1e-11 and rel_l2(qk, g["schnak_q"]) < 1e-11


def test_reference_loops_nonlinear_and_armijo(loops):
    """solve_nonlinear_equation / solve_adjoint_nonlinear_equation (helpers.py:881-1038) and armijo_line_search_ref
    (:1583-1713, the reference's own solver as callback) against outputs of the reference's own functions"""
    g = loops
    n, ns, dt = int(g["n"][0]), int(g["ns"][0]), float(g["nonlin_dt"][0])
    mesh, V = _space(n)
    nodes = V.dim()
    uk = np.zeros((ns + 1) * nodes); uk[:nodes] = g["nonlin_u0"]
    _quiet(hp.solve_nonlinear_equation, g["nonlin_c"], uk, None, V, nodes, ns, dt, mesh.dof_neighbors())
    assert rel_l2(uk, g["nonlin_u"]) < 1e-11
    pk = np.zeros_like(uk)
    _quiet(hp.solve_adjoint_nonlinear_equation, g["nonlin_u"].copy(), g["nonlin_uhat"], pk, ns * dt, V, nodes, ns, dt,
           mesh.dof_neighbors())
    assert rel_l2(pk, g["nonlin_p"]) < 1e-11
    lo, hi = g["armijo_bounds"]
    res = _quiet(hp.armijo_line_search_ref, g["nonlin_u"].copy(), g["nonlin_c"], g["armijo_d"], g["nonlin_uhat"], ns, dt, lo, hi,
                 float(g["armijo_beta"][0]), float(g["armijo_cost0"][0]), nodes, "finaltime", V,
                 nonlinear_solver=hp.solve_nonlinear_equation, dof_neighbors=mesh.dof_neighbors())
    v1, c_inc, k = res
    assert k == int(g["armijo_its"][0]) and np.array_equal(c_inc, g["armijo_c"]) and rel_l2(v1, g["armijo_u"]) < 1e-11


def test_reference_loops_chemotaxis(loops):
    """solve_chtxs_system / solve_adjoint_chtxs_system (helpers.py:1250-1581; all-time and final-time) against outputs of
    the reference's own functions"""
    g = loops
    n, ns, dt = int(g["n"][0]), int(g["ns"][0]), float(g["chtxs_dt"][0])
    mesh, V = _space(n)
    nodes = V.dim()
    mk = np.zeros((ns + 1) * nodes); fk = np.zeros_like(mk)
    mk[:nodes], fk[:nodes] = g["chtxs_m0"], g["chtxs_f0"]
    _quiet(hp.solve_chtxs_system, g["chtxs_c"], mk, fk, V, nodes, ns, dt, mesh.dof_neighbors())
    assert rel_l2(mk, g["chtxs_m"]) < 1e-12 and rel_l2(fk, g["chtxs_f"]) < 1e-12
    for optim, mh, fh, tag in (("alltime", g["chtxs_mhat"], g["chtxs_fhat"], "at"),
                               ("finaltime", g["chtxs_mhat"][ns * nodes:], g["chtxs_fhat"][ns * nodes:], "ft")):
        pk = np.zeros_like(mk); qk = np.zeros_like(mk)
        _quiet(hp.solve_adjoint_chtxs_system, g["chtxs_m"].copy(), g["chtxs_f"].copy(), mh, fh, pk, qk, g["chtxs_c"], ns * dt, V,
               nodes, ns, dt, mesh.dof_neighbors(), optim)
        assert rel_l2(pk, g[f"chtxs_p_{tag}"]) < 1e-11 and rel_l2(qk, g[f"chtxs_q_{tag}"]) < 1e-11, optim


def test_refactored_nonlinear_pgd_script_loop():
    """the projected-gradient loop of nonlinear_FCT_PDECO_refactored.py:105-210 written against the drop-in
    `hp` (same calls, same order, same fail / restart bookkeeping), against the run of the reference script's OWN source
    lines with the reference's helpers.py on oracle/fake_dolfin.py (tests/golden/ref_pgd_nonlinear.npz): costs within 1e-9,
    state / adjoint / control / direction within 1e-11, the same number of line-search trials."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "ref_pgd_nonlinear.npz")))
    n, num_steps, dt = int(g["n"][0]), int(g["ns"][0]), float(g["dt"][0])
    mesh, V = _space(n)
    nodes = V.dim()
    vertex_to_dof = vertex_to_dof_map(V)
    dof_neighbors = hp.find_node_neighbours(mesh, nodes, vertex_to_dof)
    from fem_fct_pdeco_b200.forms import TestFunction, TrialFunction, dx
    M = hp.assemble_sparse(TrialFunction(V) * TestFunction(V) * dx)
    a1, a2, T = 0, 1, num_steps * dt
    beta, c_lower, c_upper, optim, tol, max_iter_armijo, max_iter_GD = 1e-1, -1, 1, "finaltime", 1e-4, 5, 2
    u0 = hp.nonlinear_equation_IC(a1, a2, 1.0 / n, nodes, vertex_to_dof)
    assert np.array_equal(u0, g["u0"])
    uhat_T = g["uhat_T"]

    def script():
        vec_length = (num_steps + 1) * nodes
        ck = np.zeros(vec_length)
        uk = np.zeros(vec_length)
        uk[:nodes] = u0
        uk, _ = hp.solve_nonlinear_equation(ck, uk, None, V, nodes, num_steps, dt, dof_neighbors)
        pk = np.zeros(vec_length)
        pk = hp.solve_adjoint_nonlinear_equation(uk, uhat_T, pk, T, V, nodes, num_steps, dt, dof_neighbors)
        cost_fun_old = hp.cost_functional(uk, uhat_T, ck, num_steps, dt, M, beta, optim)
        cost_fun_new = (2 + tol) * cost_fun_old
        stop_crit = hp.rel_err(cost_fun_new, cost_fun_old)
        dk = np.zeros(vec_length)
        it = fail_count = fail_restart_count = 0
        fail_pass = False
        cost_fun_vals, armijo_its = [cost_fun_old], []
        while (stop_crit >= tol or fail_pass) and it < max_iter_GD:
            dk = -(beta * ck - pk)
            uk, ck, iters = hp.armijo_line_search_ref(
                uk, ck, dk, uhat_T, num_steps, dt, c_lower, c_upper, beta, cost_fun_old, nodes, optim, V,
                dof_neighbors=dof_neighbors, nonlinear_solver=hp.solve_nonlinear_equation, max_iter=max_iter_armijo)
            pk = hp.solve_adjoint_nonlinear_equation(uk, uhat_T, pk, T, V, nodes, num_steps, dt, dof_neighbors)
            if iters == max_iter_armijo:
                fail_count += 1
                fail_pass = True
                if fail_count == 3:
                    break
            elif fail_count > 0:
                fail_count = 0
                fail_restart_count += 1
                fail_pass = False
                if fail_restart_count == 5:
                    break
            cost_fun_new = hp.cost_functional(uk, uhat_T, ck, num_steps, dt, M, beta, optim)
            stop_crit = hp.rel_err(cost_fun_new, cost_fun_old)
            cost_fun_vals.append(cost_fun_new)
            armijo_its.append(iters)
            it += 1
            cost_fun_old = cost_fun_new
        return uk, pk, ck, dk, cost_fun_vals, armijo_its, it, stop_crit

    uk, pk, ck, dk, cost_fun_vals, armijo_its, it, stop_crit = _quiet(script)
    assert it == int(g["it"][0]) and armijo_its == list(g["armijo_its"])
    assert np.allclose(cost_fun_vals, g["cost"], rtol=1e-9, atol=0)
    assert abs(stop_crit - float(g["stop_crit"][0])) <= 1e-6 * abs(float(g["stop_crit"][0]))
    for name, a in (("u", uk), ("p", pk), ("c", ck), ("d", dk)):
        assert rel_l2(a, g[name]) < 1e-11, name


def _two_species_pgd(solve, solve_adj, direction, cost_kw, armijo_kw, u0, v0, V, nodes, num_steps, dt, dof_neighbors, M, beta,
                     c_lower, c_upper, optim, tol, max_iter_armijo, max_iter_GD, targets, at_least_two):
    """the projected-gradient loop shared by Schnak_FCT_PDECO_refactored.py:122-246 and
    chemotaxis_FCT_PDECO_AT_refactored.py:122-257 (same calls, same order, same fail / restart bookkeeping)"""
    uhat, vhat = targets
    vec_length = (num_steps + 1) * nodes
    ck = np.zeros(vec_length)
    uk = np.zeros(vec_length); vk = np.zeros(vec_length)
    uk[:nodes] = u0; vk[:nodes] = v0
    uk, vk = solve(ck, uk, vk, V, nodes, num_steps, dt, dof_neighbors)
    pk = np.zeros(vec_length); qk = np.zeros(vec_length)
    pk, qk = solve_adj(uk, vk, pk, qk, ck)
    cost_fun_old = hp.cost_functional(uk, uhat, ck, num_steps, dt, M, beta, optim, var2=vk, var2_target=vhat)
    cost_fun_new = (2 + tol) * cost_fun_old
    stop_crit = hp.rel_err(cost_fun_new, cost_fun_old)
    it = fail_count = fail_restart_count = 0
    fail_pass = False
    cost_fun_vals, armijo_its = [cost_fun_old], []
    while (stop_crit >= tol or fail_pass or (at_least_two and it < 2)) and it < max_iter_GD:
        dk = direction(ck, uk, pk, qk)
        uk, vk, ck, iters = hp.armijo_line_search_ref(uk, ck, dk, uhat, num_steps, dt, c_lower, c_upper, beta, cost_fun_old, nodes,
                                                      optim, V, dof_neighbors=dof_neighbors, var2=vk, var2_target=vhat,
                                                      nonlinear_solver=solve, max_iter=max_iter_armijo, **armijo_kw)
        pk, qk = solve_adj(uk, vk, pk, qk, ck)
        if iters == max_iter_armijo:
            fail_count += 1
            fail_pass = True
            if fail_count == cost_kw["fail_max"]:
                break
        elif fail_count > 0:
            fail_count = 0
            fail_restart_count += 1
            fail_pass = False
            if fail_restart_count == 5:
                break
        cost_fun_new = hp.cost_functional(uk, uhat, ck, num_steps, dt, M, beta, optim, var2=vk, var2_target=vhat)
        stop_crit = hp.rel_err(cost_fun_new, cost_fun_old)
        cost_fun_vals.append(cost_fun_new)
        armijo_its.append(iters)
        it += 1
        cost_fun_old = cost_fun_new
    return uk, vk, pk, qk, ck, cost_fun_vals, armijo_its, it


def test_refactored_two_species_pgd_script_loops():
    """The projected-gradient loops of Schnak_FCT_PDECO_refactored.py (final time, dk = -(beta ck - gamma/r pk), eight
    backtracking trials per iteration) and chemotaxis_FCT_PDECO_AT_refactored.py (all time, dk = -(beta ck - qk uk / r), line
    search with gam = 1e-5, s0 = 2) on the drop-in `hp`, against the runs of the scripts' OWN source lines with the reference's
    helpers.py on oracle/fake_dolfin.py (tests/golden/ref_pgd_two_species.npz)."""
    import os
    from conftest import GOLDEN
    from fem_fct_pdeco_b200.forms import TestFunction, TrialFunction, dx
    g = dict(np.load(os.path.join(GOLDEN, "ref_pgd_two_species.npz")))
    n = int(g["n"][0])
    mesh, V = _space(n)
    nodes = V.dim()
    vertex_to_dof = vertex_to_dof_map(V)
    dof_neighbors = hp.find_node_neighbours(mesh, nodes, vertex_to_dof)
    M = hp.assemble_sparse(TrialFunction(V) * TestFunction(V) * dx)

    # ---- Schnakenberg, final time (Schnak_FCT_PDECO_refactored.py) ----
    num_steps, dt = int(g["s_ns"][0]), float(g["s_dt"][0])
    T = num_steps * dt
    gamma = hp.get_schnak_sys_params()[4]
    uhat_T, vhat_T = g["s_uhat"], g["s_vhat"]
    u0, v0 = hp.schnak_sys_IC(0, 1, 1.0 / n, nodes, vertex_to_dof)
    assert np.array_equal(u0, g["s_u0"]) and np.array_equal(v0, g["s_v0"])
    res = _quiet(_two_species_pgd, hp.solve_schnak_system,
                 lambda uk, vk, pk, qk, ck: hp.solve_adjoint_schnak_system(uk, vk, uhat_T, vhat_T, pk, qk, T, V, nodes, num_steps,
                                                                           dt, dof_neighbors),
                 lambda ck, uk, pk, qk: -(1e-1 * ck - gamma / 1 * pk), {"fail_max": 3}, {}, u0, v0, V, nodes, num_steps, dt,
                 dof_neighbors, M, 1e-1, 0, 10, "finaltime", 1e-3, 10, 2, (uhat_T, vhat_T), False)
    uk, vk, pk, qk, ck, cost, its, it = res
    assert it == int(g["s_it"][0]) and its == list(g["s_its"])
    assert np.allclose(cost, g["s_cost"], rtol=1e-9, atol=0)
    for name, a in (("u", uk), ("v", vk), ("p", pk), ("q", qk), ("c", ck)):
        assert rel_l2(a, g["s_" + name]) < 1e-10, name

    # ---- chemotaxis, all time (chemotaxis_FCT_PDECO_AT_refactored.py) ----
    num_steps, dt = int(g["c_ns"][0]), float(g["c_dt"][0])
    T = num_steps * dt
    rescaling = 1 / 10
    uhat, vhat = g["c_uhat"], g["c_vhat"]
    u0, v0 = hp.chtxs_sys_IC(0, 1, 1.0 / n, nodes, vertex_to_dof)
    assert np.array_equal(u0, g["c_u0"]) and np.array_equal(v0, g["c_v0"])
    res = _quiet(_two_species_pgd, hp.solve_chtxs_system,
                 lambda uk, vk, pk, qk, ck: hp.solve_adjoint_chtxs_system(uk, vk, uhat, vhat, pk, qk, ck, T, V, nodes, num_steps, dt,
                                                                          dof_neighbors, "alltime", mesh=mesh, deltax=1.0 / n,
                                                                          vertex_to_dof=vertex_to_dof, rescaling=rescaling),
                 lambda ck, uk, pk, qk: -(1e-3 * ck - qk * uk / rescaling), {"fail_max": 5}, {"gam": 1e-5, "s0": 2}, u0, v0, V,
                 nodes, num_steps, dt, dof_neighbors, M, 1e-3, 0, 20, "alltime", 1e-4, 20, 2, (uhat, vhat), True)
    uk, vk, pk, qk, ck, cost, its, it = res
    assert it == int(g["c_it"][0]) and its == list(g["c_its"])
    assert np.allclose(cost, g["c_cost"], rtol=1e-9, atol=0)
    for name, a in (("u", uk), ("v", vk), ("p", pk), ("q", qk), ("c", ck)):
        assert rel_l2(a, g["c_" + name]) < 1e-10, name
