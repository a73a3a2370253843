"""GPU tests of BASELINE configs 3 and 4 at script level: the main loops of chemotaxis_mimura_FCT_PGD.py:157-260 and
Schnak_FCT_PDECO.py:190-300 written against the drop-in names (`from helpers import *`, `import mimura_data_helpers`), compared
with the trajectories the reference scripts' OWN loop source produces (tests/golden/ref_cfg3.npz, ref_cfg4.npz: executed by
tests/golden/make_golden.py with the reference's helpers.py / old_helpers.py / mimura_data_helpers.py on oracle/fake_dolfin.py),
then the scripts' line-search and cost calls, whose definitions are lost in the reference (re-specified, parity unpinned)."""
import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest
from scipy.sparse.linalg import spsolve

from conftest import GOLDEN, rel_l2
from fem_fct_pdeco_b200 import helpers as hp
from fem_fct_pdeco_b200 import mimura_data_helpers
from fem_fct_pdeco_b200.forms import (Expression, TestFunction, TrialFunction, VectorFunctionSpace, div, dot, dx, grad, project,
                                      vec_to_function)
from fem_fct_pdeco_b200.mesh import FunctionSpaceP1, RectMeshP1

pytestmark = pytest.mark.gpu


def _quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize("fixture", ["ref_cfg4.npz", "ref_cfg4_K.npz"])      # 11^2 DoF; K = the script's own 51^2-DoF mesh
def test_script_schnak_fct_pdeco_cfg4(fixture):
    g = dict(np.load(os.path.join(GOLDEN, fixture)))
    n, num_steps, dt = int(g["n"][0]), int(g["ns"][0]), float(g["dt"][0])
    a1, a2 = 0, 1
    beta, c_lower, c_upper = 0.1, -1, 1
    Du, Dv, c_a, c_b, gamma, omega1, omega2 = 1 / 100, 8.6676, 0.1, 0.9, 230.82, 100, 0.6
    T = num_steps * dt
    wind = Expression(('-(x[1]-0.5)*sin(2*pi*t)', '(x[0]-0.5)*sin(2*pi*t)'), degree=4, pi=np.pi, t=0)
    mesh = RectMeshP1(n, a1, a2)
    V = FunctionSpaceP1(mesh)
    nodes = V.dim()
    u, w = TrialFunction(V), TestFunction(V)
    W = VectorFunctionSpace(mesh, "CG", 1)
    dof_neighbors = hp.find_node_neighbours(mesh, nodes, None)
    M = hp.assemble_sparse_lil(u * w * dx)
    M_Lump = hp.row_lump(M, nodes)
    Ad = hp.assemble_sparse(dot(grad(u), grad(w)) * dx)
    vec_length = (num_steps + 1) * nodes
    uk = np.zeros(vec_length); vk = np.zeros(vec_length)
    uk[:nodes], vk[:nodes] = g["u0"], g["v0"]
    ck = g["c"].copy()
    uhat_T, vhat_T = g["uhat_T"], g["vhat_T"]

    def state_and_adjoint():
        t = 0
        uk[nodes:] = np.zeros(num_steps * nodes)
        vk[nodes:] = np.zeros(num_steps * nodes)
        for i in range(1, num_steps + 1):                                   # Schnak_FCT_PDECO.py:194-227
            start, end = i * nodes, (i + 1) * nodes
            t += dt
            wind.t = t
            u_n, v_n = uk[start - nodes:start], vk[start - nodes:start]
            u_n_fun, v_n_fun, c_np1_fun = (vec_to_function(x, V) for x in (u_n, v_n, ck[start:end]))
            A = hp.assemble_sparse(dot(wind, grad(w)) * u * dx)
            mat_u = -(Du * Ad - omega1 * A)
            rhs_u = np.asarray(hp.assemble((gamma * (c_np1_fun + u_n_fun ** 2 * v_n_fun)) * w * dx))
            uk[start:end] = hp.FCT_alg(mat_u, rhs_u, u_n, dt, nodes, M, M_Lump, dof_neighbors, source_mat=gamma * M)
            u_np1_fun = vec_to_function(uk[start:end], V)
            M_u2 = hp.assemble_sparse(u_np1_fun * u_np1_fun * u * w * dx)
            rhs_v = np.asarray(hp.assemble((gamma * c_b) * w * dx))
            vk[start:end] = spsolve(M + dt * (Dv * Ad - omega2 * A + gamma * M_u2), M @ v_n + dt * rhs_v)
        qk = np.zeros(vec_length); pk = np.zeros(vec_length)
        qk[num_steps * nodes:] = vhat_T - vk[num_steps * nodes:]
        pk[num_steps * nodes:] = uhat_T - uk[num_steps * nodes:]
        t = T
        for i in reversed(range(0, num_steps)):                              # :239-279
            start, end = i * nodes, (i + 1) * nodes
            t -= dt
            q_np1, p_np1 = qk[end:end + nodes], pk[end:end + nodes]
            p_np1_fun = vec_to_function(p_np1, V)
            u_n_fun, v_n_fun = vec_to_function(uk[start:end], V), vec_to_function(vk[start:end], V)
            wind.t = t
            wind_fun = project(wind, W)
            A = hp.assemble_sparse(div(wind_fun * u) * w * dx)
            M_u2 = hp.assemble_sparse(u_n_fun * u_n_fun * u * w * dx)
            rhs_q = np.asarray(hp.assemble(gamma * p_np1_fun * u_n_fun ** 2 * w * dx))
            qk[start:end] = spsolve(M + dt * (Dv * Ad - omega2 * A + gamma * M_u2), M @ q_np1 + dt * rhs_q)
            q_n_fun = vec_to_function(qk[start:end], V)
            mat_p = -Du * Ad + omega1 * A
            M_uv = hp.assemble_sparse(u_n_fun * v_n_fun * u * w * dx)
            rhs_p = np.asarray(hp.assemble(- 2 * gamma * u_n_fun * v_n_fun * q_n_fun * w * dx))
            pk[start:end] = hp.FCT_alg(mat_p, rhs_p, p_np1, dt, nodes, M, M_Lump, dof_neighbors,
                                       source_mat=gamma * M - 2 * gamma * M_uv)
        return pk, qk

    pk, qk = _quiet(state_and_adjoint)
    assert rel_l2(uk, g["u"]) < 1e-11 and rel_l2(vk, g["v"]) < 1e-11
    assert rel_l2(pk, g["p"]) < 1e-11 and rel_l2(qk, g["q"]) < 1e-11
    # :285-306: descent direction, line search (lost definition, re-specified), projection, cost
    cost_fun_k = _quiet(hp.cost_functional, uk, uhat_T, ck, num_steps, dt, M, beta, var2=vk, var2_target=vhat_T, optim='finaltime')
    dk = -(beta * ck - gamma * pk)
    sk, u_inc, v_inc = _quiet(hp.armijo_line_search, uk, ck, dk, uhat_T, num_steps, dt, M, c_lower, c_upper, beta, cost_fun_k,
                              nodes, V=V, optim='finaltime', dof_neighbors=dof_neighbors, example='Schnak', var2=vk,
                              var2_target=vhat_T)
    assert 0 < sk <= 1 and u_inc.shape == uk.shape and v_inc.shape == vk.shape
    ckp1 = np.clip(ck + sk * dk, c_lower, c_upper)
    cost_fun_kp1 = _quiet(hp.cost_functional, u_inc, uhat_T, ckp1, num_steps, dt, M, beta, optim='finaltime', var2=v_inc,
                          var2_target=vhat_T)
    dif = _quiet(hp.L2_norm_sq_Q, ckp1 - ck, num_steps, dt, M)
    assert sk == 2.0 ** -9 or cost_fun_kp1 - cost_fun_k <= -1e-4 / sk * dif        # Armijo condition, or max_iter reached
    with pytest.raises(ValueError):
        hp.armijo_line_search(uk, ck, dk, uhat_T, num_steps, dt, M, c_lower, c_upper, beta, cost_fun_k, nodes, V=V,
                              optim='finaltime', example='no such example')


@pytest.mark.parametrize("fixture", ["ref_cfg3.npz", "ref_cfg3_C.npz"])      # 9^2 DoF; C = the script's own 129^2 mesh, dt
def test_script_chemotaxis_mimura_pgd_cfg3(fixture):
    g = dict(np.load(os.path.join(GOLDEN, fixture)))
    n, num_steps, dt = int(g["n"][0]), int(g["ns"][0]), float(g["dt"][0])
    a1, a2 = g["box"]
    delta, Dm, Df, chi = g["params"]
    sampled = "sample" in g
    if sampled:
        # the 129^2 fixture holds every sample-th DoF of each time level and the norms of the full fields; its inputs are
        # regenerated here exactly as tests/golden/make_golden.py:ref_script_cfg3 draws them
        nodes_ = (n + 1) ** 2
        rng = np.random.default_rng(int(g["seed"][0]))
        m_ = RectMeshP1(n, a1, a2)
        g["m0"] = hp.reorder_vector_to_dof(mimura_data_helpers.m_initial_condition(a1, a2, (a2 - a1) / n).reshape(nodes_), 1, nodes_,
                                           m_.vertex_to_dof)
        assert np.array_equal(g["m0"][::int(g["sample"][0])], g["m0_s"])
        g["f0"] = 1 / 32 * np.ones(nodes_)
        g["c"] = 0.5 + rng.random((num_steps + 1) * nodes_)
        g["mhat_T"] = g["m0"] * (1 + 0.05 * rng.random(nodes_))
        g["fhat_T"] = g["f0"] * (1 + 0.05 * rng.random(nodes_))
    beta, c_lower, c_upper = 1, 0, 1.5
    T = num_steps * dt
    mesh = RectMeshP1(n, a1, a2)
    V = FunctionSpaceP1(mesh)
    nodes = V.dim()
    u, v = TrialFunction(V), TestFunction(V)
    dof_neighbors = hp.find_node_neighbours(mesh, nodes, None)
    M = hp.assemble_sparse_lil(u * v * dx)
    M_Lump = hp.row_lump(M, nodes)
    Ad = hp.assemble_sparse(dot(grad(u), grad(v)) * dx)
    Mat_fq = M + dt * (Df * Ad + delta * M)
    vec_length = (num_steps + 1) * nodes
    mk = np.zeros(vec_length); fk = np.zeros(vec_length)
    mk[:nodes], fk[:nodes] = g["m0"], g["f0"]
    ck = g["c"].copy()
    mhat_T, fhat_T = g["mhat_T"], g["fhat_T"]

    def state_and_adjoint():
        fk[nodes:] = np.zeros(num_steps * nodes)
        mk[nodes:] = np.zeros(num_steps * nodes)
        for i in range(1, num_steps + 1):                                   # chemotaxis_mimura_FCT_PGD.py:162-186
            start, end = i * nodes, (i + 1) * nodes
            m_n = mk[start - nodes:start]
            m_n_fun = vec_to_function(m_n, V)
            c_np1_fun = vec_to_function(ck[start:end], V)
            f_n_fun = vec_to_function(fk[start - nodes:start], V)
            f_rhs = hp.rhs_chtx_f(f_n_fun, m_n_fun, c_np1_fun, dt, v)
            fk[start:end] = spsolve(Mat_fq, f_rhs)
            f_np1_fun = vec_to_function(fk[start:end], V)
            A_m = mimura_data_helpers.mat_chtx_m(f_np1_fun, m_n_fun, Dm, chi, u, v)
            m_rhs = mimura_data_helpers.rhs_chtx_m(m_n_fun, v)
            mk[start:end] = hp.FCT_alg(A_m, m_rhs, m_n, dt, nodes, M, M_Lump, dof_neighbors)
        qk = np.zeros(vec_length); pk = np.zeros(vec_length)
        qk[num_steps * nodes:] = fhat_T - fk[num_steps * nodes:]
        pk[num_steps * nodes:] = mhat_T - mk[num_steps * nodes:]
        for i in reversed(range(0, num_steps)):                              # :201-225
            start, end = i * nodes, (i + 1) * nodes
            q_np1, p_np1 = qk[end:end + nodes], pk[end:end + nodes]
            p_np1_fun, q_np1_fun = vec_to_function(p_np1, V), vec_to_function(q_np1, V)
            m_n_fun, f_n_fun, c_n_fun = (vec_to_function(x[start:end], V) for x in (mk, fk, ck))
            q_rhs = hp.rhs_chtx_q(q_np1_fun, m_n_fun, p_np1_fun, chi, dt, v)
            qk[start:end] = spsolve(Mat_fq, q_rhs)
            q_n_fun = vec_to_function(qk[start:end], V)
            A_p = mimura_data_helpers.mat_chtx_p(f_n_fun, m_n_fun, Dm, chi, u, v)
            p_rhs = hp.rhs_chtx_p(c_n_fun, q_n_fun, v)
            pk[start:end] = hp.FCT_alg(A_p, p_rhs, p_np1, dt, nodes, M, M_Lump, dof_neighbors)
        return pk, qk

    pk, qk = _quiet(state_and_adjoint)
    if sampled:
        st = int(g["sample"][0])
        for name, a in (("m", mk), ("f", fk), ("p", pk), ("q", qk)):
            a = a.reshape(num_steps + 1, nodes)
            assert rel_l2(a[:, ::st], g[name + "_s"]) < 1e-11, name
            assert np.allclose(np.linalg.norm(a, axis=1), g[name + "_norm"], rtol=1e-12, atol=0), name
    else:
        assert rel_l2(mk, g["m"]) < 1e-11 and rel_l2(fk, g["f"]) < 1e-11
        assert rel_l2(pk, g["p"]) < 1e-11 and rel_l2(qk, g["q"]) < 1e-11
    # :231-256: descent direction, the lost line search and cost (re-specified), projection, stopping criteria
    dk = -(beta * ck - qk * mk)
    cost_fun_k = _quiet(hp.cost_functional_proj_FT, mk, fk, ck, dk, 0, mhat_T, fhat_T, num_steps, dt, M, c_lower, c_upper, beta)
    ref_cost = _quiet(hp.cost_functional, mk, mhat_T, np.clip(ck, c_lower, c_upper), num_steps, dt, M, beta, "finaltime", var2=fk,
                      var2_target=fhat_T)
    assert cost_fun_k == ref_cost
    m_before = mk.copy()
    sk = _quiet(hp.armijo_line_search_chtxs, mk, fk, qk, ck, dk, mhat_T, fhat_T, Mat_fq, chi, Dm, Df, num_steps, dt, nodes, M,
                M_Lump, Ad, c_lower, c_upper, beta, V, dof_neighbors)
    assert 0 < sk <= 1 and not np.array_equal(mk, m_before)                # the accepted trial's state is left in mk, fk
    ckp1 = np.clip(ck + sk * dk, c_lower, c_upper)
    cost_fun_kp1 = _quiet(hp.cost_functional_proj_FT, mk, fk, ckp1, dk, sk, mhat_T, fhat_T, num_steps, dt, M, c_lower, c_upper, beta)
    dif = _quiet(hp.L2_norm_sq_Q, ckp1 - ck, num_steps, dt, M)
    assert sk == 2.0 ** -4 or cost_fun_kp1 - cost_fun_k <= -1e-4 / sk * dif
    assert np.isfinite(np.abs(cost_fun_k - cost_fun_kp1) / np.abs(cost_fun_k))
