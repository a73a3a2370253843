"""Manufactured-solution driver advection_FCT_PDECO_alltime_exact.py on the drop-in names.

The script's problem -- du/dt - eps lap(u) + w.grad(u) = c + g, adjoint -dp/dt - eps lap(p) - w.grad(p) = uhat - u, exact
u, p, c = proj(p/beta) (advection_FCT_PDECO_alltime_exact.py:73-143) -- gives three tests:
  * convergence: with the exact control the state loop (:236-249) reproduces u_ex, with the exact state the adjoint loop
    (:255-271) reproduces p_ex, and halving dx (dt = dx^2) cuts the L2(Q) errors by more than 1.8 (second order in dx for the
    smooth solution: the limiter is inactive);
  * the same two sweeps against the numpy oracle (1e-11);
  * three projected-gradient iterations of the script's main loop (:220-324) with the re-specified legacy armijo_line_search
    and cost_functional_proj: the cost drops by orders of magnitude and the control moves towards the exact one in every
    iteration."""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest

from conftest import rel_l2
from fem_fct_pdeco_b200 import helpers as hp
from fem_fct_pdeco_b200.forms import Expression, TestFunction, TrialFunction, assemble, dot, dx, grad, vec_to_function
from fem_fct_pdeco_b200.mesh import FunctionSpaceP1, RectMeshP1
from oracle.fct_numpy import Pattern, fct_step_legacy
from oracle.p1assembly import P1Assembler
from oracle.p1mesh import RectMesh

pytestmark = pytest.mark.gpu

A1, A2, BETA, C_LO, C_UP, E1, E2, K1, K2, EPS = 0, 1, 0.001, 0, 0.5, 0.2, 0.3, 1, 1, 0.001
PI = np.pi


class Exact:
    """the script's closed forms (advection_FCT_PDECO_alltime_exact.py:73-143), T is the final time"""

    def __init__(self, T):
        self.T = T

    def u(self, t, X, Y):
        return np.exp(E1 * t) * (np.sin(K1 * PI * X) * np.sin(K1 * PI * Y)) ** 2

    def p(self, t, X, Y):
        return (np.exp(E2 * self.T) - np.exp(E2 * t)) * (np.sin(K2 * PI * X) * np.sin(K2 * PI * Y)) ** 2

    def c(self, t, X, Y):
        return np.clip(1 / BETA * self.p(t, X, Y), C_LO, C_UP)

    @staticmethod
    def wind(X, Y):
        return 2 * (Y - 0.5) * X * (1 - X), -2 * (X - 0.5) * Y * (1 - Y)

    def g(self, t, X, Y):
        wx, wy = self.wind(X, Y)
        e = np.exp(E1 * t)
        dudx = 2 * K1 * PI * e * np.sin(K1 * PI * X) * np.cos(K1 * PI * X) * np.sin(K1 * PI * Y) ** 2
        dudy = 2 * K1 * PI * e * np.sin(K1 * PI * X) ** 2 * np.sin(K1 * PI * Y) * np.cos(K1 * PI * Y)
        uxx = 2 * (PI * K1) ** 2 * e * np.cos(2 * K1 * PI * X) * np.sin(K1 * PI * Y) ** 2
        uyy = 2 * (PI * K1) ** 2 * e * np.sin(K1 * PI * X) ** 2 * np.cos(2 * K1 * PI * Y)
        return E1 * self.u(t, X, Y) - EPS * (uxx + uyy) + wx * dudx + wy * dudy - self.c(t, X, Y)

    def uhat(self, t, X, Y):
        wx, wy = self.wind(X, Y)
        f = np.exp(E2 * self.T) - np.exp(E2 * t)
        s2 = (np.sin(K2 * PI * X) * np.sin(K2 * PI * Y)) ** 2
        dpdt = -E2 * np.exp(E2 * t) * s2
        dpdx = 2 * K2 * PI * f * np.sin(K2 * PI * X) * np.cos(K2 * PI * X) * np.sin(K2 * PI * Y) ** 2
        dpdy = 2 * K2 * PI * f * np.sin(K2 * PI * X) ** 2 * np.sin(K2 * PI * Y) * np.cos(K2 * PI * Y)
        pxx = 2 * (PI * K2) ** 2 * f * np.cos(2 * K2 * PI * X) * np.sin(K2 * PI * Y) ** 2
        pyy = 2 * (PI * K2) ** 2 * f * np.sin(K2 * PI * X) ** 2 * np.cos(2 * K2 * PI * Y)
        return -dpdt - EPS * (pxx + pyy) - wx * dpdx - wy * dpdy + self.u(t, X, Y)


class Setup:
    """everything the script builds before its main loop (:60-205), with the drop-in names"""

    def __init__(self, n, T):
        self.n, self.T = n, T
        self.deltax = (A2 - A1) / n
        self.dt = self.deltax ** 2
        self.num_steps = round(T / self.dt)
        self.mesh = RectMeshP1(n, A1, A2)
        self.V = FunctionSpaceP1(self.mesh)
        self.nodes = self.V.dim()
        u, v = TrialFunction(self.V), TestFunction(self.V)
        self.v = v
        g1 = np.linspace(A1, A2, n + 1)
        X, Y = np.meshgrid(g1, g1)
        self.v2d = self.mesh.vertex_to_dof
        self.dof_neighbors = hp.find_node_neighbours(self.mesh, self.nodes, self.v2d)
        wind = Expression(('2*(x[1]-0.5)*x[0]*(1-x[0])', '-2*(x[0]-0.5)*x[1]*(1-x[1])'), degree=4)
        self.M = hp.assemble_sparse_lil(u * v * dx)
        self.M_Lump = hp.row_lump(self.M, self.nodes)
        Ad = hp.assemble_sparse(dot(grad(u), grad(v)) * dx)
        A = hp.assemble_sparse(dot(wind, grad(v)) * u * dx)
        self.A_u = A - EPS * Ad
        self.A_p = -A - EPS * Ad
        ex = Exact(T)
        ns, nodes = self.num_steps, self.nodes

        def traj(fn):
            flat = np.concatenate([fn(i * self.dt, X, Y).reshape(nodes) for i in range(ns + 1)])
            return hp.reorder_vector_to_dof_time(flat, ns + 1, nodes, self.v2d)
        self.g, self.uhat = traj(ex.g), traj(ex.uhat)
        self.u_ex, self.p_ex, self.c_ex = traj(ex.u), traj(ex.p), traj(ex.c)

    def state(self, ck, step=None):
        """:236-249"""
        step = step or (lambda A, r, un: hp.FCT_alg(A, r, un, self.dt, self.nodes, self.M, self.M_Lump, self.dof_neighbors))
        nodes, ns = self.nodes, self.num_steps
        uk = np.zeros((ns + 1) * nodes)
        uk[:nodes] = self.u_ex[:nodes]
        for i in range(1, ns + 1):
            start, end = i * nodes, (i + 1) * nodes
            rhs = self.load(self.g[start:end] + ck[start:end])
            uk[start:end] = step(self.A_u, rhs, uk[start - nodes:start])
        return uk

    def adjoint(self, uk, step=None):
        """:255-271"""
        step = step or (lambda A, r, un: hp.FCT_alg(A, r, un, self.dt, self.nodes, self.M, self.M_Lump, self.dof_neighbors))
        nodes, ns = self.nodes, self.num_steps
        pk = np.zeros((ns + 1) * nodes)
        for i in reversed(range(0, ns)):
            start, end = i * nodes, (i + 1) * nodes
            rhs = self.load(self.uhat[start:end] - uk[start:end])
            pk[start:end] = step(self.A_p, rhs, pk[end:end + nodes])
        return pk

    def load(self, nodal):
        """assemble(f v dx) for a P1 coefficient (the script's vec_to_function + assemble)"""
        return np.asarray(assemble(vec_to_function(nodal, self.V) * self.v * dx))

    def err(self, a, b):
        return np.sqrt(hp.L2_norm_sq_Q(a - b, self.num_steps, self.dt, self.M) / hp.L2_norm_sq_Q(b, self.num_steps, self.dt, self.M))


def _quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_manufactured_solution_convergence_and_oracle():
    T = 0.2
    errs = {}
    for n in (10, 20):
        s = _quiet(Setup, n, T)
        uk = _quiet(s.state, s.c_ex)
        pk = _quiet(s.adjoint, s.u_ex)
        errs[n] = (_quiet(s.err, uk, s.u_ex), _quiet(s.err, pk, s.p_ex))
        if n == 10:
            # the same sweeps through the numpy oracle (legacy sign convention, direct low-order solve)
            om = RectMesh(n, A1, A2)
            asm = P1Assembler(om)
            pat = Pattern(*om.pattern())
            Mv, Kv = asm.mass(), asm.stiffness()
            Av = asm.conv_conservative(lambda x, y: Exact.wind(x, y), degree=5)
            ML = asm.lumped(Mv)
            o_load = lambda f: asm.load_p1_product(f)      # noqa: E731
            s_o = s
            real_load = s.load
            try:
                s.load = o_load
                uo = s_o.state(s.c_ex, step=lambda A, r, un, A_=Av - EPS * Kv: fct_step_legacy(pat, A_, r, un, s.dt, Mv, ML))
                po = s_o.adjoint(s.u_ex, step=lambda A, r, un, A_=-Av - EPS * Kv: fct_step_legacy(pat, A_, r, un, s.dt, Mv, ML))
            finally:
                s.load = real_load
            assert rel_l2(uk, uo) < 1e-11 and rel_l2(pk, po) < 1e-11
    assert errs[10][0] < 0.05 and errs[10][1] < 0.05
    assert errs[20][0] < errs[10][0] / 1.8 and errs[20][1] < errs[10][1] / 1.8


def test_manufactured_solution_projected_gradient_iterations():
    s = _quiet(Setup, 10, 0.2)
    nodes, ns, dt = s.nodes, s.num_steps, s.dt
    L = (ns + 1) * nodes
    zeros_nt = np.zeros(L)
    ck = np.zeros(L)
    wk = np.zeros(L)
    wk[:nodes] = s.u_ex[:nodes]
    uk0 = np.zeros(L); uk0[:nodes] = s.u_ex[:nodes]
    steps = []
    cost = [10 * _quiet(hp.cost_functional_proj, uk0, zeros_nt, ck, zeros_nt, 0, s.uhat, ns, dt, s.M, C_LO, C_UP, BETA)]
    dist = [_quiet(s.err, ck, s.c_ex)]
    for it in range(3):                                                       # :220-324
        uk = _quiet(s.state, ck)
        pk = _quiet(s.adjoint, uk)
        dk = -(BETA * ck - pk)
        wk[nodes:] = 0.0
        for i in range(1, ns + 1):                                            # move in u, :279-293
            start, end = i * nodes, (i + 1) * nodes
            w_rhs = s.load(dk[start:end])
            wk[start:end] = _quiet(hp.FCT_alg, s.A_u, w_rhs, wk[start - nodes:start], dt, nodes, s.M, s.M_Lump, s.dof_neighbors)
        sk = _quiet(hp.armijo_line_search, uk, pk, wk, ck, dk, s.uhat, ns, dt, s.M, C_LO, C_UP, BETA)
        assert 0 < sk <= 1
        steps.append(sk)
        ckp1 = np.clip(ck + sk * dk, C_LO, C_UP)
        cost.append(_quiet(hp.cost_functional_proj, uk, wk, ckp1, dk, sk, s.uhat, ns, dt, s.M, C_LO, C_UP, BETA))
        ck = ckp1
        dist.append(_quiet(s.err, ck, s.c_ex))
    # The script evaluates its cost on u_k + s w_k, the state of the UNPROJECTED control c_k + s d_k, so the printed cost is
    # not the cost of the projected iterate and need not be monotone once the bounds are active (they are: p/beta >> c_upper);
    # what the iteration must do is leave the initial cost far behind and move the control towards the exact one every time.
    assert cost[1] < 0.05 * cost[0] and all(c < 0.05 * cost[0] for c in cost[1:]), (cost, steps, dist)
    assert all(b < a for a, b in zip(dist, dist[1:])), (cost, steps, dist)


def test_final_time_line_search_call_shape():
    """advection_FCT_PDECO_finaltime_exact.py:230, 296-297, 374-385: the final-time generation of the legacy calls --
    cost_functional_proj_FT(u, 0, c, 0, 0, uhat_T, zeros, ...), p(T) = uhat_T - u(T) with a homogeneous adjoint, and
    armijo_line_search(u, p, w, c, d, uhat_T, ..., cost_fun_k, optim='finaltime') returning (s, u + s w)."""
    s = _quiet(Setup, 10, 0.1)
    nodes, ns, dt = s.nodes, s.num_steps, s.dt
    L = (ns + 1) * nodes
    zeros_nt = np.zeros(L)
    uhat_T = s.u_ex[ns * nodes:]
    ck = np.zeros(L)
    uk = _quiet(s.state, ck)
    cost_fun_k = _quiet(hp.cost_functional_proj_FT, uk, zeros_nt, ck, zeros_nt, 0, uhat_T, np.zeros(nodes), ns, dt, s.M, C_LO, C_UP, BETA)
    assert cost_fun_k == _quiet(hp.cost_functional, uk, uhat_T, ck, ns, dt, s.M, BETA, optim="finaltime")
    pk = np.zeros(L)
    pk[ns * nodes:] = uhat_T - uk[ns * nodes:]
    for i in reversed(range(0, ns)):
        start, end = i * nodes, (i + 1) * nodes
        pk[start:end] = _quiet(hp.FCT_alg, s.A_p, np.zeros(nodes), pk[end:end + nodes], dt, nodes, s.M, s.M_Lump, s.dof_neighbors)
    dk = -(BETA * ck - pk)
    wk = np.zeros(L)
    for i in range(1, ns + 1):
        start, end = i * nodes, (i + 1) * nodes
        wk[start:end] = _quiet(hp.FCT_alg, s.A_u, s.load(dk[start:end]), wk[start - nodes:start], dt, nodes, s.M, s.M_Lump,
                               s.dof_neighbors)
    sk, u_inc = _quiet(hp.armijo_line_search, uk, pk, wk, ck, dk, uhat_T, ns, dt, s.M, C_LO, C_UP, BETA, cost_fun_k,
                       optim='finaltime')
    assert 0 < sk <= 1 and np.array_equal(u_inc, uk + sk * wk)
    ckp1 = np.clip(ck + sk * dk, C_LO, C_UP)
    cost_fun_kp1 = _quiet(hp.cost_functional, u_inc, uhat_T, ckp1, ns, dt, s.M, BETA, optim='finaltime')
    dif = _quiet(hp.L2_norm_sq_Q, ckp1 - ck, ns, dt, s.M)
    assert sk == 2.0 ** -9 or cost_fun_kp1 - cost_fun_k <= -1e-4 / sk * dif        # Armijo condition, or max_iter reached
    with pytest.raises(ValueError):
        hp.armijo_line_search(uk, pk, wk, ck, dk, uhat_T, ns, dt, s.M, C_LO, C_UP, BETA, cost_fun_k, optim='sometime')
