"""Oracle: CPU restatement of the reference's refactored state / adjoint time loops and line search.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parity status: these loops need dolfin in the reference, so they
cannot be run here; no shipped data exercises them except the chemotaxis *state* loop (pinned in
tests/test_oracle_golden.py).  For the rest the oracle is "restatement of the loop o oracle fct_step (itself pinned
on the reference's own FCT_alg_ref)": PARITY UNPINNED for the loop bodies.

    SchnakProblem       helpers.py:443-698   (solve_schnak_system, solve_adjoint_schnak_system)
    NonlinearProblem    helpers.py:835-1038  (solve_nonlinear_equation, solve_adjoint_nonlinear_equation)
    ChemotaxisAdjoint   helpers.py:1387-1581 (solve_adjoint_chtxs_system)
    armijo_ref          helpers.py:1583-1713 (armijo_line_search_ref, nonlinear_solver branch)
"""
import numpy as np
from scipy.sparse.linalg import spsolve

from .fct_numpy import cost_functional, fct_step, l2_norm_sq_q
from .p1mesh import reorder_vector_to_dof
from .pdeco_numpy import ChemotaxisProblem, _Base


def schnak_wind(x, y):
    """helpers.py:506-508"""
    return (y - 0.5) * x * (1 - x), -(x - 0.5) * y * (1 - y)


def nonlinear_wind(x, y):
    """helpers.py:876-878 (speed = 1)"""
    return 2 * (y - 0.5) * x * (1 - x), -2 * (x - 0.5) * y * (1 - y)


class SchnakProblem(_Base):
    Du, Dv, c_a, c_b, gamma, omega1, omega2 = 1 / 100, 8.6676, 0.1, 0.9, 230.82, 100, 0.6

    def initial_condition(self):
        """helpers.py:443-483"""
        X, Y = self.grid()
        con = 0.1
        u = self.c_a + self.c_b + con * np.cos(2 * np.pi * (X + Y)) + 0.01 * (sum(np.cos(2 * np.pi * X * i) for i in range(1, 9)))
        v = self.c_b / pow(self.c_a + self.c_b, 2) + con * np.cos(2 * np.pi * (X + Y)) + 0.01 * (sum(np.cos(2 * np.pi * X * i) for i in range(1, 9)))
        r = lambda a: reorder_vector_to_dof(a.reshape(self.nodes), 1, self.nodes, self.mesh.vertex_to_dof)
        return r(u), r(v)

    def state(self, control, u0, v0, num_steps, dt, rescaling=1.0):
        """helpers.py:511-597 (control reused from step 1: App. D-1)"""
        asm, pat = self.asm, self.pat
        A = asm.conv_conservative(schnak_wind, degree=5)
        c1 = np.asarray(control).reshape(num_steps + 1, self.nodes)[1]
        u = np.zeros((num_steps + 1, self.nodes)); v = np.zeros_like(u)
        u[0], v[0] = u0, v0
        Mc = pat.csr(self.M)
        for i in range(1, num_steps + 1):
            un, vn = u[i - 1], v[i - 1]
            Mat1 = self.Du * self.K - self.omega1 * A
            rhs1 = self.gamma / rescaling * asm.load_p1_product(c1) + self.gamma * asm.load_p1_product(un, un, vn)
            u[i] = fct_step(pat, Mat1, rhs1, un, dt, self.M, self.ML, S=self.gamma * self.M)
            Mu2 = asm.mass_p1_product(u[i], u[i])
            rhs2 = asm.load_constant(self.gamma * self.c_b)
            Mat2 = self.M + dt * (self.Dv * self.K - self.omega2 * A + self.gamma * Mu2)
            v[i] = spsolve(pat.csr(Mat2).tocsc(), Mc @ vn + dt * rhs2)
        return u, v

    def adjoint(self, u, v, uhat_T, vhat_T, num_steps, dt):
        """helpers.py:599-698"""
        asm, pat = self.asm, self.pat
        A = asm.conv_nonconservative(schnak_wind, degree=5)
        p = np.zeros((num_steps + 1, self.nodes)); q = np.zeros_like(p)
        p[num_steps] = uhat_T - u[num_steps]
        q[num_steps] = vhat_T - v[num_steps]
        Mc = pat.csr(self.M)
        for i in reversed(range(num_steps)):
            un, vn = u[i], v[i]
            Mu2 = asm.mass_p1_product(un, un)
            rhs_q = self.gamma * asm.load_p1_product(p[i + 1], un, un)
            Mat_q = self.M + dt * (self.Dv * self.K - self.omega2 * A + self.gamma * Mu2)
            q[i] = spsolve(pat.csr(Mat_q).tocsc(), Mc @ q[i + 1] + dt * rhs_q)
            Mat_p = self.Du * self.K - self.omega1 * A
            Muv = asm.mass_p1_product(un, vn)
            rhs_p = -2 * self.gamma * asm.load_p1_product(un, vn, q[i])
            Mat_rhs = self.gamma * self.M - 2 * self.gamma * Muv
            p[i] = fct_step(pat, Mat_p, rhs_p, p[i + 1], dt, self.M, self.ML, S=Mat_rhs)
        return p, q


class NonlinearProblem(_Base):
    eps = 1e-4

    def initial_condition(self):
        """helpers.py:835-865"""
        X, Y = self.grid()
        ic = 5 * Y * (Y - 1) * X * (X - 1) * np.sin(4 * X * np.pi)
        return reorder_vector_to_dof(ic.reshape(self.nodes), 1, self.nodes, self.mesh.vertex_to_dof)

    def state(self, control, u0, num_steps, dt):
        """helpers.py:881-966 (control reused from step 1: App. D-1)"""
        asm, pat = self.asm, self.pat
        A = asm.conv_conservative(nonlinear_wind, degree=5)
        Mat1 = A - self.eps * self.K
        c1 = np.asarray(control).reshape(num_steps + 1, self.nodes)[1]
        rhs = asm.load_p1_product(c1)
        u = np.zeros((num_steps + 1, self.nodes))
        u[0] = u0
        for i in range(1, num_steps + 1):
            un = u[i - 1]
            Mat_rhs = -self.M + 1 / 3 * asm.mass_p1_product(un, un)
            u[i] = fct_step(pat, -Mat1, rhs, un, dt, self.M, self.ML, S=Mat_rhs)
        return u

    def adjoint(self, u, uhat_T, num_steps, dt):
        """helpers.py:968-1038"""
        asm, pat = self.asm, self.pat
        A = asm.conv_conservative(nonlinear_wind, degree=5)
        Mat_p = -A - self.eps * self.K
        p = np.zeros((num_steps + 1, self.nodes))
        p[num_steps] = uhat_T - u[num_steps]
        for i in reversed(range(num_steps)):
            Mat_rhs = asm.mass_p1_product(u[i], u[i]) - self.M
            p[i] = fct_step(pat, -Mat_p, np.zeros(self.nodes), p[i + 1], dt, self.M, self.ML, S=Mat_rhs)
        return p


class ChemotaxisAdjoint(ChemotaxisProblem):
    def adjoint(self, u, v, uhat, vhat, control, num_steps, dt, optim, rescaling=0.1):
        """helpers.py:1387-1581.  uhat/vhat: final-time targets (finaltime) or whole trajectories (alltime).
        Note the reference adds the *nodal* differences uhat-u, vhat-v to the assembled right-hand sides
        (helpers.py:1509, 1535): reproduced."""
        if optim not in ("alltime", "finaltime"):
            raise ValueError(f"Invalid value for 'optim': '{optim}'. Must be one of ['alltime', 'finaltime'].")
        asm, pat = self.asm, self.pat
        control = np.asarray(control).reshape(num_steps + 1, self.nodes)
        p = np.zeros((num_steps + 1, self.nodes)); q = np.zeros_like(p)
        if optim == "finaltime":
            p[num_steps] = uhat - u[num_steps]
            q[num_steps] = vhat - v[num_steps]
        Mc = pat.csr(self.M)
        Mat_q = pat.csr(self.M + dt * (self.Df * self.K + self.delta * self.M)).tocsc()
        for i in reversed(range(num_steps)):
            un, vn = u[i], v[i]
            Aa = asm.chemotaxis_adjoint_mat(un, vn, self.eta, degree=5)
            Mat_p = self.Dm * self.K - self.chi * Aa
            rhs_p = asm.load_p1_product(control[i], q[i + 1]) / rescaling
            if optim == "alltime":
                rhs_p = rhs_p + (uhat[i] - u[i])
            p[i] = fct_step(pat, Mat_p, rhs_p, p[i + 1], dt, self.M, self.ML)
            rhs_q = asm.load_grad_pair(lambda phi, xy: self.chi * asm.at_quad(un, phi) * np.exp(-self.eta * asm.at_quad(un, phi)),
                                       p[i], 4)
            if optim == "alltime":
                rhs_q = rhs_q + (vhat[i] - v[i])
            q[i] = spsolve(Mat_q, Mc @ q[i + 1] + dt * rhs_q)
        return p, q


def armijo_ref(prob, solver, var1, c, d, var1_target, num_steps, dt, c_lower, c_upper, beta, costfun_init, optim,
               gam=1e-4, max_iter=10, s0=1, var2=None, var2_target=None):
    """helpers.py:1583-1713, nonlinear_solver branch.  `solver(c_inc) -> (var1, var2)` (flattened trajectories).
    Returns (var1, var2, c_inc, k+1)."""
    if optim not in ("alltime", "finaltime"):
        raise ValueError(f"Invalid value for 'optim': '{optim}'. Must be one of ['alltime', 'finaltime'].")
    pat, M = prob.pat, prob.M
    s = s0
    k = 0
    c_inc = c
    for k in range(max_iter):
        c_inc = np.clip(c + s * d, c_lower, c_upper)
        var1, var2 = solver(c_inc)
        cost2 = cost_functional(pat, var1, var1_target, c_inc, num_steps, dt, M, beta, optim, var2=var2,
                                var2_target=var2_target)
        armijo = cost2 - costfun_init
        dif = l2_norm_sq_q(pat, c_inc - c, num_steps, dt, M)
        if armijo <= -gam / s * dif:
            break
        s /= 2
    return var1, var2, c_inc, k + 1
