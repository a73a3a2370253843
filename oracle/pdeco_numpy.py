"""Oracle: CPU restatement of the reference's time loops / PDECO drivers around the FCT step.

TEST INFRASTRUCTURE (see oracle/__init__.py).

* SolidBodyProblem      advection_solidbody_FCT.py:25-148 (config 1; produced data/solidbody_t*_u.csv)
* ChemotaxisProblem     helpers.py:1197-1385 solve_chtxs_system (produced Chtxs_data_dx0.025_dt0.001/)
* AdvectionDriftPDECO   advection_solidbody_FCT_PDECO_alltime.py:43-303 + old_helpers.py:1-85
                        (config 2; template of the 4096^2 benchmark config 5) and
                        advection_solidbodyGaussian_FCT.py (its target generator)
"""
import numpy as np
from scipy.sparse.linalg import spsolve

from .fct_numpy import (Pattern, chebsi, cost_functional, fct_step, fct_step_legacy, l2_norm_sq_q)
from .p1assembly import P1Assembler
from .p1mesh import RectMesh, reorder_vector_to_dof


class _Base:
    def __init__(self, n, a1, a2):
        self.mesh = RectMesh(n, a1, a2)
        self.asm = P1Assembler(self.mesh)
        self.pat = Pattern(*self.mesh.pattern())
        self.nodes = self.mesh.nodes
        self.M = self.asm.mass()
        self.ML = self.asm.lumped(self.M)
        self.K = self.asm.stiffness()

    def grid(self):
        """np.arange grids exactly as the scripts build them (rounding matters for thresholded ICs,
        SURVEY.md App. D-3): advection_solidbody_FCT.py:56-58, helpers.py:466-468."""
        m = self.mesh
        dx = (m.a2 - m.a1) / m.n
        X = np.arange(m.a1, m.a2 + dx, dx)[: m.n + 1]
        return np.meshgrid(X, X)


class SolidBodyProblem(_Base):
    """Zalesak slotted disc under rotation 1/om*(-y,x) + drift 2*(1,1), forward FCT only."""

    def __init__(self, n=80, a1=-1.0, a2=1.0, slit_width=0.1, om=np.pi / 40, drift=2.0, eps=0.0):
        super().__init__(n, a1, a2)
        xy = self.mesh.dof_xy
        A = self.asm.conv_conservative_p1(-xy[:, 1] / om + drift, xy[:, 0] / om + drift)
        self.A_u = A - eps * self.K            # advection_solidbody_FCT.py:106-109
        self.slit = slit_width

    def initial_condition(self):
        X, Y = self.grid()
        R = np.sqrt(X ** 2 + (Y - 1 / 3) ** 2)
        out = ((R < 1 / 3) & ((np.abs(X) > self.slit) | (Y > 0.5))).astype(np.float64)
        return reorder_vector_to_dof(out.reshape(self.nodes), 1, self.nodes, self.mesh.vertex_to_dof)

    def forward(self, num_steps, dt, u0=None, solver="spsolve", keep=False):
        u = self.initial_condition() if u0 is None else u0.copy()
        traj = [u.copy()]
        for _ in range(num_steps):
            u = fct_step_legacy(self.pat, self.A_u, None, u, dt, self.M, self.ML, solver=solver)
            if keep:
                traj.append(u.copy())
        return np.array(traj) if keep else u


class ChemotaxisProblem(_Base):
    """helpers.py:1197-1385."""
    delta, Dm, Df, chi, gamma, eta = 100.0, 0.05, 0.05, 0.25, 100.0, 0.5

    def initial_condition(self):
        """helpers.py:1213-1248 (np.random.seed(5))."""
        N = self.mesh.n + 1
        np.random.seed(5)
        u_init = 1.5 + 0.1 * (0.5 - np.random.rand(N, N))
        u0 = reorder_vector_to_dof(u_init.reshape(self.nodes), 1, self.nodes, self.mesh.vertex_to_dof)
        return u0, u0.copy()

    def forward(self, control, m0, f0, num_steps, dt, control_const=None, rescaling=0.1):
        """Returns trajectories m[num_steps+1, nodes], f[...].  The refactored solver reuses the
        step-1 control for all steps (helpers.py:1332-1333): reproduced."""
        asm, pat = self.asm, self.pat
        Mat2 = pat.csr(self.M + dt * (self.Df * self.K + self.delta * self.M)).tocsc()
        m = np.zeros((num_steps + 1, self.nodes)); f = np.zeros_like(m)
        m[0], f[0] = m0, f0
        c_fun = None
        for i in range(1, num_steps + 1):
            if control_const is None and c_fun is None:
                c_fun = np.asarray(control).reshape(num_steps + 1, self.nodes)[i]
            if control_const is not None:
                src = control_const * asm.load_p1_product(m[i - 1])
            else:
                src = asm.load_p1_product(c_fun, m[i - 1])
            rhs2 = asm.load_p1_product(f[i - 1]) + dt * src / rescaling
            f[i] = spsolve(Mat2, rhs2)
            mn = m[i - 1]
            Aa = asm.chemotaxis_conv(f[i], lambda phi, xy: np.exp(-self.eta * asm.at_quad(mn, phi)), degree=4)
            A = self.Dm * self.K - self.chi * Aa
            m[i] = fct_step(pat, A, None, mn, dt, self.M, self.ML)
        return m, f


class AdvectionDriftPDECO(_Base):
    """All-time tracking PDECO with a scalar drift-speed control c(x,t) multiplying b = (1,1):
        du/dt + div(u c b) = 0  (eps = 0, rotation off).
    State/adjoint/gradient loops: advection_solidbody_FCT_PDECO_alltime.py:210-275;
    Armijo: old_helpers.py:1-85."""

    def __init__(self, n, a1=-1.0, a2=1.0, beta=0.01, c_lower=0.0, c_upper=5.0, bx=1.0, by=1.0, eps=0.0,
                 solver="spsolve"):
        super().__init__(n, a1, a2)
        self.beta, self.c_lower, self.c_upper = beta, c_lower, c_upper
        self.bx, self.by, self.eps = bx, by, eps
        self.solver = solver
        self.Md = self.M[self.pat.diagpos]

    def gaussian_ic(self):
        """advection_solidbody_FCT_PDECO_alltime.py:92-112 rescaled to the mesh's box:
        exp(-20((x-x1)^2 + 5 (y-y1)^2)) with (x1,y1) = (-2/3,-5/6) on [-1,1]^2."""
        m = self.mesh
        X, Y = self.grid()
        s = 2.0 / (m.a2 - m.a1)                      # map to [-1,1]
        Xr = (X - m.a1) * s - 1.0
        Yr = (Y - m.a1) * s - 1.0
        out = np.exp(-20 * ((Xr + 2 / 3) ** 2 + 5 * (Yr + 5 / 6) ** 2))
        return reorder_vector_to_dof(out.reshape(self.nodes), 1, self.nodes, m.vertex_to_dof)

    def operator(self, c):
        """A_u = -eps*Ad + Adrift1 + Adrift2 (legacy sign; :222-226)."""
        return self.asm.drift_mass(c, self.bx, self.by) + self.asm.drift_conv(c, self.bx, self.by) - self.eps * self.K

    def state(self, c_traj, u0, num_steps, dt, info=None):
        """:210-228.  c_traj[num_steps+1, nodes]; returns u[num_steps+1, nodes]."""
        c_traj = np.asarray(c_traj).reshape(num_steps + 1, self.nodes)
        u = np.zeros((num_steps + 1, self.nodes))
        u[0] = u0
        for i in range(1, num_steps + 1):
            A_u = self.operator(c_traj[i])
            u[i] = fct_step_legacy(self.pat, A_u, None, u[i - 1], dt, self.M, self.ML, solver=self.solver, info=info)
        return u

    def target(self, u0, num_steps, dt, c_const=2.0):
        """advection_solidbodyGaussian_FCT.py:63-143: forward FCT with constant wind c*(1,1)."""
        return self.state(np.full((num_steps + 1, self.nodes), c_const), u0, num_steps, dt)

    def adjoint(self, c_traj, u, uhat, num_steps, dt):
        """:235-259.  p(T) = 0; p_rhs = assemble((uhat_n - u_n) v dx) = M (uhat_n - u_n)."""
        c_traj = np.asarray(c_traj).reshape(num_steps + 1, self.nodes)
        p = np.zeros((num_steps + 1, self.nodes))
        Mc = self.pat.csr(self.M)
        for i in reversed(range(0, num_steps)):
            A_p = -self.operator(c_traj[i])
            p_rhs = Mc @ (uhat[i] - u[i])
            p[i] = fct_step_legacy(self.pat, A_p, p_rhs, p[i + 1], dt, self.M, self.ML, solver=self.solver)
        return p

    def gradient(self, c_traj, u, p, num_steps):
        """:265-275.  d_k = ChebSI(-(beta M c + assemble(p (b.grad u) v dx)))."""
        c_traj = np.asarray(c_traj).reshape(num_steps + 1, self.nodes)
        Mc = self.pat.csr(self.M)
        d = np.zeros((num_steps + 1, self.nodes))
        for i in range(num_steps + 1):
            rhs = -(self.beta * (Mc @ c_traj[i]) + self.asm.load_drift_grad(p[i], u[i], self.bx, self.by))
            d[i] = chebsi(self.pat, rhs, self.M, self.Md, 20, 0.5, 2)
        return d

    def cost(self, u, uhat, c, num_steps, dt):
        """cost_functional(..., optim='alltime') (helpers.py:383-441) on flattened trajectories."""
        return cost_functional(self.pat, np.ravel(u), np.ravel(uhat), np.ravel(c), num_steps, dt, self.M,
                               self.beta, "alltime")

    def armijo(self, u0, c, d, uhat, num_steps, dt, gam=1e-4, max_iter=5, s0=1.0):
        """old_helpers.py:1-85 (armijo_line_search_sbr_drift).  `cost_functional_proj` has no surviving
        definition in the reference: re-specified as cost_functional(u, uhat, c) on the
        already-projected control -- parity unpinned for this function."""
        c = np.asarray(c).reshape(num_steps + 1, self.nodes)
        d = np.asarray(d).reshape(num_steps + 1, self.nodes)
        s = 1.0
        k = 0
        proj = np.clip(c + s * d, self.c_lower, self.c_upper) - c
        grad_l2 = l2_norm_sq_q(self.pat, proj.ravel(), num_steps, dt, self.M)
        u_cur = self.state(c, u0, num_steps, dt)
        cost0 = self.cost(u_cur, uhat, c, num_steps, dt)
        armijo = 1e5
        u_inc = u_cur
        while armijo > -gam / s * grad_l2 and k < max_iter:
            s = s0 * (1 / 2 ** k)
            c_inc = np.clip(c + s * d, self.c_lower, self.c_upper)
            u_inc = self.state(c_inc, u0, num_steps, dt)
            cost2 = self.cost(u_inc, uhat, c_inc, num_steps, dt)
            armijo = cost2 - cost0
            grad_l2 = l2_norm_sq_q(self.pat, (c_inc - c).ravel(), num_steps, dt, self.M)
            k += 1
        return s, u_inc, k
