"""Reference-named state / adjoint time loops and the Armijo line search, running on the B200 kernels.

Same names, argument order, in-place semantics (the caller's trajectory buffers are written AND returned) and
error behaviour as the reference (KarolinaBenkova/FEM-FCT-PDECO):

    solve_schnak_system                helpers.py:511-597      solve_adjoint_schnak_system        :599-698
    solve_nonlinear_equation           helpers.py:881-966      solve_adjoint_nonlinear_equation   :968-1038
    solve_chtxs_system                 helpers.py:1250-1385    solve_adjoint_chtxs_system         :1387-1581
    armijo_line_search_ref             helpers.py:1583-1713    get_*_params / *_IC                :443-509, 835-879, 1197-1248

`V` is the dolfin-free stand-in `FunctionSpaceP1` (fem-fct-pdeco_b200/mesh.py); `control_fun` may be a number
(the reference's `Constant`) or a DoF vector (the reference's `Function`).  The six time loops run inside libfctpdeco
(fct_forward_* / fct_adjoint_*, csrc/fct_drivers.cu) on device trajectories: these functions upload the caller's arrays once,
call the driver and download the result into the caller's buffers.  Reference quirks that
parity depends on are reproduced and marked.
"""
import ctypes as C
import warnings

import numpy as np

from . import _lib
from .helpers import L2_norm_sq_Q, cost_functional, reorder_vector_to_dof

# monomial order 1,x,y,x^2,xy,y^2,x^3,x^2y,xy^2,y^3  (include/fctpdeco.h, FCT_FORM_WIND_POLY3)
_SCHNAK_WIND = np.array([0, -0.5, 0, 0.5, 1, 0, 0, -1, 0, 0,          # (y-1/2) x (1-x)     helpers.py:506
                         0, 0, 0.5, 0, -1, -0.5, 0, 0, 1, 0], float)   # -(x-1/2) y (1-y)    helpers.py:507


def _grid(V):
    m = V.mesh()
    dx = (m.a2 - m.a1) / m.n
    X = np.arange(m.a1, m.a2 + dx, dx)[: m.n + 1]
    return np.meshgrid(X, X)


# ---- parameter getters / initial conditions ----------------------------------------------------------
def get_schnak_sys_params():
    """helpers.py:485-509 (the wind is returned as FCT_FORM_WIND_POLY3 coefficients instead of a df.Expression)"""
    return 1 / 100, 8.6676, 0.1, 0.9, 230.82, 100, 0.6, _SCHNAK_WIND.copy()


def get_nonlinear_eqns_params():
    """helpers.py:867-879"""
    speed = 1
    return 1e-4, speed, 2 * speed * _SCHNAK_WIND


def get_chtxs_sys_params():
    """helpers.py:1197-1211"""
    return 100, 0.05, 0.05, 0.25, 100, 0.5


def schnak_sys_IC(a1, a2, deltax, nodes, vertex_to_dof):
    """helpers.py:443-483"""
    X = np.arange(a1, a2 + deltax, deltax)
    X, Y = np.meshgrid(X, X)
    _, _, c_a, c_b, _, _, _, _ = get_schnak_sys_params()
    con = 0.1
    u_init = c_a + c_b + con * np.cos(2 * np.pi * (X + Y)) + 0.01 * (sum(np.cos(2 * np.pi * X * i) for i in range(1, 9)))
    v_init = c_b / pow(c_a + c_b, 2) + con * np.cos(2 * np.pi * (X + Y)) + 0.01 * (sum(np.cos(2 * np.pi * X * i) for i in range(1, 9)))
    return (reorder_vector_to_dof(u_init.reshape(nodes), 1, nodes, vertex_to_dof),
            reorder_vector_to_dof(v_init.reshape(nodes), 1, nodes, vertex_to_dof))


def nonlinear_equation_IC(a1, a2, deltax, nodes, vertex_to_dof):
    """helpers.py:835-865"""
    X = np.arange(a1, a2 + deltax, deltax)
    X, Y = np.meshgrid(X, X)
    ic = 5 * Y * (Y - 1) * X * (X - 1) * np.sin(4 * X * np.pi)
    return reorder_vector_to_dof(ic.reshape(nodes), 1, nodes, vertex_to_dof)


def chtxs_sys_IC(a1, a2, deltax, nodes, vertex_to_dof):
    """helpers.py:1213-1248 (np.random.seed(5); u(0) = v(0))"""
    sqnodes = round(np.sqrt(nodes))
    np.random.seed(5)
    u_init = 1.5 + 0.1 * (0.5 - np.random.rand(sqnodes, sqnodes))
    u_init_dof = reorder_vector_to_dof(u_init.reshape(nodes), 1, nodes, vertex_to_dof)
    return u_init_dof, u_init_dof


# ---- device-side helpers ------------------------------------------------------------------------------
class _Dev:
    """context + scratch buffers of a function space (one instance per context: buffers are reused between calls)"""

    def __new__(cls, V, nodes):
        ctx = V.mesh().context()
        if ctx.n != nodes:
            raise ValueError(f"nodes={nodes} does not match the function space ({ctx.n})")
        self = getattr(ctx, "_solver_dev", None)
        if self is None:
            self = object.__new__(cls)
            self.ctx = ctx
            self.M, self.ML, self.Md, self.K = ctx.static()
            self._vec, self._mat, self._traj = {}, {}, {}
            ctx._solver_dev = self
        return self

    def traj(self, name, num_steps):
        """device trajectory of (num_steps + 1) time levels; slice(i) views one level"""
        size = (num_steps + 1) * self.ctx.n
        t = self._traj.get(name)
        if t is None or t.size != size:
            if t is not None:
                t.free()
            t = self._traj[name] = self.ctx.empty(size)
        return t

    def level(self, traj, i):
        return traj.slice(i * self.ctx.n, self.ctx.n)

    def vec(self, name):
        if name not in self._vec:
            self._vec[name] = self.ctx.empty(self.ctx.n)
        return self._vec[name]

    def mat(self, name):
        if name not in self._mat:
            self._mat[name] = self.ctx.empty(self.ctx.nnz)
        return self._mat[name]

    def load_control(self, out, control_fun, other=None, scale=1.0, accumulate=False):
        """out (+)= scale * assemble(control_fun * [other] * v * dx); control_fun: number or device vector"""
        ctx, L = self.ctx, _lib
        if np.isscalar(control_fun):
            if other is None:
                ctx.assemble_vector(L.LOAD_CONST, out, s0=float(control_fun), scale=scale, accumulate=accumulate)
            else:
                ctx.assemble_vector(L.LOAD_P1_1, out, c0=other, scale=scale * float(control_fun), accumulate=accumulate)
        elif other is None:
            ctx.assemble_vector(L.LOAD_P1_1, out, c0=control_fun, scale=scale, accumulate=accumulate)
        else:
            ctx.assemble_vector(L.LOAD_P1_2, out, c0=control_fun, c1=other, scale=scale, accumulate=accumulate)


def _hostp(a):
    """(keep-alive array, pointer) of a small host parameter array"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.c_void_p)


def _ctl(cfun):
    """(device pointer or None, constant) of a control that is a number or a device vector"""
    if cfun is None:
        return None, 0.0
    return (None, float(cfun)) if np.isscalar(cfun) else (cfun.ptr, 0.0)


def _run(what, fn, *args):
    try:
        _lib.check(fn(*args))
    except _lib.FctError as e:
        raise _lib.FctError(f"{what}: {e}") from None


def _control_slice(dev, control, control_fun, start, end):
    """the reference builds control_fun once, from the FIRST step's slice, and then reuses it
    (helpers.py:577-578, 950-951, 1332-1333) -- reproduced by the callers"""
    if control_fun is not None:
        return control_fun if np.isscalar(control_fun) else dev.ctx.array(np.asarray(control_fun, dtype=np.float64))
    return dev.ctx.array(np.asarray(control[start:end], dtype=np.float64))


def _solve(ctx, kind, mat, b, x, what):
    """second-species system (the reference: spsolve).  1e-13 on the relative residual: the recursive residual of the Krylov
    loops bottoms out around 1e-13..1e-14 for these M + dt(...) systems, and libfctpdeco accepts a stagnated iterate at that
    level; a genuine failure names the system"""
    if kind == _lib.SOLVER_PCG and ctx.n >= 50000:
        # large SPD systems: Chebyshev-polynomial preconditioned CG (7x fewer reductions at 1025^2, profiles/r2_pcg_table.txt)
        kind = _lib.SOLVER_CHEB_PCG
    try:
        its, res = ctx.solve(kind, mat, b, x, rtol=1e-13, maxit=20000)
    except _lib.FctError as e:
        raise _lib.FctError(f"linear solve for '{what}' failed: {e}") from None
    return its


# ---- Schnakenberg ---------------------------------------------------------------------------------------
def solve_schnak_system(control, var1, var2, V, nodes, num_steps, dt, dof_neighbors, control_fun=None, rescaling=1):
    """helpers.py:511-597"""
    dev = _Dev(V, nodes)
    var1[nodes:] = np.zeros(num_steps * nodes)
    var2[nodes:] = np.zeros(num_steps * nodes)
    d1, d2 = dev.traj("var1", num_steps), dev.traj("var2", num_steps)
    dev.level(d1, 0).upload(var1[:nodes]); dev.level(d2, 0).upload(var2[:nodes])
    cfun = _control_slice(dev, control, control_fun, nodes, 2 * nodes) if num_steps >= 1 else None
    _forward_schnak_dev(dev, cfun, d1, d2, num_steps, dt, rescaling)
    if num_steps >= 1:
        d1.slice(nodes, num_steps * nodes).download(var1[nodes:]); d2.slice(nodes, num_steps * nodes).download(var2[nodes:])
    return var1, var2


def _forward_schnak_dev(dev, cfun, d1, d2, num_steps, dt, rescaling=1):
    """the loop of helpers.py:560-597 on device trajectories (level 0 = initial condition); cfun: number or device vector.
    Runs in libfctpdeco (fct_forward_schnak, csrc/fct_drivers.cu): assemblies, FCT step and BiCGStab solve of every time
    level are enqueued from C."""
    Du, Dv, _, c_b, gamma, omega1, omega2, wind = get_schnak_sys_params()
    print("Solving the system of advective Schnakenberg state equations...")
    cp, cc = _ctl(cfun)
    par, ppar = _hostp([Du, Dv, c_b, gamma, omega1, omega2])
    w, pw = _hostp(wind)
    _run("solve_schnak_system", _lib.lib.fct_forward_schnak, dev.ctx.handle, cp, cc, d1.ptr, d2.ptr, int(num_steps), float(dt),
         ppar, pw, float(rescaling), None)


def solve_adjoint_schnak_system(uk, vk, uhat_T, vhat_T, pk, qk, T, V, nodes, num_steps, dt, dof_neighbors):
    """helpers.py:599-698"""
    Du, Dv, _, c_b, gamma, omega1, omega2, wind = get_schnak_sys_params()
    dev = _Dev(V, nodes)
    pk[num_steps * nodes:] = uhat_T - uk[num_steps * nodes:]
    qk[num_steps * nodes:] = vhat_T - vk[num_steps * nodes:]
    du, dv = dev.traj("var1", num_steps).upload(uk), dev.traj("var2", num_steps).upload(vk)
    dp, dq = dev.traj("adj1", num_steps), dev.traj("adj2", num_steps)
    dev.level(dp, num_steps).upload(pk[num_steps * nodes:]); dev.level(dq, num_steps).upload(qk[num_steps * nodes:])
    print("\nSolving adjoint equation...")
    par, ppar = _hostp([Du, Dv, c_b, gamma, omega1, omega2])
    w, pw = _hostp(wind)
    _run("solve_adjoint_schnak_system", _lib.lib.fct_adjoint_schnak, dev.ctx.handle, du.ptr, dv.ptr, dp.ptr, dq.ptr,
         int(num_steps), float(dt), ppar, pw, None)
    if num_steps >= 1:
        dp.slice(0, num_steps * nodes).download(pk[:num_steps * nodes]); dq.slice(0, num_steps * nodes).download(qk[:num_steps * nodes])
    return pk, qk


# ---- nonlinear advection-reaction ---------------------------------------------------------------------
def solve_nonlinear_equation(control, var1, var2, V, nodes, num_steps, dt, dof_neighbors, control_fun=None,
                             show_plots=False, vertex_to_dof=None):
    """helpers.py:881-966"""
    if var2 is not None:
        warnings.warn("Warning: 'var2' is not None. Ensure this is intentional.")
    dev = _Dev(V, nodes)
    var1[nodes:] = np.zeros(num_steps * nodes)
    d1 = dev.traj("var1", num_steps)
    dev.level(d1, 0).upload(var1[:nodes])
    cfun = _control_slice(dev, control, control_fun, nodes, 2 * nodes) if num_steps >= 1 else None
    _forward_nonlinear_dev(dev, cfun, d1, None, num_steps, dt)
    if num_steps >= 1:
        d1.slice(nodes, num_steps * nodes).download(var1[nodes:])
    return var1, None


def _forward_nonlinear_dev(dev, cfun, d1, d2, num_steps, dt, rescaling=None):
    """the loop of helpers.py:935-966 on a device trajectory (level 0 = initial condition); fct_forward_nonlinear"""
    eps, _, wind = get_nonlinear_eqns_params()
    print("\nSolving nonlinear state equation...")
    cp, cc = _ctl(cfun)
    w, pw = _hostp(wind)
    _run("solve_nonlinear_equation", _lib.lib.fct_forward_nonlinear, dev.ctx.handle, cp, cc, d1.ptr, int(num_steps), float(dt),
         float(eps), pw, None)


def solve_adjoint_nonlinear_equation(uk, uhat_T, pk, T, V, nodes, num_steps, dt, dof_neighbors):
    """helpers.py:968-1038"""
    eps, _, wind = get_nonlinear_eqns_params()
    dev = _Dev(V, nodes)
    pk[num_steps * nodes:] = uhat_T - uk[num_steps * nodes:]
    du, dp = dev.traj("var1", num_steps).upload(uk), dev.traj("adj1", num_steps)
    dev.level(dp, num_steps).upload(pk[num_steps * nodes:])
    print("\nSolving adjoint equation...")
    w, pw = _hostp(wind)
    _run("solve_adjoint_nonlinear_equation", _lib.lib.fct_adjoint_nonlinear, dev.ctx.handle, du.ptr, dp.ptr, int(num_steps),
         float(dt), float(eps), pw, None)
    if num_steps >= 1:
        dp.slice(0, num_steps * nodes).download(pk[:num_steps * nodes])
    return pk


def solve_chtxs_system(control, var1, var2, V, nodes, num_steps, dt, dof_neighbors, control_fun=None,
                       show_plots=False, vertex_to_dof=None, generation_mode=False, output_dir=None, rescaling=1 / 10):
    """helpers.py:1250-1385"""
    dev = _Dev(V, nodes)
    if generation_mode:
        if len(var1) != nodes or len(var2) != nodes or len(control) != nodes:
            raise ValueError(f"Generation mode, the input vectors should be of length {nodes}")
        # no trajectory is kept: two time levels ping-pong (helpers.py:1318-1323)
        d1, d2 = dev.traj("gen1", 1), dev.traj("gen2", 1)
        dev.level(d1, 0).upload(var1); dev.level(d2, 0).upload(var2)
        cfun = _control_slice(dev, control, control_fun, 0, nodes) if num_steps >= 1 else None

        def dump(i, u1, v1):
            if output_dir is not None and i % 100 == 0:
                t = i * dt
                u1.download().tofile(output_dir / f"chtxs_m_t{round(t, 2)}.csv", sep=",")
                v1.download().tofile(output_dir / f"chtxs_f_t{round(t, 2)}.csv", sep=",")
        _forward_chtxs_dev(dev, cfun, d1, d2, num_steps, dt, rescaling, pingpong=True, after_step=dump)
        last = num_steps % 2
        var1[:] = dev.level(d1, last).download(); var2[:] = dev.level(d2, last).download()
        return var1, var2
    var1[nodes:] = np.zeros(num_steps * nodes)
    var2[nodes:] = np.zeros(num_steps * nodes)
    d1, d2 = dev.traj("var1", num_steps), dev.traj("var2", num_steps)
    dev.level(d1, 0).upload(var1[:nodes]); dev.level(d2, 0).upload(var2[:nodes])
    cfun = _control_slice(dev, control, control_fun, nodes, 2 * nodes) if num_steps >= 1 else None
    _forward_chtxs_dev(dev, cfun, d1, d2, num_steps, dt, rescaling)
    if num_steps >= 1:
        d1.slice(nodes, num_steps * nodes).download(var1[nodes:]); d2.slice(nodes, num_steps * nodes).download(var2[nodes:])
    return var1, var2


def _forward_chtxs_dev(dev, cfun, d1, d2, num_steps, dt, rescaling=1 / 10, pingpong=False, after_step=None):
    """the loop of helpers.py:1318-1385 on device trajectories (level 0 = initial condition): fct_forward_chtxs; the
    generation mode (pingpong: two time levels only, periodic dumps) stays a Python loop over the same device calls"""
    delta, Dm, Df, chi, _, eta = get_chtxs_sys_params()
    ctx, L = dev.ctx, _lib
    print("Solving the system of chemotaxis state equations...")
    if not pingpong:
        cp, cc = _ctl(cfun)
        par, ppar = _hostp([delta, Dm, Df, chi, eta])
        _run("solve_chtxs_system", L.lib.fct_forward_chtxs, ctx.handle, cp, cc, d1.ptr, d2.ptr, int(num_steps), float(dt), ppar,
             float(rescaling), None)
        return
    Mat2 = dev.mat("Mat2"); ctx.vals_axpby(1.0 + dt * delta, dev.M, dt * Df, dev.K, Mat2)   # M + dt(Df Ad + delta M)
    A, rhs = dev.mat("A"), dev.vec("rhs1")
    for i in range(1, num_steps + 1):
        a, b = (i - 1) % 2, i % 2
        un, vn, u1, v1 = dev.level(d1, a), dev.level(d2, a), dev.level(d1, b), dev.level(d2, b)
        # var2_rhs = assemble(v_n w dx + dt * c * u_n / r * w dx)
        ctx.assemble_vector(L.LOAD_P1_1, rhs, c0=vn)
        dev.load_control(rhs, cfun, other=un, scale=dt / rescaling, accumulate=True)
        ctx.axpby(1.0, vn, 0.0, None, v1)
        _solve(ctx, L.SOLVER_PCG, Mat2, rhs, v1, "var2")
        # A_var1 = Dm*Ad - chi*Aa, Aa = exp(-eta u_n) grad(v_{n+1}).grad(w) u   (quadrature degree 4)
        ctx.assemble_matrix(L.FORM_CHTX_EXP, A, c0=v1, c1=un, s0=eta, scale=-chi)
        ctx.vals_axpby(1.0, A, Dm, dev.K, A)
        _check(ctx.step(A, un, dt, u1))
        if after_step is not None:
            after_step(i, u1, v1)


def solve_adjoint_chtxs_system(uk, vk, uhat, vhat, pk, qk, control, T, V, nodes, num_steps, dt, dof_neighbors, optim,
                               show_plots=None, vertex_to_dof=None, out_folder=None, mesh=None, deltax=None,
                               rescaling=1 / 10):
    """helpers.py:1387-1581"""
    valid_options = ["alltime", "finaltime"]
    if optim not in valid_options:
        raise ValueError(f"Invalid value for 'optim': '{optim}'. Must be one of {valid_options}.")
    delta, Dm, Df, chi, _, eta = get_chtxs_sys_params()
    dev = _Dev(V, nodes)
    if optim == "finaltime":
        pk[num_steps * nodes:] = uhat - uk[num_steps * nodes:]
        qk[num_steps * nodes:] = vhat - vk[num_steps * nodes:]
    du, dv = dev.traj("var1", num_steps).upload(uk), dev.traj("var2", num_steps).upload(vk)
    dc = dev.traj("ctl", num_steps).upload(control)
    dp, dq = dev.traj("adj1", num_steps), dev.traj("adj2", num_steps)
    dev.level(dp, num_steps).upload(pk[num_steps * nodes:(num_steps + 1) * nodes])
    dev.level(dq, num_steps).upload(qk[num_steps * nodes:(num_steps + 1) * nodes])
    uh = vh = None
    if optim == "alltime":          # the reference adds the NODAL differences (helpers.py:1509,1535)
        uh, vh = dev.traj("tgt1", num_steps).upload(uhat), dev.traj("tgt2", num_steps).upload(vhat)
    print("\nSolving adjoint equation...")
    par, ppar = _hostp([delta, Dm, Df, chi, eta])
    _run("solve_adjoint_chtxs_system", _lib.lib.fct_adjoint_chtxs, dev.ctx.handle, du.ptr, dv.ptr,
         uh.ptr if uh is not None else None, vh.ptr if vh is not None else None, dp.ptr, dq.ptr, dc.ptr, int(num_steps),
         float(dt), ppar, float(rescaling), None)
    if num_steps >= 1:
        dp.slice(0, num_steps * nodes).download(pk[:num_steps * nodes]); dq.slice(0, num_steps * nodes).download(qk[:num_steps * nodes])
    return pk, qk


def _check(info):
    """fct_step completes an unconverged Jacobi solve with BiCGStab and raises only if that fails too, so an unconverged
    flag here is an internal error"""
    if info is not None and not info.converged:
        raise _lib.FctError(f"low-order solve did not converge ({info.solver_sweeps} Jacobi sweeps)")


# ---- projected Armijo line search -------------------------------------------------------------------------
def armijo_line_search_ref(var1, c, d, var1_target, num_steps, dt, c_lower, c_upper, beta, costfun_init, nodes, optim,
                           V, gam=1e-4, max_iter=10, s0=1, nonlinear_solver=None, dof_neighbors=None, var2=None,
                           var2_target=None, w1=None, w2=None):
    """helpers.py:1583-1713.  Return shape follows the reference: (var1, var2, c_inc, k+1) if var2 is not None else
    (var1, c_inc, k+1) (App. D-11).  The linear branch (w1 given) never assembles M in the reference and fails there
    (App. D-10); here M comes from the function space in both branches."""
    valid_options = ["alltime", "finaltime"]
    if optim not in valid_options:
        raise ValueError(f"Invalid value for 'optim': '{optim}'. Must be one of {valid_options}.")
    fwd = _DEVICE_FORWARD.get(nonlinear_solver) if w1 is None else None
    if fwd is not None:
        return _armijo_ref_device(fwd, var1, c, d, var1_target, num_steps, dt, c_lower, c_upper, beta, costfun_init, nodes,
                                  optim, V, gam, max_iter, s0, var2, var2_target)
    ctx = V.mesh().context()
    M = ctx.to_scipy(ctx.static()[0].download())
    s = s0
    control_dif_L2 = 1
    armijo = float("inf")
    c_inc = c
    k = 0
    for k in range(max_iter):
        print(f"{k=}")
        c_inc = np.clip(c + s * d, c_lower, c_upper)
        if w1 is None:
            var1, var2 = nonlinear_solver(c_inc, var1, var2, V, nodes, num_steps, dt, dof_neighbors)
            cost2 = cost_functional(var1, var1_target, c_inc, num_steps, dt, M, beta, optim=optim, var2=var2,
                                    var2_target=var2_target)
        else:
            var1_inc = var1 + s * w1
            var2_inc = var2 + s * w2 if w2 is not None else None
            cost2 = cost_functional(var1_inc, var1_target, c_inc, num_steps, dt, M, beta, optim=optim, var2=var2_inc,
                                    var2_target=var2_target)
        armijo = cost2 - costfun_init
        control_dif_L2 = L2_norm_sq_Q(c_inc - c, num_steps, dt, M)
        print(f"Updated cost={cost2}, Orig. cost={costfun_init}")
        print(f"Cost difference={armijo}")
        print(f"Threshold value: {-gam/s*control_dif_L2=}")
        if armijo <= -gam / s * control_dif_L2:
            print(f"Converged in {k+1} iterations: Armijo condition satisfied.")
            break
        s /= 2
    if armijo > -gam / s * control_dif_L2:
        print(f"Stopped: Maximum number of iterations reached ({max_iter}) .")
    return (var1, var2, c_inc, k + 1) if var2 is not None else (var1, c_inc, k + 1)


# the reference-named solvers whose loop also exists on device trajectories: the line search then keeps everything on the GPU
_DEVICE_FORWARD = {solve_nonlinear_equation: _forward_nonlinear_dev, solve_schnak_system: _forward_schnak_dev,
                   solve_chtxs_system: _forward_chtxs_dev}


def _cost_device(dev, d1, d_t1, d_cinc, num_steps, dt, beta, optim, d2=None, d_t2=None):
    """cost_functional (helpers.py:383-441) on device trajectories: the same kernels and summation order as the numpy-facing
    shim, without the uploads"""
    ctx = dev.ctx
    if optim == "alltime":
        print("Calculating L^2(Q)-norm...")
        func = 0.5 * ctx.norm_sq_Q(dev.M, d1, num_steps, dt, target=d_t1)
        if d2 is not None and d_t2 is not None:
            func += 0.5 * ctx.norm_sq_Q(dev.M, d2, num_steps, dt, target=d_t2)
    else:
        print("Calculating L^2(\\Omega)-norm...")
        dif = dev.vec("dif")
        ctx.axpby(1.0, dev.level(d1, num_steps), -1.0, d_t1, dif)
        func = 0.5 * ctx.dot_M(dev.M, dif, dif)
        if d2 is not None and d_t2 is not None:
            ctx.axpby(1.0, dev.level(d2, num_steps), -1.0, d_t2, dif)
            func += 0.5 * ctx.dot_M(dev.M, dif, dif)
    func += beta / 2 * ctx.norm_sq_Q(dev.M, d_cinc, num_steps, dt)
    return func


def _armijo_ref_device(fwd, var1, c, d, var1_target, num_steps, dt, c_lower, c_upper, beta, costfun_init, nodes, optim, V,
                       gam, max_iter, s0, var2, var2_target):
    """armijo_line_search_ref for the solvers of this module: control, direction and targets go to the GPU once, every trial
    (projection, state solve, cost, ||c_inc - c||) runs on device trajectories and moves scalars only, the accepted state and
    control come back once.  The share of the host<->device copies is printed."""
    import time
    dev = _Dev(V, nodes); ctx = dev.ctx
    t_start = time.perf_counter()
    ctx.xfer_reset(True)
    size = (num_steps + 1) * nodes
    d_c, d_d, d_cinc, d_dif = (dev.traj(k, num_steps) for k in ("arm_c", "arm_d", "arm_cinc", "arm_dif"))
    d_c.upload(c); d_d.upload(d)
    d1 = dev.traj("var1", num_steps)
    dev.level(d1, 0).upload(var1[:nodes])
    d2 = d_t2 = None
    if var2 is not None:
        d2 = dev.traj("var2", num_steps)
        dev.level(d2, 0).upload(var2[:nodes])
    if optim == "alltime":
        d_t1 = dev.traj("arm_t1", num_steps).upload(var1_target)
        if var2 is not None and var2_target is not None:
            d_t2 = dev.traj("arm_t2", num_steps).upload(var2_target)
    else:
        d_t1 = dev.vec("arm_t1").upload(var1_target)
        if var2 is not None and var2_target is not None:
            d_t2 = dev.vec("arm_t2").upload(var2_target)
    s = s0
    control_dif_L2 = 1
    armijo = float("inf")
    k = 0
    for k in range(max_iter):
        print(f"{k=}")
        ctx.clip_axpy(d_c, s, d_d, c_lower, c_upper, d_cinc)                       # c_inc = clip(c + s d)
        cfun = dev.level(d_cinc, 1) if num_steps >= 1 else None                    # App. D-1: the first step's slice, reused
        fwd(dev, cfun, d1, d2, num_steps, dt)
        cost2 = _cost_device(dev, d1, d_t1, d_cinc, num_steps, dt, beta, optim, d2, d_t2)
        armijo = cost2 - costfun_init
        ctx.axpby(1.0, d_cinc, -1.0, d_c, d_dif, length=size)
        control_dif_L2 = ctx.norm_sq_Q(dev.M, d_dif, num_steps, dt)
        print(f"Updated cost={cost2}, Orig. cost={costfun_init}")
        print(f"Cost difference={armijo}")
        print(f"Threshold value: {-gam/s*control_dif_L2=}")
        if armijo <= -gam / s * control_dif_L2:
            print(f"Converged in {k+1} iterations: Armijo condition satisfied.")
            break
        s /= 2
    if armijo > -gam / s * control_dif_L2:
        print(f"Stopped: Maximum number of iterations reached ({max_iter}) .")
    c_inc = d_cinc.download()
    if num_steps >= 1:
        var1[nodes:] = 0.0
        d1.slice(nodes, num_steps * nodes).download(var1[nodes:])
        if var2 is not None:
            var2[nodes:] = 0.0
            d2.slice(nodes, num_steps * nodes).download(var2[nodes:])
    ctx.sync()
    x = ctx.xfer_reset(False)
    total = time.perf_counter() - t_start
    print(f"H2D/D2H: {x['seconds']:.4f} s of {total:.4f} s ({100 * x['seconds'] / total:.1f} %), "
          f"{x['h2d_bytes'] / 1e6:.1f} MB in (control, direction, targets, initial state), {x['d2h_bytes'] / 1e6:.1f} MB out "
          f"(accepted state, control); {k + 1} trial(s) moved scalars only")
    _LAST_ARMIJO_XFER.update(x, total_seconds=total, trials=k + 1)
    return (var1, var2, c_inc, k + 1) if var2 is not None else (var1, c_inc, k + 1)


_LAST_ARMIJO_XFER = {}


# ---- legacy drift-control line search (config 2 / 5) ------------------------------------------------------------
def cost_functional_proj(u, w, c, d, s, uhatvec, num_steps, dt, M, c_lower, c_upper, beta):
    """The reference calls this (old_helpers.py:32,73; advection_FCT_PDECO_alltime_exact.py:213,310) but no definition
    survives anywhere in the repo.  Re-specified from its call sites as the all-time cost of the
    state u + s*w and the projected control clip(c, c_lower, c_upper); with w = 0 and an already projected c -- the
    only way the drift line search uses it -- this is cost_functional(u, uhat, c, ..., 'alltime').  PARITY UNPINNED."""
    return cost_functional(np.asarray(u) + s * np.asarray(w), uhatvec, np.clip(c, c_lower, c_upper), num_steps, dt, M, beta,
                           "alltime")


def armijo_line_search_sbr_drift(u, p, c, d, uhatvec, eps, drift, num_steps, dt, nodes, M, M_Lump, Ad, Arot, c_lower,
                                 c_upper, beta, V, dof_neighbors, gam=10 ** -4, max_iter=5, s0=1, optim="alltime"):
    """old_helpers.py:1-85: projected Armijo search for the drift-speed control; every trial is a full forward solve of
    du/dt + div(u c b) - eps lap(u) = 0, run here as one device-resident loop (fct_advdrift_state).  `drift` is the
    constant vector b (a 2-tuple or forms.Constant), `Arot` must be zero (as in the script, :148).  Writes the state
    of the accepted trial into `u` in place and returns (s, u) like the reference.  optim='finaltime' needs the lost
    cost_functional_proj_FT and is not supported."""
    if optim != "alltime":
        raise NotImplementedError("armijo_line_search_sbr_drift: only optim='alltime' (cost_functional_proj_FT is lost)")
    if Arot is not None and getattr(Arot, "nnz", 0) and abs(Arot).max() != 0:
        raise NotImplementedError("armijo_line_search_sbr_drift: the rotation operator must be zero")
    b = getattr(drift, "b", drift)
    bx, by = float(b[0]), float(b[1])
    ctx = V.mesh().context()
    if ctx.n != nodes:
        raise ValueError(f"nodes={nodes} does not match the function space ({ctx.n})")
    k = 0
    s = 1
    Z = np.zeros(np.shape(u))
    grad_costfun_L2 = L2_norm_sq_Q(np.clip(c + s * d, c_lower, c_upper) - c, num_steps, dt, M)
    print(f"{grad_costfun_L2=}")
    costfun_init = cost_functional_proj(u, Z, c, d, s, uhatvec, num_steps, dt, M, c_lower, c_upper, beta)
    d_c, d_u = ctx.empty(u.size), ctx.empty(u.size)
    armijo = 10 ** 5
    while armijo > -gam / s * grad_costfun_L2 and k < max_iter:
        s = s0 * (1 / 2 ** k)
        c_inc = np.clip(c + s * d, c_lower, c_upper)
        print(f"{k =}")
        print("Solving state equations...")
        u[nodes:] = np.zeros(num_steps * nodes)
        d_c.upload(c_inc); d_u.upload(u)
        ctx.advdrift_state(d_c, d_u, num_steps, dt, bx=bx, by=by, eps=float(eps))
        d_u.download(u)
        cost2 = cost_functional_proj(u, Z, c_inc, d, s, uhatvec, num_steps, dt, M, c_lower, c_upper, beta)
        armijo = cost2 - costfun_init
        grad_costfun_L2 = L2_norm_sq_Q(c_inc - c, num_steps, dt, M)
        k += 1
    d_c.free(); d_u.free()
    print(f"Armijo exit at {k=} with {s=}")
    return s, u


# ---- lost legacy names of configs 3 and 4: re-specified, PARITY UNPINNED -------------------------------------
def cost_functional_proj_FT(var1, var2, c, d, s, var1_target, var2_target, num_steps, dt, M, c_lower, c_upper, beta):
    """Called by chemotaxis_mimura_FCT_PGD.py:138,252 as (mk, fk, ck, dk, sk, mhat_T, fhat_T, ...) and by old_helpers.py:35,76 /
    advection_FCT_PDECO_finaltime_exact.py:230 as (u, Z, c, d, s, uhat_T, z, ...) with Z, z zero arrays; no definition
    survives in the reference.  The one reading that fits both call sites: final-time tracking of BOTH states plus the
    Tikhonov term of the projected control,
        1/2 |var1(T) - var1_target|^2_M + 1/2 |var2(T) - var2_target|^2_M + beta/2 |clip(c)|^2_{L2(Q)},
    (the advection scripts pass zeros for the second species, which then contributes nothing); `d`, `s` describe the step
    that produced `c` and do not enter.  PARITY UNPINNED."""
    return cost_functional(np.asarray(var1), var1_target, np.clip(c, c_lower, c_upper), num_steps, dt, M, beta, "finaltime",
                           var2=np.asarray(var2), var2_target=var2_target)


def armijo_line_search(*args, **kwargs):
    """The reference's scripts call a function of this name in two generations, neither of which survives:

    (A) Schnak_FCT_PDECO.py:297-300, chemotaxis_FCT_PDECO.py:272-276, nonlinear_FCT_PDECO_alltime.py:229-232,
        advection_FCT_PDECO_finaltime.py:265-267:
            armijo_line_search(var1, c, d, var1_target, num_steps, dt, M, c_lower, c_upper, beta, costfun_init, nodes,
                               V=, optim=, dof_neighbors=, example='Schnak'|'chtxs'|'nonlinear', var2=, var2_target=,
                               w=|w1=, w2=, max_iter=)  ->  (s, var1_inc[, var2_inc])
    (B) advection_FCT_PDECO_alltime_exact.py:299, advection_FCT_PDECO_finaltime_exact.py:374-376 (linear problems):
            armijo_line_search(u, p, w, c, d, uhat, num_steps, dt, M, c_lower, c_upper, beta[, costfun_init, optim=])
                               ->  s   (or (s, u_inc) when costfun_init is given)

    Both are re-specified on armijo_line_search_ref (helpers.py:1583-1713), the surviving refactoring of the same search:
    projected backtracking s = s0 / 2^k on cost(clip(c + s d)) - cost0 <= -gam / s |clip(c + s d) - c|^2_{L2(Q)}, every trial
    a full forward solve with the solver `example` names (A) or the linearised state u + s w (B, and A with w / w1 given).
    PARITY UNPINNED."""
    if len(args) >= 5 and np.ndim(args[4]) == 0:
        return _armijo_legacy_a(*args, **kwargs)
    return _armijo_legacy_b(*args, **kwargs)


def _armijo_legacy_a(var1, c, d, var1_target, num_steps, dt, M, c_lower, c_upper, beta, costfun_init, nodes, V=None,
                     optim="alltime", dof_neighbors=None, example=None, var2=None, var2_target=None, w=None, w1=None, w2=None,
                     gam=1e-4, max_iter=10, s0=1):
    if w1 is None:
        w1 = w
    solver = None
    if w1 is None:
        solvers = {"Schnak": solve_schnak_system, "schnak": solve_schnak_system, "nonlinear": solve_nonlinear_equation,
                   "chtxs": solve_chtxs_system}
        if example not in solvers:
            raise ValueError(f"Invalid value for 'example': '{example}'. Must be one of {sorted(set(solvers))}.")
        solver = solvers[example]
    res = armijo_line_search_ref(var1, c, d, var1_target, num_steps, dt, c_lower, c_upper, beta, costfun_init, nodes, optim, V,
                                 gam=gam, max_iter=max_iter, s0=s0, nonlinear_solver=solver, dof_neighbors=dof_neighbors,
                                 var2=var2, var2_target=var2_target, w1=w1, w2=w2)
    its = res[-1]
    s = s0 / 2 ** (its - 1)                      # the step of the last trial, i.e. of the returned state / control
    v1 = res[0]
    v2 = res[1] if var2 is not None else None
    if w1 is not None:                           # linearised problems: the trial state is u + s w
        v1 = np.asarray(var1) + s * np.asarray(w1)
        v2 = None if var2 is None else (np.asarray(var2) + s * np.asarray(w2) if w2 is not None else var2)
    return (s, v1, v2) if var2 is not None else (s, v1)


def _armijo_legacy_b(u, p, w, c, d, uhat, num_steps, dt, M, c_lower, c_upper, beta, costfun_init=None, optim="alltime",
                     gam=1e-4, max_iter=10, s0=1):
    if optim not in ("alltime", "finaltime"):
        raise ValueError(f"Invalid value for 'optim': '{optim}'. Must be one of ['alltime', 'finaltime'].")
    u, w, c, d = (np.asarray(x, dtype=np.float64) for x in (u, w, c, d))
    nodes = u.size // (num_steps + 1)

    def cost(ui, ci):
        return cost_functional(ui, uhat, ci, num_steps, dt, M, beta, optim)
    cost0 = cost(u, np.clip(c, c_lower, c_upper)) if costfun_init is None else costfun_init
    s = s0
    for k in range(max_iter):
        s = s0 / 2 ** k
        c_inc = np.clip(c + s * d, c_lower, c_upper)
        u_inc = u + s * w
        if cost(u_inc, c_inc) - cost0 <= -gam / s * L2_norm_sq_Q(c_inc - c, num_steps, dt, M):
            break
    del nodes
    return s if costfun_init is None else (s, u + s * w)


def armijo_line_search_chtxs(mk, fk, qk, ck, dk, mhat_T, fhat_T, Mat_fq, chi, Dm, Df, num_steps, dt, nodes, M, M_Lump, Ad,
                             c_lower, c_upper, beta, V, dof_neighbors, gam=10 ** -4, max_iter=5, s0=1):
    """chemotaxis_mimura_FCT_PGD.py:238-240 calls this; no definition survives.  Re-specified on the pattern of its sibling
    armijo_line_search_sbr_drift (old_helpers.py:1-85): s = s0 / 2^k, c_inc = clip(c + s d), every trial re-runs the script's
    own state loop (chemotaxis_mimura_FCT_PGD.py:160-183: rhs_chtx_f, the f-solve with Mat_fq, mat_chtx_m, rhs_chtx_m, the
    legacy FCT_alg) into mk, fk IN PLACE, the cost is cost_functional_proj_FT; returns the accepted step s.  PARITY UNPINNED."""
    from . import mimura_data_helpers as mdh
    from .forms import TestFunction, TrialFunction, vec_to_function
    from .helpers import FCT_alg, context_for, rhs_chtx_f
    u_, v_ = TrialFunction(V), TestFunction(V)
    ctx = V.mesh().context()
    dMat = ctx.array(ctx.embed(Mat_fq))
    drhs, dx_ = ctx.empty(ctx.n), ctx.empty(ctx.n)
    Z = np.zeros_like(ck)
    k, s = 0, 1
    grad_costfun_L2 = L2_norm_sq_Q(np.clip(ck + s * dk, c_lower, c_upper) - ck, num_steps, dt, M)
    print(f"{grad_costfun_L2=}")
    costfun_init = cost_functional_proj_FT(mk, fk, ck, dk, s, mhat_T, fhat_T, num_steps, dt, M, c_lower, c_upper, beta)
    armijo = 10 ** 5
    del Z, context_for
    while armijo > -gam / s * grad_costfun_L2 and k < max_iter:
        s = s0 * (1 / 2 ** k)
        c_inc = np.clip(ck + s * dk, c_lower, c_upper)
        print(f"{k =}")
        print("Solving state equations...")
        fk[nodes:] = np.zeros(num_steps * nodes)
        mk[nodes:] = np.zeros(num_steps * nodes)
        for i in range(1, num_steps + 1):
            start, end = i * nodes, (i + 1) * nodes
            m_n = mk[start - nodes:start]
            m_n_fun = vec_to_function(m_n, V)
            c_np1_fun = vec_to_function(c_inc[start:end], V)
            f_n_fun = vec_to_function(fk[start - nodes:start], V)
            f_rhs = rhs_chtx_f(f_n_fun, m_n_fun, c_np1_fun, dt, v_)
            drhs.upload(f_rhs); dx_.upload(fk[start - nodes:start])
            _solve(ctx, _lib.SOLVER_PCG, dMat, drhs, dx_, "f (Mat_fq)")
            dx_.download(fk[start:end])
            f_np1_fun = vec_to_function(fk[start:end], V)
            A_m = mdh.mat_chtx_m(f_np1_fun, m_n_fun, Dm, chi, u_, v_)
            m_rhs = mdh.rhs_chtx_m(m_n_fun, v_)
            mk[start:end] = FCT_alg(A_m, m_rhs, m_n, dt, nodes, M, M_Lump, dof_neighbors)
        cost2 = cost_functional_proj_FT(mk, fk, c_inc, dk, s, mhat_T, fhat_T, num_steps, dt, M, c_lower, c_upper, beta)
        armijo = cost2 - costfun_init
        grad_costfun_L2 = L2_norm_sq_Q(c_inc - ck, num_steps, dt, M)
        k += 1
    for a in (dMat, drhs, dx_):
        a.free()
    print(f"Armijo exit at {k=} with {s=}")
    return s


# ---- trajectory I/O --------------------------------------------------------------------------------
def import_data_final(file_path, nodes, vertex_to_dof, num_steps=0, time_dep=False):
    """helpers.py:1874-1911.  Besides the reference's comma-separated text (`np.tofile(sep=',')`), a `.npy` file with the
    same flat DoF-ordered layout is accepted: a 4097^2 all-time trajectory is gigabytes of text otherwise."""
    from .helpers import reorder_vector_from_dof
    sqnodes = round(np.sqrt(nodes))
    if str(file_path).endswith(".npy"):
        data = np.load(file_path, mmap_mode="r")
    else:
        data = np.genfromtxt(file_path, delimiter=",")
    if time_dep:
        data = np.asarray(data[:(num_steps + 1) * nodes], dtype=np.float64)
        data_re = reorder_vector_from_dof(data, num_steps + 1, nodes, vertex_to_dof)
    else:
        data = np.asarray(data[num_steps * nodes:(num_steps + 1) * nodes], dtype=np.float64)
        data_re = reorder_vector_from_dof(data, 1, nodes, vertex_to_dof)
        data_re = data_re.reshape((sqnodes, sqnodes))
    return data_re, data


def export_trajectory(file_path, data):
    """Writer counterpart of import_data_final / extract_data: the flat DoF-ordered trajectory either as the
    reference's comma-separated text (`data.tofile(path, sep=",")`, e.g. advection_solidbody_FCT_PDECO_alltime.py:409-411) or,
    for a path ending in .npy, as a binary array with the same layout (memory-mappable; 8 B instead of ~25 B per value)."""
    data = np.ascontiguousarray(data, dtype=np.float64).ravel()
    if str(file_path).endswith(".npy"):
        np.save(file_path, data)
    else:
        data.tofile(file_path, sep=",")
    return None


def extract_data(file_path, file_name, T, dt, nodes, vertex_to_dof):
    """helpers.py:1913-1956: cut the time slice at T out of a flat trajectory file (csv as in the reference, or .npy)"""
    import os
    idx = round(T / dt)
    start_col, end_col = idx * nodes, (idx + 1) * nodes
    npy = os.path.join(file_path, f"{file_name}.npy")
    if os.path.exists(npy):
        data = np.asarray(np.load(npy, mmap_mode="r")[start_col:end_col])
        output_file = os.path.join(file_path, f"{file_name}_T{T}.npy")
        np.save(output_file, data)
    else:
        import pandas as pd
        input_file = os.path.join(file_path, f"{file_name}.csv")
        output_file = os.path.join(file_path, f"{file_name}_T{T}.csv")
        data = pd.read_csv(input_file, header=None, usecols=range(start_col, end_col), nrows=1)
        np.savetxt(output_file, data.to_numpy().flatten(), delimiter=",")
    print(f"Extracted data at {T=} into {output_file}.")
    return None
