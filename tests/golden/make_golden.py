"""Generate the committed golden fixtures under tests/golden/.

Run ONLY in the build container (needs /root/reference); the fixtures it writes are committed so
that the GPU box -- which has no reference tree -- can run the parity tests.

  python tests/golden/make_golden.py

Writes
  ref_data.npz        the reference's shipped data files, as float64 binary:
                        chtxs_m, chtxs_f  [11,1681]  (Chtxs_data_dx0.025_dt0.001/chtxs_{m,f}_t0.01.csv)
                        solidbody_t0.25, _t0.5, _t1  [6561]   (data/solidbody_t*_u.csv)
  ref_fct_cases.npz   inputs + outputs of the reference's OWN functions, imported unmodified from
                      /root/reference/helpers.py with dolfin/matplotlib stubbed (oracle/ref_loader.py):
                      FCT_alg_ref (4 cases incl. rhs / non_flux_mat / pruned zeros), ChebSI,
                      artificial_diffusion_mat, L2_norm_sq_Q, L2_norm_sq_Omega, cost_functional.
  ref_legacy.npz      inputs + outputs of the reference's OWN legacy FCT_alg (old_helpers.py:112-204, an import-less
                      fragment: its source is executed unmodified in the namespace of the reference's helpers.py, which
                      provides numpy / scipy and ChebSI / artificial_diffusion_mat / sparse_nonzero): two cases, one with
                      source_mat and a right-hand side.
  ref_armijo.npz      inputs + outputs of the reference's OWN armijo_line_search_ref (helpers.py:1583-1713), run
                      unmodified with `assemble_sparse` returning the restated mass matrix and the oracle's time
                      loops as `nonlinear_solver` callbacks (one-species final-time case, two-species all-time case,
                      a case that exhausts max_iter).
  ref_loops.npz       inputs + outputs of the reference's OWN time loops (solve_schnak_system, solve_adjoint_schnak_system,
                      solve_nonlinear_equation, solve_adjoint_nonlinear_equation, solve_chtxs_system,
                      solve_adjoint_chtxs_system all-time / final-time) and of armijo_line_search_ref with the reference's own
                      solver as callback: helpers.py runs unmodified on oracle/fake_dolfin.py, a numpy stand-in for the slice
                      of dolfin it uses, which is first checked on the shipped chemotaxis trajectory.
  ref_cfg2.npz        BASELINE config 2: inputs + outputs of the state / adjoint / gradient loops of
                      advection_solidbody_FCT_PDECO_alltime.py:206-275 -- the script's own source lines executed with the
                      reference's helpers.py (on oracle/fake_dolfin.py) and its legacy FCT_alg.
  ref_cfg3.npz        BASELINE config 3: state + adjoint loops of chemotaxis_mimura_FCT_PGD.py:157-225, executed from the
                      script's source with the reference's mimura_data_helpers.py and old_helpers.py form builders.
  ref_cfg4.npz        BASELINE config 4: state + adjoint loops of Schnak_FCT_PDECO.py:190-279 (time-dependent Expression
                      wind, project(wind, W), div(w_h u) w), executed from the script's source.
"""
import os
import sys

import numpy as np
from scipy.sparse import lil_matrix

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.p1assembly import P1Assembler          # noqa: E402
from oracle.p1mesh import RectMesh                 # noqa: E402
from oracle.fct_numpy import Pattern               # noqa: E402
from oracle.ref_loader import REFERENCE_DIR, load_reference_helpers   # noqa: E402


def ref_data():
    out = {}
    d = os.path.join(REFERENCE_DIR, "Chtxs_data_dx0.025_dt0.001")
    out["chtxs_m"] = np.genfromtxt(os.path.join(d, "chtxs_m_t0.01.csv"), delimiter=",").reshape(11, 1681)
    out["chtxs_f"] = np.genfromtxt(os.path.join(d, "chtxs_f_t0.01.csv"), delimiter=",").reshape(11, 1681)
    for t in ("0.25", "0.5", "1"):
        out[f"solidbody_t{t}"] = np.genfromtxt(os.path.join(REFERENCE_DIR, "data", f"solidbody_t{t}_u.csv"),
                                               delimiter=",")
    np.savez_compressed(os.path.join(HERE, "ref_data.npz"), **out)
    print("ref_data.npz:", {k: v.shape for k, v in out.items()})


def ref_fct_cases():
    hp = load_reference_helpers()
    rng = np.random.default_rng(5)
    out = {}

    def run_case(tag, n, a1, a2, A, rhs, u_n, dt, S, asm, pat, mesh):
        nodes = mesh.nodes
        M = asm.mass()
        Ml = lil_matrix(pat.csr(M))
        MLl = hp.row_lump(Ml, nodes)
        nb = mesh.dof_neighbors()
        Ain = pat.csr(A)
        Ain.eliminate_zeros()          # what reaches FCT_alg_ref after scipy arithmetic (App. D-5)
        Sin = None if S is None else pat.csr(S)
        u1 = hp.FCT_alg_ref(Ain, rhs, u_n, dt, nodes, Ml, MLl, nb, non_flux_mat=Sin)
        out[f"{tag}_n"] = np.array([n])
        out[f"{tag}_box"] = np.array([a1, a2], dtype=np.float64)
        out[f"{tag}_A"] = A
        out[f"{tag}_rhs"] = rhs
        out[f"{tag}_un"] = u_n
        out[f"{tag}_dt"] = np.array([dt])
        out[f"{tag}_S"] = np.zeros(0) if S is None else S
        out[f"{tag}_out"] = np.asarray(u1).ravel()
        print(tag, "done", nodes)

    # case a: solid-body rotation + drift operator (advection_solidbody_FCT.py), legacy sign -> ref sign
    n, a1, a2 = 16, -1.0, 1.0
    mesh = RectMesh(n, a1, a2); asm = P1Assembler(mesh); pat = Pattern(*mesh.pattern())
    om = np.pi / 40
    xy = mesh.dof_xy
    A = -asm.conv_conservative_p1(-xy[:, 1] / om + 2, xy[:, 0] / om + 2)
    u_n = (np.hypot(xy[:, 0], xy[:, 1] - 1 / 3) < 1 / 3).astype(np.float64)
    run_case("solid", n, a1, a2, A, np.zeros(mesh.nodes), u_n, 0.125 ** 2 / 4, None, asm, pat, mesh)

    # case b: chemotaxis operator Dm*K - chi*Aa, random fields (helpers.py:1350-1356)
    n, a1, a2 = 12, 0.0, 1.0
    mesh = RectMesh(n, a1, a2); asm = P1Assembler(mesh); pat = Pattern(*mesh.pattern())
    m = 1.5 + 0.1 * (0.5 - rng.random(mesh.nodes))
    f = 1.5 + 0.1 * (0.5 - rng.random(mesh.nodes))
    K = asm.stiffness()
    Aa = asm.chemotaxis_conv(f, lambda phi, xyq: np.exp(-0.5 * asm.at_quad(m, phi)), degree=4)
    run_case("chtxs", n, a1, a2, 0.05 * K - 0.25 * Aa, np.zeros(mesh.nodes), m, 1e-3, None, asm, pat, mesh)

    # case c: Schnakenberg-like: wind operator, rhs, non_flux_mat = gamma*M (helpers.py:579-589)
    n, a1, a2 = 10, 0.0, 1.0
    mesh = RectMesh(n, a1, a2); asm = P1Assembler(mesh); pat = Pattern(*mesh.pattern())
    wind = lambda x, y: ((y - 0.5) * x * (1 - x), -(x - 0.5) * y * (1 - y))
    Aw = asm.conv_conservative(wind, degree=5)
    K = asm.stiffness(); M = asm.mass()
    u_n = 1.0 + 0.1 * np.cos(2 * np.pi * (mesh.dof_xy[:, 0] + mesh.dof_xy[:, 1]))
    rhs = asm.load_p1_product(u_n, u_n, scale=230.82) + asm.load_constant(23.082)
    run_case("schnak", n, a1, a2, 0.01 * K - 100 * Aw, rhs, u_n, 1e-3, 230.82 * M, asm, pat, mesh)

    # case d: drift-control operator (advection_solidbody_FCT_PDECO_alltime.py:222-228), random control
    n, a1, a2 = 12, -1.0, 1.0
    mesh = RectMesh(n, a1, a2); asm = P1Assembler(mesh); pat = Pattern(*mesh.pattern())
    c = 1.0 + rng.random(mesh.nodes)
    A_u = asm.drift_mass(c, 1.0, 1.0) + asm.drift_conv(c, 1.0, 1.0)
    xy = mesh.dof_xy
    u_n = np.exp(-20 * ((xy[:, 0] + 2 / 3) ** 2 + 5 * (xy[:, 1] + 5 / 6) ** 2))
    rhs = asm.load_p1_product(rng.random(mesh.nodes))
    run_case("drift", n, a1, a2, -A_u, rhs, u_n, 0.01, None, asm, pat, mesh)

    # ChebSI / artificial_diffusion_mat / norms on the last mesh
    M = asm.mass()
    Mc = pat.csr(M)
    b = rng.random(mesh.nodes)
    out["cheb_b"] = b
    out["cheb_out"] = hp.ChebSI(b, Mc, Mc.diagonal(), 20, 0.5, 2)
    out["cheb7_out"] = hp.ChebSI(b, Mc, Mc.diagonal(), 7, 0.5, 2)
    Dref = hp.artificial_diffusion_mat(lil_matrix(pat.csr(A_u)))
    out["adm_in"] = A_u
    out["adm_out"] = pat.embed(Dref)
    ns, dt = 4, 0.05
    phi = rng.random((ns + 1) * mesh.nodes)
    tgt = rng.random((ns + 1) * mesh.nodes)
    ctl = rng.random((ns + 1) * mesh.nodes)
    out["norm_phi"] = phi
    out["norm_tgt"] = tgt
    out["norm_ctl"] = ctl
    out["norm_Q"] = np.array([hp.L2_norm_sq_Q(phi, ns, dt, Mc)])
    out["norm_Omega"] = np.array([hp.L2_norm_sq_Omega(phi[:mesh.nodes], Mc)])
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        out["cost_alltime"] = np.array([hp.cost_functional(phi, tgt, ctl, ns, dt, Mc, 0.01, "alltime")])
        out["cost_finaltime"] = np.array([hp.cost_functional(phi, tgt[:mesh.nodes], ctl, ns, dt, Mc, 0.01,
                                                             "finaltime")])
        out["cost_alltime2"] = np.array([hp.cost_functional(phi, tgt, ctl, ns, dt, Mc, 0.01, "alltime",
                                                            var2=tgt, var2_target=phi)])
    out["norm_meta"] = np.array([ns, dt, 0.01])
    np.savez_compressed(os.path.join(HERE, "ref_fct_cases.npz"), **out)
    print("ref_fct_cases.npz written:", len(out), "arrays")


def ref_legacy():
    """old_helpers.py has no imports; FCT_alg's source text is compiled as it stands."""
    import contextlib
    import io
    hp = load_reference_helpers()
    src = open(os.path.join(REFERENCE_DIR, "old_helpers.py")).read().splitlines()
    start = next(i for i, l in enumerate(src) if l.startswith("def FCT_alg("))
    ns = dict(vars(hp))
    exec(compile("\n".join(src[start:]), "old_helpers.py:FCT_alg", "exec"), ns)
    FCT_alg = ns["FCT_alg"]
    rng = np.random.default_rng(11)
    out = {}

    def run_case(tag, n, a1, a2, A_legacy, rhs, u_n, dt, S):
        mesh = RectMesh(n, a1, a2); asm = P1Assembler(mesh); pat = Pattern(*mesh.pattern())
        Ml = lil_matrix(pat.csr(asm.mass()))
        MLl = hp.row_lump(Ml, mesh.nodes)
        Ain = pat.csr(A_legacy); Ain.eliminate_zeros()
        Sin = None if S is None else pat.csr(S)
        with contextlib.redirect_stdout(io.StringIO()):
            u1 = FCT_alg(Ain, rhs, u_n, dt, mesh.nodes, Ml, MLl, mesh.dof_neighbors(), source_mat=Sin)
        out[f"{tag}_n"] = np.array([n]); out[f"{tag}_box"] = np.array([a1, a2], dtype=np.float64)
        out[f"{tag}_A"] = A_legacy; out[f"{tag}_rhs"] = rhs; out[f"{tag}_un"] = u_n; out[f"{tag}_dt"] = np.array([dt])
        out[f"{tag}_S"] = np.zeros(0) if S is None else S
        out[f"{tag}_out"] = np.asarray(u1).ravel()
        print(tag, "legacy FCT_alg done", mesh.nodes)

    # (a) solid-body rotation + drift, legacy sign as in advection_solidbody_FCT.py:106-148
    n, a1, a2 = 14, -1.0, 1.0
    mesh = RectMesh(n, a1, a2); asm = P1Assembler(mesh)
    om = np.pi / 40
    xy = mesh.dof_xy
    A_u = asm.conv_conservative_p1(-xy[:, 1] / om + 2, xy[:, 0] / om + 2)
    u_n = (np.hypot(xy[:, 0], xy[:, 1] - 1 / 3) < 1 / 3).astype(np.float64)
    run_case("lsolid", n, a1, a2, A_u, np.zeros(mesh.nodes), u_n, (2.0 / n) ** 2 / 4, None)
    # (b) Schnakenberg-like legacy call with source_mat and a right-hand side (Schnak_FCT_PDECO.py:205)
    n, a1, a2 = 10, 0.0, 1.0
    mesh = RectMesh(n, a1, a2); asm = P1Assembler(mesh)
    wind = lambda x, y: ((y - 0.5) * x * (1 - x), -(x - 0.5) * y * (1 - y))
    Aw = asm.conv_conservative(wind, degree=5)
    u_n = 1.0 + 0.1 * np.cos(2 * np.pi * (mesh.dof_xy[:, 0] + mesh.dof_xy[:, 1])) + 0.01 * rng.random(mesh.nodes)
    rhs = asm.load_p1_product(u_n, u_n, scale=230.82) + asm.load_constant(23.082)
    run_case("lschnak", n, a1, a2, -(0.01 * asm.stiffness() - 100 * Aw), rhs, u_n, 1e-3, 230.82 * asm.mass())
    np.savez_compressed(os.path.join(HERE, "ref_legacy.npz"), **out)
    print("ref_legacy.npz written:", len(out), "arrays")


def ref_armijo():
    """The reference's projected Armijo search is host control flow around a solver callback and cost_functional;
    it builds M itself through dolfin (helpers.py:1654-1660), which is the only thing patched here."""
    import contextlib
    import io
    from oracle import pdeco_systems as osys
    from oracle.fct_numpy import cost_functional as o_cost
    hp = load_reference_helpers()

    class _Sym:                       # stands in for TrialFunction / TestFunction / dx in `u*v*dx`
        def __mul__(self, other): return self
        __rmul__ = __mul__

    out = {}
    rng = np.random.default_rng(7)

    def run(tag, prob, solver_ref, var1, c, d, target, ns, dt, lo, hi, beta, cost0, optim, **kw):
        Mc = prob.pat.csr(prob.M)
        saved = (hp.df.TrialFunction, hp.df.TestFunction, hp.dx, hp.assemble_sparse)
        hp.df.TrialFunction = lambda V: _Sym()
        hp.df.TestFunction = lambda V: _Sym()
        hp.dx = _Sym()
        hp.assemble_sparse = lambda form: Mc
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                res = hp.armijo_line_search_ref(var1, c, d, target, ns, dt, lo, hi, beta, cost0, prob.nodes, optim, None,
                                                nonlinear_solver=solver_ref, dof_neighbors=None, **kw)
        finally:
            hp.df.TrialFunction, hp.df.TestFunction, hp.dx, hp.assemble_sparse = saved
        out[f"{tag}_c"] = c; out[f"{tag}_d"] = d; out[f"{tag}_target"] = target
        out[f"{tag}_meta"] = np.array([ns, dt, lo, hi, beta, cost0, kw.get("max_iter", 10)], dtype=np.float64)
        out[f"{tag}_var1"] = np.asarray(res[0]).ravel()
        out[f"{tag}_cinc"] = np.asarray(res[-2]).ravel()
        out[f"{tag}_k"] = np.array([res[-1]])
        if len(res) == 4:
            out[f"{tag}_var2"] = np.asarray(res[1]).ravel()
        print(tag, "armijo iterations:", res[-1])

    # (a) one species, final-time tracking (nonlinear_FCT_PDECO_refactored.py:148-152)
    n, ns, dt, beta = 10, 4, 2e-3, 0.1
    prob = osys.NonlinearProblem(n, 0.0, 1.0)
    u0 = prob.initial_condition()
    c = rng.random((ns + 1) * prob.nodes)
    u = prob.state(c, u0, ns, dt)
    uhat_T = u[-1] + 0.05 * rng.random(prob.nodes)
    p = prob.adjoint(u, uhat_T, ns, dt)
    d = -(beta * c - p.ravel())
    cost0 = o_cost(prob.pat, u.ravel(), uhat_T, c, ns, dt, prob.M, beta, "finaltime")
    solver = lambda ci, v1, v2, V, nodes, num_steps, dt_, nb: (prob.state(ci, u0, num_steps, dt_).ravel(), None)
    out["nl_u0"] = u0
    run("nl", prob, solver, u.ravel().copy(), c, d, uhat_T, ns, dt, 0.0, 1.0, beta, cost0, "finaltime")
    # (a'') an overshooting direction with wide bounds: the step is halved a few times before the condition holds
    run("nlbt", prob, solver, u.ravel().copy(), c, 400.0 * d, uhat_T, ns, dt, -50.0, 50.0, beta, cost0, "finaltime")
    # (a') the same search with a direction that never satisfies the condition: exhausts max_iter = 3
    run("nlmax", prob, solver, u.ravel().copy(), c, -d, uhat_T, ns, dt, 0.0, 1.0, beta, cost0, "finaltime", max_iter=3)

    # (b) two species, all-time tracking (Schnak_FCT_PDECO_refactored.py:167-178)
    n, ns, dt, beta = 8, 3, 1e-3, 0.05
    prob2 = osys.SchnakProblem(n, 0.0, 1.0)
    u0, v0 = prob2.initial_condition()
    c2 = 0.5 + rng.random((ns + 1) * prob2.nodes)
    uu, vv = prob2.state(c2, u0, v0, ns, dt)
    uhat = uu.ravel() + 0.01 * rng.random(uu.size)
    vhat = vv.ravel() + 0.01 * rng.random(vv.size)
    d2 = rng.random(c2.size) - 0.5
    cost0 = o_cost(prob2.pat, uu.ravel(), uhat, c2, ns, dt, prob2.M, beta, "alltime", var2=vv.ravel(), var2_target=vhat)

    def solver2(ci, v1, v2, V, nodes, num_steps, dt_, nb):
        a, b = prob2.state(ci, u0, v0, num_steps, dt_)
        return a.ravel(), b.ravel()
    out["sk_u0"] = u0; out["sk_v0"] = v0; out["sk_target2"] = vhat
    run("sk", prob2, solver2, uu.ravel().copy(), c2, d2, uhat, ns, dt, 0.0, 5.0, beta, cost0, "alltime",
        var2=vv.ravel().copy(), var2_target=vhat)
    np.savez_compressed(os.path.join(HERE, "ref_armijo.npz"), **out)
    print("ref_armijo.npz written:", len(out), "arrays")


def ref_loops():
    """The reference's OWN time loops and line search, helpers.py run unmodified on oracle/fake_dolfin.py (numpy P1 assembly
    behind dolfin's names): solve_schnak_system / solve_adjoint_schnak_system (helpers.py:511-698), solve_nonlinear_equation /
    solve_adjoint_nonlinear_equation (:881-1038), solve_chtxs_system / solve_adjoint_chtxs_system (:1250-1581, all-time and
    final-time), armijo_line_search_ref (:1583-1713) with the reference's own solver as callback.  Before anything is
    written the stand-in is checked end to end on the reference's shipped data: solve_chtxs_system(control_fun=Constant(100),
    rescaling=1) on the 40 x 40 mesh must reproduce Chtxs_data_dx0.025_dt0.001/chtxs_{m,f}_t0.01.csv."""
    import contextlib
    import io
    from oracle import fake_dolfin as fd
    from oracle.ref_loader import load_reference_helpers_on_fake_dolfin
    hp = load_reference_helpers_on_fake_dolfin()
    quiet = contextlib.redirect_stdout(io.StringIO())

    def setup(n, a1=0.0, a2=1.0):
        mesh = RectMesh(n, a1, a2)
        V = fd.FunctionSpace(mesh)
        return mesh, V, mesh.nodes, mesh.dof_neighbors(), np.array(mesh.vertex_to_dof)

    # ---- pin: the shipped chemotaxis trajectory through the reference's own loop ----
    d = os.path.join(REFERENCE_DIR, "Chtxs_data_dx0.025_dt0.001")
    gm = np.genfromtxt(os.path.join(d, "chtxs_m_t0.01.csv"), delimiter=",").reshape(11, 1681)
    gf = np.genfromtxt(os.path.join(d, "chtxs_f_t0.01.csv"), delimiter=",").reshape(11, 1681)
    mesh, V, nodes, nb, v2d = setup(40)
    ns, dt = 10, 1e-3
    m0, f0 = hp.chtxs_sys_IC(0.0, 1.0, 0.025, nodes, v2d)
    var1 = np.zeros((ns + 1) * nodes); var1[:nodes] = m0
    var2 = np.zeros((ns + 1) * nodes); var2[:nodes] = f0
    with quiet:
        var1, var2 = hp.solve_chtxs_system(np.zeros((ns + 1) * nodes), var1, var2, V, nodes, ns, dt, nb,
                                           control_fun=fd.Constant(100), rescaling=1)
    em = max(np.linalg.norm(var1.reshape(ns + 1, -1)[k] - gm[k]) / np.linalg.norm(gm[k]) for k in range(ns + 1))
    ef = max(np.linalg.norm(var2.reshape(ns + 1, -1)[k] - gf[k]) / np.linalg.norm(gf[k]) for k in range(ns + 1))
    print(f"reference solve_chtxs_system on fake dolfin vs shipped trajectory: rel-L2 m {em:.2e}, f {ef:.2e}")
    assert em < 1e-13 and ef < 1e-13

    out = {}
    rng = np.random.default_rng(11)
    n, ns = 10, 4
    mesh, V, nodes, nb, v2d = setup(n)
    xy = mesh.dof_xy
    L = (ns + 1) * nodes
    out["n"] = np.array([n]); out["ns"] = np.array([ns])

    # ---- Schnakenberg ----
    dt = 5e-4
    u0, v0 = hp.schnak_sys_IC(0.0, 1.0, 1.0 / n, nodes, v2d)
    c = 0.1 + 0.05 * rng.random(L)
    uk = np.zeros(L); uk[:nodes] = u0
    vk = np.zeros(L); vk[:nodes] = v0
    with quiet:
        uk, vk = hp.solve_schnak_system(c, uk, vk, V, nodes, ns, dt, nb)
    uhat_T, vhat_T = uk[ns * nodes:] + 0.1 * rng.random(nodes), vk[ns * nodes:] + 0.1 * rng.random(nodes)
    pk, qk = np.zeros(L), np.zeros(L)
    with quiet:
        pk, qk = hp.solve_adjoint_schnak_system(uk, vk, uhat_T, vhat_T, pk, qk, ns * dt, V, nodes, ns, dt, nb)
    out.update(schnak_dt=np.array([dt]), schnak_c=c, schnak_u0=u0, schnak_v0=v0, schnak_u=uk.copy(), schnak_v=vk.copy(),
               schnak_uhat=uhat_T, schnak_vhat=vhat_T, schnak_p=pk.copy(), schnak_q=qk.copy())

    # ---- nonlinear advection-reaction ----
    dt = 1e-3
    u0 = hp.nonlinear_equation_IC(0.0, 1.0, 1.0 / n, nodes, v2d)
    c = rng.random(L) - 0.5
    uk = np.zeros(L); uk[:nodes] = u0
    with quiet:
        uk, _ = hp.solve_nonlinear_equation(c, uk, None, V, nodes, ns, dt, nb)
    uhat_T = uk[ns * nodes:] + 0.05 * rng.random(nodes)
    pk = np.zeros(L)
    with quiet:
        pk = hp.solve_adjoint_nonlinear_equation(uk, uhat_T, pk, ns * dt, V, nodes, ns, dt, nb)
    out.update(nonlin_dt=np.array([dt]), nonlin_c=c, nonlin_u0=u0, nonlin_u=uk.copy(), nonlin_uhat=uhat_T, nonlin_p=pk.copy())

    # ---- line search with the reference's own solver as callback (one-species, final time) ----
    beta = 0.1
    M = hp.assemble_sparse(fd.TrialFunction(V) * fd.TestFunction(V) * fd.dx)
    with quiet:
        cost0 = hp.cost_functional(uk, uhat_T, c, ns, dt, M, beta, optim="finaltime")
    dk = -(beta * c - pk)
    var1 = uk.copy()
    with quiet:
        u_new, c_new, its = hp.armijo_line_search_ref(var1, c, dk, uhat_T, ns, dt, -0.4, 0.4, beta, cost0, nodes, "finaltime", V,
                                                      nonlinear_solver=hp.solve_nonlinear_equation, dof_neighbors=nb)
    out.update(armijo_beta=np.array([beta]), armijo_cost0=np.array([cost0]), armijo_d=dk, armijo_u=np.array(u_new),
               armijo_c=np.array(c_new), armijo_its=np.array([its]), armijo_bounds=np.array([-0.4, 0.4]))

    # ---- chemotaxis: state and both adjoints ----
    dt = 1e-3
    m0, f0 = hp.chtxs_sys_IC(0.0, 1.0, 1.0 / n, nodes, v2d)
    c = 50.0 + 20.0 * rng.random(L)
    mk = np.zeros(L); mk[:nodes] = m0
    fk = np.zeros(L); fk[:nodes] = f0
    with quiet:
        mk, fk = hp.solve_chtxs_system(c, mk, fk, V, nodes, ns, dt, nb)
    mhat, fhat = mk + 0.05 * rng.random(L), fk + 0.05 * rng.random(L)
    pk, qk = np.zeros(L), np.zeros(L)
    with quiet:
        pk, qk = hp.solve_adjoint_chtxs_system(mk, fk, mhat, fhat, pk, qk, c, ns * dt, V, nodes, ns, dt, nb, "alltime")
    out.update(chtxs_dt=np.array([dt]), chtxs_c=c, chtxs_m0=m0, chtxs_f0=f0, chtxs_m=mk.copy(), chtxs_f=fk.copy(),
               chtxs_mhat=mhat, chtxs_fhat=fhat, chtxs_p_at=pk.copy(), chtxs_q_at=qk.copy())
    pk, qk = np.zeros(L), np.zeros(L)
    with quiet:
        pk, qk = hp.solve_adjoint_chtxs_system(mk, fk, mhat[ns * nodes:], fhat[ns * nodes:], pk, qk, c, ns * dt, V, nodes, ns,
                                               dt, nb, "finaltime")
    out.update(chtxs_p_ft=pk.copy(), chtxs_q_ft=qk.copy())
    np.savez_compressed(os.path.join(HERE, "ref_loops.npz"), **out)
    print("ref_loops.npz written:", len(out), "arrays")


def ref_script_cfg2(n=10, dt=2e-3, sample=1, out_name="ref_cfg2.npz"):
    """BASELINE config 2 (n = 80, dt = 0.001 are the script's own mesh and time step; `sample` > 1 stores every sample-th DoF of
    each time level plus the norms of the full fields): the state / adjoint / gradient loops of advection_solidbody_FCT_PDECO_alltime.py (:206-275, the shape
    of the 4096^2 benchmark) -- the script's own source lines, compiled as they stand and executed in a namespace that holds
    the reference's helpers.py (on oracle/fake_dolfin.py), the legacy FCT_alg of old_helpers.py (its source, compiled as it
    stands) and the script's set-up variables on a small mesh.  Pins the drift-control operators Adrift1 / Adrift2, the
    adjoint right-hand side and the gradient ChebSI solve on the reference's code."""
    import contextlib
    import io
    from oracle import fake_dolfin as fd
    from oracle.ref_loader import load_reference_helpers_on_fake_dolfin
    hp = load_reference_helpers_on_fake_dolfin()
    ns_ = dict(vars(hp))
    old = open(os.path.join(REFERENCE_DIR, "old_helpers.py")).read().splitlines()
    start = next(i for i, l in enumerate(old) if l.startswith("def FCT_alg("))
    exec(compile("\n".join(old[start:]), "old_helpers.py:FCT_alg", "exec"), ns_)
    script = open(os.path.join(REFERENCE_DIR, "advection_solidbody_FCT_PDECO_alltime.py")).read().splitlines()
    i0 = next(i for i, l in enumerate(script) if "print('Solving state equation...')" in l)
    i1 = next(i for i, l in enumerate(script) if "4. step size control" in l) - 1          # the banner line above it
    body = "\n".join(l[4:] if l.startswith("    ") else l for l in script[i0:i1])           # the loops live inside `while`
    num_steps, beta, eps = 3, 0.01, 0
    mesh = RectMesh(n, -1.0, 1.0)
    V = fd.FunctionSpace(mesh)
    nodes = mesh.nodes
    u, v = fd.TrialFunction(V), fd.TestFunction(V)
    rng = np.random.default_rng(21)
    xy = mesh.dof_xy
    u0 = np.exp(-20 * ((xy[:, 0] + 2 / 3) ** 2 + 5 * (xy[:, 1] + 5 / 6) ** 2))
    vec_length = (num_steps + 1) * nodes
    uk = np.zeros(vec_length); uk[:nodes] = u0
    ck = 0.5 + rng.random(vec_length)
    uhat_all = np.tile(u0, num_steps + 1) * (1.0 + 0.1 * rng.random(vec_length))
    M = hp.assemble_sparse_lil(u * v * fd.dx)
    Ad = hp.assemble_sparse(fd.dot(fd.grad(u), fd.grad(v)) * fd.dx)
    ns_.update(np=np, V=V, u=u, v=v, dx=fd.dx, dot=fd.dot, grad=fd.grad, assemble=fd.assemble, nodes=nodes, num_steps=num_steps,
               dt=dt, T=num_steps * dt, beta=beta, eps=eps, drift=fd.Constant(('1', '1')), M=M, M_diag=M.diagonal(),
               M_Lump=hp.row_lump(M, nodes), Ad=Ad, Arot=0 * Ad, dof_neighbors=mesh.dof_neighbors(), vec_length=vec_length,
               uk=uk, ck=ck, uhat_all=uhat_all, pk=np.zeros(vec_length), dk=np.zeros(vec_length))
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(body, "advection_solidbody_FCT_PDECO_alltime.py:loops", "exec"), ns_)
    out = dict(n=np.array([n]), ns=np.array([num_steps]), dt=np.array([dt]), beta=np.array([beta]), u0=u0, c=ck, uhat=uhat_all,
               u=ns_["uk"].copy(), p=ns_["pk"].copy(), d=ns_["dk"].copy())
    print(out_name, "|u|, |p|, |d| =", *(float(np.linalg.norm(out[k])) for k in "upd"))
    if sample > 1:
        # u0 is a closed form of the DoF coordinates and c, uhat come from default_rng(21) (c first): the test regenerates them
        full = {k: out[k] for k in "upd"}
        out = dict(n=out["n"], ns=out["ns"], dt=out["dt"], beta=out["beta"], sample=np.array([sample]), seed=np.array([21]))
        for k, a in full.items():
            out[k + "_s"] = a.reshape(num_steps + 1, nodes)[:, ::sample].copy()
            out[k + "_norm"] = np.linalg.norm(a.reshape(num_steps + 1, nodes), axis=1)
    np.savez_compressed(os.path.join(HERE, out_name), **out)
    print(out_name, "written")


def _script_namespace(hp, fd):
    """helpers.py namespace + the legacy definitions of old_helpers.py (FCT_alg, rhs_chtx_*; compiled from their source as
    it stands) + dolfin's names on the stand-in: what `from dolfin import *; from helpers import *` gives the legacy scripts"""
    import scipy.sparse.linalg as spl
    ns_ = dict(vars(hp))
    for name in ("dx", "dot", "grad", "div", "exp", "assemble", "project", "TrialFunction", "TestFunction", "Function",
                 "Constant", "Expression", "VectorFunctionSpace"):
        ns_[name] = getattr(fd, name)
    old = open(os.path.join(REFERENCE_DIR, "old_helpers.py")).read().splitlines()
    start = next(i for i, l in enumerate(old) if l.startswith("def rhs_chtx_m("))
    exec(compile("\n".join(old[start:]), "old_helpers.py:87-204", "exec"), ns_)
    ns_.update(np=np, spsolve=spl.spsolve)
    return ns_


def _loop_source(script_name, first_marker, last_marker):
    """source lines of a legacy script from the line containing first_marker up to (not including) the banner line above the
    line containing last_marker, de-indented by the 4 spaces of the enclosing `while`"""
    script = open(os.path.join(REFERENCE_DIR, script_name)).read().splitlines()
    i0 = next(i for i, l in enumerate(script) if first_marker in l)
    i1 = next(i for i, l in enumerate(script) if last_marker in l and i > i0) - 1
    return "\n".join(l[4:] if l.startswith("    ") else l for l in script[i0:i1])


def ref_script_cfg4(n=10, name="ref_cfg4.npz"):
    """BASELINE config 4 (n = 50 is the script's own mesh, deltax = 0.02, with its own dt = 0.002 and T = 3 dt): the state and
    adjoint loops of Schnak_FCT_PDECO.py (:190-279) -- the script's own source lines, with
    the time-dependent Expression wind, project(wind, W) and the div(w_h u) w form -- executed with the reference's helpers.py
    and legacy FCT_alg on oracle/fake_dolfin.py."""
    import contextlib
    import io
    from oracle import fake_dolfin as fd
    from oracle.ref_loader import load_reference_helpers_on_fake_dolfin
    hp = load_reference_helpers_on_fake_dolfin()
    ns_ = _script_namespace(hp, fd)
    body = _loop_source("Schnak_FCT_PDECO.py", "print('Solving state equations...')", "3. choose the descent direction")
    num_steps, dt = 3, 2e-3
    mesh = RectMesh(n, 0.0, 1.0)
    V = fd.FunctionSpace(mesh)
    W = fd.VectorFunctionSpace(mesh)
    nodes = mesh.nodes
    u, w = fd.TrialFunction(V), fd.TestFunction(V)
    rng = np.random.default_rng(31)
    v2d = np.array(mesh.vertex_to_dof)
    u0, v0 = hp.schnak_sys_IC(0.0, 1.0, 1.0 / n, nodes, v2d)
    vec_length = (num_steps + 1) * nodes
    uk = np.zeros(vec_length); vk = np.zeros(vec_length)
    uk[:nodes], vk[:nodes] = u0, v0
    ck = 0.1 * np.ones(vec_length) + 0.01 * (rng.random(vec_length) - 0.5)
    uhat_T, vhat_T = u0 * (1 + 0.05 * rng.random(nodes)), v0 * (1 + 0.05 * rng.random(nodes))
    wind = fd.Expression(('-(x[1]-0.5)*sin(2*pi*t)', '(x[0]-0.5)*sin(2*pi*t)'), degree=4, pi=np.pi, t=0)
    M = hp.assemble_sparse_lil(u * w * fd.dx)
    ns_.update(V=V, W=W, u=u, w=w, nodes=nodes, num_steps=num_steps, dt=dt, T=num_steps * dt, vec_length=vec_length, wind=wind,
               Du=1 / 100, Dv=8.6676, c_a=0.1, c_b=0.9, gamma=230.82, omega1=100, omega2=0.6, M=M, M_Lump=hp.row_lump(M, nodes),
               Ad=hp.assemble_sparse(fd.dot(fd.grad(u), fd.grad(w)) * fd.dx), dof_neighbors=mesh.dof_neighbors(), uk=uk, vk=vk,
               ck=ck, uhat_T=uhat_T, vhat_T=vhat_T)
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(body, "Schnak_FCT_PDECO.py:loops", "exec"), ns_)
    out = dict(n=np.array([n]), ns=np.array([num_steps]), dt=np.array([dt]), u0=u0, v0=v0, c=ck, uhat_T=uhat_T, vhat_T=vhat_T,
               u=ns_["uk"].copy(), v=ns_["vk"].copy(), p=ns_["pk"].copy(), q=ns_["qk"].copy())
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "written; norms", *(float(np.linalg.norm(out[k])) for k in "uvpq"))


def ref_script_cfg3(n=8, a2=4.0, dt=0.02, num_steps=3, sample=1, out_name="ref_cfg3.npz"):
    """BASELINE config 3 (n = 128, a2 = 16, dt = 0.1: the script's own mesh and time step; `sample` > 1 stores every sample-th
    DoF of each time level plus the norms of the full fields, so that the 129^2 fixture stays small): the state and adjoint
    loops of chemotaxis_mimura_FCT_PGD.py (:157-225) -- the script's own source
    lines with the reference's mimura_data_helpers.py and the legacy form builders / FCT_alg of old_helpers.py -- executed on
    oracle/fake_dolfin.py."""
    import contextlib
    import importlib.util
    import io
    import types
    from oracle import fake_dolfin as fd
    from oracle.ref_loader import load_reference_helpers_on_fake_dolfin
    hp = load_reference_helpers_on_fake_dolfin()
    ns_ = _script_namespace(hp, fd)
    # the reference's mimura_data_helpers.py, unmodified (its image-library imports are stubbed: they serve data loading only)
    saved = {k: sys.modules.get(k) for k in ("dolfin", "helpers", "PIL", "PIL.Image", "matplotlib", "matplotlib.image")}
    sys.modules["dolfin"] = fd.make_module()
    sys.modules["helpers"] = hp
    for name in ("PIL", "PIL.Image", "matplotlib", "matplotlib.image"):
        if saved[name] is None:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["PIL"].Image = sys.modules["PIL.Image"]
    sys.modules["matplotlib"].image = sys.modules["matplotlib.image"]
    try:
        spec = importlib.util.spec_from_file_location("_fct_reference_mimura", os.path.join(REFERENCE_DIR, "mimura_data_helpers.py"))
        mdh = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mdh)
    finally:
        for k, v_ in saved.items():
            if v_ is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v_
    body = _loop_source("chemotaxis_mimura_FCT_PGD.py", "print('Solving state equations...')", "3. choose the descent direction")
    a1 = 0.0
    delta, Dm, Df, chi = 32, 0.0625, 1, 8.5
    mesh = RectMesh(n, a1, a2)
    V = fd.FunctionSpace(mesh)
    nodes = mesh.nodes
    u, v = fd.TrialFunction(V), fd.TestFunction(V)
    rng = np.random.default_rng(41)
    v2d = np.array(mesh.vertex_to_dof)
    m0 = hp.reorder_vector_to_dof(mdh.m_initial_condition(a1, a2, (a2 - a1) / n).reshape(nodes), 1, nodes, v2d)
    f0 = 1 / 32 * np.ones(nodes)
    vec_length = (num_steps + 1) * nodes
    mk = np.zeros(vec_length); fk = np.zeros(vec_length)
    mk[:nodes], fk[:nodes] = m0, f0
    ck = 0.5 + rng.random(vec_length)
    mhat_T, fhat_T = m0 * (1 + 0.05 * rng.random(nodes)), f0 * (1 + 0.05 * rng.random(nodes))
    M = hp.assemble_sparse_lil(u * v * fd.dx)
    Ad = hp.assemble_sparse(fd.dot(fd.grad(u), fd.grad(v)) * fd.dx)
    ns_.update(V=V, u=u, v=v, nodes=nodes, num_steps=num_steps, dt=dt, T=num_steps * dt, vec_length=vec_length, Dm=Dm, Df=Df,
               chi=chi, delta=delta, M=M, M_Lump=hp.row_lump(M, nodes), Ad=Ad, Mat_fq=M + dt * (Df * Ad + delta * M),
               dof_neighbors=mesh.dof_neighbors(), mk=mk, fk=fk, ck=ck, mhat_T=mhat_T, fhat_T=fhat_T, mimura_data_helpers=mdh)
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(body, "chemotaxis_mimura_FCT_PGD.py:loops", "exec"), ns_)
    out = dict(n=np.array([n]), ns=np.array([num_steps]), dt=np.array([dt]), box=np.array([a1, a2]), m0=m0, f0=f0, c=ck,
               mhat_T=mhat_T, fhat_T=fhat_T, m=ns_["mk"].copy(), f=ns_["fk"].copy(), p=ns_["pk"].copy(), q=ns_["qk"].copy(),
               params=np.array([delta, Dm, Df, chi], dtype=np.float64))
    print(out_name, "norms", *(float(np.linalg.norm(out[k])) for k in "mfpq"))
    if sample > 1:
        # inputs other than m0 are regenerated by the test from the same seeded generator (default_rng(41): c, then the two
        # targets); outputs are stored at every sample-th DoF of each time level, with the 2-norms of the full fields
        full = {k: out[k] for k in "mfpq"}
        out = dict(n=out["n"], ns=out["ns"], dt=out["dt"], box=out["box"], params=out["params"], sample=np.array([sample]),
                   seed=np.array([41]), m0_s=m0[::sample])
        for k, a in full.items():
            out[k + "_s"] = a.reshape(num_steps + 1, nodes)[:, ::sample].copy()
            out[k + "_norm"] = np.linalg.norm(a.reshape(num_steps + 1, nodes), axis=1)
    np.savez_compressed(os.path.join(HERE, out_name), **out)
    print(out_name, "written")


def ref_script_refactored_nonlinear():
    """the projected-gradient loop of nonlinear_FCT_PDECO_refactored.py (:105-210) -- the script's own source
    lines: initial state / adjoint solve, dk = -(beta ck - pk), armijo_line_search_ref with hp.solve_nonlinear_equation as
    the callback, adjoint solve, cost_functional / rel_err stopping test, the fail / restart bookkeeping -- executed with
    the reference's helpers.py on oracle/fake_dolfin.py for two iterations on a 10 x 10 mesh (the plotting calls of the
    script are replaced by no-ops; the target is synthetic, the script reads its own from a CSV)."""
    import contextlib
    import io
    import time
    import types
    from oracle import fake_dolfin as fd
    from oracle.ref_loader import load_reference_helpers_on_fake_dolfin
    hp = load_reference_helpers_on_fake_dolfin()
    script = open(os.path.join(REFERENCE_DIR, "nonlinear_FCT_PDECO_refactored.py")).read().splitlines()
    i0 = next(i for i, l in enumerate(script) if l.startswith("vec_length = (num_steps + 1) * nodes"))
    i1 = next(i for i, l in enumerate(script) if "Save results" in l)
    body = "\n".join(script[i0:i1])
    hp_ns = types.SimpleNamespace(**{k: v for k, v in vars(hp).items() if not k.startswith("__")})
    hp_ns.plot_nonlinear_solution = lambda *a, **k: None
    hp_ns.plot_progress = lambda *a, **k: None
    n, num_steps, dt = 10, 4, 1e-3
    a1, a2 = 0, 1
    mesh = RectMesh(n, float(a1), float(a2))
    V = fd.FunctionSpace(mesh)
    nodes = mesh.nodes
    v2d = np.array(mesh.vertex_to_dof)
    u, v = fd.TrialFunction(V), fd.TestFunction(V)
    M = hp.assemble_sparse(u * v * fd.dx)
    u0 = hp.nonlinear_equation_IC(a1, a2, 1.0 / n, nodes, v2d)
    rng = np.random.default_rng(51)
    uhat_T = u0 * (0.6 + 0.2 * rng.random(nodes))
    ns_ = dict(np=np, hp=hp_ns, time=time, a1=a1, a2=a2, dt=dt, T=num_steps * dt, T_data=num_steps * dt, num_steps=num_steps,
               nodes=nodes, V=V, M=M, u0=u0, uhat_T=uhat_T, uhat_T_re=None, vertex_to_dof=v2d, dof_neighbors=mesh.dof_neighbors(),
               beta=1e-1, c_lower=-1, c_upper=1, optim="finaltime", tol=1e-4, max_iter_armijo=5, max_iter_GD=2,
               produce_plots=False, out_folder=None)
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(body, "nonlinear_FCT_PDECO_refactored.py:105-210", "exec"), ns_)
    out = dict(n=np.array([n]), ns=np.array([num_steps]), dt=np.array([dt]), u0=u0, uhat_T=uhat_T, u=ns_["uk"].copy(),
               p=ns_["pk"].copy(), c=ns_["ck"].copy(), d=ns_["dk"].copy(), cost=np.array(ns_["cost_fun_vals"]),
               armijo_its=np.array(ns_["armijo_its"]), it=np.array([ns_["it"]]), stop_crit=np.array([ns_["stop_crit"]]))
    np.savez_compressed(os.path.join(HERE, "ref_pgd_nonlinear.npz"), **out)
    print("ref_pgd_nonlinear.npz written; cost", out["cost"], "armijo its", out["armijo_its"], "it", out["it"])


def _exec_refactored_pgd(script_name, ns_):
    """executes lines `vec_length = ...` .. `Save results` of a *_refactored.py script (its projected-gradient loop) in ns_,
    with the reference's helpers.py on oracle/fake_dolfin.py as `hp` (plotting entry points replaced by no-ops)"""
    import contextlib
    import io
    import time
    import types
    from oracle.ref_loader import load_reference_helpers_on_fake_dolfin
    hp = load_reference_helpers_on_fake_dolfin()
    script = open(os.path.join(REFERENCE_DIR, script_name)).read().splitlines()
    i0 = next(i for i, l in enumerate(script) if l.startswith("vec_length = (num_steps + 1) * nodes"))
    i1 = next(i for i, l in enumerate(script) if "Save results" in l)
    hp_ns = types.SimpleNamespace(**{k: v for k, v in vars(hp).items() if not k.startswith("__")})
    for name in ("plot_nonlinear_solution", "plot_two_var_solution", "plot_progress"):
        setattr(hp_ns, name, lambda *a, **k: None)
    ns_.update(np=np, hp=hp_ns, time=time, produce_plots=False, out_folder=None)
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile("\n".join(script[i0:i1]), script_name + ":pgd-loop", "exec"), ns_)
    return hp, ns_


def ref_script_refactored_two_species():
    """The projected-gradient loops of Schnak_FCT_PDECO_refactored.py (:122-246, final-time, dk = -(beta ck - gamma/r pk)) and
    chemotaxis_FCT_PDECO_AT_refactored.py (:122-257, all-time, dk = -(beta ck - qk uk / r), line search with gam / s0) -- the
    scripts' own source lines with the reference's helpers.py on oracle/fake_dolfin.py, two iterations on a 10 x 10 mesh,
    synthetic targets."""
    from oracle import fake_dolfin as fd
    from oracle.ref_loader import load_reference_helpers_on_fake_dolfin
    hp = load_reference_helpers_on_fake_dolfin()
    n, a1, a2 = 10, 0, 1
    mesh = RectMesh(n, float(a1), float(a2))
    V = fd.FunctionSpace(mesh)
    nodes = mesh.nodes
    v2d = np.array(mesh.vertex_to_dof)
    M = hp.assemble_sparse(fd.TrialFunction(V) * fd.TestFunction(V) * fd.dx)
    base = dict(a1=a1, a2=a2, nodes=nodes, V=V, M=M, mesh=mesh, dx=1.0 / n, vertex_to_dof=v2d, dof_neighbors=mesh.dof_neighbors())
    # ---- Schnakenberg, final time ----
    rng = np.random.default_rng(61)
    num_steps, dt = 4, 5e-4
    u0, v0 = hp.schnak_sys_IC(a1, a2, 1.0 / n, nodes, v2d)
    Du, Dv, true_control, c_b, gamma, omega1, omega2, wind = hp.get_schnak_sys_params()
    uhat_T, vhat_T = u0 * (1 + 0.1 * rng.random(nodes)), v0 * (1 + 0.1 * rng.random(nodes))
    _, ns_ = _exec_refactored_pgd("Schnak_FCT_PDECO_refactored.py", dict(
        base, dt=dt, T=num_steps * dt, T_data=num_steps * dt, num_steps=num_steps, u0=u0, v0=v0, uhat_T=uhat_T, vhat_T=vhat_T,
        uhat_T_re=None, vhat_T_re=None, gamma=gamma, rescaling=1, beta=1e-1, c_lower=0, c_upper=10, optim="finaltime", tol=1e-3,
        max_iter_armijo=10, max_iter_GD=2))
    out = dict(n=np.array([n]), s_ns=np.array([num_steps]), s_dt=np.array([dt]), s_u0=u0, s_v0=v0, s_uhat=uhat_T, s_vhat=vhat_T,
               s_u=ns_["uk"].copy(), s_v=ns_["vk"].copy(), s_p=ns_["pk"].copy(), s_q=ns_["qk"].copy(), s_c=ns_["ck"].copy(),
               s_cost=np.array(ns_["cost_fun_vals"]), s_its=np.array(ns_["armijo_its"]), s_it=np.array([ns_["it"]]))
    print("Schnak: cost", out["s_cost"], "armijo its", out["s_its"], "it", out["s_it"])
    # ---- chemotaxis, all time ----
    num_steps, dt = 4, 5e-4
    L = (num_steps + 1) * nodes
    u0, v0 = hp.chtxs_sys_IC(a1, a2, 1.0 / n, nodes, v2d)
    uhat = np.tile(u0, num_steps + 1) * (1 + 0.1 * rng.random(L))
    vhat = np.tile(v0, num_steps + 1) * (1 + 0.1 * rng.random(L))
    _, ns_ = _exec_refactored_pgd("chemotaxis_FCT_PDECO_AT_refactored.py", dict(
        base, dt=dt, T=num_steps * dt, num_steps=num_steps, u0=u0, v0=v0, uhat=uhat, vhat=vhat, uhat_re=None, vhat_re=None,
        rescaling=1 / 10, beta=1e-3, c_lower=0, c_upper=20, optim="alltime", tol=1e-4, max_iter_armijo=20, max_iter_GD=2,
        armijo_gamma=1e-5, armijo_s0=2))
    out.update(c_ns=np.array([num_steps]), c_dt=np.array([dt]), c_u0=u0, c_v0=v0, c_uhat=uhat, c_vhat=vhat,
               c_u=ns_["uk"].copy(), c_v=ns_["vk"].copy(), c_p=ns_["pk"].copy(), c_q=ns_["qk"].copy(), c_c=ns_["ck"].copy(),
               c_cost=np.array(ns_["cost_fun_vals"]), c_its=np.array(ns_["armijo_its"]), c_it=np.array([ns_["it"]]))
    print("chemotaxis AT: cost", out["c_cost"], "armijo its", out["c_its"], "it", out["c_it"])
    np.savez_compressed(os.path.join(HERE, "ref_pgd_two_species.npz"), **out)
    print("ref_pgd_two_species.npz written")


def ref_script_full_sizes():
    """configs 2, 3 and 4 on the meshes BASELINE names: 81^2 DoF on [-1,1]^2 (three time levels, sampled), 129^2 DoF on
    [0,16]^2 with dt = 0.1 (two time levels, sampled), 51^2 DoF with dt = 0.002 and the script's own three time levels"""
    ref_script_cfg2(n=80, dt=0.001, sample=4, out_name="ref_cfg2_M.npz")
    ref_script_cfg4(n=50, name="ref_cfg4_K.npz")
    ref_script_cfg3(n=128, a2=16.0, dt=0.1, num_steps=2, sample=8, out_name="ref_cfg3_C.npz")


if __name__ == "__main__":
    which = sys.argv[1:] or ["ref_data", "ref_fct_cases", "ref_legacy", "ref_armijo", "ref_loops", "ref_script_cfg2",
                             "ref_script_cfg3", "ref_script_cfg4", "ref_script_full_sizes", "ref_script_refactored_nonlinear",
                             "ref_script_refactored_two_species"]
    for name in which:
        globals()[name]()
