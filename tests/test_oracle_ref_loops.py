"""CPU tests: the oracle's restatements of the reference's time loops (oracle/pdeco_systems.py, oracle/pdeco_numpy.py)
against outputs of the reference's OWN loops -- helpers.py run unmodified on oracle/fake_dolfin.py by
tests/golden/make_golden.py::ref_loops (ref_loops.npz) -- and the form evaluator of fake_dolfin against the
hand-written element tensors of oracle/p1assembly.py.  With these the loop bodies of solve_schnak_system,
solve_adjoint_schnak_system, solve_nonlinear_equation, solve_adjoint_nonlinear_equation and
solve_adjoint_chtxs_system (helpers.py:511-698, 881-1038, 1387-1581) are pinned on reference outputs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2
from oracle import fake_dolfin as fd
from oracle import pdeco_systems as osys
from oracle.fct_numpy import cost_functional as o_cost
from oracle.p1assembly import P1Assembler
from oracle.p1mesh import RectMesh
from oracle.ref_loader import reference_available


@pytest.fixture(scope="module")
def loops():
    return dict(np.load(os.path.join(GOLDEN, "ref_loops.npz")))


def test_fake_dolfin_forms_match_element_tensors():
    """tree evaluator (UFL degree estimation + FIAT rules) vs the specialised element tensors, form by form"""
    mesh = RectMesh(7, -1.0, 1.0)
    V = fd.FunctionSpace(mesh)
    asm = P1Assembler(mesh)
    u, v = fd.TrialFunction(V), fd.TestFunction(V)
    rng = np.random.default_rng(0)
    f, g = rng.random(mesh.nodes), rng.random(mesh.nodes)
    ff, gg = fd.Function(V), fd.Function(V)
    ff.vector().set_local(f); gg.vector().set_local(g)
    mat = lambda form: fd.assemble(form).getValuesCSR()[2]
    wind = fd.Expression(("1 * (x[1] - 0.5) * x[0] * (1 - x[0])", "-1 * (x[0] - 0.5) * x[1] * (1 - x[1])"), degree=4, t=0)
    eta = 0.5
    checks = [
        (mat(u * v * fd.dx), asm.mass()),
        (mat(fd.dot(fd.grad(u), fd.grad(v)) * fd.dx), asm.stiffness()),
        (mat(fd.dot(wind, fd.grad(v)) * u * fd.dx), asm.conv_conservative(osys.schnak_wind, 5)),
        (mat(fd.dot(wind, fd.grad(u)) * v * fd.dx), asm.conv_nonconservative(osys.schnak_wind, 5)),
        (mat(ff ** 2 * u * v * fd.dx), asm.mass_p1_product(f, f)),
        (mat(ff * gg * u * v * fd.dx), asm.mass_p1_product(f, g)),
        (mat(fd.exp(-eta * ff) * fd.dot(fd.grad(gg), fd.grad(v)) * u * fd.dx),
         asm.chemotaxis_conv(g, lambda phi, xy: np.exp(-eta * asm.at_quad(f, phi)), 4)),
        (mat((1 - eta * ff) * fd.exp(-eta * ff) * fd.dot(fd.grad(u), fd.grad(gg)) * v * fd.dx),
         asm.chemotaxis_adjoint_mat(f, g, eta, 5)),
        (fd.assemble((2.0 * ff + 3.0 * (ff ** 2 * gg)) * v * fd.dx), 2 * asm.load_p1_product(f) + 3 * asm.load_p1_product(f, f, g)),
        (fd.assemble(ff * v * fd.dx + 0.1 * gg * ff / 0.5 * v * fd.dx), asm.load_p1_product(f) + 0.2 * asm.load_p1_product(g, f)),
        (fd.assemble(0.25 * ff * fd.exp(-eta * ff) * fd.dot(fd.grad(gg), fd.grad(v)) * fd.dx),
         asm.load_grad_pair(lambda phi, xy: 0.25 * asm.at_quad(f, phi) * np.exp(-eta * asm.at_quad(f, phi)), g, 4)),
        (fd.assemble(3.5 * v * fd.dx), asm.load_constant(3.5)),
    ]
    for k, (a, b) in enumerate(checks):
        assert np.abs(np.asarray(a) - b).max() <= 1e-14 * max(np.abs(b).max(), 1.0), k
    # quadrature degrees UFL would estimate for the rule-sensitive (exp) forms: SURVEY.md App. B.3
    assert (fd.exp(-eta * ff) * fd.dot(fd.grad(gg), fd.grad(v)) * u).degree() == 4
    assert ((1 - eta * ff) * fd.exp(-eta * ff) * fd.dot(fd.grad(u), fd.grad(gg)) * v).degree() == 5
    assert (0.25 * ff * fd.exp(-eta * ff) * fd.dot(fd.grad(gg), fd.grad(v))).degree() == 4


def test_schnak_restatement_vs_reference_loops(loops):
    g = loops
    n, ns, dt = int(g["n"][0]), int(g["ns"][0]), float(g["schnak_dt"][0])
    orc = osys.SchnakProblem(n, 0.0, 1.0)
    u0, v0 = orc.initial_condition()
    assert np.array_equal(u0, g["schnak_u0"]) and np.array_equal(v0, g["schnak_v0"])
    u, v = orc.state(g["schnak_c"], u0, v0, ns, dt, rescaling=1.0)
    assert rel_l2(u.ravel(), g["schnak_u"]) < 1e-12 and rel_l2(v.ravel(), g["schnak_v"]) < 1e-12
    ref_u, ref_v = g["schnak_u"].reshape(ns + 1, -1), g["schnak_v"].reshape(ns + 1, -1)
    p, q = orc.adjoint(ref_u, ref_v, g["schnak_uhat"], g["schnak_vhat"], ns, dt)
    assert rel_l2(p.ravel(), g["schnak_p"]) < 1e-12 and rel_l2(q.ravel(), g["schnak_q"]) < 1e-12


def test_nonlinear_restatement_and_armijo_vs_reference_loops(loops):
    g = loops
    n, ns, dt = int(g["n"][0]), int(g["ns"][0]), float(g["nonlin_dt"][0])
    orc = osys.NonlinearProblem(n, 0.0, 1.0)
    u0 = orc.initial_condition()
    assert np.array_equal(u0, g["nonlin_u0"])
    u = orc.state(g["nonlin_c"], u0, ns, dt)
    assert rel_l2(u.ravel(), g["nonlin_u"]) < 1e-12
    p = orc.adjoint(g["nonlin_u"].reshape(ns + 1, -1), g["nonlin_uhat"], ns, dt)
    assert rel_l2(p.ravel(), g["nonlin_p"]) < 1e-12
    # the reference's armijo_line_search_ref with its own solve_nonlinear_equation as callback
    beta, cost0 = float(g["armijo_beta"][0]), float(g["armijo_cost0"][0])
    assert abs(o_cost(orc.pat, g["nonlin_u"], g["nonlin_uhat"], g["nonlin_c"], ns, dt, orc.M, beta, "finaltime") / cost0 - 1) < 1e-12
    lo, hi = g["armijo_bounds"]
    solver = lambda ci: (orc.state(ci, u0, ns, dt).ravel(), None)
    v1, _, c_inc, k = osys.armijo_ref(orc, solver, g["nonlin_u"], g["nonlin_c"], g["armijo_d"], g["nonlin_uhat"], ns, dt, lo, hi,
                                      beta, cost0, "finaltime")
    assert k == int(g["armijo_its"][0]) and np.array_equal(c_inc, g["armijo_c"]) and rel_l2(v1, g["armijo_u"]) < 1e-12


def test_chemotaxis_restatement_vs_reference_loops(loops):
    g = loops
    n, ns, dt = int(g["n"][0]), int(g["ns"][0]), float(g["chtxs_dt"][0])
    orc = osys.ChemotaxisAdjoint(n, 0.0, 1.0)
    m, f = orc.forward(g["chtxs_c"], g["chtxs_m0"], g["chtxs_f0"], ns, dt)
    assert rel_l2(m.ravel(), g["chtxs_m"]) < 1e-12 and rel_l2(f.ravel(), g["chtxs_f"]) < 1e-12
    mr, fr = g["chtxs_m"].reshape(ns + 1, -1), g["chtxs_f"].reshape(ns + 1, -1)
    mh, fh = g["chtxs_mhat"].reshape(ns + 1, -1), g["chtxs_fhat"].reshape(ns + 1, -1)
    p, q = orc.adjoint(mr, fr, mh, fh, g["chtxs_c"], ns, dt, "alltime")
    assert rel_l2(p.ravel(), g["chtxs_p_at"]) < 1e-12 and rel_l2(q.ravel(), g["chtxs_q_at"]) < 1e-12
    p, q = orc.adjoint(mr, fr, mh[-1], fh[-1], g["chtxs_c"], ns, dt, "finaltime")
    assert rel_l2(p.ravel(), g["chtxs_p_ft"]) < 1e-12 and rel_l2(q.ravel(), g["chtxs_q_ft"]) < 1e-12


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container only)")
def test_reference_chtxs_loop_on_fake_dolfin_reproduces_shipped_data(ref_data):
    """the stand-in itself, end to end: the reference's unmodified solve_chtxs_system against its shipped trajectory"""
    import contextlib
    import io
    from oracle.ref_loader import load_reference_helpers_on_fake_dolfin
    hp = load_reference_helpers_on_fake_dolfin()
    mesh = RectMesh(40, 0.0, 1.0)
    V = fd.FunctionSpace(mesh)
    nodes, ns, dt = mesh.nodes, 3, 1e-3
    m0, f0 = hp.chtxs_sys_IC(0.0, 1.0, 0.025, nodes, np.array(mesh.vertex_to_dof))
    var1 = np.zeros((ns + 1) * nodes); var1[:nodes] = m0
    var2 = np.zeros((ns + 1) * nodes); var2[:nodes] = f0
    with contextlib.redirect_stdout(io.StringIO()):
        hp.solve_chtxs_system(np.zeros((ns + 1) * nodes), var1, var2, V, nodes, ns, dt, mesh.dof_neighbors(),
                              control_fun=fd.Constant(100), rescaling=1)
    for k in range(ns + 1):
        assert rel_l2(var1.reshape(ns + 1, -1)[k], ref_data["chtxs_m"][k]) < 1e-13
        assert rel_l2(var2.reshape(ns + 1, -1)[k], ref_data["chtxs_f"][k]) < 1e-13


@pytest.mark.parametrize("fixture", ["ref_cfg2.npz", "ref_cfg2_M.npz"])      # 11^2 DoF; M = the script's own 81^2 mesh, dt
def test_config2_drift_loops_vs_reference_script(fixture):
    """BASELINE config 2 / 5: oracle AdvectionDriftPDECO and its C/OpenMP twin against the loops of
    advection_solidbody_FCT_PDECO_alltime.py:206-275 executed from the script's own source"""
    from conftest import cfg2_inputs, golden_field_error
    from oracle import pdeco_numpy as drv
    from oracle.fct_c import CDriftProblem
    g = dict(np.load(os.path.join(GOLDEN, fixture)))
    n, ns, dt = int(g["n"][0]), int(g["ns"][0]), float(g["dt"][0])
    orc = drv.AdvectionDriftPDECO(n, -1.0, 1.0, beta=float(g["beta"][0]))
    u0, c, uhat = cfg2_inputs(g, orc.mesh.dof_xy)
    U = orc.state(c, u0, ns, dt)
    assert golden_field_error(g, "u", U) < 1e-13
    P = orc.adjoint(c, U, uhat.reshape(ns + 1, -1), ns, dt)
    assert golden_field_error(g, "p", P) < 1e-12
    assert golden_field_error(g, "d", orc.gradient(c, U, P, ns)) < 1e-12
    cprob = CDriftProblem(n, -1.0, 1.0)               # the CPU baseline / full-size parity checker of bench.py
    uc, _ = cprob.state(c, u0, ns, dt)
    assert golden_field_error(g, "u", uc) < 1e-12
