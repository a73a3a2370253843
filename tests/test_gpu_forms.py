"""GPU tests of the UFL-like front end (fem-fct-pdeco_b200/forms.py): the reference's `assemble_sparse(form)` /
`assemble(form)` call sites, written as in the reference scripts, against the oracle's dolfin restatement."""
import numpy as np
import pytest

from conftest import rel_l2
from fem_fct_pdeco_b200 import helpers as hp
from fem_fct_pdeco_b200.forms import (Constant, Expression, TestFunction, TrialFunction, assemble, assemble_sparse,
                                      assemble_sparse_lil, dot, dx, exp, grad, vec_to_function)
from fem_fct_pdeco_b200.mesh import FunctionSpaceP1, RectMeshP1, vertex_to_dof_map
from oracle import pdeco_numpy as drv
from oracle import pdeco_systems as osys
from oracle.fct_numpy import Pattern
from oracle.p1assembly import P1Assembler
from oracle.p1mesh import RectMesh

pytestmark = pytest.mark.gpu


def _relmax(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


def test_form_catalogue():
    n = 10
    mesh = RectMeshP1(n, -1.0, 1.0)
    V = FunctionSpaceP1(mesh)
    om = RectMesh(n, -1.0, 1.0)
    asm, pat = P1Assembler(om), Pattern(*om.pattern())
    rng = np.random.default_rng(8)
    f = [1.0 + rng.random(V.dim()) for _ in range(4)]
    F = [vec_to_function(x, V) for x in f]
    u, v = TrialFunction(V), TestFunction(V)
    csr = lambda a: assemble_sparse(a)
    M = csr(u * v * dx)
    assert np.array_equal(M.indptr, pat.rowptr) and np.array_equal(M.indices, pat.colidx)      # explicit zeros kept
    assert _relmax(M.data, asm.mass()) < 1e-13
    assert _relmax(assemble_sparse_lil(dot(grad(u), grad(v)) * dx).tocsr().toarray(), pat.csr(asm.stiffness()).toarray()) < 1e-13
    # weighted masses (helpers.py:591, 692, 953)
    assert _relmax(csr(F[0] ** 2 * u * v * dx).data, asm.mass_p1_product(f[0], f[0])) < 1e-13
    assert _relmax(csr(F[0] * F[1] * u * v * dx).data, asm.mass_p1_product(f[0], f[1])) < 1e-13
    assert _relmax(csr((4 - 2 * F[0]) * u * v * dx).data, 4 * asm.mass() - 2 * asm.mass_p1_product(f[0])) < 1e-13
    # winds: solid-body rotation + drift (advection_solidbody_FCT.py:77-80,106), Schnakenberg wind (helpers.py:506-508)
    omg = np.pi / 40
    wind = 1 / omg * Expression(('-x[1]', 'x[0]'), degree=4) + Expression(('2', '2'), degree=4)
    ref = asm.conv_conservative(lambda x, y: (-y / omg + 2, x / omg + 2), degree=5)
    assert _relmax(csr(dot(wind, grad(v)) * u * dx).data, ref) < 1e-13
    sw = Expression(("1 * (x[1] - 0.5) * x[0] * (1 - x[0])", "-1 * (x[0] - 0.5) * x[1] * (1 - x[1])"), degree=4, t=0)
    assert _relmax(csr(dot(sw, grad(v)) * u * dx).data, asm.conv_conservative(osys.schnak_wind, degree=5)) < 1e-13
    assert _relmax(csr(dot(sw, grad(u)) * v * dx).data, asm.conv_nonconservative(osys.schnak_wind, degree=5)) < 1e-13
    # drift-control forms (advection_solidbody_FCT_PDECO_alltime.py:222-223)
    drift = Constant(('1', '1'))
    assert _relmax(csr(dot(drift, grad(F[0])) * u * v * dx).data, asm.drift_mass(f[0], 1.0, 1.0)) < 1e-13
    assert _relmax(csr(dot(drift, grad(v)) * F[0] * u * dx).data, asm.drift_conv(f[0], 1.0, 1.0)) < 1e-13
    # chemotaxis forms (old_helpers.py:102; helpers.py:1350-1351, 1499-1500)
    eta = 0.5
    assert _relmax(csr(dot(grad(F[1]), grad(v)) * u * dx).data, asm.chemotaxis_conv(f[1])) < 1e-13
    ref = asm.chemotaxis_conv(f[1], lambda phi, xy: np.exp(-eta * asm.at_quad(f[0], phi)), degree=4)
    assert _relmax(csr(exp(-eta * F[0]) * dot(grad(F[1]), grad(v)) * u * dx).data, ref) < 1e-13
    ref = asm.chemotaxis_adjoint_mat(f[0], f[1], eta)
    assert _relmax(csr((1 - eta * F[0]) * exp(-eta * F[0]) * dot(grad(u), grad(F[1])) * v * dx).data, ref) < 1e-13
    # sums of forms and scalar factors: A_u = -eps*Ad + Adrift1 + Adrift2
    A_u = csr(-0.001 * dot(grad(u), grad(v)) * dx + dot(drift, grad(F[0])) * u * v * dx + dot(drift, grad(v)) * F[0] * u * dx)
    ref = -0.001 * asm.stiffness() + asm.drift_mass(f[0], 1.0, 1.0) + asm.drift_conv(f[0], 1.0, 1.0)
    assert _relmax(A_u.data, ref) < 1e-13
    # linear forms (helpers.py:584-585, 594, 1339-1340, 1531-1532; advection_solidbody_FCT_PDECO_alltime.py:255, 273)
    gamma, dt = 230.82, 1e-3
    b = np.asarray(assemble((gamma / 1 * F[2] + gamma * (F[0] ** 2 * F[1])) * v * dx))
    assert _relmax(b, gamma * asm.load_p1_product(f[2]) + gamma * asm.load_p1_product(f[0], f[0], f[1])) < 1e-13
    assert _relmax(assemble((gamma * 0.9) * v * dx), asm.load_constant(gamma * 0.9)) < 1e-13
    b = assemble(F[1] * v * dx + dt * Constant(100) * F[0] / 0.1 * v * dx)
    assert _relmax(b, asm.load_p1_product(f[1]) + dt * 100 / 0.1 * asm.load_p1_product(f[0])) < 1e-13
    assert _relmax(assemble((F[3] - F[0]) * v * dx), pat.csr(asm.mass()) @ (f[3] - f[0])) < 1e-13
    assert _relmax(assemble(F[2] * dot(drift, grad(F[0])) * v * dx), asm.load_drift_grad(f[2], f[0], 1.0, 1.0)) < 1e-13
    chi = 0.25
    ref = asm.load_grad_pair(lambda phi, xy: chi * asm.at_quad(f[0], phi) * np.exp(-eta * asm.at_quad(f[0], phi)), f[2], 4)
    assert _relmax(assemble(chi * F[0] * exp(-eta * F[0]) * dot(grad(F[2]), grad(v)) * dx), ref) < 1e-13
    with pytest.raises(NotImplementedError):
        assemble_sparse(F[0] * F[1] * F[2] * F[3] * u * v * dx)


def test_script_advection_solidbody_FCT():
    """the main loop of advection_solidbody_FCT.py:84-148 written against the drop-in names, 5 steps"""
    a1, a2, deltax = -1, 1, 0.1 / 2 / 2
    intervals_line = round((a2 - a1) / deltax)
    slit_width, om, eps, dt = 0.1, np.pi / 40, 0, 0.001
    mesh = RectMeshP1(intervals_line, a1, a2)
    V = FunctionSpaceP1(mesh)
    nodes = V.dim()
    u, v = TrialFunction(V), TestFunction(V)
    X = np.arange(a1, a2 + deltax, deltax)
    X, Y = np.meshgrid(X, X)
    R = np.sqrt(X ** 2 + (Y - 1 / 3) ** 2)
    out = ((R < 1 / 3) & ((np.abs(X) > slit_width) | (Y > 0.5))).astype(float)
    wind = 1 / om * Expression(('-x[1]', 'x[0]'), degree=4) + Expression(('2', '2'), degree=4)
    vertextodof = vertex_to_dof_map(V)
    dof_neighbors = hp.find_node_neighbours(mesh, nodes, vertextodof)
    M = hp.assemble_sparse_lil(u * v * dx)
    M_Lump = hp.row_lump(M, nodes)
    Ad = hp.assemble_sparse(dot(grad(u), grad(v)) * dx)
    A = hp.assemble_sparse(dot(wind, grad(v)) * u * dx)
    A_u = A - eps * Ad
    u0 = hp.reorder_vector_to_dof_time(out.reshape(nodes), 1, nodes, vertextodof)
    num_steps = 5
    uk = np.zeros((num_steps + 1) * nodes)
    uk[:nodes] = u0
    for i in range(1, num_steps + 1):
        start, end = i * nodes, (i + 1) * nodes
        uk_n = uk[start - nodes:start]
        uk[start:end] = hp.FCT_alg(A_u, np.zeros(nodes), uk_n, dt, nodes, M, M_Lump, dof_neighbors)
    orc = drv.SolidBodyProblem(intervals_line, a1, a2, slit_width=slit_width)
    ref = orc.forward(num_steps, dt, keep=True)
    for i in range(1, num_steps + 1):
        assert rel_l2(uk[i * nodes:(i + 1) * nodes], ref[i]) < 1e-12 * i


def test_mimura_legacy_form_builders():
    """config 3 (chemotaxis_mimura_FCT_PGD.py:175-225): the legacy form builders of old_helpers.py:87-111 and
    mimura_data_helpers.py:65-109, by name, against the oracle's assembler"""
    from fem_fct_pdeco_b200 import mimura_data_helpers as mdh
    n = 9
    mesh = RectMeshP1(n, 0.0, 16.0)
    V = FunctionSpaceP1(mesh)
    om = RectMesh(n, 0.0, 16.0)
    asm, pat = P1Assembler(om), Pattern(*om.pattern())
    rng = np.random.default_rng(12)
    f, m, c, q, p = [1.0 + rng.random(V.dim()) for _ in range(5)]
    F, Mf, Cf, Q, P = [vec_to_function(x, V) for x in (f, m, c, q, p)]
    u, v = TrialFunction(V), TestFunction(V)
    dt, chi, Dm = 0.1, 8.5, 0.0625
    Mmat = pat.csr(asm.mass())
    K = asm.stiffness()
    assert _relmax(mdh.rhs_chtx_m(Mf, v), asm.load_p1_product(m, m) - asm.load_p1_product(m, m, m)) < 1e-13
    assert _relmax(mdh.rhs_chtx_f(F, Mf, dt, v), Mmat @ f + dt * (Mmat @ m)) < 1e-13
    assert _relmax(hp.rhs_chtx_f(F, Mf, Cf, dt, v), Mmat @ f + dt * asm.load_p1_product(m, c)) < 1e-13
    assert _relmax(hp.rhs_chtx_p(Cf, Q, v), asm.load_p1_product(c, q)) < 1e-13
    Aa_exp = asm.chemotaxis_conv(f, lambda phi, xy: np.exp(-0.5 * asm.at_quad(m, phi)), degree=4)
    Am = mdh.mat_chtx_m(F, Mf, Dm, chi, u, v)
    assert _relmax(pat.embed(Am), -Dm * K + chi * Aa_exp) < 1e-13
    Ap = mdh.mat_chtx_p(F, Mf, Dm, chi, u, v)
    assert _relmax(pat.embed(Ap), -Dm * K - chi * asm.chemotaxis_conv(f)) < 1e-13
    # rhs_chtx_q: q v + dt chi (grad m . grad p) v, checked cell by cell
    xy, cells = om.dof_xy, om.cells
    load = np.zeros(V.dim())
    for T in cells:
        (x0, y0), (x1, y1), (x2, y2) = xy[T]
        det = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
        gx = np.array([y1 - y2, y2 - y0, y0 - y1]) / det
        gy = np.array([x2 - x1, x0 - x2, x1 - x0]) / det
        gm = np.array([gx @ m[T], gy @ m[T]]); gp = np.array([gx @ p[T], gy @ p[T]])
        load[T] += (gm @ gp) * abs(det) / 6.0
    assert _relmax(hp.rhs_chtx_q(Q, Mf, P, chi, dt, v), Mmat @ q + dt * chi * load) < 1e-12
    # the same-named builders of old_helpers.py:87-111 (star-imported by chemotaxis_FCT_PDECO.py:189-266): other reaction terms
    Mw = asm.mass_p1_product(m)
    assert _relmax(hp.rhs_chtx_m(Mf, v), 4 * (Mmat @ m)) < 1e-13
    assert _relmax(pat.embed(hp.mat_chtx_m(F, Mf, Dm, chi, u, v)), -Dm * K + chi * asm.chemotaxis_conv(f) + Mw) < 1e-13
    assert _relmax(pat.embed(hp.mat_chtx_p(F, Mf, Dm, chi, u, v)),
                   -Dm * K - chi * asm.chemotaxis_conv(f) + 4 * asm.mass() - 2 * Mw) < 1e-13
    # norm_true_control (helpers.py:1958-2001; nonlinear_FCT_PDECO_refactored.py:235)
    T, dts = 0.3, 0.1
    cv = np.sin(2 * np.pi * om.dof_xy[:, 0]) * np.sin(2 * np.pi * om.dof_xy[:, 1])
    Msp = hp.assemble_sparse(u * v * dx)
    exp_nl = sum(w * (cv @ (Mmat @ cv)) for w in (0.5, 1.0, 1.0, 0.5)) * dts
    assert abs(hp.norm_true_control("nonlinear", T, dts, Msp, V) / exp_nl - 1) < 1e-12
    ones = np.ones(V.dim())
    assert abs(hp.norm_true_control("Schnak", T, dts, Msp, V, c_a=0.1) / (0.01 * 3 * dts * (ones @ (Mmat @ ones))) - 1) < 1e-12
    # (m_initial_condition is pure numpy and is compared with the reference's own function on the CPU:
    # tests/test_host_shims_vs_reference.py)
