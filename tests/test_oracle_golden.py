"""CPU tests: the oracle (numpy restatement) against the reference's golden data and against
outputs of the reference's own functions (fixtures from tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle.fct_numpy import (Pattern, artificial_diffusion, chebsi, cost_functional, fct_step,
                              fct_step_legacy, l2_norm_sq_omega, l2_norm_sq_q)
from oracle.p1assembly import P1Assembler
from oracle.p1mesh import RectMesh, reorder_vector_from_dof, reorder_vector_to_dof, vertex_to_dof_rect
from oracle import pdeco_numpy as drv


def _setup(n, a1, a2):
    mesh = RectMesh(n, a1, a2)
    return mesh, P1Assembler(mesh), Pattern(*mesh.pattern())


def test_vertex_to_dof_closed_form_examples():
    # SURVEY.md App. B.2 examples recovered from the shipped chemotaxis trajectory
    v2d = vertex_to_dof_rect(40)
    N = 41
    assert v2d[0 * N + 0] == 820 and v2d[0 * N + 40] == 1680
    assert v2d[40 * N + 0] == 0 and v2d[40 * N + 1] == 2
    assert np.array_equal(np.sort(v2d), np.arange(N * N))


def test_pattern_counts():
    for n in (1, 2, 5, 40):
        mesh = RectMesh(n)
        rowptr, colidx = mesh.pattern()
        N = n + 1
        edges = 2 * N * (N - 1) + (N - 1) ** 2
        assert colidx.size == N * N + 2 * edges
        assert np.all(np.diff(rowptr) <= 7)
        for i in range(N * N):
            r = colidx[rowptr[i]:rowptr[i + 1]]
            assert np.all(np.diff(r) > 0) and i in r


def test_reorder_roundtrip():
    mesh = RectMesh(6)
    v = np.random.default_rng(0).random(3 * mesh.nodes)
    d = reorder_vector_to_dof(v, 3, mesh.nodes, mesh.vertex_to_dof)
    assert np.array_equal(reorder_vector_from_dof(d, 3, mesh.nodes, mesh.vertex_to_dof), v)


@pytest.mark.parametrize("tag", ["solid", "chtxs", "schnak", "drift"])
def test_fct_step_vs_reference_function(ref_cases, tag):
    c = ref_cases
    n = int(c[f"{tag}_n"][0]); a1, a2 = c[f"{tag}_box"]
    mesh, asm, pat = _setup(n, a1, a2)
    M = asm.mass(); ML = asm.lumped(M)
    S = c[f"{tag}_S"] if c[f"{tag}_S"].size else None
    out = fct_step(pat, c[f"{tag}_A"], c[f"{tag}_rhs"], c[f"{tag}_un"], float(c[f"{tag}_dt"][0]), M, ML, S=S)
    assert rel_l2(out, c[f"{tag}_out"]) < 5e-15
    outj = fct_step(pat, c[f"{tag}_A"], c[f"{tag}_rhs"], c[f"{tag}_un"], float(c[f"{tag}_dt"][0]), M, ML, S=S,
                    solver="jacobi")
    assert rel_l2(outj, c[f"{tag}_out"]) < 1e-13


def test_chebsi_adm_norms_vs_reference_functions(ref_cases):
    c = ref_cases
    mesh, asm, pat = _setup(12, -1.0, 1.0)
    M = asm.mass()
    assert rel_l2(chebsi(pat, c["cheb_b"], M, M[pat.diagpos]), c["cheb_out"]) < 1e-15
    assert rel_l2(chebsi(pat, c["cheb_b"], M, M[pat.diagpos], 7), c["cheb7_out"]) < 1e-15
    assert np.abs(artificial_diffusion(pat, c["adm_in"]) - c["adm_out"]).max() < 1e-15
    ns, dt, beta = int(c["norm_meta"][0]), c["norm_meta"][1], c["norm_meta"][2]
    phi, tgt, ctl = c["norm_phi"], c["norm_tgt"], c["norm_ctl"]
    assert abs(l2_norm_sq_q(pat, phi, ns, dt, M) / c["norm_Q"][0] - 1) < 1e-14
    assert abs(l2_norm_sq_omega(pat, phi[:mesh.nodes], M) / c["norm_Omega"][0] - 1) < 1e-14
    assert abs(cost_functional(pat, phi, tgt, ctl, ns, dt, M, beta, "alltime") / c["cost_alltime"][0] - 1) < 1e-14
    assert abs(cost_functional(pat, phi, tgt[:mesh.nodes], ctl, ns, dt, M, beta, "finaltime")
               / c["cost_finaltime"][0] - 1) < 1e-14
    assert abs(cost_functional(pat, phi, tgt, ctl, ns, dt, M, beta, "alltime", var2=tgt, var2_target=phi)
               / c["cost_alltime2"][0] - 1) < 1e-14
    with pytest.raises(ValueError):
        cost_functional(pat, phi, tgt, ctl, ns, dt, M, beta, "sometime")


def test_chemotaxis_golden_trajectory(ref_data):
    """Chtxs_data_dx0.025_dt0.001/chtxs_{m,f}_t0.01.csv: IC + 10 steps of solve_chtxs_system
    (helpers.py:1250-1385) with control_fun=Constant(100), rescaling=1."""
    gm, gf = ref_data["chtxs_m"], ref_data["chtxs_f"]
    prob = drv.ChemotaxisProblem(40, 0.0, 1.0)
    m0, f0 = prob.initial_condition()
    assert np.array_equal(m0, gm[0]) and np.array_equal(f0, gf[0])
    m, f = prob.forward(None, m0, f0, 10, 1e-3, control_const=100.0, rescaling=1.0)
    for s in range(1, 11):
        assert rel_l2(m[s], gm[s]) < 1e-14, s
        assert rel_l2(f[s], gf[s]) < 1e-14, s


def test_solidbody_golden_t025(ref_data):
    """data/solidbody_t0.25_u.csv = advection_solidbody_FCT.py with slit 0.05, dt = deltax**2, 400 steps."""
    prob = drv.SolidBodyProblem(80, -1.0, 1.0, slit_width=0.05)
    u = prob.forward(400, 0.025 ** 2)
    assert rel_l2(u, ref_data["solidbody_t0.25"]) < 1e-13


# ---- the reference's own projected Armijo search (helpers.py:1583-1713) --------------------------------------
@pytest.fixture(scope="module")
def ref_armijo():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_armijo.npz"))


@pytest.mark.parametrize("tag,expect_k", [("nl", 1), ("nlbt", 6), ("nlmax", 3), ("sk", 10)])
def test_armijo_ref_vs_reference_function(ref_armijo, tag, expect_k):
    """oracle/pdeco_systems.armijo_ref against outputs of the reference's armijo_line_search_ref (run unmodified in the
    build container by tests/golden/make_golden.py:ref_armijo with the oracle's time loops as solver callbacks): same
    number of trials, bit-identical accepted control, same state trajectories."""
    from oracle import pdeco_systems as osys
    from oracle.fct_numpy import cost_functional as o_cost  # noqa: F401
    g = ref_armijo
    ns, dt, lo, hi, beta, cost0, max_iter = g[f"{tag}_meta"]
    ns, max_iter = int(ns), int(max_iter)
    assert int(g[f"{tag}_k"][0]) == expect_k
    if tag == "sk":
        prob = osys.SchnakProblem(8, 0.0, 1.0)
        u0, v0 = g["sk_u0"], g["sk_v0"]

        def solver(ci):
            a, b = prob.state(ci, u0, v0, ns, dt)
            return a.ravel(), b.ravel()
        v1, v2, cinc, k = osys.armijo_ref(prob, solver, None, g["sk_c"], g["sk_d"], g["sk_target"], ns, dt, lo, hi, beta, cost0,
                                          "alltime", max_iter=max_iter, var2=np.zeros(1), var2_target=g["sk_target2"])
        assert rel_l2(v2, g["sk_var2"]) < 1e-13
    else:
        prob = osys.NonlinearProblem(10, 0.0, 1.0)
        u0 = g["nl_u0"]
        solver = lambda ci: (prob.state(ci, u0, ns, dt).ravel(), None)
        v1, v2, cinc, k = osys.armijo_ref(prob, solver, None, g[f"{tag}_c"], g[f"{tag}_d"], g[f"{tag}_target"], ns, dt, lo, hi,
                                          beta, cost0, "finaltime", max_iter=max_iter)
        assert v2 is None
    assert k == expect_k
    assert np.array_equal(cinc, g[f"{tag}_cinc"])
    assert rel_l2(v1, g[f"{tag}_var1"]) < 1e-13


# ---- the reference's own legacy FCT_alg (old_helpers.py:112-204) -------------------------------------------------
@pytest.mark.parametrize("tag", ["lsolid", "lschnak"])
def test_fct_step_legacy_vs_reference_legacy_function(tag):
    """The legacy sign convention (M du/dt = A u - S u + r) pinned on the legacy function itself, executed unmodified from
    old_helpers.py by tests/golden/make_golden.py:ref_legacy -- not only through FCT_alg(A, S) == FCT_alg_ref(-A, S)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_legacy.npz"))
    n = int(g[f"{tag}_n"][0]); a1, a2 = g[f"{tag}_box"]
    mesh, asm, pat = _setup(n, a1, a2)
    M = asm.mass()
    S = g[f"{tag}_S"] if g[f"{tag}_S"].size else None
    out = fct_step_legacy(pat, g[f"{tag}_A"], g[f"{tag}_rhs"], g[f"{tag}_un"], float(g[f"{tag}_dt"][0]), M, asm.lumped(M),
                          source=S)
    assert rel_l2(out, g[f"{tag}_out"]) < 1e-13
    # and the identity the drop-in relies on
    out2 = fct_step(pat, -g[f"{tag}_A"], g[f"{tag}_rhs"], g[f"{tag}_un"], float(g[f"{tag}_dt"][0]), M, asm.lumped(M), S=S)
    assert np.array_equal(out, out2)
