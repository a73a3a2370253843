"""CPU tests: the C/OpenMP twin of the oracle (oracle/fct_c.c, used by bench.py's CPU baseline at full size) against the
numpy oracle, which is itself pinned on the reference's goldens (tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import pdeco_numpy as drv
from oracle.fct_c import CDriftProblem
from oracle.p1mesh import RectMesh, transpose_positions


@pytest.mark.parametrize("n", [1, 7, 40])
def test_mesh_numbering_pattern_bit_exact(n):
    c = CDriftProblem(n, -1.0, 1.0)
    m = RectMesh(n, -1.0, 1.0)
    assert np.array_equal(c.vertex_to_dof, m.vertex_to_dof)
    assert np.array_equal(c.cells, m.cells)
    assert np.array_equal(c.dof_xy, m.dof_xy)
    rp, ci = m.pattern()
    assert np.array_equal(c.rowptr, rp) and np.array_equal(c.colidx, ci)
    assert np.array_equal(c.tpos, transpose_positions(rp, ci))


def test_matrices_and_state_loop_vs_numpy_oracle():
    n, ns = 32, 4
    orc = drv.AdvectionDriftPDECO(n, 0.0, 1.0, solver="jacobi")
    c = CDriftProblem(n, 0.0, 1.0)
    assert rel_l2(c.M, orc.M) < 1e-14 and rel_l2(c.ML, orc.ML) < 1e-14
    rng = np.random.default_rng(4)
    ctl = 1.0 + rng.random((ns + 1, orc.nodes))
    assert np.abs(c.drift_operator(ctl[1]) - orc.operator(ctl[1])).max() < 1e-13 * np.abs(orc.operator(ctl[1])).max()
    h = 1.0 / n
    dt = 0.25 * h / (2 * np.sqrt(2))
    u0 = orc.gaussian_ic()
    u_o = orc.state(ctl, u0, ns, dt)
    u_c, sweeps = c.state(ctl, u0, ns, dt)
    assert sweeps >= 2 * ns
    for i in range(1, ns + 1):
        assert rel_l2(u_c[i], u_o[i]) < 1e-12 * i


def test_threads_reported():
    assert CDriftProblem(2).threads() >= 1
