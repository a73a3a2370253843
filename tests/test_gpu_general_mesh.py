"""GPU parity on a mesh that is NOT the structured dolfin RectangleMesh: a criss-cross triangulation (alternating
diagonals) with a scrambled DoF numbering.  Rows have 5 or 9 entries, so the generic code paths run (shared-memory
staged assembly instead of the 8-entry register path, generic row loops, general CSR pattern handed in by the
caller) -- the library takes rowptr/colidx/cells/coordinates as inputs precisely so that any P1 mesh can be used."""
import numpy as np
import pytest
import scipy.sparse as sp

import fem_fct_pdeco_b200 as fp
from fem_fct_pdeco_b200 import _lib
from conftest import rel_l2
from oracle.fct_numpy import Pattern, chebsi, fct_step
from oracle.p1assembly import P1Assembler

pytestmark = pytest.mark.gpu


class CrissCrossMesh:
    def __init__(self, n, seed=3):
        N = n + 1
        lin = np.linspace(0.0, 1.0, N)
        X, Y = np.meshgrid(lin, lin)
        # a mild interior perturbation: genuinely non-uniform element shapes
        rng = np.random.default_rng(seed)
        h = 1.0 / n
        X[1:-1, 1:-1] += 0.2 * h * (rng.random((N - 2, N - 2)) - 0.5)
        Y[1:-1, 1:-1] += 0.2 * h * (rng.random((N - 2, N - 2)) - 0.5)
        xy = np.stack([X.ravel(), Y.ravel()], axis=1)
        cells = []
        for iy in range(n):
            for ix in range(n):
                v0 = iy * N + ix; v1 = v0 + 1; v2 = v0 + N; v3 = v2 + 1
                if (ix + iy) % 2 == 0:
                    cells += [(v0, v1, v3), (v0, v3, v2)]
                else:
                    cells += [(v0, v1, v2), (v1, v3, v2)]
        cells = np.array(cells, dtype=np.int64)
        perm = rng.permutation(N * N)                 # vertex -> DoF
        self.nodes = N * N
        self.cells = perm[cells].astype(np.int32)
        self.dof_xy = np.empty_like(xy)
        self.dof_xy[perm] = xy
        c = self.cells.astype(np.int64)
        P = sp.coo_matrix((np.ones(9 * len(c), dtype=np.int8), (np.repeat(c, 3, axis=1).ravel(), np.tile(c, (1, 3)).ravel())),
                          shape=(self.nodes, self.nodes)).tocsr()
        P.sum_duplicates(); P.sort_indices()
        self._pattern = (P.indptr.astype(np.int32), P.indices.astype(np.int32))

    def pattern(self):
        return self._pattern


def _relmax(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


def test_general_mesh_assembly_and_fct_step():
    mesh = CrissCrossMesh(14)
    rowptr, colidx = mesh.pattern()
    assert np.diff(rowptr).max() == 9
    asm, pat = P1Assembler(mesh), Pattern(rowptr, colidx)
    ctx = fp.FctContext(rowptr, colidx)
    ctx.set_mesh(mesh.cells, mesh.dof_xy)
    ctx.assemble_static()
    M, ML, Md, K = ctx.static()
    Mo = asm.mass()
    assert _relmax(M.download(), Mo) < 1e-13 and _relmax(K.download(), asm.stiffness()) < 1e-13
    assert _relmax(ML.download(), asm.lumped(Mo)) < 1e-13
    rng = np.random.default_rng(9)
    c = 1.0 + rng.random(mesh.nodes)
    m = 1.0 + rng.random(mesh.nodes)
    dcv, dmv = ctx.array(c), ctx.array(m)
    A = ctx.empty(ctx.nnz)
    ctx.assemble_matrix(_lib.FORM_DRIFT, A, c0=dcv, s0=1.0, s1=0.5)
    Ao = asm.drift_mass(c, 1.0, 0.5) + asm.drift_conv(c, 1.0, 0.5)
    assert _relmax(A.download(), Ao) < 1e-13
    ctx.assemble_matrix(_lib.FORM_CHTX_EXP, A, c0=dcv, c1=dmv, s0=0.5)
    ref = asm.chemotaxis_conv(c, lambda phi, xy: np.exp(-0.5 * asm.at_quad(m, phi)), degree=4)
    assert _relmax(A.download(), ref) < 1e-13
    out = ctx.empty(ctx.n)
    ctx.assemble_vector(_lib.LOAD_P1_2, out, c0=dcv, c1=dmv)
    assert _relmax(out.download(), asm.load_p1_product(c, m)) < 1e-13
    # ChebSI and a full FCT step (legacy sign) with the drift operator, rhs and a non-flux matrix
    b = rng.random(mesh.nodes)
    y = ctx.empty(ctx.n)
    ctx.chebsi(M, Md, ctx.array(b), y, 20)
    assert rel_l2(y.download(), chebsi(pat, b, Mo, Mo[pat.diagpos])) < 1e-14
    u_n = np.exp(-30 * ((mesh.dof_xy[:, 0] - 0.4) ** 2 + (mesh.dof_xy[:, 1] - 0.5) ** 2))
    rhs = asm.load_p1_product(rng.random(mesh.nodes))
    S = 3.0 * Mo
    dt = 2e-3
    ctx.assemble_matrix(_lib.FORM_DRIFT, A, c0=dcv, s0=1.0, s1=0.5)
    un1 = ctx.empty(ctx.n)
    info = ctx.step(A, ctx.array(u_n), dt, un1, sign=-1.0, S=ctx.array(S), rhs=ctx.array(rhs))
    assert info.converged
    ref = fct_step(pat, -Ao, rhs, u_n, dt, Mo, asm.lumped(Mo), S=S)
    assert rel_l2(un1.download(), ref) < 1e-12
    # the host-buffer entry point and an unsymmetric pattern is rejected
    out2, info2 = ctx.step_host(A.download(), u_n, dt, sign=-1.0, S_vals=S, rhs=rhs)
    assert info2.converged and rel_l2(out2, un1.download()) < 1e-13
    bad_cols = colidx.copy()
    k = rowptr[0]
    first_off = [j for j in range(rowptr[0], rowptr[1]) if colidx[j] != 0][0]
    with pytest.raises(fp.FctError):
        # drop the transposed partner of entry (0, j): no longer structurally symmetric
        j = int(colidx[first_off])
        keep = np.ones(colidx.size, dtype=bool)
        row_j = np.arange(rowptr[j], rowptr[j + 1])
        keep[row_j[colidx[row_j] == 0]] = False
        rp2 = np.concatenate([[0], np.cumsum(np.bincount(np.repeat(np.arange(mesh.nodes), np.diff(rowptr))[keep],
                                                        minlength=mesh.nodes))]).astype(np.int32)
        fp.FctContext(rp2, colidx[keep])
