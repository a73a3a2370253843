"""Oracle: vectorised numpy/scipy restatement of the reference's FCT step and its helpers.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the reference lines it
follows (/root/reference = KarolinaBenkova/FEM-FCT-PDECO).  All sparse operands are *value
arrays on one fixed CSR pattern* (rowptr, colidx) -- the reference indexes M[i,j], D[i,j] by
dof_neighbors, which is the same thing (helpers.py:1818-1822).

Parity pinning: checked against the reference's own FCT_alg_ref/ChebSI/artificial_diffusion_mat
(imported unmodified, oracle/ref_loader.py) to <= 1e-15 per step and against the shipped golden
trajectories (tests/test_oracle_golden.py).
"""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve

from .p1mesh import transpose_positions


class Pattern:
    """Fixed CSR pattern + derived index arrays."""

    def __init__(self, rowptr, colidx):
        self.rowptr = np.array(rowptr, dtype=np.int32)     # private copies
        self.colidx = np.array(colidx, dtype=np.int32)
        self.n = self.rowptr.size - 1
        self.nnz = self.colidx.size
        self.rows = np.repeat(np.arange(self.n, dtype=np.int32), np.diff(self.rowptr))
        self.tpos = transpose_positions(self.rowptr, self.colidx)
        self.offdiag = self.rows != self.colidx
        self.diagpos = np.flatnonzero(~self.offdiag)
        assert self.diagpos.size == self.n, "pattern must contain the full diagonal"

    def csr(self, vals):
        # index arrays are copied: scipy's in-place ops (eliminate_zeros, ...) must not touch the pattern
        return sp.csr_matrix((np.array(vals, dtype=np.float64), self.colidx.copy(), self.rowptr.copy()),
                             shape=(self.n, self.n))

    def embed(self, mat):
        """Values of a scipy sparse matrix on this pattern (explicit zeros re-inserted: scipy's
        +,-,* prune them).  Raises if `mat` has entries outside the pattern."""
        coo = sp.coo_matrix(mat)
        key = coo.row.astype(np.int64) * self.n + coo.col.astype(np.int64)
        own = self.rows.astype(np.int64) * self.n + self.colidx.astype(np.int64)
        pos = np.searchsorted(own, key)
        ok = (pos < self.nnz)
        ok[ok] = own[pos[ok]] == key[ok]
        if not ok.all():
            if np.any(coo.data[~ok] != 0.0):
                raise ValueError("matrix has nonzeros outside the fixed P1 pattern")
        return np.bincount(pos[ok], weights=coo.data[ok], minlength=self.nnz)

    def rowsum(self, vals):
        return np.bincount(self.rows, weights=vals, minlength=self.n)


def artificial_diffusion(pat, mat):
    """helpers.py:206-242.  d_ij = max(0, -m_ij, -m_ji) (i != j), d_ii = -sum_j d_ij."""
    m = np.asarray(mat, dtype=np.float64)
    d = np.maximum(np.maximum(-m, 0.0), np.maximum(-m[pat.tpos], 0.0))
    d[~pat.offdiag] = 0.0
    d[pat.diagpos] = -pat.rowsum(d)
    return d


def chebsi(pat, vec, M, Md, cheb_iter=20, lmin=0.5, lmax=2):
    """helpers.py:143-185, literally (fixed iteration count; omega special-cased at k == 2)."""
    Mc = pat.csr(M) if not sp.issparse(M) else M
    ymid = np.zeros_like(vec)
    yold = np.zeros_like(ymid)
    omega = 0
    rho = (lmax - lmin) / (lmax + lmin)
    Md = (lmin + lmax) / 2 * Md
    ynew = ymid
    for k in range(1, cheb_iter + 1):
        if k == 2:
            omega = 1 / (1 - rho ** 2 / 2)
        else:
            omega = 1 / (1 - (omega * rho ** 2) / 4)
        r = vec - Mc @ ymid
        z = r / Md
        ynew = omega * (z + ymid - yold) + yold
        yold = ymid
        ymid = ynew
    return ynew


def jacobi_solve(pat, L, b, x0, rtol=1e-14, maxit=200):
    """Jacobi sweeps x <- x + D^-1 (b - L x), stopped on ||dx||_inf <= rtol*||x||_inf.
    Not in the reference (it uses SuperLU spsolve, helpers.py:1782); this is the CPU twin of the
    GPU solver, used only for the CPU baseline at sizes where a direct solve is impractical."""
    Lc = pat.csr(L)
    dinv = 1.0 / L[pat.diagpos]
    x = x0.copy()
    its = 0
    for its in range(1, maxit + 1):
        dx = dinv * (b - Lc @ x)
        x += dx
        if np.max(np.abs(dx)) <= rtol * np.max(np.abs(x)):
            break
    return x, its


def fct_step(pat, A, rhs, u_n, dt, M, ML, S=None, solver="spsolve", info=None):
    """One FCT step, FCT_alg_ref sign convention (helpers.py:1715-1872):
    [M + dt (A + S)] u+ = M u^n + dt rhs, flux matrix K = -A.

    A, M, S: values on `pat`; rhs, u_n, ML: vectors.  solver = "spsolve" (reference: SuperLU,
    helpers.py:1782) or "jacobi" (CPU twin of the GPU solver)."""
    A = np.asarray(A, dtype=np.float64)
    rows, cols = pat.rows, pat.colidx
    rhs = np.zeros(pat.n) if rhs is None else np.asarray(rhs, dtype=np.float64)
    # 1. D = artificial_diffusion_mat(-A)                                   (:1769)
    D = artificial_diffusion(pat, -A)
    # 2. low-order system                                                    (:1775-1782)
    L = dt * (A - D)
    L[pat.diagpos] = ML + L[pat.diagpos]
    if S is not None:
        L = L + dt * np.asarray(S, dtype=np.float64)
    b = ML * u_n + dt * rhs
    if solver == "spsolve":
        u_low = spsolve(pat.csr(L).tocsc(), b)
        its = 0
    else:
        u_low, its = jacobi_solve(pat, L, b, u_n)
    if info is not None:
        info["solver_its"] = its
        info["min_rowsum_L"] = float(pat.rowsum(L).min())
    # 4. du/dt by 20 Chebyshev iterations; S is *not* included             (:1814-1815)
    g = -(pat.csr(A) @ u_low) + rhs
    udot = chebsi(pat, g, M, M[pat.diagpos], 20, 0.5, 2)
    # 5. raw antidiffusive fluxes                                            (:1818-1822)
    F = M * (udot[rows] - udot[cols]) + D * (u_low[rows] - u_low[cols])
    F[pat.diagpos] = 0.0
    # 6. P+-, Q+-                                                            (:1827-1843)
    p_pos = pat.rowsum(np.maximum(F, 0.0))
    p_neg = pat.rowsum(np.minimum(F, 0.0))
    ul_c = u_low[cols]
    q_pos = np.maximum.reduceat(ul_c, pat.rowptr[:-1]) - u_low
    q_neg = np.minimum.reduceat(ul_c, pat.rowptr[:-1]) - u_low
    # 7. R+-                                                                 (:1846-1851)
    r_pos = np.ones(pat.n)
    r_neg = np.ones(pat.n)
    mp = p_pos != 0
    mn = p_neg != 0
    r_pos[mp] = np.minimum(1, ML[mp] * q_pos[mp] / (dt * p_pos[mp]))
    r_neg[mn] = np.minimum(1, ML[mn] * q_neg[mn] / (dt * p_neg[mn]))
    # 8. limited fluxes                                                      (:1860-1866)
    pos = F > 0
    alpha = np.where(pos, np.minimum(r_pos[rows], r_neg[cols]), np.minimum(r_neg[rows], r_pos[cols]))
    Fbar = pat.rowsum(alpha * F)
    # 9. explicit correction                                                 (:1870)
    return u_low + dt * Fbar / ML


def fct_step_legacy(pat, A, rhs, u_n, dt, M, ML, source=None, **kw):
    """Legacy FCT_alg (old_helpers.py:115-204): M du/dt = A u - S u + r  ==  FCT_alg_ref(-A, S)."""
    return fct_step(pat, -np.asarray(A, dtype=np.float64), rhs, u_n, dt, M, ML, S=source, **kw)


# --- norms and cost functional -------------------------------------------------------------
def l2_norm_sq_omega(pat, phi, M):
    """helpers.py:362-381: phi^T M phi."""
    return float(phi @ (pat.csr(M) @ phi))


def l2_norm_sq_q(pat, phi, num_steps, dt, M):
    """helpers.py:330-360: trapezoid in time, mass matrix in space."""
    Mc = pat.csr(M)
    sl = np.split(np.asarray(phi, dtype=np.float64), num_steps + 1)
    w = np.ones(num_steps + 1)
    w[0] = 0.5
    w[-1] = 0.5
    return sum([w[i] * sl[i].transpose() @ Mc @ sl[i] for i in range(num_steps + 1)]) * dt


def cost_functional(pat, var1, var1_target, control, num_steps, dt, M, beta, optim,
                    var2=None, var2_target=None):
    """helpers.py:383-441."""
    valid = ["alltime", "finaltime"]
    if optim not in valid:
        raise ValueError(f"Invalid value for 'optim': '{optim}'. Must be one of {valid}.")
    if optim == "alltime":
        func = 0.5 * l2_norm_sq_q(pat, var1 - var1_target, num_steps, dt, M)
        if var2 is not None and var2_target is not None:
            func += 0.5 * l2_norm_sq_q(pat, var2 - var2_target, num_steps, dt, M)
    else:
        nodes = var1_target.shape[0]
        func = 0.5 * l2_norm_sq_omega(pat, var1[num_steps * nodes:] - var1_target, M)
        if var2 is not None and var2_target is not None:
            func += 0.5 * l2_norm_sq_omega(pat, var2[num_steps * nodes:] - var2_target, M)
    func += beta / 2 * l2_norm_sq_q(pat, control, num_steps, dt, M)
    return func
