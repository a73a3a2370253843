"""Oracle: numpy restatement of the dolfin P1 assembly calls on the reference's hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference assembles with
``assemble_sparse`` (helpers.py:87-104: dolfin assemble -> PETSc AIJ -> csr_matrix, explicit
zeros kept, columns ascending) and ``assemble(linear form)``.  dolfin / FFC / FIAT / PETSc are
third-party and absent from /root/reference (no pinned version anywhere in that repo; the
shipped byte-code is cpython-38 => FEniCS-legacy 2019.x).  What is restated here is the
published algorithm: per-cell P1 element tensors by FIAT's default triangle quadrature of the
UFL-estimated degree, summed cell by cell (dolfin cell order) into the fixed CSR pattern.

All matrices are returned as *value arrays on the fixed pattern* of ``RectMesh.pattern()``
(``to_csr`` wraps one as scipy CSR).  Index convention: i = test function, j = trial function.

Form catalogue covered (reference call sites cited per function).
"""
import numpy as np
import scipy.sparse as sp

# --- FIAT "default" triangle schemes on the reference triangle (weights sum to 1/2) ----------
# degree 1: centroid; degree 2: 3-point; degree 4: Strang-Fix 6-point (literal 15-digit constants:
# the chemotaxis golden depends on them); degree 5: Strang-Fix 7-point.
_Q = {}
_Q[1] = (np.array([[1.0 / 3.0, 1.0 / 3.0]]), np.array([0.5]))
_Q[2] = (np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]), np.array([1 / 6, 1 / 6, 1 / 6]))
_a1, _b1 = 0.816847572980459, 0.091576213509771
_a2, _b2 = 0.108103018168070, 0.445948490915965
_Q[4] = (np.array([[_b1, _b1], [_a1, _b1], [_b1, _a1], [_b2, _b2], [_a2, _b2], [_b2, _a2]]),
         np.array([0.109951743655322 / 2] * 3 + [0.223381589678011 / 2] * 3))
_c1, _d1 = 0.79742698535308720, 0.10128650732345633
_c2, _d2 = 0.05971587178976981, 0.47014206410511505
_Q[5] = (np.array([[1 / 3, 1 / 3], [_d1, _d1], [_c1, _d1], [_d1, _c1], [_d2, _d2], [_c2, _d2], [_d2, _c2]]),
         np.array([0.225 / 2] + [0.12593918054482717 / 2] * 3 + [0.13239415278850616 / 2] * 3))
_Q[3] = _Q[4]   # any exact rule reproduces polynomial integrands to rounding
_Q[0] = _Q[1]


def quad_rule(degree):
    """(points[q,2] on the reference triangle, weights[q]) exact for polynomials of `degree`."""
    if degree > 5:
        raise ValueError("quadrature degree > 5 is not on the reference's hot path")
    pts, w = _Q[degree]
    phi = np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], axis=1)   # phi[q, a]
    return pts, w, phi


class P1Assembler:
    """Element tensors + scatter into the fixed CSR pattern of a ``RectMesh``-like mesh.

    ``mesh`` needs: cells (int32[nc,3], DoF-indexed), dof_xy (float64[n,2]), nodes, pattern().
    """

    def __init__(self, mesh):
        self.mesh = mesh
        self.n = mesh.nodes
        self.rowptr, self.colidx = mesh.pattern()
        self.nnz = self.colidx.size
        c = mesh.cells.astype(np.int64)
        self.cells = c
        p = mesh.dof_xy[c]                                   # [nc,3,2]
        x0, y0 = p[:, 0, 0], p[:, 0, 1]
        x1, y1 = p[:, 1, 0], p[:, 1, 1]
        x2, y2 = p[:, 2, 0], p[:, 2, 1]
        det = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
        self.detJ = np.abs(det)
        self.area = 0.5 * self.detJ
        g1 = np.stack([(y2 - y0) / det, -(x2 - x0) / det], axis=1)
        g2 = np.stack([-(y1 - y0) / det, (x1 - x0) / det], axis=1)
        g0 = -(g1 + g2)
        self.G = np.stack([g0, g1, g2], axis=1)              # G[c,a,:] = grad phi_a
        self.P = p
        # slot[c,i,j]: CSR position of (cells[c,i], cells[c,j])
        rows = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(self.rowptr))
        own = rows * self.n + self.colidx.astype(np.int64)
        key = (c[:, :, None] * self.n + c[:, None, :]).ravel()
        self.slot = np.searchsorted(own, key)
        assert np.array_equal(own[self.slot], key)
        self.rows = rows
        self.diagpos = np.flatnonzero(rows == self.colidx)

    # -- helpers -----------------------------------------------------------------------
    def scatter_matrix(self, local):
        """local[nc,3,3] -> values on the pattern (cell-order summation, explicit zeros kept)."""
        return np.bincount(self.slot, weights=local.ravel(), minlength=self.nnz)

    def scatter_vector(self, local):
        return np.bincount(self.cells.ravel(), weights=local.ravel(), minlength=self.n)

    def to_csr(self, vals):
        return sp.csr_matrix((np.array(vals, dtype=np.float64), self.colidx.copy(), self.rowptr.copy()),
                             shape=(self.n, self.n))

    def at_quad(self, f, phi):
        """P1 field f (DoF vector) at quadrature points: [nc,q]."""
        return f[self.cells] @ phi.T

    def xy_quad(self, phi):
        return np.einsum('qa,cad->cqd', phi, self.P)

    # -- static matrices ---------------------------------------------------------------
    def mass(self):
        """u*v*dx (helpers.py:553, 1305; every script)."""
        Me = (self.area / 12.0)[:, None, None] * (np.ones((3, 3)) + np.eye(3))[None]
        return self.scatter_matrix(Me)

    def stiffness(self):
        """dot(grad(u),grad(v))*dx (helpers.py:555, 1307)."""
        Ke = self.area[:, None, None] * np.einsum('cid,cjd->cij', self.G, self.G)
        return self.scatter_matrix(Ke)

    def lumped(self, Mvals):
        """row_lump (helpers.py:309-328): diag(M 1)."""
        return np.bincount(self.rows, weights=Mvals, minlength=self.n)

    # -- weighted masses ---------------------------------------------------------------
    def weighted_mass(self, coef_q_fn, degree):
        """g*u*v*dx with g given at quadrature points by coef_q_fn(phi, xyq) -> [nc,q].

        Sites: helpers.py:591 (u_np1**2), :683, :692 (u*v), :953, :1032; old_helpers.py:103,110.
        """
        pts, w, phi = quad_rule(degree)
        gq = coef_q_fn(phi, self.xy_quad(phi))
        Me = np.einsum('cq,q,qi,qj->cij', gq, w, phi, phi) * self.detJ[:, None, None]
        return self.scatter_matrix(Me)

    def mass_p1_product(self, *fields):
        """(f1*f2*...)*u*v*dx for P1 fields (quadrature degree = number of fields + 2)."""
        deg = len(fields) + 2
        return self.weighted_mass(lambda phi, xy: np.prod([self.at_quad(f, phi) for f in fields], axis=0), deg)

    # -- convection operators ----------------------------------------------------------
    def conv_conservative(self, wind_fn, degree=5):
        """dot(w,grad(v))*u*dx with analytic wind w(x,y)->(wx,wy)  (helpers.py:581, 933, 1015;
        advection_solidbody_FCT.py:106).  A[i,j] = grad phi_i . int w phi_j.  Polynomial winds of
        degree <= 4 are integrated exactly (dolfin interpolates Expression(degree=4) to P4 first)."""
        pts, w, phi = quad_rule(degree)
        xy = self.xy_quad(phi)
        wx, wy = wind_fn(xy[..., 0], xy[..., 1])
        wx = np.broadcast_to(wx, xy.shape[:2])
        wy = np.broadcast_to(wy, xy.shape[:2])
        Wx = np.einsum('cq,q,qj->cj', wx, w, phi) * self.detJ[:, None]
        Wy = np.einsum('cq,q,qj->cj', wy, w, phi) * self.detJ[:, None]
        Ae = self.G[:, :, 0][:, :, None] * Wx[:, None, :] + self.G[:, :, 1][:, :, None] * Wy[:, None, :]
        return self.scatter_matrix(Ae)

    def conv_conservative_p1(self, wx, wy):
        """Same form with a P1 (nodal) wind: closed form A_e[i,j] = grad phi_i . area/12 (sum_k w_k + w_j)
        (SURVEY.md App. F.2).  Exact for linear winds such as the solid-body rotation."""
        c = self.cells
        sx = wx[c].sum(axis=1)[:, None] + wx[c]
        sy = wy[c].sum(axis=1)[:, None] + wy[c]
        Wx = (self.area / 12.0)[:, None] * sx
        Wy = (self.area / 12.0)[:, None] * sy
        Ae = self.G[:, :, 0][:, :, None] * Wx[:, None, :] + self.G[:, :, 1][:, :, None] * Wy[:, None, :]
        return self.scatter_matrix(Ae)

    def conv_nonconservative(self, wind_fn, degree=5):
        """dot(w,grad(u))*v*dx (helpers.py:681): transpose of the element tensor above."""
        pts, w, phi = quad_rule(degree)
        xy = self.xy_quad(phi)
        wx, wy = wind_fn(xy[..., 0], xy[..., 1])
        wx = np.broadcast_to(wx, xy.shape[:2])
        wy = np.broadcast_to(wy, xy.shape[:2])
        Wx = np.einsum('cq,q,qi->ci', wx, w, phi) * self.detJ[:, None]
        Wy = np.einsum('cq,q,qi->ci', wy, w, phi) * self.detJ[:, None]
        Ae = self.G[:, :, 0][:, None, :] * Wx[:, :, None] + self.G[:, :, 1][:, None, :] * Wy[:, :, None]
        return self.scatter_matrix(Ae)

    def drift_mass(self, c, bx, by):
        """dot(b,grad(c))*u*v*dx, b constant, c P1 (advection_solidbody_FCT_PDECO_alltime.py:222,252;
        old_helpers.py:62): (b.grad c)|_e * area/12 (1+delta_ij)."""
        gc = np.einsum('ca,cad->cd', c[self.cells], self.G)
        s = bx * gc[:, 0] + by * gc[:, 1]
        Me = (s * self.area / 12.0)[:, None, None] * (np.ones((3, 3)) + np.eye(3))[None]
        return self.scatter_matrix(Me)

    def drift_conv(self, c, bx, by):
        """dot(b,grad(v))*c*u*dx (advection_solidbody_FCT_PDECO_alltime.py:223,253; old_helpers.py:63):
        (b.grad phi_i) * area/12 (sum_k c_k + c_j)."""
        ce = c[self.cells]
        W = (self.area / 12.0)[:, None] * (ce.sum(axis=1)[:, None] + ce)
        bg = bx * self.G[:, :, 0] + by * self.G[:, :, 1]
        return self.scatter_matrix(bg[:, :, None] * W[:, None, :])

    def chemotaxis_conv(self, f, coef_q_fn=None, degree=1):
        """coef * dot(grad(f),grad(v)) * u * dx, f P1:
        coef = 1 (old_helpers.py:102,108; mimura_data_helpers.py), or exp(-eta*m) at quadrature degree 4
        (helpers.py:1350-1351).  A[i,j] = (grad f . grad phi_i) int coef phi_j."""
        pts, w, phi = quad_rule(degree)
        gf = np.einsum('ca,cad->cd', f[self.cells], self.G)
        s = np.einsum('cd,cid->ci', gf, self.G)
        if coef_q_fn is None:
            cq = np.ones((self.cells.shape[0], w.size))
        else:
            cq = coef_q_fn(phi, self.xy_quad(phi))
        Wj = np.einsum('cq,q,qj->cj', cq, w, phi) * self.detJ[:, None]
        return self.scatter_matrix(s[:, :, None] * Wj[:, None, :])

    def chemotaxis_adjoint_mat(self, u, vn, eta, degree=5):
        """(1-eta*u)*exp(-eta*u)*dot(grad(u_trial),grad(v_n))*w*dx (helpers.py:1499-1500):
        A[i,j] = (grad phi_j . grad v_n) int coef phi_i, quadrature degree 5 (UFL estimate)."""
        pts, w, phi = quad_rule(degree)
        gv = np.einsum('ca,cad->cd', vn[self.cells], self.G)
        s = np.einsum('cd,cjd->cj', gv, self.G)
        uq = self.at_quad(u, phi)
        cq = (1.0 - eta * uq) * np.exp(-eta * uq)
        Wi = np.einsum('cq,q,qi->ci', cq, w, phi) * self.detJ[:, None]
        return self.scatter_matrix(Wi[:, :, None] * s[:, None, :])

    # -- load vectors ------------------------------------------------------------------
    def load(self, coef_q_fn, degree):
        """f*v*dx with f at quadrature points (helpers.py:584-585, 594, 684, 693, 956, 1339-1340, 1505)."""
        pts, w, phi = quad_rule(degree)
        fq = coef_q_fn(phi, self.xy_quad(phi))
        be = np.einsum('cq,q,qi->ci', fq, w, phi) * self.detJ[:, None]
        return self.scatter_vector(be)

    def load_p1_product(self, *fields, scale=1.0):
        deg = len(fields) + 1
        return scale * self.load(lambda phi, xy: np.prod([self.at_quad(f, phi) for f in fields], axis=0), deg)

    def load_constant(self, value):
        """Constant(value)*v*dx (helpers.py:594): value*area/3 per cell vertex."""
        return self.scatter_vector(np.repeat((value * self.area / 3.0)[:, None], 3, axis=1))

    def load_grad_pair(self, coef_q_fn, p, degree):
        """coef * dot(grad(p),grad(w)) * dx, e.g. chi*u*exp(-eta*u) (helpers.py:1531-1532)."""
        pts, w, phi = quad_rule(degree)
        gp = np.einsum('ca,cad->cd', p[self.cells], self.G)
        s = np.einsum('cd,cid->ci', gp, self.G)
        cq = coef_q_fn(phi, self.xy_quad(phi))
        integ = np.einsum('cq,q->c', cq, w) * self.detJ
        return self.scatter_vector(s * integ[:, None])

    def load_drift_grad(self, p, u, bx, by):
        """p*dot(b,grad(u))*v*dx (advection_solidbody_FCT_PDECO_alltime.py:273): (b.grad u)|_e * M_e p_e."""
        gu = np.einsum('ca,cad->cd', u[self.cells], self.G)
        s = bx * gu[:, 0] + by * gu[:, 1]
        pe = p[self.cells]
        be = (s * self.area / 12.0)[:, None] * (pe.sum(axis=1)[:, None] + pe)
        return self.scatter_vector(be)
