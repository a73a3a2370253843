"""Oracle: dolfin-free restatement of the mesh / DoF bookkeeping the reference gets from dolfin.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, in numpy:

* ``RectangleMesh(Point(a1,a1), Point(a2,a2), n, n)`` with the default "right" diagonal
  (every reference script, e.g. advection_solidbody_FCT.py:48, chemotaxis_adjoint_equations.py:57-58):
  vertex id = iy*(n+1)+ix, each square split into (v0,v1,v3),(v0,v2,v3).
* ``vertex_to_dof_map(FunctionSpace(mesh,'CG',1))`` (advection_solidbody_FCT.py:82): the
  anti-diagonal numbering recovered from the reference's shipped chemotaxis trajectory
  (SURVEY.md App. B.2).
* ``find_node_neighbours`` (helpers.py:271-307): DoF-indexed neighbour lists, self last.
* the CSR pattern of ``assemble_sparse`` (helpers.py:87-104): full P1 cell coupling pattern,
  columns ascending, explicit zeros kept.

The pattern here is derived from the *cell list* (generic, via scipy COO->CSR); the product
library derives it in closed form per vertex -- the two are compared bit-for-bit in the tests.
"""
import numpy as np
import scipy.sparse as sp


class RectMesh:
    """Structured P1 triangulation of [a1,a2]^2 with n x n squares ("right" diagonals)."""

    def __init__(self, n, a1=0.0, a2=1.0):
        self.n = int(n)
        self.a1 = float(a1)
        self.a2 = float(a2)
        self.nodes = (self.n + 1) ** 2
        self.ncells = 2 * self.n * self.n
        self.h = (self.a2 - self.a1) / self.n
        N = self.n + 1
        # vertex coordinates, vertex order (iy major)
        lin = self.a1 + self.h * np.arange(N, dtype=np.float64)
        X, Y = np.meshgrid(lin, lin)          # X[iy,ix]
        self.vertex_xy = np.stack([X.ravel(), Y.ravel()], axis=1)
        # cells in dolfin order: per square (iy major, ix minor): (v0,v1,v3) then (v0,v2,v3)
        ix, iy = np.meshgrid(np.arange(self.n), np.arange(self.n))
        v0 = (iy * N + ix).ravel()
        v1 = v0 + 1
        v2 = v0 + N
        v3 = v2 + 1
        cells = np.empty((self.ncells, 3), dtype=np.int64)
        cells[0::2] = np.stack([v0, v1, v3], axis=1)
        cells[1::2] = np.stack([v0, v2, v3], axis=1)
        self.cells_vertex = cells
        self.vertex_to_dof = vertex_to_dof_rect(self.n)
        self.cells = self.vertex_to_dof[cells].astype(np.int32)   # DoF-indexed cells
        xy = np.empty_like(self.vertex_xy)
        xy[self.vertex_to_dof] = self.vertex_xy
        self.dof_xy = xy                                           # coordinates in DoF order
        self._pattern = None

    # -- CSR pattern -------------------------------------------------------------------
    def pattern(self):
        """(rowptr int32[n+1], colidx int32[nnz]) of the P1 coupling pattern, columns sorted."""
        if self._pattern is None:
            c = self.cells.astype(np.int64)
            rows = np.repeat(c, 3, axis=1).ravel()
            cols = np.tile(c, (1, 3)).ravel()
            P = sp.coo_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)),
                              shape=(self.nodes, self.nodes)).tocsr()
            P.sum_duplicates()
            P.sort_indices()
            self._pattern = (P.indptr.astype(np.int32), P.indices.astype(np.int32))
        return self._pattern

    def dof_neighbors(self):
        """List-of-lists as produced by helpers.py:271-307 (neighbours first, own index last)."""
        rowptr, colidx = self.pattern()
        out = []
        for i in range(self.nodes):
            cols = colidx[rowptr[i]:rowptr[i + 1]]
            out.append([int(j) for j in cols if j != i] + [i])
        return out


def vertex_to_dof_rect(n):
    """Closed-form vertex_to_dof_map of CG1 on the n x n "right" RectangleMesh.

    d = ix - iy + n in [0, 2n]; len(d) = min(d, 2n-d)+1; start(d) = sum_{k<d} len(k);
    pos = ix if d <= n else iy; dof = start(d) + pos.   vec_dof[v2d[i]] = vec_vertex[i]
    (helpers.py:33-38).
    """
    N = n + 1
    ix, iy = np.meshgrid(np.arange(N, dtype=np.int64), np.arange(N, dtype=np.int64))
    d = ix - iy + n
    lens = np.minimum(np.arange(2 * n + 1), 2 * n - np.arange(2 * n + 1)) + 1
    start = np.concatenate([[0], np.cumsum(lens)[:-1]])
    pos = np.where(d <= n, ix, iy)
    return (start[d] + pos).ravel()


def transpose_positions(rowptr, colidx):
    """tpos[k] = position of entry (j,i) for entry k=(i,j); requires a structurally symmetric pattern."""
    n = rowptr.size - 1
    nnz = colidx.size
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    key = colidx.astype(np.int64) * n + rows           # key of the transposed entry
    own = rows * n + colidx.astype(np.int64)           # sorted ascending by construction
    tpos = np.searchsorted(own, key)
    if not (tpos < nnz).all() or not np.array_equal(own[tpos], key):
        raise ValueError("pattern is not structurally symmetric")
    return tpos.astype(np.int32)


def reorder_vector_to_dof(vec, num_steps, nodes, vertex_to_dof):
    """Vectorised restatement of helpers.py:13-39."""
    v = np.asarray(vec, dtype=np.float64).reshape(num_steps, nodes)
    out = np.zeros_like(v)
    out[:, np.asarray(vertex_to_dof, dtype=np.int64)] = v
    return out.reshape(-1)


def reorder_vector_from_dof(vec_dof, num_steps, nodes, vertex_to_dof):
    """Vectorised restatement of helpers.py:41-67."""
    v = np.asarray(vec_dof, dtype=np.float64).reshape(num_steps, nodes)
    return v[:, np.asarray(vertex_to_dof, dtype=np.int64)].reshape(-1)
