"""dolfin-free stand-ins for the mesh / function-space objects the reference scripts pass around.

``RectangleMesh(Point(a1,a1), Point(a2,a2), n, n)`` + ``FunctionSpace(mesh,'CG',1)`` +
``vertex_to_dof_map(V)`` (advection_solidbody_FCT.py:48-50,82) are replaced by closed-form host code in
libfctpdeco (fct_mesh_rect_build): "right" diagonals, anti-diagonal CG1 DoF numbering.
"""
import ctypes as C

import numpy as np

from ._lib import check, lib
from .context import FctContext, _hp


class RectMeshP1:
    """P1 triangulation of [a1,a2]^2 with n x n squares, in dolfin's vertex/cell/DoF numbering."""

    def __init__(self, n, a1=0.0, a2=1.0):
        self.n = int(n)
        self.a1, self.a2 = float(a1), float(a2)
        nodes, cells, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.fct_mesh_rect_sizes(self.n, C.byref(nodes), C.byref(cells), C.byref(nnz)))
        self.nodes, self.ncells, self.nnz = nodes.value, cells.value, nnz.value
        self.vertex_to_dof = np.empty(self.nodes, dtype=np.int32)
        self.cells = np.empty((self.ncells, 3), dtype=np.int32)
        self.dof_xy = np.empty((self.nodes, 2), dtype=np.float64)
        self.rowptr = np.empty(self.nodes + 1, dtype=np.int32)
        self.colidx = np.empty(self.nnz, dtype=np.int32)
        check(lib.fct_mesh_rect_build(self.n, self.a1, self.a2, _hp(self.vertex_to_dof), _hp(self.cells),
                                      _hp(self.dof_xy), _hp(self.rowptr), _hp(self.colidx)))
        self._ctx = None

    # dolfin-like accessors -------------------------------------------------------------------
    def num_vertices(self):
        return self.nodes

    def pattern(self):
        return self.rowptr, self.colidx

    def dof_neighbors(self):
        """helpers.py:271-307 find_node_neighbours: DoF neighbours of each DoF, own index last."""
        out = []
        rp, ci = self.rowptr, self.colidx
        for i in range(self.nodes):
            row = ci[rp[i]:rp[i + 1]]
            out.append([int(j) for j in row if j != i] + [i])
        return out

    def context(self, device=0):
        """the (cached) GPU context of this mesh with M, M_L, K assembled on the device"""
        if self._ctx is None:
            ctx = FctContext(self.rowptr, self.colidx, device=device)
            ctx.set_mesh(self.cells, self.dof_xy)
            ctx.set_rect(self.n, 0)
            ctx.assemble_static()
            self._ctx = ctx
        return self._ctx


class FunctionSpaceP1:
    """stand-in for dolfin.FunctionSpace(mesh, 'CG', 1)"""

    def __init__(self, mesh):
        self._mesh = mesh

    def dim(self):
        return self._mesh.nodes

    def mesh(self):
        return self._mesh


def vertex_to_dof_map(V):
    return V.mesh().vertex_to_dof.copy()
