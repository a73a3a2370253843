#!/usr/bin/env python
"""Minimal driver for ncu captures of the tile kernels: fused Jacobi launches (K = 4) and one ChebSI (20 iterations)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_fct_pdeco_b200.mesh import RectMeshP1  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mesh = RectMeshP1(cells, 0.0, 1.0)
ctx = mesh.context(device=0)
n, nnz = mesh.nodes, mesh.nnz
dt = 0.25 / cells / (2 * np.sqrt(2))
xy = mesh.dof_xy
u0 = np.exp(-20 * ((2 * xy[:, 0] - 1 + 2 / 3) ** 2 + 5 * (2 * xy[:, 1] - 1 + 5 / 6) ** 2))
c0 = 1.0 + 0.25 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1])
d_c, d_u = ctx.array(c0), ctx.array(u0)
A = ctx.empty(nnz)
ctx.assemble_matrix(2, A, c0=d_c, s0=1.0, s1=1.0, scale=-1.0)
M, _, Md, _ = ctx.static()
b, y = ctx.array(u0), ctx.empty(n)
ctx.bench_jacobi_fused(A, d_u, dt, sweeps=4, reps=1)
ctx.chebsi(M, Md, b, y, 20)
ctx.sync()
