#!/usr/bin/env python
"""Micro-benchmark of the overlapped-tile kernels against the per-pass kernels at 4097^2 (CUDA events on the library's
stream): ms per Jacobi sweep for K = 2, 3, 4 fused sweeps, ms per ChebSI call (20 iterations), one state FCT step."""
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
print("tiles active:", ctx.tiles_active(), "templates:", ctx.template_count())
print("per-sweep kernel: %.4f ms/sweep" % min(ctx.bench_jacobi_sweeps(A, d_u, dt, reps=20) for _ in range(3)))
if ctx.tiles_active():
    for k in (2, 3, 4):
        print("tile kernel K=%d: %.4f ms/sweep" % (k, min(ctx.bench_jacobi_fused(A, d_u, dt, sweeps=k, reps=6) for _ in range(3))))
M, _, Md, _ = ctx.static()
b, y = ctx.array(u0), ctx.empty(n)
e0, e1 = ctx.event(), ctx.event()
ctx.chebsi(M, Md, b, y, 20)
ctx.record(e0)
for _ in range(5):
    ctx.chebsi(M, Md, b, y, 20)
ctx.record(e1)
print("ChebSI(20): %.4f ms" % (ctx.elapsed_ms(e0, e1) / 5))
nt = 4
dcc = ctx.array(np.tile(c0, nt + 1))
utr = np.zeros((nt + 1) * n); utr[:n] = u0
du = ctx.array(utr)
sw = ctx.advdrift_state(dcc, du, nt, dt)
ctx.record(e0)
sw = ctx.advdrift_state(dcc, du, nt, dt)
ctx.record(e1)
print("state FCT step: %.4f ms (%d sweeps/step)" % (ctx.elapsed_ms(e0, e1) / nt, sw // nt))
