#!/usr/bin/env python
"""ncu driver: one state FCT step + one adjoint FCT step + one gradient slice of BASELINE config 5 (4096^2 cells) inside a
cudaProfilerStart/Stop range, after one untimed pass.  Run as
    ncu --set full --clock-control none --import-source on --profile-from-start off -o prof python tools/kprof_step.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_fct_pdeco_b200 import _lib  # noqa: E402
from fem_fct_pdeco_b200.mesh import RectMeshP1  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nt = 1
mesh = RectMeshP1(cells, 0.0, 1.0)
ctx = mesh.context(device=0)
n = mesh.nodes
dt = 0.25 / cells / (2 * np.sqrt(2))
xy = mesh.dof_xy
u0 = np.exp(-20 * ((2 * xy[:, 0] - 1 + 2 / 3) ** 2 + 5 * (2 * xy[:, 1] - 1 + 5 / 6) ** 2))
c0 = 1.0 + 0.25 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1])
L = (nt + 1) * n
d_c = ctx.array(np.tile(c0, nt + 1))
utr = np.zeros(L); utr[:n] = u0
d_u, d_uhat = ctx.array(utr), ctx.array(np.tile(u0, nt + 1))
d_p, d_d = ctx.empty(L), ctx.empty(L)
M = ctx.static()[0]


def one_pass():
    ctx.advdrift_state(d_c, d_u, nt, dt)
    ctx.advdrift_adjoint(d_c, d_u, d_uhat, d_p, nt, dt)
    ctx.advdrift_gradient(d_c, d_u, d_p, d_d, nt, 0.01)
    return ctx.norm_sq_Q(M, d_u, nt, dt, target=d_uhat)


one_pass()
ctx.sync()
_lib.lib.fct_profiler_range(1)
one_pass()
ctx.sync()
_lib.lib.fct_profiler_range(0)
