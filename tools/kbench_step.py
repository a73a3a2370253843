#!/usr/bin/env python
"""Times one state FCT step of BASELINE config 5 kernel by kernel group (CUDA events): assembly, whole step."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_fct_pdeco_b200.mesh import RectMeshP1  # noqa: E402
cells = 4096
mesh = RectMeshP1(cells, 0.0, 1.0)
ctx = mesh.context(device=0)
n, nnz = mesh.nodes, mesh.nnz
dt = 0.25 / cells / (2 * np.sqrt(2))
xy = mesh.dof_xy
u0 = np.exp(-20 * ((2 * xy[:, 0] - 1 + 2 / 3) ** 2 + 5 * (2 * xy[:, 1] - 1 + 5 / 6) ** 2))
c0 = 1.0 + 0.25 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1])
nt = 4
d_c = ctx.array(np.tile(c0, nt + 1))
utr = np.zeros((nt + 1) * n); utr[:n] = u0
d_u = ctx.array(utr)
e0, e1 = ctx.event(), ctx.event()
ctx.advdrift_state(d_c, d_u, nt, dt)
best = 1e9
for _ in range(3):
    ctx.record(e0)
    ctx.advdrift_state(d_c, d_u, nt, dt)
    ctx.record(e1)
    best = min(best, ctx.elapsed_ms(e0, e1) / nt)
print(json.dumps({"knobs": {k: v for k, v in os.environ.items() if k.startswith("FCT_")}, "state_step_ms": best}))
