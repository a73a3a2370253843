#!/usr/bin/env python
"""Kernel micro-bench on the 4096^2 mesh (tuning aid, not the bench contract): times the Jacobi sweep of the low-order
solve and one Chebyshev iteration with CUDA events on the library's stream.  Env knobs are read at context creation, so
run one process per variant:  FCT_NST_JTPL=4 python tools/kbench.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_fct_pdeco_b200.mesh import RectMeshP1  # noqa: E402


def main():
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    mesh = RectMeshP1(cells, 0.0, 1.0)
    ctx = mesh.context(device=0)
    n, nnz = mesh.nodes, mesh.nnz
    h = 1.0 / cells
    dt = 0.25 * h / (2 * np.sqrt(2))
    xy = mesh.dof_xy
    u0 = np.exp(-20 * ((2 * xy[:, 0] - 1 + 2 / 3) ** 2 + 5 * (2 * xy[:, 1] - 1 + 5 / 6) ** 2))
    c0 = 1.0 + 0.25 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1])
    d_c, d_u = ctx.array(c0), ctx.array(u0)
    A = ctx.empty(nnz)
    ctx.assemble_matrix(2, A, c0=d_c, s0=1.0, s1=1.0, scale=-1.0)
    jac = min(ctx.bench_jacobi_sweeps(A, d_u, dt, reps=20) for _ in range(3))
    M, _, Md, _ = ctx.static()
    b, y = ctx.array(u0), ctx.empty(n)
    e0, e1 = ctx.event(), ctx.event()
    ctx.chebsi(M, Md, b, y, 20)
    best = 1e9
    for _ in range(3):
        ctx.record(e0)
        for _ in range(5):
            ctx.chebsi(M, Md, b, y, 20)
        ctx.record(e1)
        t20 = ctx.elapsed_ms(e0, e1)
        ctx.record(e0)
        for _ in range(5):
            ctx.chebsi(M, Md, b, y, 1)
        ctx.record(e1)
        t1 = ctx.elapsed_ms(e0, e1)
        best = min(best, (t20 - t1) / (5 * 19))
    best20 = 1e9
    for _ in range(3):
        ctx.record(e0)
        for _ in range(5):
            ctx.chebsi(M, Md, b, y, 20)
        ctx.record(e1)
        best20 = min(best20, ctx.elapsed_ms(e0, e1) / 5)
    try:
        fused = min(ctx.bench_jacobi_fused(A, d_u, dt, sweeps=14, reps=5) for _ in range(3))
    except Exception as e:  # noqa: BLE001
        fused = str(e)[:60]
    knobs = {k: v for k, v in os.environ.items() if k.startswith("FCT_")}
    print(json.dumps({"knobs": knobs, "jacobi_sweep_ms": jac, "jacobi_fused_ms_per_sweep": fused, "cheb_iter_ms": best,
                      "chebsi20_ms": best20, "templates": ctx.template_count()}))


if __name__ == "__main__":
    main()
