#!/usr/bin/env python
"""the second-species system of the chemotaxis problem, M + dt (Df Ad + delta M) on [0,16]^2 with the
reference's parameters (chemotaxis_mimura_FCT_PGD.py:39-55; the solve is helpers.py:1342), solved to 1e-13 by Jacobi-PCG and by
Chebyshev-polynomial preconditioned CG: outer iterations, matrix passes and time against the mesh size.
    python tools/pcg_table.py [n ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_fct_pdeco_b200 import _lib  # noqa: E402
from fem_fct_pdeco_b200.mesh import RectMeshP1  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [80, 320, 1024, 2048]
    dt, Df, delta = 0.1, 1.0, 32.0
    print(f"{'n':>6} {'DoF':>9} | {'Jacobi-PCG its':>14} {'ms':>8} | {'Cheb(8)-PCG its':>15} {'passes':>7} {'ms':>8} | rel. diff")
    for n in sizes:
        mesh = RectMeshP1(n, 0.0, 16.0)
        ctx = mesh.context()
        M, _, _, K = ctx.static()
        A = ctx.empty(mesh.nnz)
        ctx.vals_axpby(1.0 + dt * delta, M, dt * Df, K, A)
        rng = np.random.default_rng(n)
        xt = ctx.array(rng.standard_normal(mesh.nodes))
        b = ctx.empty(mesh.nodes)
        ctx.spmv(A, xt, b)
        res = {}
        for kind in (_lib.SOLVER_PCG, _lib.SOLVER_CHEB_PCG):
            for rep in range(2):
                x = ctx.array(np.zeros(mesh.nodes))
                ctx.sync()
                t0 = time.perf_counter()
                its, _ = ctx.solve(kind, A, b, x, rtol=1e-13, maxit=20000)
                ctx.sync()
                ms = (time.perf_counter() - t0) * 1e3
            res[kind] = (its, ms, x.download())
        a, c = res[_lib.SOLVER_PCG], res[_lib.SOLVER_CHEB_PCG]
        diff = np.linalg.norm(a[2] - c[2]) / np.linalg.norm(a[2])
        print(f"{n:6d} {mesh.nodes:9d} | {a[0]:14d} {a[1]:8.2f} | {c[0]:15d} {c[0] * 9:7d} {c[1]:8.2f} | {diff:.1e}")
        ctx.close()


if __name__ == "__main__":
    main()
