#!/usr/bin/env python
"""Determinism stress of the fused tile kernels (the stand-in for compute-sanitizer racecheck, which the GPU pool does not
allow): the fused launches (mbarrier / cp.async pipelines across tiles, shared-memory ping-pong inside a tile) are repeated on
the same input, with the CTA count capped to different values so that tile-to-CTA assignment and pipeline depth change, and
every repetition must reproduce the one-launch-per-pass kernels bit for bit.  A data race shows up as a mismatch in some run.
    python tools/stress_tiles.py [reps]      -> one line per (n, grid cap); exit code 1 on any mismatch"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_fct_pdeco_b200.mesh import RectMeshP1  # noqa: E402


def run(ctx, mesh, b, u0, c, dt, fused):
    M, _, Md, _ = ctx.static()
    y = ctx.empty(mesh.nodes)
    ctx.chebsi(M, Md, ctx.array(b), y, 20)
    A = ctx.empty(mesh.nnz)
    ctx.assemble_matrix(2, A, c0=ctx.array(c), s0=1.0, s1=1.0, scale=-1.0)
    x = ctx.empty(mesh.nodes)
    ctx.debug_jacobi_fixed(A, ctx.array(u0), dt, 16, fused, x)
    return y.download(), x.download()


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    bad = 0
    for n, caps in ((300, (0, 1, 5, 33)), (1024, (0, 37, 100)), (2048, (0, 61))):
        dt = 0.25 / n / (2 * np.sqrt(2))
        mesh = RectMeshP1(n, 0.0, 1.0)
        rng = np.random.default_rng(n)
        b = rng.random(mesh.nodes) - 0.5
        u0 = rng.random(mesh.nodes)
        c = 1.0 + rng.random(mesh.nodes)
        os.environ["FCT_NO_TILES"] = "1"
        ref = run(RectMeshP1(n, 0.0, 1.0).context(), mesh, b, u0, c, dt, 0)
        os.environ["FCT_NO_TILES"] = "0"
        for cap in caps:
            os.environ["FCT_TILE_GRID"] = str(cap)
            ctx = RectMeshP1(n, 0.0, 1.0).context()
            assert ctx.tiles_active()
            mism = 0
            for _ in range(reps):
                y, x = run(ctx, mesh, b, u0, c, dt, 4)
                mism += int(not np.array_equal(y, ref[0])) + int(not np.array_equal(x, ref[1]))
            print(f"n={n:5d} grid cap={cap:3d}: {reps} repetitions of ChebSI(20, fused 5) + Jacobi(16, fused 4), mismatches vs "
                  f"per-pass kernels: {mism}")
            bad += mism
            ctx.close()
    print("STRESS", "FAILED" if bad else "PASSED")
    return 1 if bad else 0


if __name__ == "__main__":
    raise SystemExit(main())
