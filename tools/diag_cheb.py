#!/usr/bin/env python
"""diagnostic: ChebSI tile kernel vs per-iteration kernel, where do they differ?"""
import os, subprocess, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
if len(sys.argv) > 3:      # child
    from fem_fct_pdeco_b200.mesh import RectMeshP1
    mesh = RectMeshP1(n, 0.0, 1.0)
    ctx = mesh.context()
    rng = np.random.default_rng(5)
    b = rng.random(mesh.nodes) - 0.5
    M, _, Md, _ = ctx.static()
    y = ctx.empty(mesh.nodes)
    ctx.chebsi(M, Md, ctx.array(b), y, iters)
    np.save(sys.argv[3], y.download())
    sys.exit(0)
outs = []
for tiles in ("1", "0"):
    env = dict(os.environ, FCT_NO_TILES=tiles)
    f = f"/tmp/diag_{tiles}.npy"
    subprocess.check_call([sys.executable, __file__, str(n), str(iters), f], env=env)
    outs.append(np.load(f))
a, bb = outs
bad = np.flatnonzero(a != bb)
lens = np.array([min(d, 2 * n - d) + 1 for d in range(2 * n + 1)])
start = np.concatenate([[0], np.cumsum(lens)])
print("n", n, "iters", iters, "mismatches", len(bad), "of", a.size, "max rel", np.abs(a - bb).max() / np.abs(a).max())
dd = np.searchsorted(start, bad, side="right") - 1
pos = bad - start[dd]
for i in range(min(30, len(bad))):
    print(int(bad[i]), "d", int(dd[i]), "pos", int(pos[i]), "len", int(lens[dd[i]]), a[bad[i]], bb[bad[i]])
if len(bad):
    print("d range", dd.min(), dd.max(), "pos-from-end min", (lens[dd] - 1 - pos).min(), "pos min", pos.min())
    import collections
    print("by d (first 20):", sorted(collections.Counter(dd.tolist()).items())[:20])
