"""CPU test of the overlapped-tile scheme of csrc/fct_tile.cu (no GPU): a numpy emulation of what one fused launch does --
region = tile interior + K-wide frame in (diagonal, position) space, pass s updates the rows at least s away from the
region's edge, only the interior is written back -- must reproduce K global Jacobi sweeps on every owned row, for the
library's own tile lists (fct_debug_tile_list, host code), single block and row-block partitions with truncated halo
rows.  It also asserts the two facts the kernel relies on: every neighbour of (d, pos) lies in (d +- 1, pos +- 1), and
every owned row is the interior row of exactly one tile."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from fem_fct_pdeco_b200._lib import check, lib
from fem_fct_pdeco_b200.distributed import LocalProblem
from oracle.p1mesh import RectMesh


def tile_list(n, g0, rb, re, K):
    cnt, geom = C.c_int32(), (C.c_int32 * 2)()
    check(lib.fct_debug_tile_list(n, g0, rb, re, K, None, 0, C.byref(cnt), geom))
    out = np.zeros(2 * max(cnt.value, 1), dtype=np.int32)
    check(lib.fct_debug_tile_list(n, g0, rb, re, K, out.ctypes.data_as(C.c_void_p), cnt.value, C.byref(cnt), geom))
    return out[: 2 * cnt.value].reshape(-1, 2), geom[0], geom[1]


def numbering(n):
    """(d, pos) of every global row of the anti-diagonal numbering, and start(d)"""
    lens = np.array([min(d, 2 * n - d) + 1 for d in range(2 * n + 1)])
    start = np.concatenate([[0], np.cumsum(lens)])
    d_of = np.repeat(np.arange(2 * n + 1), lens)
    pos_of = np.arange(start[-1]) - start[d_of]
    return d_of, pos_of, start, lens


def emulate_launch(n, K, g0, rowptr, colidx, vals, b, x, rb, re):
    """one fused launch on the local block (rows g0 .. g0+nloc): returns x_out on the owned rows [rb, re) and how often
    each owned row was written"""
    nloc = len(rowptr) - 1
    d_of, pos_of, start, lens = numbering(n)
    tiles, ND, NP = tile_list(n, g0, rb, re, K)
    out = np.full(nloc, np.nan)
    written = np.zeros(nloc, dtype=int)
    for d0, p0 in tiles:
        dlo, plo = d0 - K, p0 - K
        reg = -np.ones((ND, NP), dtype=np.int64)           # local row of every region slot, -1 = does not exist
        for dl in range(ND):
            d = dlo + dl
            if d < 0 or d > 2 * n:
                continue
            for pl in range(NP):
                pos = plo + pl
                if 0 <= pos < lens[d]:
                    r = start[d] + pos - g0
                    if 0 <= r < nloc:
                        reg[dl, pl] = r
        cur = np.where(reg >= 0, x[np.maximum(reg, 0)], np.nan)
        for s in range(1, K + 1):
            new = cur.copy()
            for dl in range(s, ND - s):
                for pl in range(s, NP - s):
                    r = reg[dl, pl]
                    if r < 0:
                        continue
                    acc = 0.0
                    for k in range(rowptr[r], rowptr[r + 1]):
                        cgl = colidx[k] + g0
                        ddl, ppl = d_of[cgl] - dlo, pos_of[cgl] - plo
                        assert abs(d_of[cgl] - d_of[r + g0]) <= 1 and abs(pos_of[cgl] - pos_of[r + g0]) <= 1
                        assert 0 <= ddl < ND and 0 <= ppl < NP and reg[ddl, ppl] == colidx[k]
                        acc += vals[k] * cur[ddl, ppl]
                    new[dl, pl] = b[r] - acc
            cur = new
        for dl in range(K, ND - K):
            for pl in range(K, NP - K):
                r = reg[dl, pl]
                if r >= 0 and rb <= r < re:
                    out[r] = cur[dl, pl]
                    written[r] += 1
    return out, written


def jacobi_system(n, seed=0):
    """a row-scaled strictly diagonally dominant system on the P1 pattern (zero diagonal slot, like k_low_build writes it)"""
    m = RectMesh(n, 0.0, 1.0)
    rowptr, colidx = m.pattern()
    rng = np.random.default_rng(seed)
    vals = -rng.random(len(colidx)) / 8.0
    rows = np.repeat(np.arange(m.nodes), np.diff(rowptr))
    vals[colidx == rows] = 0.0
    return m, rowptr.astype(np.int64), colidx.astype(np.int64), vals, rng.random(m.nodes), rng.random(m.nodes)


@pytest.mark.parametrize("n,K", [(3, 2), (9, 4), (41, 4), (41, 5), (50, 3)])
def test_tiles_reproduce_global_sweeps_single_block(n, K):
    m, rowptr, colidx, vals, b, x = jacobi_system(n)
    A = sp.csr_matrix((vals, colidx, rowptr), shape=(m.nodes, m.nodes))
    ref = x.copy()
    for _ in range(K):
        ref = b - A @ ref
    out, written = emulate_launch(n, K, 0, rowptr, colidx, vals, b, x, 0, m.nodes)
    assert np.all(written == 1)
    assert np.allclose(out, ref, rtol=0, atol=1e-15)


@pytest.mark.parametrize("n,world,K", [(40, 2, 4), (40, 3, 2), (64, 4, 4)])
def test_tiles_on_row_blocks_with_truncated_halo_rows(n, world, K):
    """depth-K halo rings: after K fused passes on the local (truncated) block every OWNED row equals the global sweeps"""
    m, rowptr, colidx, vals, b, x = jacobi_system(n, seed=1)
    A = sp.csr_matrix((vals, colidx, rowptr), shape=(m.nodes, m.nodes))
    ref = x.copy()
    for _ in range(K):
        ref = b - A @ ref
    for rank in range(world):
        lp = LocalProblem(rowptr.astype(np.int32), colidx.astype(np.int32), m.cells, m.dof_xy, rank, world, depth=K, rect_n=n)
        lv = lp.scatter_values(vals)
        out, written = emulate_launch(n, K, lp.G0, lp.rowptr.astype(np.int64), lp.colidx.astype(np.int64), lv,
                                      lp.scatter(b), lp.scatter(x), lp.row_begin, lp.row_end)
        own = slice(lp.row_begin, lp.row_end)
        assert np.all(written[own] == 1) and written.sum() == lp.row_end - lp.row_begin
        assert np.allclose(out[own], ref[lp.R0:lp.R1], rtol=0, atol=1e-15)


@pytest.mark.parametrize("n,world,K", [(40, 2, 4), (48, 3, 3)])
def test_two_fused_launches_share_one_exchange_on_deep_halos(n, world, K):
    """halo depth 2K (csrc/fct_tile.cu: tile_out_range): the first launch writes ring depth-K = every row whose K-ring lies
    inside the local range, the second one follows WITHOUT an exchange and must give the owned rows of 2K global sweeps.
    Rows the first launch does not write stay NaN here, so any read of a stale row by a row that matters would show."""
    m, rowptr, colidx, vals, b, x = jacobi_system(n, seed=2)
    A = sp.csr_matrix((vals, colidx, rowptr), shape=(m.nodes, m.nodes))
    ref = x.copy()
    for _ in range(2 * K):
        ref = b - A @ ref
    for rank in range(world):
        lp = LocalProblem(rowptr.astype(np.int32), colidx.astype(np.int32), m.cells, m.dof_xy, rank, world, depth=2 * K, rect_n=n)
        lr, lc, lv = lp.rowptr.astype(np.int64), lp.colidx.astype(np.int64), lp.scatter_values(vals)
        mid_rb, mid_re = int(lp.ring_lo[K]), int(lp.ring_hi[K])              # ring depth-K of a depth-2K halo
        mid, _ = emulate_launch(n, K, lp.G0, lr, lc, lv, lp.scatter(b), lp.scatter(x), mid_rb, mid_re)
        assert not np.isnan(mid[mid_rb:mid_re]).any()
        out, written = emulate_launch(n, K, lp.G0, lr, lc, lv, lp.scatter(b), mid, lp.row_begin, lp.row_end)
        own = slice(lp.row_begin, lp.row_end)
        assert np.all(written[own] == 1)
        assert np.allclose(out[own], ref[lp.R0:lp.R1], rtol=0, atol=1e-15)
