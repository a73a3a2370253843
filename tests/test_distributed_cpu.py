"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU host logic: row-block partition, local patterns with
one-ring halo, send/recv ranges, scatter/gather of vectors and value arrays, and the rendezvous plumbing that
carries the NCCL unique id.  The exchanges are replayed with torch.distributed (gloo) send/recv on CPU tensors
using exactly the ranges the CUDA library is given."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from fem_fct_pdeco_b200.distributed import LocalProblem, column_ranges, partition_rows, torch_broadcaster
from fem_fct_pdeco_b200.mesh import RectMeshP1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_rows_and_ranges():
    b = partition_rows(10, 3)
    assert list(b) == [0, 4, 7, 10]
    m = RectMeshP1(40)
    bounds = partition_rows(m.nodes, 4)
    g0, g1 = column_ranges(m.rowptr, m.colidx, bounds)
    for r in range(4):
        assert g0[r] <= bounds[r] and g1[r] >= bounds[r + 1]
        # halo is about one mesh diagonal on each side
        assert bounds[r] - g0[r] <= 43 and g1[r] - bounds[r + 1] <= 43
    with pytest.raises(ValueError):
        column_ranges(m.rowptr, m.colidx, partition_rows(m.nodes, 200))     # blocks thinner than the bandwidth


@pytest.mark.parametrize("world", [2, 3])
def test_local_problems_reproduce_global_spmv(world):
    m = RectMeshP1(14, -1.0, 1.0)
    rng = np.random.default_rng(0)
    vals = rng.standard_normal(m.nnz)
    A = sp.csr_matrix((vals, m.colidx, m.rowptr), shape=(m.nodes, m.nodes))
    x = rng.standard_normal(m.nodes)
    y = A @ x
    covered = np.zeros(m.nodes, dtype=int)
    for rank in range(world):
        lp = LocalProblem(m.rowptr, m.colidx, m.cells, m.dof_xy, rank, world)
        Al = sp.csr_matrix((lp.scatter_values(vals), lp.colidx, lp.rowptr), shape=(lp.n, lp.n))
        yl = Al @ lp.scatter(x)
        assert np.array_equal(lp.owned(yl), y[lp.R0:lp.R1])          # same entries, same order: bit-identical
        covered[lp.R0:lp.R1] += 1
        # every local row keeps its diagonal; owned rows are complete
        for r in range(lp.n):
            assert r in lp.colidx[lp.rowptr[r]:lp.rowptr[r + 1]]
        # local cells: exactly the cells touching an owned vertex, all vertices local
        gc = m.cells.astype(np.int64)
        assert lp.cells.shape[0] == int(((gc >= lp.R0) & (gc < lp.R1)).any(axis=1).sum())
        assert lp.cells.min() >= 0 and lp.cells.max() < lp.n
        # trajectories scatter slice-wise
        tr = rng.standard_normal((3, m.nodes))
        assert np.array_equal(lp.scatter(tr.ravel()).reshape(3, -1), tr[:, lp.G0:lp.G1])
    assert (covered == 1).all()


def _worker(rank, world, port, n_cells, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = RectMeshP1(n_cells)
        lp = LocalProblem(m.rowptr, m.colidx, m.cells, m.dof_xy, rank, world)
        # the unique-id rendezvous used by init_comm
        uid = torch_broadcaster()(bytes(range(128)) if rank == 0 else None)
        assert uid == bytes(range(128))
        # halo exchange replayed with the library's ranges: owned entries valid, halo entries poisoned
        x = np.arange(m.nodes, dtype=np.float64) * 1.5 + 0.25
        loc = lp.scatter(x).copy()
        loc[:lp.row_begin] = np.nan
        loc[lp.row_end:] = np.nan
        t = torch.from_numpy(loc)
        reqs = []
        if rank > 0:
            reqs.append(dist.isend(t[lp.send_lo[0]:lp.send_lo[1]].clone(), rank - 1))
            reqs.append(dist.irecv(t[0:lp.row_begin], rank - 1))
        if rank + 1 < world:
            reqs.append(dist.isend(t[lp.send_hi[0]:lp.send_hi[1]].clone(), rank + 1))
            reqs.append(dist.irecv(t[lp.row_end:lp.n], rank + 1))
        for q in reqs:
            q.wait()
        assert np.array_equal(t.numpy(), x[lp.G0:lp.G1])
        # gather of the owned parts rebuilds the global vector on rank 0
        parts = [None] * world
        dist.all_gather_object(parts, lp.owned(t.numpy()).copy())
        assert np.array_equal(np.concatenate(parts), x)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_ranges_gloo(world):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, 24, ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world))


@pytest.mark.parametrize("depth", [2, 4, 8])
def test_deep_halo_rings(depth):
    """depth-k halos: nested ring ranges, rows of ring k-1 complete, local cells = cells touching ring k-1,
    send ranges cover the neighbour's (k rings wide) halo with owned rows only"""
    m = RectMeshP1(30)
    world = 3
    lps = [LocalProblem(m.rowptr, m.colidx, m.cells, m.dof_xy, r, world, depth=depth) for r in range(world)]
    A = sp.csr_matrix((np.ones(m.nnz), m.colidx, m.rowptr), shape=(m.nodes, m.nodes))
    for r, lp in enumerate(lps):
        assert lp.ring_lo[0] == lp.row_begin and lp.ring_hi[0] == lp.row_end
        assert lp.ring_lo[depth] == 0 and lp.ring_hi[depth] == lp.n
        for j in range(1, depth + 1):
            assert lp.ring_lo[j] <= lp.ring_lo[j - 1] and lp.ring_hi[j] >= lp.ring_hi[j - 1]
            # ring j = exactly the columns referenced by ring j-1 (graph distance j from the owned rows)
            rows = np.arange(lp.ring_lo[j - 1], lp.ring_hi[j - 1]) + lp.G0
            cols = A[rows].indices
            assert cols.min() == lp.ring_lo[j] + lp.G0 and cols.max() + 1 == lp.ring_hi[j] + lp.G0
        # rows of ring depth-1 keep all their entries
        for row in (lp.ring_lo[depth - 1], lp.ring_hi[depth - 1] - 1):
            assert lp.rowptr[row + 1] - lp.rowptr[row] == m.rowptr[row + lp.G0 + 1] - m.rowptr[row + lp.G0]
        if r > 0:      # what I send down is exactly the lower neighbour's upper halo, all of it rows I own
            prev = lps[r - 1]
            assert lp.send_lo[0] == lp.row_begin and lp.send_lo[1] - lp.send_lo[0] == prev.n - prev.row_end
            assert lp.send_lo[1] <= lp.row_end
        if r + 1 < world:
            nxt = lps[r + 1]
            assert lp.send_hi[1] == lp.row_end and lp.send_hi[1] - lp.send_hi[0] == nxt.row_begin
            assert lp.send_hi[0] >= lp.row_begin
