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
