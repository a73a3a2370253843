"""Multi-GPU layer: contiguous row-block partition of the DoFs with a one-ring halo, one process per GPU.

The reference is single-process; this module is what the B200 build adds for BASELINE config 5 at
2/4/8 GPUs.  dolfin's CG1 DoF order is an anti-diagonal numbering, so the P1 matrix is block tridiagonal over
the diagonals and a contiguous DoF block only references a short contiguous range on either side
(SURVEY.md 8e): rank r owns global rows [R0,R1) and stores the contiguous range [G0,G1) = every column its
rows reference.  Local index = global - G0; [0,row_begin) is the halo from rank r-1, [row_end,n) from r+1.

Host logic only (numpy + torch.distributed for the rendezvous); the exchanges themselves are NCCL send/recv
issued by libfctpdeco (fct_ctx_init_comm / fct_halo_exchange).
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

from . import _lib
from ._lib import check, lib
from .context import FctContext, _hp


def partition_rows(n, world):
    """equal contiguous row blocks: bounds[r]..bounds[r+1]"""
    base, rem = divmod(int(n), int(world))
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def ring_ranges(rowptr, colidx, a, b, depth):
    """nested row ranges [lo_j, hi_j), j = 0..depth: ring 0 = [a,b), ring j = every column referenced by the rows of
    ring j-1 (columns are sorted per row, and the DoF order makes these sets contiguous ranges)"""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    lo, hi = [int(a)], [int(b)]
    for _ in range(depth):
        first = colidx[rowptr[lo[-1]:hi[-1]]]
        last = colidx[rowptr[lo[-1] + 1:hi[-1] + 1] - 1]
        lo.append(min(int(first.min()), lo[-1]))
        hi.append(max(int(last.max()) + 1, hi[-1]))
    return lo, hi


def column_ranges(rowptr, colidx, bounds, depth=1):
    """[G0,G1) per rank: the depth-ring closure of the rank's owned rows"""
    world = len(bounds) - 1
    g0 = np.empty(world, dtype=np.int64)
    g1 = np.empty(world, dtype=np.int64)
    for r in range(world):
        a, b = int(bounds[r]), int(bounds[r + 1])
        if b <= a:
            raise ValueError(f"rank {r} owns no rows: too many ranks for this mesh")
        lo, hi = ring_ranges(rowptr, colidx, a, b, depth)
        g0[r], g1[r] = lo[-1], hi[-1]
    for r in range(world):
        if r > 0 and g0[r] < bounds[r - 1]:
            raise ValueError("halo reaches beyond the neighbouring rank: row blocks are thinner than the bandwidth")
        if r + 1 < world and g1[r] > bounds[r + 2]:
            raise ValueError("halo reaches beyond the neighbouring rank: row blocks are thinner than the bandwidth")
    return g0, g1


class LocalProblem:
    """Everything rank `rank` needs to build its context from the global mesh description."""

    def __init__(self, rowptr, colidx, cells, dof_xy, rank, world, depth=1, rect_n=None):
        n = len(rowptr) - 1
        self.rect_n = rect_n          # cells per side when the global numbering is RectMeshP1's (enables the tile kernels)
        self.n_global = n
        self.rank, self.world = int(rank), int(world)
        self.depth = int(depth)
        self.bounds = partition_rows(n, world)
        self.g0_all, self.g1_all = column_ranges(rowptr, colidx, self.bounds, self.depth)
        R0, R1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        G0, G1 = int(self.g0_all[rank]), int(self.g1_all[rank])
        lo, hi = ring_ranges(rowptr, colidx, R0, R1, self.depth)
        self.ring_lo = np.array([x - G0 for x in lo], dtype=np.int32)      # local row ranges of rings 0..depth
        self.ring_hi = np.array([x - G0 for x in hi], dtype=np.int32)
        self.R0, self.R1, self.G0, self.G1 = R0, R1, G0, G1
        self.n = G1 - G0
        self.row_begin, self.row_end = R0 - G0, R1 - G0
        rp = np.asarray(rowptr, dtype=np.int64)
        k0, k1 = int(rp[G0]), int(rp[G1])
        cols = np.asarray(colidx[k0:k1], dtype=np.int64)
        rows = np.repeat(np.arange(G0, G1, dtype=np.int64), np.diff(rp[G0:G1 + 1]))
        keep = (cols >= G0) & (cols < G1)            # the outermost ring's rows are truncated to the local range
        inner = (rows >= lo[-2]) & (rows < hi[-2])   # rows of ring depth-1 (ring 0 = owned when depth == 1)
        if not keep[inner].all():
            raise AssertionError("rows of ring depth-1 must lie completely inside [G0,G1)")
        self.keep = keep                             # mask into global entries k0..k1 (value-array scatter)
        self.k0, self.k1 = k0, k1
        lrows = rows[keep] - G0
        self.colidx = (cols[keep] - G0).astype(np.int32)
        counts = np.bincount(lrows, minlength=self.n)
        self.rowptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        # cells with at least one vertex in ring depth-1 (all their vertices are neighbours of it, hence local):
        # rows of ring depth-1 are then assembled completely, rows of the outermost ring towards ring depth-1
        cells = np.asarray(cells, dtype=np.int64).reshape(-1, 3)
        own_c = ((cells >= lo[-2]) & (cells < hi[-2])).any(axis=1)
        lc = cells[own_c]
        if lc.size and (lc.min() < G0 or lc.max() >= G1):
            raise AssertionError("a cell of an owned vertex leaves the local range")
        self.cells = (lc - G0).astype(np.int32)
        self.dof_xy = np.ascontiguousarray(np.asarray(dof_xy, dtype=np.float64).reshape(-1, 2)[G0:G1])
        # what the neighbours need from this rank (local indices)
        self.send_lo = (0, 0)
        self.send_hi = (0, 0)
        if rank > 0:            # rank-1's upper halo is the global range [R0, G1[rank-1])
            self.send_lo = (R0 - G0, int(self.g1_all[rank - 1]) - G0)
        if rank + 1 < world:    # rank+1's lower halo is the global range [G0[rank+1], R1)
            self.send_hi = (int(self.g0_all[rank + 1]) - G0, R1 - G0)

    # global <-> local vectors --------------------------------------------------------------------
    def scatter(self, vec_global):
        """local copy (owned + halo entries) of a global DoF vector or time-major trajectory"""
        v = np.asarray(vec_global, dtype=np.float64)
        if v.size == self.n_global:
            return np.ascontiguousarray(v[self.G0:self.G1])
        v = v.reshape(-1, self.n_global)
        return np.ascontiguousarray(v[:, self.G0:self.G1]).ravel()

    def owned(self, vec_local):
        """owned part of a local vector / trajectory"""
        v = np.asarray(vec_local)
        if v.size == self.n:
            return v[self.row_begin:self.row_end]
        return v.reshape(-1, self.n)[:, self.row_begin:self.row_end]

    def scatter_values(self, vals_global):
        """local value array of a global value array on the global pattern"""
        return np.ascontiguousarray(np.asarray(vals_global, dtype=np.float64)[self.k0:self.k1][self.keep])

    def make_context(self, device):
        ctx = FctContext(self.rowptr, self.colidx, device=device, row_begin=self.row_begin, row_end=self.row_end)
        ctx.set_mesh(self.cells, self.dof_xy)
        if self.world > 1:
            check(lib.fct_ctx_set_rings(ctx.handle, self.depth, self.ring_lo.ctypes.data_as(C.POINTER(C.c_int32)),
                                        self.ring_hi.ctypes.data_as(C.POINTER(C.c_int32))))
        if self.rect_n:
            ctx.set_rect(self.rect_n, self.G0)
        return ctx


def init_comm(ctx, lp, broadcast_bytes):
    """create the NCCL communicator of a context.  `broadcast_bytes(b: bytes|None) -> bytes` must return rank 0's
    bytes on every rank (e.g. via torch.distributed.broadcast_object_list)."""
    buf = (C.c_ubyte * 128)()
    if lp.rank == 0:
        check(lib.fct_nccl_unique_id(buf))
    uid = broadcast_bytes(bytes(buf) if lp.rank == 0 else None)
    buf2 = (C.c_ubyte * 128).from_buffer_copy(uid)
    check(lib.fct_ctx_init_comm(ctx.handle, buf2, lp.rank, lp.world, lp.send_lo[0], lp.send_lo[1], lp.send_hi[0],
                                lp.send_hi[1]))


def init_p2p(ctx, lp, all_gather_bytes):
    """NVLink peer-memory mailboxes (fct_p2p_*): `all_gather_bytes(b) -> [bytes of rank 0, rank 1, ...]`."""
    max_halo = int(max(np.max(lp.bounds[:-1] - lp.g0_all), np.max(lp.g1_all - lp.bounds[1:]), 1))
    buf = (C.c_ubyte * 64)()
    check(lib.fct_p2p_create(ctx.handle, lp.rank, lp.world, max_halo, buf))
    handles = all_gather_bytes(bytes(buf))
    blob = (C.c_ubyte * (64 * lp.world)).from_buffer_copy(b"".join(handles))
    check(lib.fct_p2p_connect(ctx.handle, blob))


def p2p_error(ctx):
    e = C.c_int32()
    check(lib.fct_p2p_error(ctx.handle, C.byref(e)))
    return e.value


def torch_all_gatherer():
    import torch.distributed as dist

    def ag(b):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, b)
        return out
    return ag


def torch_broadcaster():
    import torch.distributed as dist

    def bc(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    return bc


def setup_rank(mesh, rank, world, local_rank, depth=None):
    """LocalProblem + context (mesh set, communicator initialised, static matrices assembled) for this rank.
    depth: halo depth in mesh rings (default 8, FCT_HALO_DEPTH overrides; 1 = exchange after every pass)"""
    if depth is None:
        depth = int(os.environ.get("FCT_HALO_DEPTH", "8"))
    lp = LocalProblem(mesh.rowptr, mesh.colidx, mesh.cells, mesh.dof_xy, rank, world, depth=depth if world > 1 else 1,
                      rect_n=getattr(mesh, "n", None))
    ctx = lp.make_context(local_rank)
    if world > 1:
        init_comm(ctx, lp, torch_broadcaster())
        if os.environ.get("FCT_NO_P2P", "0") != "1":
            init_p2p(ctx, lp, torch_all_gatherer())
    ctx.assemble_static()
    return lp, ctx


# ------------------------------------------------------------------------------------------------------
# N ranks against one rank: the row-wise summation order does not depend on the partition, so fields must agree bit for bit
# ------------------------------------------------------------------------------------------------------
def mgpu_parity_check(cells, ns, rank, world, local_rank, depth=None, seed=5):
    """Every rank runs its row block of the drift-control state / adjoint / gradient loops on a `cells`^2 mesh (deep halos,
    peer-memory or NCCL exchanges, all-rank Jacobi stopping test); rank 0 also runs the same problem on a single-GPU
    context and compares the gathered owned trajectories.  Needs an initialised torch.distributed process group.
    Returns on rank 0 a dict (u_equal, p_equal, d_equal bit-identity flags, J_rel, host_path_equal, p2p_error, ok), on
    the other ranks {"ok": <the same verdict>}."""
    import torch
    import torch.distributed as dist
    from .mesh import RectMeshP1

    mesh = RectMeshP1(cells, 0.0, 1.0)
    lp, ctx = setup_rank(mesh, rank, world, local_rank, depth=depth)
    dt = 0.25 * (1.0 / cells) / (2 * np.sqrt(2))
    xy = mesh.dof_xy
    x, y = 2 * xy[:, 0] - 1, 2 * xy[:, 1] - 1
    u0 = np.exp(-20 * ((x + 2 / 3) ** 2 + 5 * (y + 5 / 6) ** 2))
    rng = np.random.default_rng(seed)
    c = 1.0 + rng.random((ns + 1, mesh.nodes))
    uhat = np.array([np.exp(-20 * ((x - 0.1 * k + 2 / 3) ** 2 + 5 * (y - 0.1 * k + 5 / 6) ** 2)) for k in range(ns + 1)])

    def run(cx, scatter):
        utr = np.zeros((ns + 1, mesh.nodes)); utr[0] = u0
        dc, du, duh = cx.array(scatter(c.ravel())), cx.array(scatter(utr.ravel())), cx.array(scatter(uhat.ravel()))
        dp, dd = cx.empty(du.size), cx.empty(du.size)
        sw = cx.advdrift_state(dc, du, ns, dt)
        sw = (sw, cx.advdrift_adjoint(dc, du, duh, dp, ns, dt))
        cx.advdrift_gradient(dc, du, dp, dd, ns, 0.01)
        M = cx.static()[0]
        J = 0.5 * cx.norm_sq_Q(M, du, ns, dt, target=duh) + 0.005 * cx.norm_sq_Q(M, dc, ns, dt)
        out = du.download(), dp.download(), dd.download(), J, sw
        for a in (dc, du, duh, dp, dd):
            a.free()
        return out

    u, p, d, J, sw = run(ctx, lp.scatter)
    # host-trajectory entry point on the partitioned context: same local trajectory, bit for bit
    utr0 = np.zeros((ns + 1, mesh.nodes)); utr0[0] = u0
    uh = np.ascontiguousarray(lp.scatter(utr0.ravel()))
    ctx.advdrift_state_host(np.ascontiguousarray(lp.scatter(c.ravel())), uh, ns, dt)
    host_ok = bool(np.array_equal(uh, u))
    perr = p2p_error(ctx) if world > 1 else 0
    parts = [None] * world
    dist.all_gather_object(parts, [np.ascontiguousarray(lp.owned(a)) for a in (u, p, d)])
    flags = torch.tensor([1 if host_ok else 0, 1 if perr == 0 else 0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    res = {"ok": True}
    if rank == 0:
        ug, pg, dg = [np.concatenate([parts[r][i] for r in range(world)], axis=1) for i in range(3)]
        ctx1 = mesh.context(device=local_rank)
        u1, p1, d1, J1, sw1 = run(ctx1, lambda a: a)
        u1, p1, d1 = [a.reshape(ns + 1, -1) for a in (u1, p1, d1)]
        res = {"cells": cells, "time_steps": ns, "ranks": world, "halo_depth": lp.depth,
               "u_equal": bool(np.array_equal(ug, u1)), "p_equal": bool(np.array_equal(pg, p1)),
               "d_equal": bool(np.array_equal(dg, d1)), "J_rel": abs(J / J1 - 1), "sweeps": [int(sw[0]), int(sw1[0])],
               "adjoint_sweeps": [int(sw[1]), int(sw1[1])],
               "halo_rows": [int(lp.row_begin), int(lp.n - lp.row_end)],
               "host_path_equal": bool(int(flags[0].item())), "p2p_error_free": bool(int(flags[1].item()))}
        # The row arithmetic does not depend on the partition, so N ranks and one GPU give the same bits whenever their
        # low-order solves stop after the same sweeps -- the default configuration (peer mailboxes, fused tile sweeps of the
        # same depth on both sides).  A fallback configuration that tests convergence at other sweep counts (NCCL + graph:
        # per-sweep kernels; halo depth < 4: shallower fused launches) may stop elsewhere: then the fields agree to the solver
        # tolerance instead, and the check says which of the two it verified.
        rel = [float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) for a, b in ((ug, u1), (pg, p1), (dg, d1))]
        res["max_rel"] = max(rel)
        res["bit_identical"] = bool(res["u_equal"] and res["p_equal"] and res["d_equal"])
        # same solver path on both sides?  fused tile sweeps of the same depth (N ranks: only with peer mailboxes or without the
        # CUDA-graph loop, and never deeper than the halo), or per-sweep kernels on both
        kj = int(os.environ.get("FCT_TILE_KJ", "4"))
        fused1 = bool(ctx1.tiles_active())
        fusedN = bool(ctx.tiles_active()) and (os.environ.get("FCT_NO_P2P", "0") != "1" or os.environ.get("FCT_NO_GRAPH", "0") == "1")
        same_path = (fused1 and fusedN and min(kj, lp.depth) == kj) or (not fused1 and not fusedN)
        res["same_solver_path"] = bool(same_path)
        fields_ok = res["bit_identical"] if same_path else max(rel) <= 1e-12
        res["verified"] = "bit-identical" if res["bit_identical"] else "1e-12 (different stopping points)"
        res["ok"] = bool(fields_ok and res["J_rel"] < 1e-13 and res["host_path_equal"] and res["p2p_error_free"])
        if not res["ok"]:          # where do the fields differ?  (time level, global row, anti-diagonal, position, |diff|)
            n_c = cells
            lens = np.array([min(dd, 2 * n_c - dd) + 1 for dd in range(2 * n_c + 1)])
            start = np.concatenate([[0], np.cumsum(lens)])
            for name, a, b in (("u", ug, u1), ("p", pg, p1), ("d", dg, d1)):
                bad = np.argwhere(a != b)
                info = []
                for k, r in bad[:6]:
                    dd = int(np.searchsorted(start, r, side="right") - 1)
                    info.append([int(k), int(r), dd, int(r - start[dd]), float(abs(a[k, r] - b[k, r]))])
                res[f"{name}_mismatch"] = {"count": int(len(bad)), "first": info,
                                           "max_rel": float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))}
        ctx1.close()
        mesh._ctx = None
    v = torch.tensor([1 if res["ok"] else 0], device="cuda")
    dist.broadcast(v, src=0)
    res["ok"] = bool(int(v.item()))
    ctx.sync()
    dist.barrier()
    ctx.close()
    return res


# ------------------------------------------------------------------------------------------------------
# bench.py, N > 1: strong scaling of BASELINE config 5 (one process per GPU, launched by torchrun)
# ------------------------------------------------------------------------------------------------------
def bench_multi(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from .mesh import RectMeshP1

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    import bench as bench_mod            # the repo-root bench.py (synthetic problem + JSON helpers)

    # N ranks vs 1 rank, bit for bit, on a mesh the single-GPU twin finishes in a moment (driver-visible parity of the
    # multi-GPU path; the full-size single-GPU run carries the parity against the oracle)
    mg = mgpu_parity_check(min(args.cells, 1024), 3, rank, world, local_rank)

    n_cells, nt = args.cells, args.nt
    h = 1.0 / n_cells
    dt = 0.25 * h / (2 * np.sqrt(2))
    mesh = RectMeshP1(n_cells, 0.0, 1.0)
    lp, ctx = setup_rank(mesh, rank, world, local_rank)
    u0, c0 = bench_mod.synth_fields(mesh.dof_xy)
    n_glob, nnz_glob, ncell = mesh.nodes, mesh.nnz, mesh.ncells
    n = lp.n
    L = (nt + 1) * n
    d_c, d_u, d_uhat = ctx.empty(L), ctx.empty(L), ctx.empty(L)
    xy_loc = np.ascontiguousarray(mesh.dof_xy[lp.G0:lp.G1])
    c_loc, u0_loc = lp.scatter(c0), lp.scatter(u0)
    for k in range(nt + 1):
        d_c.slice(k * n, n).upload(c_loc)
        d_uhat.slice(k * n, n).upload(bench_mod.synth_target_slice(xy_loc, k, dt))
    d_u.slice(0, n).upload(u0_loc)
    del mesh
    it = bench_mod.GradientIteration(ctx, n, nt, dt, d_c, d_u, d_uhat)

    for _ in range(args.warmup):
        it()
    ctx.sync()
    dist.barrier()
    torch.cuda.synchronize()
    sampler = bench_mod.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    del it.sweeps[:]
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(args.steps):
        J = it()
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1)
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = ctx.launch_count() - l0
    lt = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    xch = ctx.exchange_count()
    # e2e: the state sweep through the host-buffer entry point, every rank streaming its own local range (pinned host
    # memory; control slices H2D, state slices D2H inside the timed region); wall clock between barriers, max over ranks
    nt_e2e = min(nt, args.e2e_nt)
    Le = (nt_e2e + 1) * n
    hc, hu = ctx.pinned(Le), ctx.pinned(Le)
    hc.reshape(nt_e2e + 1, n)[:] = c_loc
    hu[:] = 0.0
    hu[:n] = u0_loc
    ctx.advdrift_state_host(hc, hu, nt_e2e, dt)          # warm-up
    reps_e2e = max(1, args.steps // 2)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps_e2e):
        ctx.advdrift_state_host(hc, hu, nt_e2e, dt)      # synchronises before returning
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, float(8 * n)], dtype=torch.float64, device="cuda")
    dist.all_reduce(te[:1], op=dist.ReduceOp.MAX)
    dist.all_reduce(te[1:], op=dist.ReduceOp.SUM)
    e2e_value = nt_e2e * reps_e2e / float(te[0].item())
    e2e_bytes = int(te[1].item())
    # what the box's host<->device path gives the N ranks together when they do nothing but the copies of the e2e sweep (one
    # H2D and one D2H of a local slice per time level, two streams, pinned memory): the ceiling of `e2e` at this N
    th_in, th_out = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
    td_in, td_out = torch.empty(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    for rep in range(2):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(nt_e2e):
            with torch.cuda.stream(s_in):
                td_in.copy_(th_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                th_out.copy_(td_out, non_blocking=True)
        torch.cuda.synchronize()
        probe_s = time.perf_counter() - t0
    tp = torch.tensor([probe_s], dtype=torch.float64, device="cuda")
    dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    probe_steps_per_s = nt_e2e / float(tp.item())
    del th_in, th_out, td_in, td_out
    perr = p2p_error(ctx)
    pe = torch.tensor([perr], device="cuda")
    dist.all_reduce(pe, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    rc = 0
    if rank == 0:
        fct_steps = it.fct_steps_per_pass * args.steps
        value = fct_steps / (ms * 1e-3)
        k_mean = float(np.mean([sum(s) / (3.0 * nt) for s in it.sweeps]))
        peak, peak_src = bench_mod._peaks()
        step_gb = bench_mod.step_bytes(n_glob, nnz_glob, ncell, k_mean) / 1e9
        line = {
            "metric": "FCT steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": bench_mod.workload_string(n_cells),
                       "partition": f"row blocks over {world} GPUs, depth-{lp.depth} halo rings, NVLink peer mailboxes",
                       "bench_step": f"one projected-gradient iteration over {nt} time levels: state + adjoint sweeps + gradient + "
                                     f"one Armijo trial + cost = {it.fct_steps_per_pass} FCT steps, device-resident",
                       "time_levels": nt, "dt": dt, "jacobi_sweeps_per_step": k_mean, "cost_functional": J,
                       "halo_rows": [int(lp.row_begin), int(lp.n - lp.row_end)],
                       "l2_flush": ("per-GPU working set per FCT step >> L2" if 8 * nnz_glob / world > 200e6 else
                                    "per-GPU value arrays fit the 126 MB L2 at this N: kernels run partly from L2")},
            "roofline": {"bound": "hbm", "kernel": "whole FCT step (all kernels), SURVEY App. E accounting bytes",
                         "achieved": step_gb * value / world, "peak": peak, "unit": "GB/s",
                         "frac": step_gb * value / world / peak, "traffic": None, "peak_source": peak_src,
                         "note": "per-GPU: App. E accounting bytes of an FCT step x steps/s / N (the single-GPU line carries "
                                 "the per-kernel roofline on actual bytes)"},
            "mgpu_parity": mg, "mgpu_bit_identical": bool(mg.get("ok") and mg.get("bit_identical")),
            "exchanges_per_fct_step": xch / max(1, (args.warmup + args.steps) * it.fct_steps_per_pass) if xch else None,
            "launches_per_fct_step_per_rank": lt.item() / world / fct_steps,
            "p2p_error": int(pe.item()),
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": e2e_bytes, "d2h_bytes_per_step": e2e_bytes,
                    "time_levels": nt_e2e, "host_link_GBs": e2e_value * 2 * e2e_bytes / 1e9,
                    "copy_only_steps_per_s": probe_steps_per_s, "copy_only_GBs": probe_steps_per_s * 2 * e2e_bytes / 1e9,
                    "copy_only_note": "the same per-level H2D + D2H of every rank's local slice without any compute (pinned "
                                      "memory, two streams per rank): the host-side ceiling of e2e at this N",
                    "call": "fct_advdrift_state_host on every rank (state sweep, pinned host trajectories of the rank's "
                            "local range)"},
            "gpu_launches": int(lt.item()),
            "clocks": clocks,
        }
        print(json.dumps(line))
        if not mg.get("ok") or int(pe.item()):
            sys.stderr.write("bench.py: multi-GPU parity check FAILED\n")
            rc = 3
    dist.barrier()
    dist.destroy_process_group()
    if rc:
        raise SystemExit(rc)
