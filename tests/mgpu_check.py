"""Multi-GPU parity check, launched as one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/mgpu_check.py [cells]

Every rank runs its row block of the drift-control state / adjoint / gradient loops (NCCL halo exchange,
all-reduced Jacobi stopping test); rank 0 also runs the same problem on a single-GPU context and compares the
gathered trajectories: the row-wise summation order does not depend on the partition, so fields must agree to the
last bit; the cost functional (a reduction) to 1e-13."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from fem_fct_pdeco_b200.distributed import setup_rank
    from fem_fct_pdeco_b200.mesh import RectMeshP1

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    ns = 4
    mesh = RectMeshP1(cells, 0.0, 1.0)
    lp, ctx = setup_rank(mesh, rank, world, local_rank)
    dt = 0.25 * (1.0 / cells) / (2 * np.sqrt(2))
    xy = mesh.dof_xy
    x, y = 2 * xy[:, 0] - 1, 2 * xy[:, 1] - 1
    u0 = np.exp(-20 * ((x + 2 / 3) ** 2 + 5 * (y + 5 / 6) ** 2))
    rng = np.random.default_rng(5)
    c = 1.0 + rng.random((ns + 1, mesh.nodes))
    uhat = np.array([np.exp(-20 * ((x - 0.1 * k + 2 / 3) ** 2 + 5 * (y - 0.1 * k + 5 / 6) ** 2)) for k in range(ns + 1)])

    def run(ctx, scatter):
        utr = np.zeros((ns + 1, mesh.nodes)); utr[0] = u0
        dc, du, duh = ctx.array(scatter(c.ravel())), ctx.array(scatter(utr.ravel())), ctx.array(scatter(uhat.ravel()))
        dp, dd = ctx.empty(du.size), ctx.empty(du.size)
        sw = ctx.advdrift_state(dc, du, ns, dt)
        ctx.advdrift_adjoint(dc, du, duh, dp, ns, dt)
        ctx.advdrift_gradient(dc, du, dp, dd, ns, 0.01)
        M = ctx.static()[0]
        J = 0.5 * ctx.norm_sq_Q(M, du, ns, dt, target=duh) + 0.005 * ctx.norm_sq_Q(M, dc, ns, dt)
        return du.download(), dp.download(), dd.download(), J, sw

    u, p, d, J, sw = run(ctx, lp.scatter)
    # host-trajectory entry point on the partitioned context: same local trajectory, bit for bit
    utr0 = np.zeros((ns + 1, mesh.nodes)); utr0[0] = u0
    uh = np.ascontiguousarray(lp.scatter(utr0.ravel()))
    ctx.advdrift_state_host(np.ascontiguousarray(lp.scatter(c.ravel())), uh, ns, dt)
    host_ok = bool(np.array_equal(uh, u))
    parts = [None] * world
    dist.all_gather_object(parts, [np.ascontiguousarray(lp.owned(a)) for a in (u, p, d)])
    ok = True
    if rank == 0:
        ug, pg, dg = [np.concatenate([parts[r][i] for r in range(world)], axis=1) for i in range(3)]
        ctx1 = mesh.context(device=local_rank)
        u1, p1, d1, J1, sw1 = run(ctx1, lambda a: a)
        u1, p1, d1 = [a.reshape(ns + 1, -1) for a in (u1, p1, d1)]
        res = {"u_equal": bool(np.array_equal(ug, u1)), "p_equal": bool(np.array_equal(pg, p1)),
               "d_equal": bool(np.array_equal(dg, d1)), "J_rel": abs(J / J1 - 1), "sweeps": (sw, sw1),
               "halo": (lp.row_begin, lp.n - lp.row_end)}
        print("MGPU_CHECK", world, res)
        ok = res["u_equal"] and res["p_equal"] and res["d_equal"] and res["J_rel"] < 1e-13
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    hflag = torch.tensor([1 if host_ok else 0], device="cuda")
    dist.all_reduce(hflag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_CHECK host-trajectory path equal on all ranks:", bool(int(hflag.item())))
    flag = flag * hflag
    dist.barrier()
    dist.destroy_process_group()
    if not int(flag.item()):
        raise SystemExit(1)
    if rank == 0:
        print("MGPU_CHECK PASSED")


if __name__ == "__main__":
    main()
