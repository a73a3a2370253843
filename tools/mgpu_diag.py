#!/usr/bin/env python
"""Timing split of the multi-GPU FCT step (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/mgpu_diag.py [cells] [nt] [reps]
prints ms per FCT state step (max over ranks) and the launch / exchange counts.  Environment: FCT_HALO_DEPTH, FCT_P2P_DRY
(1: exchange kernels launched but empty, 2: not launched -- results wrong, timing only), FCT_TILE_KJ / FCT_TILE_KC."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from fem_fct_pdeco_b200.distributed import setup_rank
    from fem_fct_pdeco_b200.mesh import RectMeshP1
    import bench as bench_mod

    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    nt = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    mesh = RectMeshP1(cells, 0.0, 1.0)
    dt = 0.25 / cells / (2 * np.sqrt(2))
    lp, ctx = setup_rank(mesh, rank, world, local_rank)
    u0, c0 = bench_mod.synth_fields(mesh.dof_xy)
    n = lp.n
    L = (nt + 1) * n
    d_c, d_u = ctx.empty(L), ctx.empty(L)
    c_loc = lp.scatter(c0)
    for k in range(nt + 1):
        d_c.slice(k * n, n).upload(c_loc)
    d_u.slice(0, n).upload(lp.scatter(u0))
    for _ in range(2):
        sw = ctx.advdrift_state(d_c, d_u, nt, dt)
    ctx.sync()
    dist.barrier()
    torch.cuda.synchronize()
    l0, x0 = ctx.launch_count(), ctx.exchange_count()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(reps):
        sw = ctx.advdrift_state(d_c, d_u, nt, dt)
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"cells": cells, "ranks": world, "depth": lp.depth, "dry": os.environ.get("FCT_P2P_DRY", "0"),
                          "ms_per_fct_step": float(t.item()) / (reps * nt), "sweeps_per_step": sw / nt,
                          "launches_per_step": (ctx.launch_count() - l0) / (reps * nt),
                          "exchanges_per_step": (ctx.exchange_count() - x0) / (reps * nt), "local_rows": int(n)}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
