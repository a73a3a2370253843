"""Multi-GPU parity check, launched as one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/mgpu_check.py [cells] [halo depth]

Every rank runs its row block of the drift-control state / adjoint / gradient loops (deep halos, peer-memory halo
exchange, all-rank Jacobi stopping test); rank 0 also runs the same problem on a single-GPU context and compares the
gathered trajectories: the row-wise summation order does not depend on the partition, so fields must agree to the
last bit whenever both sides stop their low-order solves after the same sweeps (always in the default configuration;
a fallback configuration that tests convergence at other sweep counts must agree to 1e-12); the cost functional
(a reduction) to 1e-13; no peer wait may have timed out (fct_p2p_error == 0).
The comparison itself is fem-fct-pdeco_b200/distributed.py:mgpu_parity_check, which bench.py --gpus N also runs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from fem_fct_pdeco_b200.distributed import mgpu_parity_check

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else None
    res = mgpu_parity_check(cells, 4, rank, world, local_rank, depth=depth)
    if rank == 0:
        print("MGPU_CHECK", world, json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()
    if not res["ok"]:
        raise SystemExit(1)
    if rank == 0:
        print("MGPU_CHECK PASSED")


if __name__ == "__main__":
    main()
