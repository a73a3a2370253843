# tests/mgpu_check.py on 2 GPUs across the fallback configurations (NCCL instead of peer mailboxes, no CUDA graph, shallow
# halos, fixed launch depth, no tiles): bit-identical where both sides stop after the same sweeps, 1e-12 otherwise.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
port=29620
run() { # cells depth env...
  cells=$1; depth=$2; shift 2
  port=$((port+1))
  echo "== $* : cells $cells, halo depth $depth"
  timeout 300 env "$@" $TR $port tests/mgpu_check.py $cells $depth 2>gpurun_out/mgpu_variants.err | grep MGPU_CHECK | python tools/mgpu_variants_filter.py
}
run 96 8 FCT_NO_P2P=1
run 256 3 FCT_NO_P2P=1
run 256 3 FCT_NO_P2P=1 FCT_NO_GRAPH=1
run 256 3 FCT_HALO_DEPTH=3
run 96 8 FCT_TILE_ADAPT=0
run 96 5 FCT_NO_TILES=1
