# parity + timing split of the multi-GPU step (see tools/mgpu_diag.py); usage: bash tools/mgpu_diag.sh [nranks]
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
port=29520
for cd in "96 4" "96 8" "1024 4" "1024 8"; do
  port=$((port+1))
  timeout 300 $TR $port tests/mgpu_check.py $cd 2>gpurun_out/mgpu_check.err | grep MGPU_CHECK | tail -n 2
done
run() { # name, cells, env...
  name=$1; cells=$2; shift 2
  port=$((port+1))
  timeout 300 env "$@" $TR $port tools/mgpu_diag.py $cells 10 3 2>gpurun_out/diag_$name.err | tail -n 1 > gpurun_out/diag_$name.json
  cat gpurun_out/diag_$name.json
}
run g4096 4096 FCT_HALO_DEPTH=8
run g2048 2048 FCT_HALO_DEPTH=8
