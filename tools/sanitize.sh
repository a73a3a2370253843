# compute-sanitizer passes over the GPU test suite on the small meshes (full-size and multi-process tests deselected) and over
# tests/mgpu_check.py when two GPUs are visible; logs -> gpurun_out/sanitize_*.log (summaries are copied to profiles/).
#   bash tools/sanitize.sh [tools...]      default: memcheck racecheck synccheck initcheck
TOOLS=${@:-memcheck racecheck synccheck initcheck}
SEL='not 4096 and not multi_gpu and not 1024'
for tool in $TOOLS; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no"
  timeout 1500 compute-sanitizer --tool $tool $extra --error-exitcode 99 --print-limit 20 --log-file gpurun_out/sanitize_$tool.log \
    python -m pytest tests -m gpu -x -q -k "$SEL" -p no:cacheprovider > gpurun_out/sanitize_$tool.out 2>&1
  echo "$tool rc=$? $(tail -n 1 gpurun_out/sanitize_$tool.out)"
  grep -c "=========     at\|Hazard\|Invalid\|Uninitialized" gpurun_out/sanitize_$tool.log
  tail -n 2 gpurun_out/sanitize_$tool.log
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 99 --print-limit 20 \
      --log-file gpurun_out/sanitize_mgpu_${tool}_%p.log python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
      --master-addr 127.0.0.1 --master-port 29561 tests/mgpu_check.py 64 4 > gpurun_out/sanitize_mgpu_$tool.out 2>&1
    echo "mgpu $tool rc=$? $(grep MGPU_CHECK gpurun_out/sanitize_mgpu_$tool.out | tail -n 1 | cut -c1-60)"
    for f in gpurun_out/sanitize_mgpu_${tool}_*.log; do tail -n 1 $f; done
  done
fi
