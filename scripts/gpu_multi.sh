#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL parity of the sharded meta-step / HierarchicalRNN step (N >= 2), then the
# weak- and strong-scaling bench lines at N GPUs (and N=1 on the same box for the ratio).
tag=${1:-r02m}
N=${2:-2}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/${tag}_smi.txt 2>&1
if [ "$N" -ge 2 ]; then
  timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_hrnn_gpu.py -q -k "sharded" --timeout 500 > $out/${tag}_dist_tests.log 2>&1
  echo "dist tests rc=$?"; grep -E "passed|failed|meta sharded|hrnn sharded" $out/${tag}_dist_tests.log | tail -5
fi
run() {  # n scaling extra...
  n=$1; sc=$2; shift 2
  if [ "$n" -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --scaling $sc --no-also --no-cpu-baseline "$@"
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $n --steps 10 --warmup 3 --scaling $sc --no-also --no-cpu-baseline "$@"
  fi
}
for sc in weak strong; do
  for n in 1 2 4 8; do
    [ $n -gt $N ] && continue
    run $n $sc > $out/${tag}_${sc}_n$n.json 2> $out/${tag}_${sc}_n$n.err
    python - $out/${tag}_${sc}_n$n.json $sc $n <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "N=%s" % sys.argv[3], "value %.4g ms %.2f" % (d["value"], d["ms_per_step"]), d["config"].get("theta_check"), "coords/gpu", d["config"]["coords_per_gpu"])
except Exception as e:
    print(sys.argv[2], sys.argv[3], "ERR", e)
PY
  done
done
