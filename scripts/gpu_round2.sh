#!/bin/bash
# Diagnostic GPU visit: BPTT v2 timelines (lazy vs eager dW issue), accuracy at scale (v1 / v2), failing tests under
# engine variants, per-workload graph-capture check.
tag=${1:-r02c}
out=gpurun_out
mkdir -p $out
for b in tc_bwd2_prof tc_bwd2_prof_eager; do
  timeout 120 build/bin/$b > $out/${tag}_$b.txt 2>&1; echo "rc=$?" >> $out/${tag}_$b.txt
done
head -3 $out/${tag}_tc_bwd2_prof.txt; head -3 $out/${tag}_tc_bwd2_prof_eager.txt
{
  for n in 65536 1000000; do
    timeout 300 python scripts/tc_accuracy_large.py $n 100
    L2O_BWD_V1=1 timeout 300 python scripts/tc_accuracy_large.py $n 100
  done
} > $out/${tag}_accuracy.txt 2>&1
cat $out/${tag}_accuracy.txt
K1="tests/test_parity_configs_gpu.py::test_lasso_full_size_B128_unrolls"
K2="tests/test_parity_configs_gpu.py::test_rnnprop_mlp_784_100_10_training_trajectory"
K3="tests/test_dist_gpu.py"
{
  echo "== lasso_full v1"; L2O_BWD_V1=1 timeout 300 python -m pytest $K1 -x -q 2>&1 | grep -E "passed|failed|assert [0-9]"
  echo "== rnnprop TC_AUTO=0"; L2O_TC_AUTO=0 timeout 400 python -m pytest $K2 -x -q 2>&1 | grep -E "passed|failed|assert [0-9]"
  echo "== rnnprop STAGE=0"; L2O_STEP_STAGE=0 timeout 400 python -m pytest $K2 -x -q 2>&1 | grep -E "passed|failed|assert [0-9]"
  echo "== dist v1"; L2O_BWD_V1=1 timeout 300 python -m pytest $K3 -x -q 2>&1 | grep -E "passed|failed|meta sharded"
  echo "== dist TC_AUTO=0"; L2O_TC_AUTO=0 timeout 300 python -m pytest $K3 -x -q 2>&1 | grep -E "passed|failed|meta sharded"
} > $out/${tag}_variants.txt 2>&1
cat $out/${tag}_variants.txt
for w in rnnprop_mlp lasso quadratic mlp; do
  timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-also --no-cpu-baseline > $out/${tag}_b_$w.json 2> $out/${tag}_b_$w.err
  echo "$w rc=$? capture-warnings: $(grep -c 'capture of the unroll failed' $out/${tag}_b_$w.err)"
  grep -B2 -A6 "capture of the unroll failed" $out/${tag}_b_$w.err | head -30
done
timeout 600 python -m pytest tests/test_kernel_net_gpu.py tests/test_kernels_gpu.py -q -k "kernel_net or lasso_grad or convolutional" --timeout 300 2>&1 | tail -15
