# A/B harness (GPU box): accuracy report + short default bench for the main library and each variant library named
# on the command line (suffixes of open_l2o_b200/csrc/libl2o_b200<suffix>.so), then the tcgen05 parity tests.
mkdir -p gpurun_out
for v in "" "$@"; do
  export L2O_LIB=/root/repo/open_l2o_b200/csrc/libl2o_b200$v.so
  echo "=== $L2O_LIB"
  timeout 300 python scripts/tc_accuracy.py 2>&1 | grep "tc "
  timeout 300 python bench.py --no-cpu-baseline --no-also --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('ms/step',round(d['ms_per_step'],2),'fwd',round(r['fwd_kernel']['ms'],2),'bwd',round(r['bwd_ms'],2))
"
  timeout 600 python -m pytest tests/test_tc_gpu.py -x -q 2>&1 | tail -1
done
