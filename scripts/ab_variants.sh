mkdir -p gpurun_out
for v in ""; do
  export L2O_LIB=/root/repo/open_l2o_b200/csrc/libl2o_b200$v.so
  echo "=== $L2O_LIB"
  timeout 300 python scripts/tc_accuracy.py 2>&1 | tail -6
  timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('ms/step',round(d['ms_per_step'],2),'fwd',round(r['fwd_kernel']['ms'],2),'bwd',round(r['bwd_ms'],2))
"
done
unset L2O_LIB
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_meta_gpu.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
