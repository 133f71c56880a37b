#!/bin/bash
# quick GPU visit: BPTT timeline, tensor-core tests, headline + mlp + lasso + rnnprop benches (no cpu legs)
tag=${1:-r02f}
out=gpurun_out
mkdir -p $out
timeout 120 build/bin/tc_bwd2_prof_v21 > $out/${tag}_bwd2_prof.txt 2>&1; echo "rc=$?" >> $out/${tag}_bwd2_prof.txt
head -16 $out/${tag}_bwd2_prof.txt | cut -c1-700
timeout 200 python -m pytest tests/test_tc_gpu.py tests/test_parity_configs_gpu.py -x -q --timeout 300 > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -3 $out/${tag}_tests.log
timeout 60 python scripts/tc_accuracy_large.py 1000000 100 2>&1 | tail -1
for w in rastrigin mlp lasso rnnprop_mlp; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-also --no-cpu-baseline > $out/${tag}_b_$w.json 2> $out/${tag}_b_$w.err
  python - $tag $w <<'PY'
import json,sys
d=json.loads(open("gpurun_out/%s_b_%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1]); r=d.get("roofline") or {}
print(sys.argv[2], "value %.4g ms %.2f" % (d["value"], d["ms_per_step"]), "bwd_ms", r.get("bwd_ms"), "fwd_ms", (r.get("fwd_kernel") or {}).get("ms"), "step_us", r.get("step_us"), "bptt_ms", (r.get("bptt") or {}).get("ms"), "e2e %.4g" % d["e2e"]["value"])
PY
done
