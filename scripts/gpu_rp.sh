#!/bin/bash
# RNNProp tensor-core BPTT (two single-chain passes) + step-kernel look-ahead loads: parity tests, then the affected bench lines.
tag=${1:-rp}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_tc_gpu.py -q -x --timeout 300 -k "rnnprop" > $out/${tag}_rp_pytest.log 2>&1; echo "pytest rnnprop rc=$?"; tail -15 $out/${tag}_rp_pytest.log
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_parity_configs_gpu.py tests/test_meta_gpu.py -q --maxfail=5 --timeout 500 > $out/${tag}_tc_pytest.log 2>&1; echo "pytest tc+parity+meta rc=$?"; tail -8 $out/${tag}_tc_pytest.log
for w in rnnprop_mlp mlp; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err; echo "bench $w rc=$?"
  python - $out/${tag}_bench_$w.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("  value %.4g ms %.3f step_us %s bptt_ms %s frac %s" % (d["value"], d["ms_per_step"], r.get("step_us"), (r.get("bptt") or {}).get("ms"), r.get("frac")))
except Exception as e:
    print("ERR", e)
PY
done
