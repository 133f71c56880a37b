#!/bin/bash
# quick A/B visit: tensor-core parity tests, accuracy at T=100, headline line (fwd / BPTT ms) and the target-line workload
tag=${1:-q}
out=gpurun_out
mkdir -p $out
timeout 420 python -m pytest tests/test_tc_gpu.py tests/test_meta_gpu.py tests/test_parity_configs_gpu.py -q -x --timeout 90 > $out/${tag}_tc_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/${tag}_tc_pytest.log
timeout 300 python scripts/tc_accuracy.py 2>&1 | tail -5
timeout 240 python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline > $out/${tag}_bench_head.json 2> $out/${tag}_bench_head.err; echo "bench rc=$?"
timeout 240 python bench.py --workload mlp --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_mlp.json 2> $out/${tag}_bench_mlp.err; echo "bench mlp rc=$?"
python - $tag <<'PY'
import json,sys
for f in ("bench_head","bench_mlp"):
    try:
        d=json.loads(open("gpurun_out/%s_%s.json" % (sys.argv[1], f)).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "bwd_ms", r.get("bwd_ms"), "fwd_ms", (r.get("fwd_kernel") or {}).get("ms"), "step_us", r.get("step_us"), "bptt", (r.get("bptt") or {}).get("ms"), "frac", r.get("frac"), "infer", (d.get("infer") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
