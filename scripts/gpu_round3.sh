#!/bin/bash
# GPU visit: BPTT v2.1 timeline + accuracy at scale, full GPU suite, benches.
tag=${1:-r02d}
out=gpurun_out
mkdir -p $out
timeout 120 build/bin/tc_bwd2_prof_v21 > $out/${tag}_bwd2_prof.txt 2>&1; echo "rc=$?" >> $out/${tag}_bwd2_prof.txt
cat $out/${tag}_bwd2_prof.txt | cut -c1-900
timeout 150 python -m pytest tests/test_tc_gpu.py -x -q -k "bwd or bptt" > $out/${tag}_v2_tc.log 2>&1
v2rc=$?; echo "v2 tc tests rc=$v2rc"; tail -3 $out/${tag}_v2_tc.log
if [ $v2rc -ne 0 ]; then export L2O_BWD_V1=1; echo "FALLING BACK TO V1"; fi
{ for n in 65536 1000000; do timeout 300 python scripts/tc_accuracy_large.py $n 100; done; } > $out/${tag}_accuracy.txt 2>&1
cat $out/${tag}_accuracy.txt
timeout 1500 python -m pytest tests -m gpu --maxfail=10 -q --durations=12 --timeout 400 > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
grep -E 'passed|failed|FAILED|ERROR|rc=' $out/${tag}_pytest.log | tail -20
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench rc=$? capture-warnings: $(grep -c 'capture of the unroll failed' $out/${tag}_bench_n1.err)"
timeout 300 python bench.py --workload mlp --steps 10 --warmup 3 --no-also > $out/${tag}_bench_mlp.json 2> $out/${tag}_bench_mlp.err
timeout 300 python bench.py --workload lasso --steps 10 --warmup 3 --no-also > $out/${tag}_bench_lasso.json 2> $out/${tag}_bench_lasso.err
python - $tag <<'PY'
import json,sys
for f in ("bench_n1","bench_mlp","bench_lasso"):
    try:
        d=json.loads(open("gpurun_out/%s_%s.json" % (sys.argv[1], f)).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, "value %.4g" % d["value"], "ms %.2f" % d.get("ms_per_step",0), "frac", r.get("frac"), "bwd_ms", r.get("bwd_ms"), "fwd_ms", (r.get("fwd_kernel") or {}).get("ms"), "e2e %.4g" % (d.get("e2e") or {}).get("value",0), "step_us", r.get("step_us"), "bptt_ms", (r.get("bptt") or {}).get("ms"))
        for a in d.get("also") or []:
            rr=a.get("roofline") or {}
            print("   also: %-52s value %.4g ms %.2f step_us %s bptt_ms %s cpu %s" % (a.get("workload","")[:52], a.get("value",0), a.get("ms_per_step",0), rr.get("step_us"), (rr.get("bptt") or {}).get("ms"), (a.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print(f, "ERR", e)
PY
