#!/bin/bash
# GPU visit: BPTT v2.1 in the library: tc tests, accuracy at scale, full suite, benches, then ONE ncu --set full capture.
tag=${1:-r02e}
out=gpurun_out
mkdir -p $out
timeout 150 python -m pytest tests/test_tc_gpu.py -x -q -k "bwd or bptt" > $out/${tag}_v2_tc.log 2>&1
v2rc=$?; echo "v2 tc tests rc=$v2rc"; tail -3 $out/${tag}_v2_tc.log
if [ $v2rc -ne 0 ]; then export L2O_BWD_V1=1; echo "FALLING BACK TO V1"; fi
{ for n in 65536 1000000; do timeout 300 python scripts/tc_accuracy_large.py $n 100; done; } > $out/${tag}_accuracy.txt 2>&1
cat $out/${tag}_accuracy.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline > $out/${tag}_bench_quick.json 2> $out/${tag}_bench_quick.err
python - $tag <<'PY'
import json,sys
d=json.loads(open("gpurun_out/%s_bench_quick.json" % sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print("QUICK value %.4g ms %.2f bwd_ms %.2f fwd_ms %.2f frac %.4f" % (d["value"], d["ms_per_step"], r["bwd_ms"], r["fwd_kernel"]["ms"], r["frac"]))
PY
timeout 1500 python -m pytest tests -m gpu --maxfail=10 -q --durations=12 --timeout 400 > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
grep -E 'passed|failed|FAILED|ERROR|rc=' $out/${tag}_pytest.log | tail -20
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench rc=$? capture-warnings: $(grep -c 'capture of the unroll failed' $out/${tag}_bench_n1.err)"
python - $tag <<'PY'
import json,sys
for f in ("bench_n1",):
    try:
        d=json.loads(open("gpurun_out/%s_%s.json" % (sys.argv[1], f)).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, "value %.4g" % d["value"], "ms %.2f" % d.get("ms_per_step",0), "frac", r.get("frac"), "bwd_ms", r.get("bwd_ms"), "fwd_ms", (r.get("fwd_kernel") or {}).get("ms"), "e2e %.4g" % (d.get("e2e") or {}).get("value",0))
        for a in d.get("also") or []:
            rr=a.get("roofline") or {}
            print("   also: %-52s value %.4g ms %.2f step_us %s bptt_ms %s cpu %s" % (a.get("workload","")[:52], a.get("value",0), a.get("ms_per_step",0), rr.get("step_us"), (rr.get("bptt") or {}).get("ms"), (a.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print(f, "ERR", e)
PY
# ncu: one full capture of the BPTT + forward kernels (303,104 coords x T=20 like round 1's captures)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'unroll_(fwd|bwd2)_kernel' -c 2 -f -o $out/${tag}_tc_fwd_bwd2 \
  python bench.py --steps 1 --warmup 0 --coords 303104 --unroll 20 --no-cpu-baseline --no-also > $out/${tag}_ncu_bench.log 2>&1
echo "ncu rc=$?"
ncu -i $out/${tag}_tc_fwd_bwd2.ncu-rep --page raw --csv > $out/${tag}_tc_fwd_bwd2.raw.csv 2>/dev/null
ls -la $out/${tag}_tc_fwd_bwd2.*
