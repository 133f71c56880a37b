#!/bin/bash
# End-of-round GPU visit on the final tree: smoke, full GPU suite, accuracy at scale, both bench arms, launch list + one
# ncu --set full capture of the forward and BPTT kernels.  Outputs under gpurun_out/<tag>_*; copy the summaries to profiles/.
tag=${1:-r02z}
out=gpurun_out
mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log
timeout 700 python -m pytest tests -m gpu --maxfail=10 -q --durations=12 --timeout 150 > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
grep -E 'passed|failed|FAILED|ERROR|rc=' $out/${tag}_pytest.log | tail -12
{ for n in 65536 1000000; do timeout 300 python scripts/tc_accuracy_large.py $n 100; done; timeout 300 python scripts/tc_accuracy.py; } > $out/${tag}_accuracy.txt 2>&1
cat $out/${tag}_accuracy.txt
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench rc=$? capture-warnings: $(grep -c 'capture of the unroll failed' $out/${tag}_bench_n1.err)"
python - $tag <<'PY'
import json,sys
for f in ("bench_n1","bench_ref"):
    try:
        d=json.loads(open("gpurun_out/%s_%s.json" % (sys.argv[1], f)).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, "value %.4g" % d["value"], "ms %.2f" % d.get("ms_per_step",0), "frac", r.get("frac"), "bwd_ms", r.get("bwd_ms"), "fwd_ms", (r.get("fwd_kernel") or {}).get("ms"), "e2e %.4g" % (d.get("e2e") or {}).get("value",0), "cpu", (d.get("cpu_baseline") or {}).get("value"))
        for a in d.get("also") or []:
            rr=a.get("roofline") or {}
            print("   also: %-52s value %.4g ms %.2f step_us %s bptt_ms %s frac %s cpu %s" % (a.get("workload","")[:52], a.get("value",0), a.get("ms_per_step",0), rr.get("step_us"), (rr.get("bptt") or {}).get("ms"), rr.get("frac"), (a.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print(f, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > $out/${tag}_launch_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'unroll_(fwd|bwd2)_kernel' -c 2 -f -o $out/${tag}_tc_fwd_bwd2 \
  python bench.py --steps 1 --warmup 0 --coords 303104 --unroll 20 --no-cpu-baseline --no-also > $out/${tag}_ncu_bench.log 2>&1
echo "ncu rc=$?"
ncu -i $out/${tag}_tc_fwd_bwd2.ncu-rep --page raw --csv > $out/${tag}_tc_fwd_bwd2.raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:'unroll_fwd_kernel' -s 4 -c 1 -f -o $out/${tag}_tc_step \
  python bench.py --workload mlp --steps 1 --warmup 0 --no-cpu-baseline --no-also > $out/${tag}_ncu_step.log 2>&1
echo "ncu step rc=$?"
ncu -i $out/${tag}_tc_step.ncu-rep --page raw --csv > $out/${tag}_tc_step.raw.csv 2>/dev/null
ls -la $out/${tag}_*.csv $out/${tag}_*.ncu-rep 2>/dev/null
