#!/bin/bash
# One GPU-box visit: layout probe, BPTT v2 bring-up checks (short timeouts), GPU parity tests, benches.
#   gpurun --timeout 2700 -- 'bash scripts/gpu_round.sh r02a'
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_smi.txt 2>&1
nproc > $out/${tag}_nproc.txt
# 1. bf16 MN-major SW128 staging probe + SS-mode MMA timing
for n in 48 32; do timeout 60 build/bin/umma_probe_bf16 $n 256 >> $out/${tag}_probe.txt 2>&1; echo "probe N=$n rc=$?" >> $out/${tag}_probe.txt; done
cat $out/${tag}_probe.txt
# 2. BPTT v2 bring-up: the tensor-core BPTT tests alone, hard 150 s limit (a deadlock must not eat the budget)
timeout 150 python -m pytest tests/test_tc_gpu.py -x -q -k "bwd or bptt" > $out/${tag}_v2_tc.log 2>&1
v2rc=$?
echo "v2 tc tests rc=$v2rc" | tee -a $out/${tag}_v2_tc.log
tail -15 $out/${tag}_v2_tc.log
if [ $v2rc -ne 0 ]; then
  echo "BPTT v2 failed its tests: rest of this visit runs the first-generation kernel (L2O_BWD_V1=1)"
  export L2O_BWD_V1=1
fi
# 2b. TMA-staged l2o_step (T = 1) bring-up
timeout 150 python -m pytest tests/test_tc_gpu.py -x -q -k "step_operator" > $out/${tag}_stage.log 2>&1
strc=$?
echo "staged step tests rc=$strc" | tee -a $out/${tag}_stage.log
tail -8 $out/${tag}_stage.log
if [ $strc -ne 0 ]; then
  echo "staged l2o_step failed: rest of this visit runs with L2O_STEP_STAGE=0"
  export L2O_STEP_STAGE=0
fi
# 2c. BPTT v2 timeline
timeout 120 build/bin/tc_bwd2_prof > $out/${tag}_bwd2_prof.txt 2>&1; echo "prof rc=$?" >> $out/${tag}_bwd2_prof.txt
cat $out/${tag}_bwd2_prof.txt
# 3. full GPU suite
timeout 1500 python -m pytest tests -m gpu --maxfail=10 -q --durations=15 --timeout 400 > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
grep -E 'passed|failed|FAILED|ERROR' $out/${tag}_pytest.log | tail -30
# 4. benches
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench rc=$?"
L2O_BWD_V1=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline > $out/${tag}_bench_n1_v1.json 2> $out/${tag}_bench_n1_v1.err
echo "bench v1 rc=$?"
timeout 300 python bench.py --workload mlp --steps 10 --warmup 3 --no-also > $out/${tag}_bench_mlp.json 2> $out/${tag}_bench_mlp.err
echo "bench mlp rc=$?"
L2O_STEP_STAGE=0 L2O_BWD_V1=1 timeout 300 python bench.py --workload mlp --steps 10 --warmup 3 --no-also --no-cpu-baseline > $out/${tag}_bench_mlp_v1.json 2> $out/${tag}_bench_mlp_v1.err
echo "bench mlp (v1 kernels) rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
echo "bench ref rc=$?"
python - $tag <<'PY'
import json,sys
for f in ("bench_n1","bench_n1_v1","bench_mlp","bench_mlp_v1","bench_ref"):
    try:
        d=json.loads(open("gpurun_out/%s_%s.json" % (sys.argv[1], f)).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, "value %.4g" % d["value"], "ms %.2f" % d.get("ms_per_step",0), "frac", r.get("frac"), "bwd_ms", r.get("bwd_ms"), "e2e", (d.get("e2e") or {}).get("value"), "step_us", r.get("step_us"), "bptt_ms", (r.get("bptt") or {}).get("ms"))
        for a in d.get("also") or []:
            print("   also:", a.get("workload","")[:50], "value %.4g" % a.get("value",0), "ms %.2f" % a.get("ms_per_step",0), "err" if "error" in a else "", (a.get("roofline") or {}).get("frac"), (a.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
