#!/bin/bash
# HierarchicalRNN per-parameter kernel: parity tests with the tcgen05 kernel, then step time (tcgen05 vs FFMA) at 32 M coordinates.
tag=${1:-hr}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_hrnn_gpu.py -m gpu -x -q --timeout 300 > $out/${tag}_hrnn_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $out/${tag}_hrnn_pytest.log
timeout 300 python scripts/hrnn_profile.py 2>&1 | tail -2
L2O_HRNN_FFMA=1 timeout 300 python scripts/hrnn_profile.py 2>&1 | tail -2
if [ "$2" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'coord_tc_kernel' -s 3 -c 1 -f -o $out/${tag}_hrnn_coord_tc python scripts/hrnn_profile.py > $out/${tag}_ncu_hrnn.log 2>&1; echo "ncu rc=$?"
ncu -i $out/${tag}_hrnn_coord_tc.ncu-rep --page raw --csv > $out/${tag}_hrnn_coord_tc.raw.csv 2>/dev/null
fi
