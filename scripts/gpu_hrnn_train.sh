#!/bin/bash
tag=${1:-ht}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests/test_hrnn_train_gpu.py tests/test_hrnn_gpu.py -q -x --timeout 120 > $out/${tag}_hrnn_train.log 2>&1; echo "pytest rc=$?"; tail -40 $out/${tag}_hrnn_train.log
