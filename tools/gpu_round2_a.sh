#!/bin/bash
# round-2 GPU check A: exact tensor-core mode parity (+ the multi-query bit-identity tests of the bf16 mode)
set -u
mkdir -p gpurun_out
LOG=gpurun_out/round2_a.log
: > $LOG
: > gpurun_out/parity_report.txt
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | tail -40 | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
step 400 python -m pytest tests/test_gpu_exact_tc.py -q --timeout 300 -s
step 200 python -m pytest tests/test_gpu_pipeline.py -q --timeout 150 -k "multi_query"   # (was tests/test_experimental.py when this ran)
