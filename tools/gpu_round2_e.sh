#!/bin/bash
# round-2 GPU check E: exact tensor-core key build + exact-mode parity incl. the full-size test + default bench
set -u
mkdir -p gpurun_out
LOG=gpurun_out/round2_e.log
: > $LOG
: > gpurun_out/parity_report.txt
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | tail -${TAILN:-40} | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
step 400 python -m pytest tests/test_gpu_exact_tc.py -q --timeout 300
step 400 python bench.py --steps 10 --warmup 3 --no-secondary
