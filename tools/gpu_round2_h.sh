#!/bin/bash
# round-2 GPU check H: whole suite after the single-pass ray generation, sanitizer over the round-2 kernels, default bench
set -u
mkdir -p gpurun_out
LOG=gpurun_out/round2_h.log
: > $LOG
: > gpurun_out/parity_report.txt
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | tail -${TAILN:-40} | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
step 600 python -m pytest tests -q -m gpu --timeout 400
step 900 bash tools/sanitize.sh
step 400 python bench.py --steps 10 --warmup 3
