#!/bin/bash
# round-2 GPU check I: the f16f8 variant (e4m3 cross terms): parity tests, sustained probe, bench
set -u
mkdir -p gpurun_out
LOG=gpurun_out/round2_i.log
: > $LOG
: > gpurun_out/parity_report.txt
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | tail -${TAILN:-40} | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
step 300 python -m pytest tests/test_gpu_exact_tc.py -q --timeout 250 -x -k "f16f8"
step 300 python -m pytest tests/test_gpu_exact_tc.py -q --timeout 250 -k "full_size"
step 200 python tools/mq_probe.py --rays 12000000 --seconds 3 --batches 8,1 --variants f16x2_mq,f16f8_mq
step 400 python bench.py --steps 10 --warmup 3 --score-impl tc_f16f8 --no-secondary --no-cpu-baseline
