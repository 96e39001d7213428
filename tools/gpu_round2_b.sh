#!/bin/bash
# round-2 GPU check B: full -m gpu suite, default bench, sustained probe of the score kernels, ncu capture of the exact kernel
set -u
mkdir -p gpurun_out
LOG=gpurun_out/round2_b.log
: > $LOG
: > gpurun_out/parity_report.txt
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | tail -${TAILN:-30} | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
step 600 python -m pytest tests -q -m gpu --timeout 400 -x
step 400 python bench.py --steps 10 --warmup 3
TAILN=40 step 200 python tools/mq_probe.py --rays 12000000 --seconds 3
step 300 ncu --set full --clock-control none --import-source on -k regex:score_tc_mq -c 4 -o gpurun_out/score_mq_r2 \
    python bench.py --gaussians 200000 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-graph --batch 4
