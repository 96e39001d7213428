#!/bin/bash
# round-2 GPU check J: ncu --set full of the BATCHED score launches (8 queries per sweep), exact and fast variants
set -u
mkdir -p gpurun_out
LOG=gpurun_out/round2_j.log
: > $LOG
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | grep -v "^==PROF==" | tail -${TAILN:-6} | cut -c1-400 | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
COMMON="--gaussians 200000 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-graph --no-latency --no-breakdown"
step 300 ncu --set full --clock-control none --import-source on -k regex:score_tc_mq -c 2 -o gpurun_out/score_mq8_f16x2_r2 python bench.py $COMMON
step 300 ncu --set full --clock-control none --import-source on -k regex:score_tc_mq -c 2 -o gpurun_out/score_mq8_f16f8_r2 python bench.py $COMMON --score-impl tc_f16f8
