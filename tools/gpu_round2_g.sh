#!/bin/bash
# round-2 GPU check G: key-build timing, launch lists (N=1 step; rank-0 share of an 8-rank run), c2 config
set -u
mkdir -p gpurun_out
LOG=gpurun_out/round2_g.log
: > $LOG
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | tail -${TAILN:-30} | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
step 200 python tools/build_probe.py 8000000
COMMON="--steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-secondary --no-latency --no-breakdown"
step 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2_n1.csv python bench.py $COMMON
step 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2_shard8.csv python bench.py $COMMON --emulate-shard 8 --batch 1
step 300 python bench.py --gaussians 100000 --height 400 --width 400 --steps 20 --warmup 5 --no-secondary
