#!/bin/bash
# First GPU call for the code that was written without a GPU (gpurun --timeout 900 -- bash tools/validate_experimental.sh):
# every step runs under its own `timeout -k` because a tcgen05 kernel with a barrier mistake hangs instead of failing,
# the validated suite runs first so a regression there is seen independently, and everything is logged to gpurun_out/.
set -u
mkdir -p gpurun_out
LOG=gpurun_out/validate_experimental.log
: > $LOG
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | tail -25 | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
step 300 python -m pytest tests -q -m gpu -x --timeout 280
export SIXDGS_EXPERIMENTAL=1
step 120 python -m pytest tests/test_experimental.py -q -x --timeout 100 -k "multi_query_score_kernel and 100-256"
step 240 python -m pytest tests/test_experimental.py -q --timeout 200
step 120 python tools/mq_probe.py --rays 12000000 --batch 8 --seconds 4
step 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --multi-query
step 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --fused-topk
