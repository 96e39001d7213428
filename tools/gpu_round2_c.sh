#!/bin/bash
# round-2 GPU check C (2 GPUs): new pipeline tests, then the 2-rank bench in both solve modes
set -u
mkdir -p gpurun_out
LOG=gpurun_out/round2_c.log
: > $LOG
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | tail -${TAILN:-30} | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
step 400 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_exact_tc.py -q --timeout 300 -x -k "not full_size"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
step 400 $TR bench.py --gpus 2 --steps 10 --warmup 3
step 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --solve weighted_ls --no-secondary
