#!/bin/bash
# round-2 GPU check D (8 GPUs): c3/c4 (1M Gaussians) in both solve modes, then c5 (5M Gaussians, 32 queries per step)
set -u
mkdir -p gpurun_out
LOG=gpurun_out/round2_d.log
: > $LOG
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" | tail -${TAILN:-12} | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
step 300 $TR bench.py --gpus 8 --steps 20 --warmup 5
step 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --solve weighted_ls --no-secondary
step 400 $TR bench.py --gpus 8 --config c5 --steps 5 --warmup 2
