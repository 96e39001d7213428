#!/bin/bash
# round-2 final check on N GPUs (N = $1): N = 1 runs the whole -m gpu suite, smoke() and the default bench with its
# secondary figures; N > 1 runs the default bench under torchrun
set -u
N=${1:-1}
mkdir -p gpurun_out
LOG=gpurun_out/round2_k_n$N.log
: > $LOG
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" | tail -${TAILN:-12} | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
if [ "$N" = "1" ]; then
  : > gpurun_out/parity_report.txt
  step 600 python -m pytest tests -q -m gpu --timeout 400
  step 200 python -c "import __graft_entry__ as g; g.smoke()"
  step 500 python bench.py --steps 20 --warmup 5
else
  step 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5
fi
