#!/bin/bash
# multi-rank bench with progress tracing and a tight timeout (N = $1, timeout $2 s)
set -u
N=${1:-2}; TO=${2:-200}
mkdir -p gpurun_out
LOG=gpurun_out/round2_l_n$N.log
export SIXDGS_BENCH_TRACE=1
timeout -k 10 $TO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > $LOG 2>&1
echo "--- exit $?" >> $LOG
grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" $LOG | tail -60 | cut -c1-600
