#!/bin/bash
# First GPU call after this tree: everything DESIGN.md §11 lists as not yet executed on a B200, with tight timeouts.
#   gpurun --timeout 900 -- bash tools/gpu_next_first_call.sh
set -u
mkdir -p gpurun_out
LOG=gpurun_out/next_first_call.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout -k 10 "$@" >> $LOG 2>&1; echo "--- exit $?" >> $LOG; }
# 1. the kernel route of the training-mode MLP: the xfail marks are lifted so a failure shows as a failure
run 300 python -m pytest tests/test_zz_gpu_training_mlp.py -q --runxfail --timeout 250
# 2. the whole GPU suite on the final tree
run 600 python -m pytest tests -q -m gpu -x --timeout 300
# 3. smoke + the default bench line (watchdogs armed)
run 300 python -c "import __graft_entry__ as g; g.smoke()"
run 400 python bench.py
grep -v "^$" $LOG | tail -80 | cut -c1-400
