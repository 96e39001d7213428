#!/bin/bash
# round-2 GPU check F: new pipeline tests (backward kernels, eval driver, PLY, top-k ties, SH degrees) + the whole suite
set -u
mkdir -p gpurun_out
LOG=gpurun_out/round2_f.log
: > $LOG
: > gpurun_out/parity_report.txt
step() { echo "=== $*" | tee -a $LOG; timeout -k 10 "$@" 2>&1 | tail -${TAILN:-60} | tee -a $LOG; echo "--- exit ${PIPESTATUS[0]}" | tee -a $LOG; }
step 600 python -m pytest tests -q -m gpu --timeout 400
step 200 python -c "import __graft_entry__ as g; g.smoke()"
