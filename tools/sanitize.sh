#!/bin/bash
# compute-sanitizer pass over the kernel families (run on the GPU box: gpurun -- bash tools/sanitize.sh).
# memcheck on everything; racecheck on the shared-memory heavy SIMT kernels (ray generation table, top-k,
# pose tail).  The tcgen05 kernels are exercised under memcheck only (racecheck does not model TMEM / TMA).
set -u
OUT=gpurun_out/sanitizer_r1.log
: > $OUT
run() { echo "=== $*" | tee -a $OUT; timeout -k 10 200 "$@" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Error|error:|hazard|Invalid" | head -20 | tee -a $OUT; }
K="degrade or quadricell or sym_eig or knn or generate_rays_vs_reference or topk or line_intersection or pose_tail or edge_cases"
run compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$K" --timeout 550 -x
run compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "score_tc_vs_torch and (100-256 or 256-7) or tf32" --timeout 550 -x
run compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "quadricell or topk_against or pose_tail or line_intersection" --timeout 550 -x
