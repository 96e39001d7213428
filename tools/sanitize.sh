#!/bin/bash
# compute-sanitizer pass over the kernel families (run on the GPU box: gpurun -- bash tools/sanitize.sh).
# memcheck on everything; racecheck on the shared-memory heavy SIMT kernels (ray generation table, top-k, pose tail,
# score backward).  The tcgen05 kernels are exercised under memcheck only (racecheck does not model TMEM / TMA).
set -u
OUT=gpurun_out/sanitizer_r2.log
: > $OUT
run() { echo "=== $*" | tee -a $OUT; timeout -k 10 300 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|error:|hazard|Invalid|done|raygen|kernels" | head -20 | tee -a $OUT; }
run compute-sanitizer --tool memcheck --error-exitcode 0 python tools/sanitize_cases.py
run compute-sanitizer --tool racecheck --error-exitcode 0 python tools/sanitize_cases.py simt
K="degrade or quadricell or sym_eig or knn_normals or generate_rays_vs_reference or line_intersection or pose_tail or edge_cases"
run compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$K" --timeout 280 -x
