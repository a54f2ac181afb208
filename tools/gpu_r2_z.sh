#!/bin/bash
# round 2, run z: racecheck (shared-memory hazards) on the eigensolver's small-matrix kernels; memcheck on the CTA-pair
# kernel, the repair path and the HVG / LOWESS kernels
mkdir -p gpurun_out
( timeout 420 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 8 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tridiag or eig_topk_small" 2>&1 | tail -12 | cut -c1-300
SCF_KNN_PAIR=1 SCF_KNN_FLAGS=0 timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python tools/knn_probe.py 3000 100 21 2>&1 | tail -3 | cut -c1-300
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "guard_failures or lowess or fused_hvg" 2>&1 | tail -4 | cut -c1-300
) 2>&1 | tee gpurun_out/r2_z.log
