#!/bin/bash
# round 2, run y: compute-sanitizer memcheck over the kernels rewritten this round (small cases)
mkdir -p gpurun_out
( timeout 700 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tridiag or jacobi or eig_topk_small or eig_topk_on_device or eig_topk_centred" 2>&1 | tail -8 | cut -c1-300
echo "rc $?"
SCF_KNN_FLAGS=0 timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python tools/knn_probe.py 5000 25 11 3000 100 21 2>&1 | tail -6 | cut -c1-300
) 2>&1 | tee gpurun_out/r2_y.log
