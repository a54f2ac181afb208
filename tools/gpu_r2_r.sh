#!/bin/bash
# round 2, run r: one or two MMA-issuing warps
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or chain" 2>&1 | tail -3 | cut -c1-300
for is in 1 2; do
echo "== issuers=$is: C3/8, C2, C1-like (5k, D 25), C4/8"
SCF_KNN_ISSUERS=$is KNN_PROBE_NQ=125000 timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1 | cut -c1-200
SCF_KNN_ISSUERS=$is timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -1 | cut -c1-200
SCF_KNN_ISSUERS=$is timeout 300 python tools/knn_probe.py 20000 25 11 2>&1 | tail -1 | cut -c1-200
SCF_KNN_ISSUERS=$is KNN_PROBE_NQ=500000 timeout 300 python tools/knn_probe.py 4000000 100 11 2>&1 | tail -1 | cut -c1-200
done
) 2>&1 | tee gpurun_out/r2_r.log
