#!/bin/bash
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
for f in 2 0; do
  echo "== SCF_KNN_FLAGS=$f"
  SCF_KNN_FLAGS=$f timeout 300 python tools/knn_probe.py 100000 50 11 400000 100 21 2>&1 | tail -2
done | tee gpurun_out/knn_flags.log
