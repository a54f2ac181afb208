#!/bin/bash
# round 2, run k: skeleton timings of the CTA-pair kNN kernel
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1 KNN_PROBE_NQ=125000
( for pair in 1 0; do for d in 160 32 8; do
  export SCARF_B200_LIB=$PWD/tools/build/libscarf_b200_dbg$d.so
  echo "== pair=$pair SCF_KNN_DEBUG=$d"
  SCF_KNN_PAIR=$pair timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1 | cut -c1-200
done; done ) 2>&1 | tee gpurun_out/r2_k.log
