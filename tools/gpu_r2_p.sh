#!/bin/bash
# round 2, run p: hit-buffer depth of the kNN epilogue
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
( for hb in 8 16 24; do
  if [ $hb = 8 ]; then unset SCARF_B200_LIB; else export SCARF_B200_LIB=$PWD/tools/build/libscarf_b200_hb$hb.so; fi
  echo "== HB=$hb: C3/8, C2"
  KNN_PROBE_NQ=125000 timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1 | cut -c1-200
  timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -1 | cut -c1-200
done ) 2>&1 | tee gpurun_out/r2_p.log
