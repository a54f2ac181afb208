#!/bin/bash
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
for d in ${VARIANTS:-0 8 32 64 96}; do
  if [ $d = 0 ]; then unset SCARF_B200_LIB; else export SCARF_B200_LIB=$PWD/scarf_b200/csrc/build/libscarf_b200_dbg$d.so; fi
  echo "== SCF_KNN_DEBUG=$d"
  timeout 300 python tools/knn_probe.py ${SHAPES:-100000 50 11 400000 100 21} 2>&1 | tail -3
done | tee -a gpurun_out/knn_variants.log
