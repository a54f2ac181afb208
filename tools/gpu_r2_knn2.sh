#!/bin/bash
# where does the kNN step time go: wait back-off on/off, then the timing skeletons (SCF_KNN_DEBUG builds)
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
for f in 2 0; do
  echo "== SCF_KNN_FLAGS=$f"
  SCF_KNN_FLAGS=$f timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -1
  SCF_KNN_FLAGS=$f KNN_PROBE_NQ=125000 timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1
done | tee gpurun_out/r2_knn_flags.log
for d in 8 32 64 96 160; do
  export SCARF_B200_LIB=$PWD/tools/build/libscarf_b200_dbg$d.so
  for f in 2 0; do
    echo "== SCF_KNN_DEBUG=$d SCF_KNN_FLAGS=$f"
    SCF_KNN_FLAGS=$f timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -1
    SCF_KNN_FLAGS=$f KNN_PROBE_NQ=125000 timeout 300 python tools/knn_probe.py 1000000 100 21 2>&1 | tail -1
  done
done | tee -a gpurun_out/r2_knn_flags.log
