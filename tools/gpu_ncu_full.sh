#!/bin/bash
# ncu --set full of selected kernels of one C2 step (one launch each): NCU_KERNELS = regex, NCU_COUNT = launches
mkdir -p gpurun_out
K=${NCU_KERNELS:-"knn_tc_kernel|project_tc_kernel|gene_stats_win_kernel|gram_tc_kernel|hvg_compact_kernel|hvg_dense_scale_kernel"}
C=${NCU_COUNT:-7}
OUT=${NCU_OUT:-c2_step_full}
timeout ${NCU_TIMEOUT:-900} ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"$K" -c $C -o gpurun_out/$OUT -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --profiler-range > gpurun_out/ncu_step_full.log 2>&1
tail -2 gpurun_out/ncu_step_full.log; ls -la gpurun_out/$OUT.ncu-rep
